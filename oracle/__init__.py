"""TEST INFRASTRUCTURE ONLY (oracle + compiled reference wrappers).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this package.  The product path (sph-erosion_b200/) never does.
"""
