import sys, json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    t=d["per_kernel_ms"]; print(d["particles"], d["h"], ("%.1f"%d["mean_neighbours"]) if d["mean_neighbours"] is not None else "-", "cap", d.get("list_capacity"), "ms %.3f density %.3f force %.3f"%(d["ms_per_step"], t["density"], t["force"]))
