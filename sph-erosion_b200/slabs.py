"""Multi-GPU driver: 1-D spatial slab decomposition along x.

One process per GPU (torchrun).  Each rank owns the global cell columns [x0, x1) of the neighbour
grid; per step

    pack      (CUDA, slab.cu)   drop ghosts, keep owned, fill the left/right send buffers with
                                migrants + the 2-layer halo; record 0 of a buffer = payload count
    records   (P2P)             ONE group of sends/receives of 32-byte records over NCCL
                                (NVLink/NVSwitch); message sizes follow last step's counts, which both
                                ends of a link know, so no count has to reach the host first
    unpack    (CUDA)            received records join the local arrays as owned or ghost; then the
                                step's only host sync (the new particle count sizes the launches)
    step      (CUDA)            the ordinary single-GPU step over owned + ghost particles

There is no collective on the data path: interactions are local, so only x-neighbours talk
(torch.distributed batch_isend_irecv = ncclSend/ncclRecv inside one group).  The reference has no
multi-process code; correctness is "k slabs == 1 GPU" (tests/test_gpu_slabs.py) and the exchange
protocol is covered on CPU with gloo (tests/test_slabs_gloo.py).

The driver is written against two small interfaces so the protocol can be exercised without a GPU:
  backend : pack() / unpack() / step()     -- GpuSlabBackend here, a numpy double in tests
  comm    : swap_records()                 -- TorchComm (nccl on GPUs, gloo on CPU)
"""
import numpy as np

RECORD_FLOATS = 8  # (x, y, z, sediment) (vx, vy, vz, id bits)
HALO = 2


def partition_columns(gnx, world, boundaries_x=None, gmin_x=None, cell=None):
    """Column ranges [x0, x1) per rank.  boundaries_x: world-1 ascending x coordinates (e.g. particle
    count quantiles); default = equal column counts.  Every slab must be at least 2*HALO wide."""
    if boundaries_x is None:
        cuts = [int(round(gnx * r / world)) for r in range(1, world)]
    else:
        cuts = [int(np.floor((np.float32(b) - np.float32(gmin_x)) / np.float32(cell))) for b in boundaries_x]
    edges = [0] + cuts + [gnx]
    out = []
    for r in range(world):
        x0, x1 = edges[r], edges[r + 1]
        if x1 - x0 < 2 * HALO:
            raise ValueError("slab %d = [%d,%d) is narrower than %d columns" % (r, x0, x1, 2 * HALO))
        out.append((x0, x1))
    return out


def balanced_cuts(hist, world, old_cuts=None, max_shift=None, min_width=None):
    """Column cuts that give every slab about the same number of particles ("slab boundaries by
    particle-count quantiles of cell-x, re-cut every K steps").  hist[c] = particles in global cell column c (summed
    over the ranks); returns world-1 ascending cut columns.  Every slab keeps at least 2*HALO columns; with old_cuts
    every cut stays HALO columns inside the span of its two old neighbours (one exchange moves data one slab over, halo
    included; a large imbalance converges over a few re-cuts) and, if given, moves by at
    most max_shift columns.  min_width: columns every slab keeps (default 2*HALO; a shared terrain needs room for the two
    boundary zones of a slab, TerrainWindowShare.min_columns)."""
    hist = np.asarray(hist, np.int64)
    gnx = hist.shape[0]
    mw = max(2 * HALO, int(min_width or 0))
    if gnx < world * mw:
        raise ValueError("%d columns cannot hold %d slabs of at least %d columns" % (gnx, world, mw))
    cum = np.concatenate([[0], np.cumsum(hist)])
    total = int(cum[-1])
    cuts = []
    for r in range(1, world):
        c = int(np.searchsorted(cum, total * r / world, side="left"))
        cuts.append(c)
    if old_cuts is not None:
        # Everything a slab needs right after the re-cut -- its new columns AND its halo -- must already be held by
        # itself or by an x-neighbour, because one exchange only moves data one slab over: every cut stays at least
        # HALO columns inside the span of its two old neighbours.
        edges = [0] + list(old_cuts) + [gnx]
        cuts = [int(min(max(c, edges[i] + HALO), edges[i + 2] - HALO)) for i, c in enumerate(cuts)]
        if max_shift is not None:
            cuts = [int(min(max(c, o - int(max_shift)), o + int(max_shift))) for c, o in zip(cuts, old_cuts)]
    # minimum widths, left to right then right to left
    lo = 0
    for i in range(world - 1):
        cuts[i] = max(cuts[i], lo + mw); lo = cuts[i]
    hi = gnx
    for i in range(world - 2, -1, -1):
        cuts[i] = min(cuts[i], hi - mw); hi = cuts[i]
    return cuts


def rebalance(backend, dist_reduce, rank, world, cols, max_shift=None, min_width=None):
    """Re-cuts the slabs by particle count.  backend: column_histogram(gnx) + reconfigure(x0, x1, far_x0);
    dist_reduce(array) sums an int64 numpy array over the ranks in place.  Call it BETWEEN steps with nothing in
    flight (after drain()); the next exchange migrates the particles that changed owner.  Returns the new column
    ranges.  A run that shares a terrain moves its row windows afterwards (TerrainWindowShare.recut)."""
    gnx = cols[-1][1]
    hist = np.asarray(backend.column_histogram(gnx), np.int64)
    dist_reduce(hist)
    cuts = balanced_cuts(hist, world, [c[0] for c in cols[1:]], max_shift, min_width)
    edges = [0] + cuts + [gnx]
    new_cols = [(edges[r], edges[r + 1]) for r in range(world)]
    backend.reconfigure(new_cols[rank][0], new_cols[rank][1], new_cols[-1][0])
    return new_cols


def rebalance_local(backends, cols, max_shift=None, min_width=None):
    """The same for K slabs driven by one process (LocalPeerGroup / tests): the histogram sum is a plain sum."""
    gnx = cols[-1][1]
    hist = sum(np.asarray(b.column_histogram(gnx), np.int64) for b in backends)
    cuts = balanced_cuts(hist, len(backends), [c[0] for c in cols[1:]], max_shift, min_width)
    edges = [0] + cuts + [gnx]
    new_cols = [(edges[r], edges[r + 1]) for r in range(len(backends))]
    for r, b in enumerate(backends):
        b.reconfigure(new_cols[r][0], new_cols[r][1], new_cols[-1][0])
    return new_cols


def ring_links(rank, world):
    """(left, right, wrap_left, wrap_right) of a slab.  3 or more slabs close into a ring: the reference's box
    response moves a particle that sits exactly on the -x wall to the +x wall (collisionS, fluid_system.h:375-382),
    i.e. from the first slab straight to the last one, so those two are linked as well (include/sphe.h
    sphe_slab_ring; a wrap link carries those particles only, no halo).  With 2 slabs the last slab already is
    the first one's neighbour."""
    if world >= 3:
        return (rank - 1) % world, (rank + 1) % world, rank == 0, rank == world - 1
    return (rank - 1 if rank > 0 else None), (rank + 1 if rank < world - 1 else None), False, False


# --------------------------------------------------------------------------- communication
class TorchComm:
    """x-neighbour P2P over torch.distributed (backend nccl on GPUs, gloo in the CPU tests)."""

    def __init__(self, rank, world):
        import torch.distributed as dist
        self.dist = dist
        self.rank, self.world = rank, world
        self.left, self.right, _, _ = ring_links(rank, world)

    def swap_records(self, send_l, out_l, send_r, out_r, recv_l, in_l, recv_r, in_r):
        """One group of sends/receives.  Buffers are float32 tensors; out_*/in_* are the payload sizes
        (records, header excluded) both ends of a link agreed on; None skips that transfer."""
        d, P = self.dist, self.dist.P2POp
        F = RECORD_FLOATS
        ops = []
        if self.left is not None:
            if out_l is not None: ops.append(P(d.isend, send_l[:(out_l + 1) * F], self.left))
            if in_l is not None: ops.append(P(d.irecv, recv_l[:(in_l + 1) * F], self.left))
        if self.right is not None:
            if out_r is not None: ops.append(P(d.isend, send_r[:(out_r + 1) * F], self.right))
            if in_r is not None: ops.append(P(d.irecv, recv_r[:(in_r + 1) * F], self.right))
        if ops:
            for w in d.batch_isend_irecv(ops):
                w.wait()


# --------------------------------------------------------------------------- GPU backend
class GpuSlabBackend:
    """The CUDA side of a slab: a FluidSystemSPH handle in slab mode + torch-owned exchange buffers
    (cap + 1 records each: record 0 is the header).  The handle runs on torch's current stream so NCCL
    and our kernels are ordered by the stream."""

    def __init__(self, sim, device, cap_records, has_left=True, has_right=True):
        import torch
        self.sim = sim
        self.cap = int(cap_records)
        self.has_left, self.has_right = has_left, has_right
        self.wrap = (False, False)     # set by make_gpu_slab for the first / last slab of a ring
        f32 = dict(dtype=torch.float32, device=device)
        n = (self.cap + 1) * RECORD_FLOATS
        self.send_l = torch.zeros(n, **f32); self.send_r = torch.zeros(n, **f32)
        self.recv_l = torch.zeros(n, **f32); self.recv_r = torch.zeros(n, **f32)
        sim.set_stream(torch.cuda.current_stream(device).cuda_stream)

    def pack(self):
        self.sim.slab_pack(self.send_l.data_ptr(), self.send_r.data_ptr(), self.cap, 2 * self.cap)

    def unpack(self, buf_l, max_l, buf_r, max_r):
        return self.sim.slab_unpack(buf_l.data_ptr() if (self.has_left and buf_l is not None) else None, max_l,
                                    buf_r.data_ptr() if (self.has_right and buf_r is not None) else None, max_r)

    def unpack_async(self, buf_l, max_l, buf_r, max_r):
        """No host sync: returns a ticket; result(ticket) gives the counts once that step has finished."""
        return self.sim.slab_unpack_async(buf_l.data_ptr() if (self.has_left and buf_l is not None) else None, max_l,
                                          buf_r.data_ptr() if (self.has_right and buf_r is not None) else None, max_r)

    def result(self, ticket, wait=True):
        return self.sim.slab_result(ticket, wait)

    def step(self):
        self.sim.Run()

    # ---- load balancing (slabs.rebalance): only calls the driver already makes elsewhere
    def column_histogram(self, gnx):
        """Owned particles per global cell column, counted on the device (sphe_slab_column_histogram): only gnx ints
        cross PCIe, so a re-cut costs a small all-reduce + re-windowing the grid, not a download of the slab."""
        return self.sim.slab_column_histogram(gnx).astype(np.int64)

    def reconfigure(self, x0, x1, far_x0):
        s = self.sim
        s.slab_configure(x0, x1, self.has_left, self.has_right)
        if self.wrap[0] or self.wrap[1]:
            s.slab_ring(self.wrap[0], self.wrap[1], far_x0)


# --------------------------------------------------------------------------- the per-step protocol
def next_size(count, cap, floor=1024):
    """Message size (payload records) both ends of a link derive from the count they both saw last."""
    return int(min(cap, max(floor, count + count // 2 + 2048)))


class SlabDriver:
    """pack -> one P2P group -> unpack -> step.

    lag = 0 (synchronous): unpack ends with the step's only host sync; message sizes follow the previous
    step's counts, which both ends of a link know; a count that outgrew its message makes both ends
    repeat that link at full capacity.

    lag >= 1 (asynchronous, the bench default): NO host sync in the step.  unpack leaves the exact particle
    count on the device and returns a ticket; the counts of step k are only read back at step k + lag
    (by then long finished), and the message sizes of step k derive from the counts of step k - lag --
    still something both ends of a link know, so they keep agreeing without talking.  The first
    `sync_steps` steps run synchronously (the initial distribution migrates particles in bulk).  The host runs
    ahead of the GPU by up to `lag` steps, which hides launch and NCCL enqueue latency.  The price: an
    overflow (halo grew by more than the 50 % + 2048 slack within `lag` steps) is detected after the
    fact and is an error instead of a re-send."""

    def __init__(self, backend, comm, lag=0, sync_steps=4):
        self.b, self.c = backend, comm
        cap = backend.cap
        self.lag = int(lag)
        self.sync_left = int(sync_steps) if lag else 0   # start-up transients (initial migration) run synchronously
        self.out_l = self.out_r = self.in_l = self.in_r = cap
        self.last = None
        self.resends = 0
        self.pending = []   # (ticket, out_l, out_r, in_l, in_r) of steps whose counts are not folded yet

    def _check(self, info, out_l, out_r, in_l, in_r):
        b, c = self.b, self.c
        if max(info["to_left"], info["to_right"], info["from_left"], info["from_right"]) > b.cap:
            raise RuntimeError("slab exchange overflow: %r > capacity %d records" % (info, b.cap))
        redo_l = c.left is not None and (info["to_left"] > out_l or info["from_left"] > in_l)
        redo_r = c.right is not None and (info["to_right"] > out_r or info["from_right"] > in_r)
        return redo_l, redo_r

    def _resize(self, info):
        b = self.b
        self.out_l, self.in_l = next_size(info["to_left"], b.cap), next_size(info["from_left"], b.cap)
        self.out_r, self.in_r = next_size(info["to_right"], b.cap), next_size(info["from_right"], b.cap)

    def exchange(self):
        b, c = self.b, self.c
        if self.lag and self.sync_left <= 0:
            return self._exchange_async()
        self.sync_left -= 1
        b.pack()
        c.swap_records(b.send_l, self.out_l, b.send_r, self.out_r, b.recv_l, self.in_l, b.recv_r, self.in_r)
        info = b.unpack(b.recv_l, self.in_l, b.recv_r, self.in_r)
        redo_l, redo_r = self._check(info, self.out_l, self.out_r, self.in_l, self.in_r)
        if redo_l or redo_r:
            self.resends += 1
            if redo_l: self.out_l = self.in_l = b.cap
            if redo_r: self.out_r = self.in_r = b.cap
            c.swap_records(b.send_l, b.cap if redo_l else None, b.send_r, b.cap if redo_r else None,
                           b.recv_l, b.cap if redo_l else None, b.recv_r, b.cap if redo_r else None)
            info = b.unpack(b.recv_l, self.in_l, b.recv_r, self.in_r)
        self._resize(info)
        self.last = info
        return info

    def _exchange_async(self):
        b, c = self.b, self.c
        # fold the step that is `lag` steps old: its counts size this step's messages on both ends
        if len(self.pending) >= self.lag:
            t, ol, orr, il, ir = self.pending.pop(0)
            info = b.result(t, True)
            redo_l, redo_r = self._check(info, ol, orr, il, ir)
            if redo_l or redo_r:
                raise RuntimeError("slab halo outgrew its message within %d steps (%r vs sizes %d/%d/%d/%d): "
                                   "use lag=0 or more slack" % (self.lag, info, ol, orr, il, ir))
            self._resize(info)
            self.last = info
        b.pack()
        c.swap_records(b.send_l, self.out_l, b.send_r, self.out_r, b.recv_l, self.in_l, b.recv_r, self.in_r)
        t = b.unpack_async(b.recv_l, self.in_l, b.recv_r, self.in_r)
        self.pending.append((t, self.out_l, self.out_r, self.in_l, self.in_r))
        return self.last

    def drain(self):
        """Fold every outstanding ticket (end of a run, before reading state)."""
        while self.pending:
            t, ol, orr, il, ir = self.pending.pop(0)
            info = self.b.result(t, True)
            if any(self._check(info, ol, orr, il, ir)):
                raise RuntimeError("slab halo outgrew its message (%r)" % (info,))
            self._resize(info)
            self.last = info
        return self.last

    def step(self):
        self.exchange()
        self.b.step()


class LocalSlabGroup:
    """K slabs driven by ONE process (all handles on the same GPU): the same backend calls as
    SlabDriver with the P2P replaced by reading the neighbour's send buffer.  Used to test the slab
    kernels on a single GPU.  async_=True uses unpack_async (no host sync between steps)."""

    def __init__(self, backends, async_=False):
        self.bs = list(backends)
        self.async_ = async_
        self.last = []
        self.tickets = []

    def step(self):
        K = len(self.bs)
        for b in self.bs:
            b.pack()
        self.last = []
        for r, b in enumerate(self.bs):
            l, rt, _, _ = ring_links(r, K)
            left = self.bs[l].send_r if l is not None else None
            right = self.bs[rt].send_l if rt is not None else None
            if self.async_:
                self.tickets.append((b, b.unpack_async(left, b.cap, right, b.cap)))
            else:
                self.last.append(b.unpack(left, b.cap, right, b.cap))
        for b in self.bs:
            b.step()

    def drain(self):
        out = [b.result(t, True) for b, t in self.tickets[-len(self.bs):]]
        self.tickets = []
        return out


# --------------------------------------------------------------------------- peer-memory exchange (the GPU default)
def device_int32_view(ptr, n, device):
    """Zero-copy torch view of an int32 device array owned by the C library (CUDA array interface)."""
    import torch

    class _Arr:
        pass
    a = _Arr()
    a.__cuda_array_interface__ = {"shape": (int(n),), "typestr": "<i4", "data": (int(ptr), False), "version": 2}
    return torch.as_tensor(a, device=device)


class TerrainShare:
    """Several ranks eroding replicas of ONE terrain: the integer accumulators are summed over the ranks
    between the phases of a step (sphe_step_phase), so every replica stays bit-identical to a single-GPU run.
    reduce = callable(tensor) performing the in-place sum (dist.all_reduce on GPUs)."""

    def __init__(self, grid, device, reduce):
        w, d, n = grid.accumulators()
        self.grid = grid
        self.want = device_int32_view(w, n, device)
        self.delta = device_int32_view(d, n, device)
        self.reduce = reduce

    def step(self, sim):
        if not self.grid.erosion.enabled:
            sim.Run(self.grid)
            return
        sim.step_phase(self.grid, 0)
        self.reduce(self.want)
        sim.step_phase(self.grid, 1)
        self.reduce(self.delta)
        sim.step_phase(self.grid, 2)


def terrain_row_cuts(grid_info, cols, origin_x, scale):
    """Terrain rows at the slab boundaries: rows map to x (H(x, z), grid.h:104-107), slab r owns the neighbour-grid
    columns cols[r] = [x0, x1), i.e. x in [gmin_x + x0*cell, gmin_x + x1*cell).  Returns the world-1 cut rows."""
    return [int(round((grid_info.gmin[0] + c[0] * grid_info.cell - origin_x) / scale)) for c in cols[1:]]


def terrain_margin_rows(cell, scale):
    """Rows an owned particle can touch beyond its slab's own rows within one step: HALO + 1 neighbour-grid columns of
    travel (a particle moving further than the halo per step breaks the slab exchange long before) + the 4 x 4 vertex
    neighbourhood of the contact search.  Checked at run time: sphe_terrain_window_violations must stay 0."""
    return int(np.ceil((HALO + 1) * cell / scale)) + 4


class TerrainWindowShare:
    """Slab-local terrain (the multi-GPU default).  Every rank holds the whole heightfield array but keeps only the
    rows under its slab + a margin of W rows current (sphe_terrain_set_window): the terrain kernels of a rank touch
    that window only, and instead of all-reducing the whole accumulator arrays (33 MB each at 8 x 1024 rows) the
    rank sums the 2W rows around each slab boundary with the ONE neighbour that can also touch them -- two small
    NCCL P2P groups per step.  Integer sums, so the result is bit-identical to the single-GPU run on every row a
    rank owns.  swap = callable(array, zone_left, zone_right) adding the neighbours' rows in place."""

    def __init__(self, grid, device, rank, world, cuts, margin, swap=None, dist=None, peer=False):
        w, d, n = grid.accumulators()
        rows, cols = grid.shape()
        self.grid, self.rank, self.world, self.device = grid, rank, world, device
        self.want = device_int32_view(w, n, device)
        self.delta = device_int32_view(d, n, device)
        self.W = int(margin)
        self._place(cuts)
        self.zone_bytes = 2 * self.W * cols * 4
        self.swap = swap      # False: the caller sums the zones itself (LocalPeerGroup)
        # peer = True: the zones travel through the slab mailboxes (sphe_slab_zone_sum: remote stores over NVLink + device
        # flags, one launch per sum) instead of two NCCL P2P groups per step; PeerSlabDriver reserves the room
        self.peer = bool(peer)
        if self.peer:
            self.swap = None
        elif swap is None:
            import torch
            self.dist = dist
            self.buf_l = torch.empty(2 * self.W * cols, dtype=torch.int32, device=device) if rank > 0 else None
            self.buf_r = torch.empty(2 * self.W * cols, dtype=torch.int32, device=device) if rank < world - 1 else None
            self.swap = self._swap_p2p

    def _place(self, cuts):
        """Owned rows, window and boundary zones for the cut rows `cuts` (world - 1 ascending rows)."""
        rows, cols = self.grid.shape()
        rank, world, W = self.rank, self.world, self.W
        if world > 1 and min(b - a for a, b in zip([0] + list(cuts), list(cuts) + [rows])) < 2 * W:
            raise ValueError("a slab covers fewer than %d terrain rows: the boundary zones of its two neighbours would overlap" % (2 * W))
        self.cuts = [int(c) for c in cuts]
        self.own = (0 if rank == 0 else self.cuts[rank - 1], rows if rank == world - 1 else self.cuts[rank])
        self.window = (max(self.own[0] - W, 0), min(self.own[1] + W, rows))
        self.grid.set_window(*self.window)
        # the rows both this rank and a neighbour can touch: [cut - W, cut + W) around each interior boundary
        self.zone_l = slice((self.own[0] - W) * cols, (self.own[0] + W) * cols) if rank > 0 else None
        self.zone_r = slice((self.own[1] - W) * cols, (self.own[1] + W) * cols) if rank < world - 1 else None

    def bind_columns(self, grid_info, origin_x, scale):
        """Remember how neighbour-grid columns map to terrain rows, so a re-cut of the slabs can move the windows."""
        self._map = (grid_info, float(origin_x), float(scale))
        return self

    def rows_of(self, cols):
        gi, ox, sc = self._map
        return terrain_row_cuts(gi, cols, ox, sc)

    def min_columns(self):
        """Neighbour-grid columns a slab must keep so that it covers its two boundary zones (2 W rows) with room to spare."""
        gi, ox, sc = self._map
        return int(np.ceil((2 * self.W + 2) * sc / gi.cell)) + 1

    # Re-cut (between two steps, nothing in flight).  A rank's replica is current only inside its window, so before the
    # windows move every row is brought up to date from its OWNER: each rank contributes its owned rows, zero elsewhere,
    # and the integer sum over the ranks is the whole current heightfield (rare: once per re-cut, 4 bytes per vertex).
    def recut_owned_rows(self):
        """This rank's contribution to the sum: its heights on the rows it owns, 0 elsewhere (a new device tensor)."""
        rows, cols = self.grid.shape()
        h = device_int32_view(self.grid.heights_device()[0], rows * cols, self.device)
        out = h.clone()
        out[:self.own[0] * cols] = 0
        out[self.own[1] * cols:] = 0
        return out

    def recut_install(self, full, cuts):
        """full = the sum of every rank's recut_owned_rows(); installs it, moves the window to `cuts` and rebuilds the cull map."""
        rows, cols = self.grid.shape()
        h = device_int32_view(self.grid.heights_device()[0], rows * cols, self.device)
        h.copy_(full)
        self._place(cuts)
        self.grid.refresh()

    def recut(self, cuts, reduce):
        """reduce(tensor): in-place sum over the ranks (dist.all_reduce)."""
        import torch
        t = self.recut_owned_rows()
        reduce(t)
        torch.cuda.synchronize()
        self.recut_install(t, cuts)

    @staticmethod
    def recut_local(shares, cuts):
        """The same for K replicas driven by one process (LocalPeerGroup / tests)."""
        parts = [sh.recut_owned_rows() for sh in shares]
        full = parts[0]
        for p in parts[1:]:
            full = full + p
        for sh in shares:
            sh.recut_install(full, cuts)

    def _swap_p2p(self, arr, zl, zr):
        d, P = self.dist, self.dist.P2POp
        ops = []
        if zl is not None:
            ops += [P(d.isend, arr[zl], self.rank - 1), P(d.irecv, self.buf_l, self.rank - 1)]
        if zr is not None:
            ops += [P(d.isend, arr[zr], self.rank + 1), P(d.irecv, self.buf_r, self.rank + 1)]
        if ops:
            for w in d.batch_isend_irecv(ops):
                w.wait()
        if zl is not None: arr[zl] += self.buf_l
        if zr is not None: arr[zr] += self.buf_r

    def step(self, sim):
        if not self.grid.erosion.enabled:
            sim.Run(self.grid)
            return
        sim.step_phase(self.grid, 0)
        self.sum_zones(sim, 0)
        sim.step_phase(self.grid, 1)
        self.sum_zones(sim, 1)
        sim.step_phase(self.grid, 2)

    def sum_zones(self, sim, which):
        """Adds the x-neighbours' copies of the boundary zones into `want` (which = 0) or `delta` (1)."""
        if self.peer:
            sim.slab_zone_sum(self.grid, which, self.zone_l.start if self.zone_l is not None else -1,
                              self.zone_r.start if self.zone_r is not None else -1, self.zone_bytes // 4)
        else:
            self.swap(self.delta if which else self.want, self.zone_l, self.zone_r)

    def own_total_fx(self):
        return self.grid.total_fx(self.own)


class PeerSlabDriver:
    """send -> recv -> step with NO transport library and NO host sync on the data path: the pack kernel
    stores migrants + halo straight into the neighbours' mailboxes over NVLink (peer memory mapped through
    CUDA IPC), the append kernel waits on the device for its mailbox flags (include/sphe.h "Peer-memory
    exchange").  torch.distributed is only used once, to hand the 64-byte mailbox handles to the
    neighbours (and for the terrain sums when a terrain is shared)."""

    def __init__(self, sim, rank, world, cap_records, reserve_particles, terrain=None):
        self.sim, self.rank, self.world = sim, rank, world
        self.cap = int(cap_records)
        self.terrain = terrain
        self.tickets = []
        self.last = None
        if getattr(terrain, "peer", False):
            sim.slab_peer_setup_zones(self.cap, int(reserve_particles), terrain.zone_bytes // 4)
        else:
            sim.slab_peer_setup(self.cap, int(reserve_particles))

    def connect(self, dist):
        """Exchange the mailbox handles (all ranks call this; it also orders every setup before any send)."""
        handles = [None] * self.world
        dist.all_gather_object(handles, self.sim.slab_peer_handle())
        l, r, _, _ = ring_links(self.rank, self.world)
        left = handles[l] if l is not None else None
        right = handles[r] if r is not None else None
        self.sim.slab_peer_connect(left, right)
        dist.barrier()

    def exchange(self):
        self.sim.slab_send()
        self.tickets.append(self.sim.slab_recv())
        if len(self.tickets) > 4:
            self.tickets.pop(0)

    def step(self):
        self.exchange()
        if self.terrain is not None:
            self.terrain.step(self.sim)
        else:
            self.sim.Run()

    def drain(self):
        """Blocks until the last exchange has finished; returns its counts (and raises on overflow/timeout)."""
        if self.tickets:
            self.last = self.sim.slab_result(self.tickets[-1], True)
            self.tickets = []
        return self.last

    def rebalance(self, dist, cols, backend, max_shift=None):
        """Re-cut the slabs by particle count (slabs.rebalance) between two steps; every rank calls it.  backend: the
        GpuSlabBackend make_gpu_slab returned for this sim.  With a slab-local terrain (TerrainWindowShare bound to the
        columns with bind_columns) the row windows move with the cuts.  Returns the new column ranges."""
        import torch
        t = self.terrain
        if t is not None and not hasattr(t, "_map"):
            raise RuntimeError("re-cutting a run that shares a terrain needs TerrainWindowShare.bind_columns(grid_info, origin_x, scale)")
        self.drain()
        on_gpu = dist.get_backend() == "nccl"

        def reduce(hist):
            x = torch.from_numpy(hist).to(torch.device("cuda", torch.cuda.current_device())) if on_gpu else torch.from_numpy(hist)
            dist.all_reduce(x)
            hist[:] = x.cpu().numpy()

        new_cols = rebalance(backend, reduce, self.rank, self.world, cols, max_shift, t.min_columns() if t is not None else None)
        if t is not None:
            t.recut(t.rows_of(new_cols), lambda x: dist.all_reduce(x))
        return new_cols


class LocalPeerGroup:
    """K slabs in ONE process on one GPU, exchanging through each other's mailboxes (connect_local): the
    same kernels and calls as PeerSlabDriver.  grid: one terrain shared by all slabs -- every slab runs
    phase 0, then every slab phase 1, then one phase 2, which is the K-rank sum without a collective."""

    def __init__(self, sims, cap_records, reserve_particles, grid=None, shares=None):
        """shares: one TerrainWindowShare per slab (each slab its own terrain replica, slab-local windows); their
        boundary-zone sums are done here with plain tensor adds between the replicas."""
        self.sims = list(sims)
        self.grid = grid
        self.shares = shares
        for s in self.sims:
            s.slab_peer_setup(cap_records, reserve_particles)
        K = len(self.sims)
        for r, s in enumerate(self.sims):
            l, rt, _, _ = ring_links(r, K)
            s.slab_peer_connect_local(self.sims[l] if l is not None else None, self.sims[rt] if rt is not None else None)
        self.tickets = []

    def step(self):
        for s in self.sims:
            s.slab_send()
        self.tickets = [s.slab_recv() for s in self.sims]
        g = self.grid
        if self.shares is not None:
            sh = self.shares
            for name, phase in (("want", 0), ("delta", 1)):
                for s, t in zip(self.sims, sh):
                    s.step_phase(t.grid, phase)
                for a, b in zip(sh[:-1], sh[1:]):     # the zone right of a == the zone left of b (same rows)
                    x, y = getattr(a, name)[a.zone_r], getattr(b, name)[b.zone_l]
                    tot = x + y
                    x.copy_(tot); y.copy_(tot)
            for s, t in zip(self.sims, sh):
                s.step_phase(t.grid, 2)
        elif g is not None and g.erosion.enabled:
            for s in self.sims:
                s.step_phase(g, 0)
            for s in self.sims:
                s.step_phase(g, 1)
            self.sims[0].step_phase(g, 2)
        else:
            for s in self.sims:
                s.Run(g)

    def drain(self):
        return [s.slab_result(t, True) for s, t in zip(self.sims, self.tickets)]


# --------------------------------------------------------------------------- scenes
def channel_block(n_axis, world, rank, jitter, spacing=0.025, layout="tiled"):
    """Rank `rank`'s share of the weak-scaling scene: `world` copies of the single-GPU scene (bench.scaled_dam_break:
    a block of n_axis^3 particles at the left end of a box of half-extent L = 0.02*n_axis) side by side in a channel
    of half-extents (world*L, L, L) without the walls in between, each block centred in its compartment, so every
    rank sees the same collapse (the blocks spread into the gaps between them) and the per-GPU work really is
    constant -- the compartment boundaries are symmetry planes, a slab neither gains nor loses particles on average.
    layout = "contiguous": see below.
    Returns (pos, ids, box_half, boundaries_x) with global ids = lattice index."""
    L = 0.02 * n_axis
    Lx = L * world
    i = np.arange(n_axis)
    # block r sits in the MIDDLE of the r-th box-sized compartment: the compartment boundaries (and the two end walls)
    # are mirror planes of the initial state, so no compartment gains particles at the expense of another
    x = (-Lx + rank * 2 * L + (L - 0.5 * (n_axis - 1) * spacing) + i * spacing).astype(np.float32)
    if layout == "contiguous":
        # ONE block `world` times as long at the left end of the channel (the single-GPU scene stretched along x): fluid
        # on both sides of every slab boundary from the first step on, i.e. a full halo exchange every step
        x = (-Lx + (rank * n_axis + i) * spacing).astype(np.float32)
    y = (-L / 4 + i * spacing).astype(np.float32)
    z = (-0.75 * L + i * spacing).astype(np.float32)
    pos = np.empty((x.size, y.size, z.size, 3), np.float32)
    pos[..., 0] = x[:, None, None]; pos[..., 1] = y[None, :, None]; pos[..., 2] = z[None, None, :]
    pos = pos.reshape(-1, 3)
    if jitter:
        rng = np.random.default_rng(0x5EED + rank)
        pos += rng.uniform(-0.2 * spacing, 0.2 * spacing, pos.shape).astype(np.float32)
    per = n_axis ** 3
    ids = (rank * per + np.arange(per)).astype(np.int32)
    bounds = [-Lx + r * 2 * L - 0.5 * spacing for r in range(1, world)]
    if layout == "contiguous":
        bounds = [-Lx + (r * n_axis - 0.5) * spacing for r in range(1, world)]
    return pos, ids, (Lx, L, L), bounds


def make_gpu_slab(pkg, device, rank, world, box_half, params, bounds_x, cap_records, variant=(6, 3)):
    """Configured FluidSystemSPH handle + backend for rank `rank` of `world` x-slabs."""
    sim = pkg.FluidSystemSPH(device=device)
    p = sim.params
    for k, v in params.items():
        if k == "g":
            p.g[0], p.g[1], p.g[2] = v
        else:
            setattr(p, k, v)
    sim.set_box(box_half)
    sim.set_variant(*variant)
    info = sim.slab_info()  # global grid
    gi = sim.grid_info()
    cols = partition_columns(info["gnx"], world, bounds_x, gi.gmin[0], gi.cell)
    x0, x1 = cols[rank]
    left, right, wrap_l, wrap_r = ring_links(rank, world)
    sim.slab_configure(x0, x1, left is not None, right is not None)
    if wrap_l or wrap_r:
        sim.slab_ring(wrap_l, wrap_r, cols[-1][0])
    backend = GpuSlabBackend(sim, device, cap_records, left is not None, right is not None)
    backend.wrap = (wrap_l, wrap_r)
    return sim, backend, cols


def parity_gate_multi(pkg, dist, dev, sim, drv, grid, tshare, windowed, box, params, variant, attach, world, rank):
    """In-bench parity at FULL size for N > 1: the state the timed region left is gathered on rank 0, every rank takes
    one more step through the ordinary exchange, and rank 0 takes the SAME step with ONE handle holding the whole scene
    (all N slabs' particles, the whole terrain) on its own GPU.  Bar: the N-slab result equals the single-domain result
    BIT FOR BIT (positions, velocities, densities, carried sediment per particle, every terrain row a rank owns); a
    particle that crossed more than one slab in a step is forwarded a step late and may differ -- those are counted,
    everything else must be exact.  The single-domain path itself is held to the oracle by the N = 1 gate
    (bench.parity_gate).  `attach` builds a fresh terrain replica."""
    import time
    t0 = time.perf_counter()

    def snapshot():
        drv.drain()
        ids, pos, vel, rho, sed = sim.slab_download()
        rows = None
        if grid is not None:
            h = grid.heights_fx()
            own = tshare.own if windowed else ((0, h.shape[0]) if rank == 0 else (0, 0))
            rows = (own, h[own[0]:own[1]].copy())
        out = [None] * world if rank == 0 else None
        dist.gather_object((ids, pos, vel, rho, sed.view(np.int32), rows), out, dst=0)
        # rank 0 is still unpickling ~1 GB at N = 8 when the others are done sending: nobody may enter the next exchange
        # before it has, or its neighbours' device-side flag waits run into their timeout
        dist.barrier()
        return out

    before = snapshot()
    drv.step()
    after = snapshot()
    tr = sim.slab_transit()
    moving = [None] * world if rank == 0 else None
    dist.gather_object(tr["to_left"] + tr["to_right"], moving, dst=0)
    if rank != 0:
        return None

    def by_id(parts, k, n):
        ids = np.concatenate([p[0] for p in parts])
        a = np.concatenate([p[k] for p in parts])
        out = np.zeros((n,) + a.shape[1:], a.dtype)
        out[ids] = a
        return ids, out

    ids0 = np.concatenate([p[0] for p in before])
    n = int(ids0.shape[0])
    res = {"particles": n, "reference": "one handle holding all %d slabs' particles and the whole terrain, same step, on rank 0's GPU" % world}
    res["every_particle_owned_once"] = bool(n == int(ids0.max()) + 1 and np.array_equal(np.sort(ids0), np.arange(n)))
    if not res["every_particle_owned_once"]:
        return False, res
    _, pos0 = by_id(before, 1, n); _, vel0 = by_id(before, 2, n); _, sed0 = by_id(before, 4, n)
    one = pkg.FluidSystemSPH(device=dev.index)
    for k, v in params.items():
        if k == "g":
            one.params.g[0], one.params.g[1], one.params.g[2] = v
        else:
            setattr(one.params, k, v)
    one.set_box(box); one.set_variant(*variant)
    one.upload_state(pos0, vel0)
    g1 = None
    if grid is not None:
        g1 = attach()
        rows_total = g1.shape()[0]
        hfx = np.zeros((rows_total, g1.shape()[1]), np.int32)
        for p in before:
            (a, b), r = p[5]
            hfx[a:b] = r
        g1.set_heights(hfx.astype(np.float32) / np.float32(4096.0))      # exact: |fx| < 2^24
        assert np.array_equal(g1.heights_fx(), hfx)
        one.set_sediment_fx(sed0)
    one.Run(g1)
    exact = {}
    for name, k in (("pos", 1), ("vel", 2), ("density", 3)):
        _, got = by_id(after, k, n)
        want = one.download(name)
        same = got.view(np.uint32).reshape(n, -1) == want.view(np.uint32).reshape(n, -1)
        exact[name] = int((~same.all(axis=1)).sum())
        scale = max(float(np.abs(want).max()), 1e-30)
        res[name + "_max_err_over_scale"] = float(np.abs(got.astype(np.float64) - want).max() / scale)
    res["particles_not_bit_equal"] = exact
    res["records_forwarded_late"] = int(sum(moving))
    worst = max(exact.values())
    ok = worst <= 2 * res["records_forwarded_late"]
    if grid is not None:
        _, seda = by_id(after, 4, n)
        want = np.rint(one.download("sediment").astype(np.float64) * 4096.0).astype(np.int32)
        res["sediment_particles_differing"] = int((seda != want).sum())
        h1 = g1.heights_fx()
        bad = 0
        for p in after:
            (a, b), r = p[5]
            bad += int((h1[a:b] != r).sum())
        res["terrain_vertices_differing"] = bad
        ok = ok and bad == 0 and res["sediment_particles_differing"] <= 2 * res["records_forwarded_late"]
    res["seconds"] = round(time.perf_counter() - t0, 2)
    return bool(ok), res


# --------------------------------------------------------------------------- bench (called by bench.py)
def bench_multi(args, pkg, n_axis, jitter, desc, METRIC, UNIT, terrain=False):
    import json
    import time
    import torch
    import torch.distributed as dist
    from bench import ClockSampler, measured_peak, scene_gravity, attach_terrain, emit, near_gpu, ALGO_BYTES, SPACING

    rank, world = dist.get_rank(), dist.get_world_size()
    local = torch.cuda.current_device()
    dev = torch.device("cuda", local)
    # no settle phase (c2): one long block, so the timed steps carry a full halo from step one.  With a settle phase (c3)
    # a long block compresses far beyond the single-GPU scene; tiled compartments keep the per-GPU work constant and
    # the settle phase brings the fluid across every slab boundary before the timed steps.
    layout = args.layout if args.layout != "auto" else ("tiled" if terrain else "contiguous")
    pos, ids, box, bounds = channel_block(n_axis, world, rank, jitter, layout=layout)
    gy = scene_gravity(n_axis, args.gravity_unscaled)
    n_local = pos.shape[0]
    layer = int(n_axis * n_axis * (0.0457 * 1.001 / SPACING + 1))      # particles per cell layer
    cap = max(2 * HALO * layer, 1 << 14)
    if getattr(args, "rebalance_every", 0):
        cap *= 2      # a re-cut moves a boundary by up to HALO columns: those particles migrate in ONE exchange, on top of the halo
    params = dict(len=box[1], dt=0.01, g=(0.0, gy, 0.0))
    sim, backend, cols = make_gpu_slab(pkg, local, rank, world, box, params, bounds, cap,
                                       (args.density_variant, args.force_variant))
    sim.slab_upload(pos, np.zeros_like(pos), ids)
    grid, tinfo, tshare = None, {}, None
    if terrain:
        # every rank holds a replica of the whole terrain; the integer erosion accumulators are summed over
        # the ranks between the phases of a step (TerrainShare), so the replicas stay bit-identical
        grid, tinfo = attach_terrain(pkg, box[1], n_axis, nx_mult=world)
        if args.terrain_share == "window":
            cuts = terrain_row_cuts(sim.grid_info(), cols, tinfo["terrain_origin"][0], tinfo["terrain_cell"])
            tshare = TerrainWindowShare(grid, dev, rank, world, cuts, terrain_margin_rows(sim.grid_info().cell, tinfo["terrain_cell"]), dist=dist,
                                        peer=(args.exchange == "peer" and getattr(args, "zone_sums", "peer") == "peer"))
            tshare.bind_columns(sim.grid_info(), tinfo["terrain_origin"][0], tinfo["terrain_cell"])
        else:
            tshare = TerrainShare(grid, dev, lambda t: dist.all_reduce(t))
    if args.exchange == "peer":
        drv = PeerSlabDriver(sim, rank, world, cap, int(n_local * 1.3) + 6 * cap, tshare)
        drv.connect(dist)
        exchange_desc = ("peer memory: the pack kernel stores migrants + 2-layer halo straight into the x-neighbours' mailboxes over NVLink "
                         "(CUDA IPC mapping), the append kernel waits on device flags; no transport library, no host sync, no collective")
    else:
        if terrain:
            raise SystemExit("--exchange nccl has no shared-terrain support; use the default peer exchange")
        drv = SlabDriver(backend, TorchComm(rank, world), lag=args.slab_lag)
        exchange_desc = "NCCL P2P (one batch_isend_irecv group per step) with the x-neighbours, counts ride in the record headers, no collective"
    windowed = terrain and args.terrain_share == "window"
    if windowed:
        exchange_desc += ("; terrain slab-local: each rank keeps the rows under its slab + %d margin rows current and sums the erosion accumulators "
                          "of the %d rows around each slab boundary with that x-neighbour (%d KB per sum, int32, no collective; transport: %s)"
                          % (tshare.window[1] - tshare.own[1] if rank < world - 1 else tshare.own[0] - tshare.window[0],
                             2 * (tshare.own[0] - tshare.window[0] if rank else tshare.window[1] - tshare.own[1]), tshare.zone_bytes // 1024,
                             "the slab mailboxes: remote stores over NVLink + device flags, one launch per sum" if tshare.peer else "2 NCCL P2P groups per step"))
    elif terrain:
        exchange_desc += "; terrain replicated, per-vertex erosion accumulators summed with 2 NCCL all-reduces (int32) per step"

    def terrain_total():
        if not windowed:
            return grid.total_fx()
        t = torch.tensor([tshare.own_total_fx()], device=dev, dtype=torch.int64)
        dist.all_reduce(t)
        return int(t.item())

    def sync_all():
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()

    def sediment_all():
        t = torch.tensor([sim.sediment_total_fx()], device=dev, dtype=torch.int64)
        dist.all_reduce(t)
        return int(t.item())

    # re-cuts by particle count: plain slabs, or slabs with a slab-local terrain (the row windows move with the cuts)
    recut = getattr(args, "rebalance_every", 0) if args.exchange == "peer" and (not terrain or args.terrain_share == "window") else 0
    recuts = 0
    if grid is not None:
        tot0 = terrain_total() + sediment_all()
        for k in range(args.settle):
            if recut and k and k % recut == 0:
                cols = drv.rebalance(dist, cols, backend); recuts += 1
            drv.step()
        drv.drain()
        tinfo["settle_steps"] = args.settle
    for _ in range(args.warmup):
        drv.step()
    if grid is not None:
        drv.drain()
        grid.contacts(reset=True)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    e0.record()
    for k in range(args.steps):
        if recut and k and k % recut == 0:
            cols = drv.rebalance(dist, cols, backend); recuts += 1      # inside the timed region: its cost is part of the step budget
        drv.step()
    e1.record()
    sync_all()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    drv.drain()
    if grid is not None:
        c = torch.tensor([grid.contacts()], device=dev, dtype=torch.int64)
        dist.all_reduce(c)
        tinfo["terrain_contacts_per_step"] = int(c.item()) / args.steps
        sed_now = sediment_all()
        tinfo["sediment_in_flight_fx"] = sed_now
        tinfo["conservation_exact"] = bool(terrain_total() + sed_now == tot0)
        hfx = grid.heights_fx()
        if windowed:
            # every row a rank keeps current must agree with the neighbour that keeps it current too
            W = tshare.own[0] - tshare.window[0] if rank else tshare.window[1] - tshare.own[1]
            zones = [None] * world
            dist.all_gather_object(zones, (hfx[tshare.own[0] - W:tshare.own[0] + W] if rank else None,
                                           hfx[tshare.own[1] - W:tshare.own[1] + W] if rank < world - 1 else None))
            tinfo["terrain_boundary_rows_identical"] = bool(all(np.array_equal(zones[r][1], zones[r + 1][0]) for r in range(world - 1)))
            v = torch.tensor([grid.window_violations()], device=dev, dtype=torch.int64)
            dist.all_reduce(v)
            tinfo["terrain_window_violations"] = int(v.item())
            tinfo["terrain_window_rows"] = list(tshare.window)
        else:
            h = torch.from_numpy(hfx.astype(np.int64).ravel()).to(dev)
            hs = torch.stack([h.sum(), (h * torch.arange(1, h.numel() + 1, device=dev) % 1000003).sum()])
            hmin, hmax = hs.clone(), hs.clone()
            dist.all_reduce(hmin, op=dist.ReduceOp.MIN); dist.all_reduce(hmax, op=dist.ReduceOp.MAX)
            tinfo["terrain_replicas_identical"] = bool(torch.equal(hmin, hmax))
    # per-kernel times: instrumented steps AFTER the timed region (an event pair around every launch costs several per
    # cent of a step, so the headline loop above carries none)
    pk_steps = max(1, min(args.steps, 20))
    sim.kernel_timing(True)
    for _ in range(pk_steps):
        drv.step()
    drv.drain()
    sync_all()
    per_kernel, launches = sim.kernel_times()
    per_kernel = {k: v * args.steps / pk_steps for k, v in per_kernel.items()}     # scaled to the timed region's step count
    launches = launches * args.steps // pk_steps
    sim.kernel_timing(False)
    per_rank = [None] * world
    dist.all_gather_object(per_rank, {"particles": sim.slab_info()["n_total"], "owned": sim.slab_info()["n_owned"],
                                      **{k: round(v / args.steps, 4) for k, v in per_kernel.items()}})
    clocks = sampler.stop() if rank == 0 else None
    tr = sim.slab_transit()
    owned = torch.tensor([sim.slab_info()["n_owned"], drv.last["to_left"] + drv.last["to_right"], tr["to_left"] + tr["to_right"], tr["forwarded"]],
                         device=dev, dtype=torch.int64)
    dist.all_reduce(owned, op=dist.ReduceOp.SUM)
    n_total = int(owned[0].item())
    ms_step = float(ms.item()) / args.steps
    value = n_total / (ms_step * 1e-3)

    gate = None
    if not getattr(args, "no_parity_gate", False):
        gate = parity_gate_multi(pkg, dist, dev, sim, drv, grid, tshare, windowed, box, params, (args.density_variant, args.force_variant),
                                 (lambda: attach_terrain(pkg, box[1], n_axis, nx_mult=world)[0]) if terrain else None, world, rank)

    # end to end: every step uploads this rank's slab from pinned host memory, exchanges, steps and
    # downloads the owned particles back to pinned host memory
    e2e_steps = max(3, min(args.steps, 10))
    o_ids, o_pos, o_vel, o_rho, _ = sim.slab_download()
    m = o_ids.shape[0]
    capn = int(m * 1.25) + 1024
    with near_gpu(local) as numa:      # pinned buffers on the NUMA node next to this rank's GPU
        hp = torch.zeros((capn, 3), dtype=torch.float32).pin_memory(); hv = torch.zeros_like(hp).pin_memory()
        hi = torch.zeros(capn, dtype=torch.int32).pin_memory(); hr = torch.zeros(capn, dtype=torch.float32).pin_memory()
    hp[:m] = torch.from_numpy(o_pos); hv[:m] = torch.from_numpy(o_vel); hi[:m] = torch.from_numpy(o_ids)
    h2d = d2h = 0
    sync_all()
    t = time.perf_counter()
    for _ in range(e2e_steps):
        sim.slab_upload_ptr(m, hp.data_ptr(), hv.data_ptr(), hi.data_ptr())
        h2d += 28 * m
        drv.step()
        drv.drain()
        m = sim.slab_download_ptr(capn, hi.data_ptr(), hp.data_ptr(), hv.data_ptr(), hr.data_ptr())
        d2h += 32 * m
    sync_all()
    e2e_dt = torch.tensor([(time.perf_counter() - t) / e2e_steps], device=dev, dtype=torch.float64)
    dist.all_reduce(e2e_dt, op=dist.ReduceOp.MAX)
    io = torch.tensor([h2d // e2e_steps, d2h // e2e_steps], device=dev, dtype=torch.int64)
    dist.all_reduce(io, op=dist.ReduceOp.SUM)
    if rank != 0:
        return
    peak, peak_src = measured_peak()
    dom = max(("density", "force", "terrain"), key=lambda k: per_kernel[k])
    t_dom = per_kernel[dom] / args.steps * 1e-3
    n_rank0 = sim.slab_info()["n_total"]
    achieved = ALGO_BYTES[dom] * n_rank0 / t_dom / 1e9
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": args.workload, "description": desc + " -- repeated %d x along x, one x-slab per GPU" % world,
                       "layout": layout + (": one block %d x as long at the left end of a channel %d x as long" % (world, world) if layout == "contiguous"
                                           else ": %d blocks, each centred in its own box-sized compartment of the channel" % world),
                       "particles": n_total, "particles_per_gpu": n_local, "particles_conserved": bool(n_total + int(owned[2].item()) == n_local * world),
                       "records_forwarded_beyond_the_neighbour": int(owned[3].item()), "h": 0.0457, "spacing": SPACING, "dt": 0.01,
                       "box_half_extents": list(box), "gravity_y": gy, "slab_columns": cols,
                       "halo_records_per_step_all_ranks": int(owned[1].item()),
                       "rebalance_every": recut, "recuts_done": recuts,
                       "exchange": exchange_desc, "exchange_mode": args.exchange,
                       "exchange_resends": getattr(drv, "resends", 0), "exchange_lag": args.slab_lag if args.exchange == "nccl" else None,
                       "l2": "working set per GPU (%.0f MB of particle arrays + neighbour lists) exceeds L2" % (n_local * 400 / 1e6),
                       "density_variant": args.density_variant, "force_variant": args.force_variant, **tinfo},
            "e2e": {"value": n_total / float(e2e_dt.item()), "unit": UNIT, "h2d_bytes_per_step": int(io[0].item()),
                    "d2h_bytes_per_step": int(io[1].item()), "ms_per_step": float(e2e_dt.item()) * 1e3, "steps": e2e_steps,
                    "api": "sphe_slab_upload (pinned host) -> pack/exchange/append -> sphe_step -> sphe_slab_download (pinned host), per rank",
                    "pinned_buffers": ("allocated from the %d CPUs NVML lists as local to the rank's GPU" % len(numa.cpus)) if numa.cpus else "default placement"},
            "parity_sampled": gate[0] if gate else None, "parity_gate": gate[1] if gate else None,
            "gpu_launches": launches * world, "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": "k_%s" % dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
                         "algorithmic_bytes_per_particle": ALGO_BYTES[dom], "rank": 0,
                         "per_kernel_timing": "the %d steps after the timed region, with a CUDA-event pair around every launch (the timed region has none)" % pk_steps,
                         "per_kernel_ms_per_step": {k: v / args.steps for k, v in per_kernel.items()},
                         "per_rank": per_rank},
            "cpu_baseline": None}
    emit(line)
