"""Test double for the CUDA side of a slab (sph-erosion_b200/slabs.py GpuSlabBackend): the same
pack / unpack / step interface in numpy, with the oracle as the step.  It restates
slab.cu's classification rules (k_slab_classify / k_slab_append) so the exchange PROTOCOL -- who sends
what to whom, count swap, sized record swap, halo width -- can run under gloo without a GPU."""
import numpy as np
import torch

from oracle import port

GHOST = 0x40000000
HALO = 2
F = 8


def cell_x(x, gmin_x, cell, gnx):
    v = np.floor((x.astype(np.float32) - np.float32(gmin_x)) / np.float32(cell))
    return np.clip(v, 0, gnx - 1).astype(np.int64)


class NumpySlabBackend:
    def __init__(self, P, G, x0, x1, has_left, has_right, cap, wrap_left=False, wrap_right=False, far_x0=1 << 30):
        self.P, self.G = P, G
        self.x0, self.x1, self.hl, self.hr = x0, x1, has_left, has_right
        self.wl, self.wr, self.far_x0 = wrap_left, wrap_right, far_x0     # ring closure (sphe_slab_ring)
        self.cap = cap
        self.gnx = int(G.dim[0])
        self.pos = np.zeros((0, 3), np.float32); self.vel = np.zeros((0, 3), np.float32)
        self.ids = np.zeros(0, np.int64); self.rho = np.zeros(0, np.float32)
        self.send_l = torch.zeros((cap + 1) * F); self.send_r = torch.zeros((cap + 1) * F)
        self.recv_l = torch.zeros((cap + 1) * F); self.recv_r = torch.zeros((cap + 1) * F)
        self.n_owned = 0
        self.transit = [np.zeros((0, F), np.float32), np.zeros((0, F), np.float32)]   # records heading further left / right
        self.forwarded = 0

    def upload(self, pos, vel, ids):
        self.pos = np.ascontiguousarray(pos, np.float32); self.vel = np.ascontiguousarray(vel, np.float32)
        self.ids = np.asarray(ids, np.int64).copy(); self.n_owned = len(ids)

    def _own(self, cx):
        return ((cx >= self.x0) | (not self.hl)) & ((cx < self.x1) | (not self.hr))

    def _records(self, m):
        r = np.zeros((int(m.sum()), F), np.float32)
        r[:, 0:3] = self.pos[m]; r[:, 4:7] = self.vel[m]
        r[:, 7] = self.ids[m].astype(np.int32).view(np.float32)
        return r

    def pack(self):
        real = (self.ids & GHOST) == 0
        self.pos, self.vel, self.ids = self.pos[real], self.vel[real], self.ids[real]
        cx = cell_x(self.pos[:, 0], self.G.gmin[0], self.G.cell, self.gnx)
        own = self._own(cx)
        far = (cx >= self.far_x0) & self.wl       # clamped from the -x wall to the +x wall: goes to the last slab
        to_l = (far if self.wl else (cx < self.x0 + HALO)) & self.hl
        to_r = (cx >= self.x1 - HALO) & self.hr & (not self.wr) & ~far
        live = own | (~far & (cx >= self.x0 - HALO) & (cx < self.x1 + HALO))
        for d, (buf, m, has) in enumerate(((self.send_l, to_l, self.hl), (self.send_r, to_r, self.hr))):
            r = self._records(m)
            if has and len(self.transit[d]):      # k_slab_forward: records in transit join the buffer of their direction
                r = np.concatenate([r, self.transit[d]]); self.forwarded += len(self.transit[d])
            self.transit[d] = np.zeros((0, F), np.float32)
            assert len(r) <= self.cap
            buf[0] = float(np.array([len(r)], np.int32).view(np.float32)[0])   # header: payload count
            buf[F:F + r.size] = torch.from_numpy(r.reshape(-1))
        self.sent = (int(self.send_l[0:1].numpy().view(np.int32)[0]), int(self.send_r[0:1].numpy().view(np.int32)[0]))
        ids = np.where(own, self.ids, self.ids | GHOST)
        self.pos, self.vel, self.ids = self.pos[live], self.vel[live], ids[live]
        self.kept = (self.pos.copy(), self.vel.copy(), self.ids.copy(), int(own.sum()))

    def unpack(self, buf_l, max_l, buf_r, max_r):
        self.pos, self.vel, self.ids, self.n_owned = self.kept[0], self.kept[1], self.kept[2], self.kept[3]
        got = []
        self.transit = [np.zeros((0, F), np.float32), np.zeros((0, F), np.float32)]
        for side, (buf, mx, has) in enumerate(((buf_l, max_l, self.hl), (buf_r, max_r, self.hr))):
            if not has or buf is None:
                got.append(0); continue
            cnt = int(buf[0:1].numpy().view(np.int32)[0])
            got.append(cnt)
            m = min(cnt, mx)
            if m:
                r = buf[F:F + m * F].numpy().reshape(m, F).copy()
                ids = r[:, 7].copy().view(np.int32).astype(np.int64) & (GHOST - 1)
                cx = cell_x(r[:, 0], self.G.gmin[0], self.G.cell, self.gnx)
                own = self._own(cx)
                # owner further along the direction of travel: hand the record on at the next exchange (k_slab_append)
                onward = ~own & ((cx >= self.x1) & self.hr & (not self.wr) if side == 0 else (cx < self.x0) & self.hl & (not self.wl))
                if onward.any():
                    t = r[onward].copy(); t[:, 7] = ids[onward].astype(np.int32).view(np.float32)
                    self.transit[1 - side] = np.concatenate([self.transit[1 - side], t])
                self.pos = np.concatenate([self.pos, r[:, 0:3]]); self.vel = np.concatenate([self.vel, r[:, 4:7]])
                self.ids = np.concatenate([self.ids, np.where(own, ids, ids | GHOST)])
                self.n_owned += int(own.sum())
        self.last_info = dict(n_total=len(self.ids), n_owned=self.n_owned, to_left=self.sent[0], to_right=self.sent[1],
                              from_left=got[0], from_right=got[1])
        return self.last_info

    def unpack_async(self, buf_l, max_l, buf_r, max_r):
        self._results = getattr(self, "_results", [])
        self._results.append(self.unpack(buf_l, max_l, buf_r, max_r))
        return len(self._results) - 1

    def result(self, ticket, wait=True):
        return self._results[ticket]

    def column_histogram(self, gnx):
        m = (self.ids & GHOST) == 0
        return np.bincount(cell_x(self.pos[m, 0], self.G.gmin[0], self.G.cell, self.gnx), minlength=gnx)

    def reconfigure(self, x0, x1, far_x0):
        self.x0, self.x1 = x0, x1
        if self.wl:
            self.far_x0 = far_x0

    def step(self):
        # the oracle sums neighbours in ascending array order: present the particles in global-id order
        o = np.argsort(self.ids & (GHOST - 1), kind="stable")
        self.pos, self.vel, self.ids = self.pos[o], self.vel[o], self.ids[o]
        S = port.State(self.pos, self.vel)
        port.step_grid(self.P, self.G, S)
        self.pos, self.vel, self.rho = S.pos, S.vel, S.density

    def owned(self):
        m = (self.ids & GHOST) == 0
        return self.ids[m], self.pos[m], self.vel[m], self.rho[m]
