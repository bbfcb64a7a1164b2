// Multi-GPU x-slab support (the reference is single-process, nothing to match).
//
// A slab owns the global cell columns cx in [x0, x1).  Before every step it
//   * drops last step's ghosts,
//   * keeps the particles it still owns, keeps particles that just left but are still inside its
//     halo zone as ghosts (the new owner has the authoritative copy),
//   * packs everything a neighbour needs -- migrants AND the HALO outermost cell layers -- into one
//     send buffer per side (one exchange per step carries both).
// Records are 32 bytes: (x, y, z, sediment) (vx, vy, vz, id bits); record 0 of a buffer is a header
// carrying the payload count, so buffers can be sent with a size both sides already agree on and the
// host never has to learn a count before posting the transfer.  The receiver decides owned/ghost from
// the record's own cell column, so sender and receiver never disagree.
//
// HALO = 2 cell layers: the density of a ghost in the first layer is recomputed locally from the
// second layer, so no second exchange of densities is needed (8e option "2-layer halo").
//
// Slot allocation uses warp-aggregated atomics (one atomicAdd per warp per destination).  The
// resulting storage order is arbitrary; the binning pass re-establishes the canonical (cell, id) order.
#include <algorithm>
#include "common.cuh"
#include "sim.h"

namespace sphe {

__device__ __forceinline__ int warp_append(bool flag, int* counter) {
    unsigned m = __ballot_sync(SPHE_FULL, flag);
    if (m == 0) return -1;
    unsigned lane = threadIdx.x & 31;
    int leader = __ffs(m) - 1;
    int base = 0;
    if ((int)lane == leader) base = atomicAdd(counter, __popc(m));
    base = __shfl_sync(SPHE_FULL, base, leader);
    return flag ? base + __popc(m & ((1u << lane) - 1u)) : -1;
}

// Block-aggregated slot allocation for up to 4 output streams at once: every warp ballots its flags,
// ONE thread per block reserves the block's share of each stream with a single atomicAdd, and every
// flagged thread gets base + (flagged threads before it in the block).  Same-address atomics with a
// return value serialise in L2 (~1 ns each): per-warp atomics on 1M particles cost 70 us, per-block 9 us.
// Must be called by all threads of a 256-thread block.  slot[k] = -1 where flag k is not set.
__device__ __forceinline__ void block_append4(const bool flag[4], int* const counter[4], int slot[4]) {
    __shared__ int wsum[4][8];
    __shared__ int bbase[4];
    const unsigned lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    unsigned m[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        m[k] = __ballot_sync(SPHE_FULL, flag[k]);
        if (lane == 0) wsum[k][w] = __popc(m[k]);
    }
    __syncthreads();
    if (threadIdx.x < 4) {
        int k = threadIdx.x, tot = 0;
#pragma unroll
        for (int j = 0; j < 8; j++) { int c = wsum[k][j]; wsum[k][j] = tot; tot += c; }
        bbase[k] = tot ? atomicAdd(counter[k], tot) : 0;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 4; k++) slot[k] = flag[k] ? bbase[k] + wsum[k][w] + __popc(m[k] & ((1u << lane) - 1u)) : -1;
}

// counters: [0] kept (owned + retained ghosts), [1] to left, [2] to right, [3] owned, [8] extent of the arrays (append base)
// IN PLACE: nothing is copied.  A particle that stays keeps its slot; last step's ghosts and particles that left the halo
// zone are marked DEAD (id = -1: the ghost bit is set, so every owned-only pass skips them) and the next binning -- which
// gathers every live particle into the sorted arrays anyway -- drops them (k_hash gives them no cell).  Only the ~5 % of
// the particles in the halo zones are read in full and written, as 32-byte records.  (Round 1 rewrote all arrays here:
// 80 B/particle, 0.04-0.09 ms per step.)
// REMOTE: send_left / send_right point into the NEIGHBOUR GPU's mailbox (peer memory over NVLink, see the
// peer-memory exchange below): the records are stored where they will be consumed, no staging copy and no
// transport call, and every storing thread fences at system scope so the stores are performed at the peer
// before this grid completes and k_slab_publish raises the flag.
#define SPHE_DEAD_ID (-1)
template <bool REMOTE>
__global__ void __launch_bounds__(256) k_slab_classify(int n_hi, const int* __restrict__ n_dev, const float4* __restrict__ posq, const float4* __restrict__ velv,
                                                       int* __restrict__ ids, const float* __restrict__ sed, GridP G, SlabP S,
                                                       float4* __restrict__ send_left, float4* __restrict__ send_right,
                                                       int cap_records, int* __restrict__ counters) {
    const int n = n_dev ? __ldg(n_dev) : n_hi;  // exact extent from the device, else the launch bound
    // Persistent grid, grid-stride over whole 256-particle tiles: the kept / owned COUNTS are accumulated per thread and
    // cost one atomic per block at the end (same-address atomics with a return value serialise in L2 at ~1 ns each: one
    // per 256-particle block was 0.03 ms of this pass at 4M particles); record slots are only allocated by the few
    // tiles that hold halo-zone particles.
    int n_live = 0, n_own = 0;
    const int tiles = (n + 255) >> 8;
    if (blockIdx.x == 0 && threadIdx.x == 0) counters[8] = n;
    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const int i = (tile << 8) + threadIdx.x;
        bool live = false, own = false, to_l = false, to_r = false;
        float4 p = make_float4(0, 0, 0, 0);
        int id = 0;
        if (i < n) {
            id = ids[i];
            if (!(id & SPHE_GHOST_BIT)) {
                p = posq[i];
                int cx = cell_axis(p.x, G.gx, G.cell, G.gnx);
                own = (cx >= S.x0 || !S.has_left) && (cx < S.x1 || !S.has_right);
                // a particle exactly on the -x wall is clamped to the +x wall (collisionS, fluid_system.h:375-382:
                // x == -len takes the else branch): it leaves the first slab for the LAST one, over the wrap link
                const bool far = S.wrap_left && cx >= S.far_x0;
                to_l = S.has_left && (S.wrap_left ? far : cx < S.x0 + S.halo);
                to_r = S.has_right && !S.wrap_right && !far && cx >= S.x1 - S.halo;
                // left the slab but still within the halo zone: stays here as a ghost
                live = own || (!far && cx >= S.x0 - S.halo && cx < S.x1 + S.halo);
            }
            const int nid = live ? (own ? id : (id | SPHE_GHOST_BIT)) : SPHE_DEAD_ID;
            if (nid != id) ids[i] = nid;
        }
        n_live += live ? 1 : 0; n_own += own ? 1 : 0;
        if (__syncthreads_or(to_l || to_r)) {   // block-uniform: tiles without halo-zone particles skip the slot allocation
            const bool flag[4] = {false, to_l, to_r, false};
            int* const ctr[4] = {&counters[0], &counters[1], &counters[2], &counters[3]};
            int slot[4];
            block_append4(flag, ctr, slot);
            const int l = slot[1], r = slot[2];
            if (to_l || to_r) {
                const float4 v = velv[i];
                const float sd = sed[i];
                if (to_l && l < cap_records) {
                    send_left[2 * (l + 1)] = make_float4(p.x, p.y, p.z, sd);
                    send_left[2 * (l + 1) + 1] = make_float4(v.x, v.y, v.z, __int_as_float(id));
                }
                if (to_r && r < cap_records) {
                    send_right[2 * (r + 1)] = make_float4(p.x, p.y, p.z, sd);
                    send_right[2 * (r + 1) + 1] = make_float4(v.x, v.y, v.z, __int_as_float(id));
                }
                if (REMOTE) __threadfence_system();
            }
        }
    }
    // counts: warp reduce, one atomic per warp leader... per block
    __shared__ int red[2][8];
    n_live = __reduce_add_sync(SPHE_FULL, n_live); n_own = __reduce_add_sync(SPHE_FULL, n_own);
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = n_live; red[1][threadIdx.x >> 5] = n_own; }
    __syncthreads();
    if (threadIdx.x < 2) {
        int t = 0;
#pragma unroll
        for (int w = 0; w < 8; w++) t += red[threadIdx.x][w];
        if (t) atomicAdd(&counters[threadIdx.x ? 3 : 0], t);
    }
}

__device__ __forceinline__ int ld_acquire_sys(const int* p) {
    int v;
    asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(int* p, int v) {
    asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// Record 0 of every exchange buffer is a header: int[0] = number of payload records that follow.
// Peer-memory exchange: the buffers are the neighbours' mailboxes; after the headers, flag_* (also in the
// neighbour's memory) is set to the step's sequence number with release semantics at system scope.  The
// classify grid has completed (stream order) and its threads fenced their remote stores, so a consumer
// that acquires the flag sees the whole payload.
__global__ void k_slab_headers(int* __restrict__ counters, float4* __restrict__ send_left, float4* __restrict__ send_right,
                               int* flag_left, int* flag_right, int seq) {
    if (threadIdx.x == 0) {
        if (send_left) send_left[0] = make_float4(__int_as_float(counters[1]), 0.f, 0.f, 0.f);
        if (send_right) send_right[0] = make_float4(__int_as_float(counters[2]), 0.f, 0.f, 0.f);
        counters[4] = 0; counters[5] = 0; counters[6] = 0;   // the unpack counters of this step
        if (counters[15]) counters[7] = 1;                    // a terrain zone sum of the last step timed out (k_zone_sum)
        if (flag_left || flag_right) {
            __threadfence_system();
            if (flag_left) st_release_sys(flag_left, seq);
            if (flag_right) st_release_sys(flag_right, seq);
        }
    }
}

// Appends the payload of both received buffers behind the current extent of the arrays (counters[8]).  All counts are read from
// device memory (the kept count from the pack counters, the payload counts from the headers), so the
// host does not have to know them before this launch.  counters[4] += owned among the appended,
// counters[5] = records taken from the left buffer, counters[6] = from the right buffer.
// PEER: rec_l / rec_r are this GPU's own mailbox, filled by the neighbours' k_slab_classify<true> over
// NVLink.  Thread 0 of every block waits (acquire, system scope) until the mailbox flags carry this step's
// sequence number; the payload is then read with L2-only loads (the L1 / read-only path is not coherent with
// stores that arrive while the kernel runs).  A flag that does not arrive within `timeout` clock cycles sets
// counters[7] (reported as an error by the host) instead of hanging the GPU.
template <bool PEER>
__global__ void __launch_bounds__(256) k_slab_append(int max_l, int max_r, const float4* rec_l, const float4* rec_r,
                                                     const int* flag_l, const int* flag_r, int seq, long long timeout,
                                                     GridP G, SlabP S, int cap_particles,
                                                     float4* __restrict__ posq, float4* __restrict__ velv,
                                                     int* __restrict__ ids, float* __restrict__ sed, int* __restrict__ counters,
                                                     int* __restrict__ n_out, float4* __restrict__ transit_l,
                                                     float4* __restrict__ transit_r, int* __restrict__ transit_n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (PEER) {
        __shared__ int arrived;
        if (threadIdx.x == 0) {
            bool ok = true;
            const long long t0 = clock64();
            if (rec_l) while (ld_acquire_sys(flag_l) < seq) { if (clock64() - t0 > timeout) { ok = false; break; } __nanosleep(200); }
            if (rec_r) while (ok && ld_acquire_sys(flag_r) < seq) { if (clock64() - t0 > timeout) { ok = false; break; } __nanosleep(200); }
            if (!ok) atomicExch(&counters[7], 1);
            arrived = ok;
        }
        __syncthreads();
        if (!arrived) { rec_l = nullptr; rec_r = nullptr; }
    }
    const int kept = counters[8];   // append base: the extent k_slab_classify saw (dead entries included)
    const int head_l = rec_l ? __float_as_int(__ldcg(&rec_l[0]).x) : 0;
    const int head_r = rec_r ? __float_as_int(__ldcg(&rec_r[0]).x) : 0;
    int from_l = min(head_l, max_l), from_r = min(head_r, max_r);
    if (from_l < 0) from_l = 0;
    if (from_r < 0) from_r = 0;
    if (i == 0) {
        counters[5] = head_l;
        counters[6] = head_r;
        // the exact particle count of the coming step, for kernels launched before the host knows it
        *n_out = min(kept + from_l + from_r, cap_particles);
    }
    bool own = false;
    const float4* rec = nullptr;
    int j = 0;
    if (i < from_l) { rec = rec_l; j = i; }
    else if (i < from_l + from_r) { rec = rec_r; j = i - from_l; }
    if (rec && kept + i < cap_particles) {
        float4 a = __ldcg(&rec[2 * (j + 1)]), b = __ldcg(&rec[2 * (j + 1) + 1]);
        int id = __float_as_int(b.w) & SPHE_ID_MASK;
        int cx = cell_axis(a.x, G.gx, G.cell, G.gnx);
        own = (cx >= S.x0 || !S.has_left) && (cx < S.x1 || !S.has_right);
        posq[kept + i] = make_float4(a.x, a.y, a.z, 0.f);
        velv[kept + i] = make_float4(b.x, b.y, b.z, 0.f);
        sed[kept + i] = a.w;
        ids[kept + i] = own ? id : (id | SPHE_GHOST_BIT);
        // A record whose owner lies FURTHER along its direction of travel (a particle that crossed more than
        // one slab in a step -- the contact response of the reference can eject particles at hundreds of
        // box units per second) is handed on at the next exchange; here it stays a ghost for one step (binned
        // at the window edge, too far from everything to be anybody's neighbour).
        const bool from_left = i < from_l;
        const bool onward = !own && (from_left ? (cx >= S.x1 && S.has_right && !S.wrap_right) : (cx < S.x0 && S.has_left && !S.wrap_left));
        if (onward) {
            const int k = atomicAdd(&transit_n[from_left ? 1 : 0], 1);
            if (k < SPHE_TRANSIT_CAP) {
                float4* t = from_left ? transit_r : transit_l;
                t[2 * k] = a; t[2 * k + 1] = make_float4(b.x, b.y, b.z, __int_as_float(id));
            } else atomicExch(&counters[7], 2);
        }
    }
    warp_append(own, &counters[4]);
}

// Terrain zone sums over peer memory (slab-local terrain): the 2W accumulator rows around a slab
// boundary are touched by both neighbours; each rank STORES its copy of the zone into the neighbour's mailbox over NVLink,
// raises the neighbour's flag when the whole grid has stored (last block), waits for the neighbour's own flag and adds
// what arrived -- one launch, no transport library, no host sync.  Every element is read (step 1) and updated (step 3) by
// the same thread, so the phases of different blocks may overlap.  arr: the accumulator array (`want` or `delta`);
// off_* < 0: no neighbour on that side.  err: set to 1 when a flag does not arrive within `timeout` clock cycles.
__global__ void __launch_bounds__(256) k_zone_sum(int* __restrict__ arr, int off_l, int off_r, int n, int* out_l, int* out_r,
                                                  int* flag_out_l, int* flag_out_r, const int* in_l, const int* in_r,
                                                  const int* flag_in_l, const int* flag_in_r, int seq, long long timeout,
                                                  int* __restrict__ done, int* __restrict__ err) {
    const int i0 = blockIdx.x * blockDim.x + threadIdx.x, stride = gridDim.x * blockDim.x;
    for (int i = i0; i < n; i += stride) {
        if (off_l >= 0) out_l[i] = arr[off_l + i];
        if (off_r >= 0) out_r[i] = arr[off_r + i];
    }
    __threadfence_system();
    __shared__ int last, arrived;
    __syncthreads();
    if (threadIdx.x == 0) last = (atomicAdd(done, 1) == (int)gridDim.x - 1);
    __syncthreads();
    if (threadIdx.x == 0) {
        if (last) {
            *done = 0;
            __threadfence_system();
            if (off_l >= 0) st_release_sys(flag_out_l, seq);
            if (off_r >= 0) st_release_sys(flag_out_r, seq);
        }
        bool ok = true;
        const long long t0 = clock64();
        if (off_l >= 0) while (ld_acquire_sys(flag_in_l) < seq) { if (clock64() - t0 > timeout) { ok = false; break; } __nanosleep(100); }
        if (off_r >= 0) while (ok && ld_acquire_sys(flag_in_r) < seq) { if (clock64() - t0 > timeout) { ok = false; break; } __nanosleep(100); }
        if (!ok) atomicExch(err, 1);
        arrived = ok;
    }
    __syncthreads();
    if (!arrived) return;
    for (int i = i0; i < n; i += stride) {
        if (off_l >= 0) arr[off_l + i] += __ldcg(&in_l[i]);
        if (off_r >= 0) arr[off_r + i] += __ldcg(&in_r[i]);
    }
}

// Records in transit (see k_slab_append) join the send buffer of their direction; runs between k_slab_classify
// and k_slab_headers, one block.  transit_n: [0] heading left, [1] heading right, [3] total forwarded so far.
template <bool REMOTE>
__global__ void __launch_bounds__(256) k_slab_forward(const float4* __restrict__ transit_l, const float4* __restrict__ transit_r,
                                                      int* __restrict__ transit_n, float4* __restrict__ send_left,
                                                      float4* __restrict__ send_right, int cap_records, int* __restrict__ counters) {
    int moved = 0;
    for (int dir = 0; dir < 2; dir++) {
        const int n = min(transit_n[dir], SPHE_TRANSIT_CAP);
        const float4* t = dir ? transit_r : transit_l;
        float4* send = dir ? send_right : send_left;
        if (!send) continue;
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            const int slot = atomicAdd(&counters[1 + dir], 1);
            if (slot < cap_records) { send[2 * (slot + 1)] = t[2 * i]; send[2 * (slot + 1) + 1] = t[2 * i + 1]; }
            if (REMOTE) __threadfence_system();
        }
        moved += n;
    }
    __syncthreads();
    if (threadIdx.x == 0) { transit_n[3] += moved; transit_n[0] = 0; transit_n[1] = 0; }
}

// owned particles only, storage order, packed xyz (tests, checkpoints, rendering hand-off)
__global__ void __launch_bounds__(256) k_slab_gather_owned(int n_hi, const int* __restrict__ n_dev, const float4* __restrict__ posq, const float4* __restrict__ velv,
                                                           const float* __restrict__ rho, const float* __restrict__ sed,
                                                           const int* __restrict__ ids, int* __restrict__ counter,
                                                           int* __restrict__ out_ids, float* __restrict__ out_pos,
                                                           float* __restrict__ out_vel, float* __restrict__ out_rho,
                                                           float* __restrict__ out_sed) {
    const int n = n_dev ? __ldg(n_dev) : n_hi;
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    bool own = (i < n) && !(ids[i] & SPHE_GHOST_BIT);   // dead entries (id = -1) carry the ghost bit too
    int k = warp_append(own, counter);
    if (!own) return;
    float4 p = posq[i], v = velv[i];
    out_ids[k] = ids[i];
    out_pos[3 * (size_t)k] = p.x; out_pos[3 * (size_t)k + 1] = p.y; out_pos[3 * (size_t)k + 2] = p.z;
    out_vel[3 * (size_t)k] = v.x; out_vel[3 * (size_t)k + 1] = v.y; out_vel[3 * (size_t)k + 2] = v.z;
    out_rho[k] = rho[i];
    out_sed[k] = sed[i];
}

// Owned particles per GLOBAL cell column (load balancing: slabs.rebalance re-cuts by particle-count quantiles): one
// shared-memory histogram per block, flushed with one atomic per non-empty column.  hist: gnx ints, zeroed by the caller.
__global__ void __launch_bounds__(256) k_slab_column_hist(int n_hi, const int* __restrict__ n_dev, const float4* __restrict__ posq,
                                                          const int* __restrict__ ids, GridP G, int* __restrict__ hist) {
    extern __shared__ int sh[];
    const int n = n_dev ? __ldg(n_dev) : n_hi;
    for (int c = threadIdx.x; c < G.gnx; c += blockDim.x) sh[c] = 0;
    __syncthreads();
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        if (!(ids[i] & SPHE_GHOST_BIT)) atomicAdd(&sh[cell_axis(posq[i].x, G.gx, G.cell, G.gnx)], 1);
    __syncthreads();
    for (int c = threadIdx.x; c < G.gnx; c += blockDim.x) if (sh[c]) atomicAdd(&hist[c], sh[c]);
}

__global__ void k_pack_state_ids(int n, const float* __restrict__ pos, const float* __restrict__ vel, const int* __restrict__ ids_in,
                                 float4* __restrict__ posq, float4* __restrict__ velv, int* __restrict__ ids, float* __restrict__ sed) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    posq[i] = make_float4(pos[3 * (size_t)i], pos[3 * (size_t)i + 1], pos[3 * (size_t)i + 2], 0.f);
    velv[i] = make_float4(vel[3 * (size_t)i], vel[3 * (size_t)i + 1], vel[3 * (size_t)i + 2], 0.f);
    ids[i] = ids_in[i] & SPHE_ID_MASK;
    sed[i] = 0.f;
}

void launch_slab_classify(cudaStream_t st, int n, const int* n_dev, const float4* posq, const float4* velv, int* ids, const float* sed,
                          const GridP& G, const SlabP& S, float4* send_left, float4* send_right, int cap_records, int* counters, bool remote) {
    if (n <= 0) return;
    const int blocks = std::min((n + 255) / 256, 148 * 8);
    if (remote)
        k_slab_classify<true><<<blocks, 256, 0, st>>>(n, n_dev, posq, velv, ids, sed, G, S, send_left, send_right, cap_records, counters);
    else
        k_slab_classify<false><<<blocks, 256, 0, st>>>(n, n_dev, posq, velv, ids, sed, G, S, send_left, send_right, cap_records, counters);
}
void launch_zone_sum(cudaStream_t st, int* arr, int off_l, int off_r, int n, int* out_l, int* out_r, int* flag_out_l, int* flag_out_r,
                     const int* in_l, const int* in_r, const int* flag_in_l, const int* flag_in_r, int seq, long long timeout, int* done, int* err) {
    if (n <= 0 || (off_l < 0 && off_r < 0)) return;
    const int blocks = std::min((n + 1023) / 1024, 96);
    k_zone_sum<<<blocks, 256, 0, st>>>(arr, off_l, off_r, n, out_l, out_r, flag_out_l, flag_out_r, in_l, in_r, flag_in_l, flag_in_r, seq, timeout, done, err);
}
void launch_slab_forward(cudaStream_t st, const float4* transit_l, const float4* transit_r, int* transit_n, float4* send_left,
                         float4* send_right, int cap_records, int* counters, bool remote) {
    if (remote) k_slab_forward<true><<<1, 256, 0, st>>>(transit_l, transit_r, transit_n, send_left, send_right, cap_records, counters);
    else k_slab_forward<false><<<1, 256, 0, st>>>(transit_l, transit_r, transit_n, send_left, send_right, cap_records, counters);
}
void launch_slab_headers(cudaStream_t st, int* counters, float4* send_left, float4* send_right, int* flag_left, int* flag_right, int seq) {
    k_slab_headers<<<1, 32, 0, st>>>(counters, send_left, send_right, flag_left, flag_right, seq);
}
void launch_slab_append(cudaStream_t st, int max_l, int max_r, const float4* rec_l, const float4* rec_r, const GridP& G,
                        const SlabP& S, int cap_particles, float4* posq, float4* velv, int* ids, float* sed, int* counters, int* n_out,
                        float4* transit_l, float4* transit_r, int* transit_n,
                        const int* flag_l, const int* flag_r, int seq, long long timeout_cycles) {
    int m = max_l + max_r;
    if (m < 1) m = 1;
    if (flag_l || flag_r)
        k_slab_append<true><<<(m + 255) / 256, 256, 0, st>>>(max_l, max_r, rec_l, rec_r, flag_l, flag_r, seq, timeout_cycles, G, S,
                                                            cap_particles, posq, velv, ids, sed, counters, n_out, transit_l, transit_r, transit_n);
    else
        k_slab_append<false><<<(m + 255) / 256, 256, 0, st>>>(max_l, max_r, rec_l, rec_r, nullptr, nullptr, 0, 0, G, S, cap_particles,
                                                             posq, velv, ids, sed, counters, n_out, transit_l, transit_r, transit_n);
}
void launch_slab_gather_owned(cudaStream_t st, int n, const int* n_dev, const float4* posq, const float4* velv, const float* rho, const float* sed,
                              const int* ids, int* counter, int* out_ids, float* out_pos, float* out_vel, float* out_rho,
                              float* out_sed) {
    if (n > 0)
        k_slab_gather_owned<<<(n + 255) / 256, 256, 0, st>>>(n, n_dev, posq, velv, rho, sed, ids, counter, out_ids, out_pos, out_vel,
                                                            out_rho, out_sed);
}
int launch_slab_column_hist(cudaStream_t st, int n, const int* n_dev, const float4* posq, const int* ids, const GridP& G, int* hist) {
    if (n <= 0) return 0;
    const size_t smem = (size_t)G.gnx * sizeof(int);
    if (smem > 200 * 1024) return -1;
    if (smem > 48 * 1024) cudaFuncSetAttribute(k_slab_column_hist, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k_slab_column_hist<<<std::min((n + 255) / 256, 592), 256, smem, st>>>(n, n_dev, posq, ids, G, hist);
    return 0;
}
void launch_pack_state_ids(cudaStream_t st, int n, const float* pos, const float* vel, const int* ids_in, float4* posq, float4* velv,
                           int* ids, float* sed) {
    if (n > 0) k_pack_state_ids<<<(n + 255) / 256, 256, 0, st>>>(n, pos, vel, ids_in, posq, velv, ids, sed);
}

}  // namespace sphe
