// Binning: cell hash -> per-cell counts -> exclusive prefix scan (cell-start table) -> counting-sort
// scatter -> deterministic in-cell ranking by particle id -> physical reorder of the SoA arrays.
//
// The result is the canonical order "sorted by (cell id, particle id)", bit-identical to
// oracle/sph_oracle.c so_bin() regardless of storage history or atomic scheduling.
// All kernels here are HBM-bandwidth bound (DESIGN.md section "kernels").
#include "common.cuh"
#include "sim.h"

namespace sphe {

// ---------------------------------------------------------------- hash + count
// One thread per particle (storage order).  Storage order is last step's sorted order, so
// consecutive lanes mostly share a cell: counts are aggregated per run of equal cells inside the
// warp (shfl + ballot) and only the run head issues the atomic.
__global__ void __launch_bounds__(256) k_hash(int n_hi, const int* __restrict__ n_dev, const float4* __restrict__ posq, const int* __restrict__ ids, GridP G,
                                              uint32_t* __restrict__ cell, int* __restrict__ count) {
    const int n = n_dev ? __ldg(n_dev) : n_hi;  // exact count from the device in slab mode, else the launch bound
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned lane = threadIdx.x & 31;
    uint32_t c = 0xffffffffu;
    if (i < n) {
        if (!ids || ids[i] != -1) {   // slab mode: an entry the exchange dropped has no cell
            float4 p = posq[i];
            int cx, cy, cz;
            cell_coords(G, p.x, p.y, p.z, cx, cy, cz);
            c = (uint32_t)((cx * G.ny + cy) * G.nz + cz);
        }
        cell[i] = c;
    }
    uint32_t prev = __shfl_up_sync(SPHE_FULL, c, 1);
    bool head = (lane == 0) || (c != prev);
    unsigned heads = __ballot_sync(SPHE_FULL, head);
    if (head && c != 0xffffffffu) {
        unsigned above = (lane == 31) ? 0u : (heads & ~((2u << lane) - 1u));
        int next = above ? (__ffs(above) - 1) : 32;
        atomicAdd(&count[c], next - (int)lane);
    }
}

// ---------------------------------------------------------------- exclusive scan over cells
// Reduce-then-scan in two launches: per-tile sums, then per-tile scan (each block sums the tile sums before it).
// The final pass also primes the scatter cursor and clears the counts for the next step.
#ifndef SCAN_THREADS_
#define SCAN_THREADS_ 256
#define SCAN_ITEMS_ 16
#endif
constexpr int SCAN_THREADS = SCAN_THREADS_;
constexpr int SCAN_ITEMS = SCAN_ITEMS_;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ int warp_incl_scan(int v) {
    unsigned lane = threadIdx.x & 31;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(SPHE_FULL, v, o);
        if (lane >= (unsigned)o) v += t;
    }
    return v;
}

// inclusive block scan of one value per thread; returns inclusive prefix, total in *total
__device__ __forceinline__ int block_incl_scan(int v, int* total) {
    __shared__ int ws[32];
    unsigned lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int inc = warp_incl_scan(v);
    if (lane == 31) ws[w] = inc;
    __syncthreads();
    if (w == 0) {
        int nw = (blockDim.x + 31) >> 5;
        int x = (int)lane < nw ? ws[lane] : 0;
        x = warp_incl_scan(x);
        ws[lane] = x;
    }
    __syncthreads();
    int off = w ? ws[w - 1] : 0;
    *total = ws[((blockDim.x + 31) >> 5) - 1];
    __syncthreads();
    return inc + off;
}

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_reduce(long long ncells, const int* __restrict__ count,
                                                               int* __restrict__ tile_sum) {
    long long base = (long long)blockIdx.x * SCAN_TILE;
    int s = 0;
    // vectorised int4 loads: count is 16-byte aligned and SCAN_TILE is a multiple of 4
    const int4* c4 = reinterpret_cast<const int4*>(count + base);
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS / 4; k++) {
        long long e = base + ((long long)k * SCAN_THREADS + threadIdx.x) * 4;
        if (e + 3 < ncells) {
            int4 v = c4[k * SCAN_THREADS + threadIdx.x];
            s += v.x + v.y + v.z + v.w;
        } else {
            for (int q = 0; q < 4; q++) if (e + q < ncells) s += count[e + q];
        }
    }
    int total;
    block_incl_scan(s, &total);
    if (threadIdx.x == 0) tile_sum[blockIdx.x] = total;
}

__global__ void __launch_bounds__(1024) k_scan_tiles(int ntiles, int* __restrict__ tile_sum) {
    __shared__ int carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (int b = 0; b < ntiles; b += 1024) {
        int i = b + threadIdx.x;
        int v = i < ntiles ? tile_sum[i] : 0;
        int total;
        int inc = block_incl_scan(v, &total);
        int carry = carry_s;
        if (i < ntiles) tile_sum[i] = carry + inc - v;  // exclusive
        __syncthreads();
        if (threadIdx.x == 0) carry_s = carry + total;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_final(long long ncells, int* __restrict__ count,
                                                             const int* __restrict__ tile_off,
                                                             int* __restrict__ cell_start, int* __restrict__ cursor) {
    long long base = (long long)blockIdx.x * SCAN_TILE;
    // each thread owns SCAN_ITEMS consecutive cells (blocked arrangement)
    long long e0 = base + (long long)threadIdx.x * SCAN_ITEMS;
    int v[SCAN_ITEMS];
    int s = 0;
    if (e0 + SCAN_ITEMS <= ncells) {
        const int4* c4 = reinterpret_cast<const int4*>(count + e0);
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS / 4; k++) {
            int4 t = c4[k];
            v[4 * k] = t.x; v[4 * k + 1] = t.y; v[4 * k + 2] = t.z; v[4 * k + 3] = t.w;
        }
    } else {
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; k++) v[k] = (e0 + k < ncells) ? count[e0 + k] : 0;
    }
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) s += v[k];
    // offset of this tile = sum of the raw tile sums before it, computed by every block for itself (a few KB from L2):
    // no separate single-CTA launch for the scan of the tile sums
    int pre = 0;
    for (int i = threadIdx.x; i < (int)blockIdx.x; i += SCAN_THREADS) pre += tile_off[i];
    int tile_base;
    block_incl_scan(pre, &tile_base);
    int total;
    int inc = block_incl_scan(s, &total);
    int run = tile_base + inc - s;
    if (e0 + SCAN_ITEMS <= ncells) {
        int o[SCAN_ITEMS];
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; k++) { o[k] = run; run += v[k]; }
        int4* s4 = reinterpret_cast<int4*>(cell_start + e0);
        int4* u4 = reinterpret_cast<int4*>(cursor + e0);
        int4* z4 = reinterpret_cast<int4*>(count + e0);
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS / 4; k++) {
            int4 t = make_int4(o[4 * k], o[4 * k + 1], o[4 * k + 2], o[4 * k + 3]);
            s4[k] = t; u4[k] = t; z4[k] = make_int4(0, 0, 0, 0);
        }
    } else {
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; k++) {
            if (e0 + k < ncells) { cell_start[e0 + k] = run; cursor[e0 + k] = run; count[e0 + k] = 0; run += v[k]; }
        }
    }
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) cell_start[ncells] = tile_base + total;   // number of live particles
}

// ---------------------------------------------------------------- the same scan in ONE launch (decoupled look-back)
// Round 1 scanned in two launches (tile sums, then every block re-summed the tile sums before it): 17 us at 0.78M cells,
// pure latency.  Here every tile publishes its aggregate, looks back over its predecessors' status words a warp at a time
// until it meets one that already knows its inclusive prefix, and publishes its own (Merrill & Garland).  Status words carry
// the launch's epoch, so nothing has to be cleared between steps; tiles are handed out by a ticket so a tile's predecessors
// always run no later than it does.
//   word = epoch << 34 | status << 32 | value;  status 1 = aggregate of the tile, 2 = inclusive prefix up to the tile
__device__ __forceinline__ unsigned long long scan_pack(unsigned epoch, unsigned status, int value) {
    return ((unsigned long long)epoch << 34) | ((unsigned long long)status << 32) | (unsigned)value;
}

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_onepass(long long ncells, int* __restrict__ count, int* __restrict__ cell_start,
                                                               int* __restrict__ cursor, volatile unsigned long long* state,
                                                               unsigned* __restrict__ ticket, unsigned ticket_base, unsigned epoch) {
    __shared__ int s_tile, s_prefix;
    if (threadIdx.x == 0) s_tile = (int)(atomicAdd(ticket, 1u) - ticket_base);
    __syncthreads();
    const int tile = s_tile;
    const long long e0 = (long long)tile * SCAN_TILE + (long long)threadIdx.x * SCAN_ITEMS;
    int v[SCAN_ITEMS];
    int s = 0;
    if (e0 + SCAN_ITEMS <= ncells) {
        const int4* c4 = reinterpret_cast<const int4*>(count + e0);
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS / 4; k++) {
            int4 t = c4[k];
            v[4 * k] = t.x; v[4 * k + 1] = t.y; v[4 * k + 2] = t.z; v[4 * k + 3] = t.w;
        }
    } else {
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; k++) v[k] = (e0 + k < ncells) ? count[e0 + k] : 0;
    }
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) s += v[k];
    int total;
    const int inc = block_incl_scan(s, &total);
    if (threadIdx.x < 32) {
        const unsigned lane = threadIdx.x;
        if (tile == 0) {
            if (lane == 0) { state[0] = scan_pack(epoch, 2u, total); s_prefix = 0; }
        } else {
            if (lane == 0) { state[tile] = scan_pack(epoch, 1u, total); __threadfence(); }
            // warp-wide look-back: lane l inspects tile (p - l); stop at the nearest tile that has its inclusive prefix
            int run = 0;
            for (int p = tile - 1;;) {
                const int q = p - (int)lane;
                unsigned long long w = 0;
                unsigned st = 3u;                                  // 3 = before the first tile: contributes nothing, ends the walk
                if (q >= 0) {
                    w = state[q];
                    st = ((unsigned)(w >> 34) == epoch) ? (unsigned)((w >> 32) & 3u) : 0u;
                }
                const unsigned invalid = __ballot_sync(SPHE_FULL, st == 0u);
                const unsigned done = __ballot_sync(SPHE_FULL, st >= 2u);
                // usable lanes: those nearer than the first invalid one
                const unsigned first_invalid = invalid ? (unsigned)(__ffs((int)invalid) - 1) : 32u;
                const unsigned first_done = done ? (unsigned)(__ffs((int)done) - 1) : 32u;
                if (first_done < first_invalid) {
                    // sum lanes 0 .. first_done (inclusive; a "before the first tile" lane adds 0)
                    const int c = (lane <= first_done && st != 3u) ? (int)(unsigned)w : 0;
                    run += __reduce_add_sync(SPHE_FULL, c);
                    break;
                }
                if (first_invalid == 32u) {   // 32 aggregates, none with a prefix yet: take them all and look further back
                    run += __reduce_add_sync(SPHE_FULL, (int)(unsigned)w);
                    p -= 32;
                }
                // else: a predecessor has not published yet -> look again
            }
            if (lane == 0) { s_prefix = run; __threadfence(); state[tile] = scan_pack(epoch, 2u, run + total); }
        }
    }
    __syncthreads();
    const int tile_base = s_prefix;
    int run = tile_base + inc - s;
    if (e0 + SCAN_ITEMS <= ncells) {
        int o[SCAN_ITEMS];
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; k++) { o[k] = run; run += v[k]; }
        int4* s4 = reinterpret_cast<int4*>(cell_start + e0);
        int4* u4 = reinterpret_cast<int4*>(cursor + e0);
        int4* z4 = reinterpret_cast<int4*>(count + e0);
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS / 4; k++) {
            int4 t = make_int4(o[4 * k], o[4 * k + 1], o[4 * k + 2], o[4 * k + 3]);
            s4[k] = t; u4[k] = t; z4[k] = make_int4(0, 0, 0, 0);
        }
    } else {
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; k++) {
            if (e0 + k < ncells) { cell_start[e0 + k] = run; cursor[e0 + k] = run; count[e0 + k] = 0; run += v[k]; }
        }
    }
    if (tile == (int)gridDim.x - 1 && threadIdx.x == 0) cell_start[ncells] = tile_base + total;   // number of live particles
}

// ---------------------------------------------------------------- counting-sort scatter
// slot = cursor[cell]++ (run-aggregated).  The order INSIDE a cell is whatever the atomics give;
// k_rank_reorder makes it canonical.
#ifndef SCATTER_ROUNDS
#define SCATTER_ROUNDS 4
#endif
// A block takes SCATTER_ROUNDS x 256 consecutive particles, one round per 256: the loads of all rounds are issued first
// and the cursor atomics of all rounds are in flight together (the atomic's return is the kernel's critical path:
// ncu long_scoreboard 45 warps per issue with one particle per thread).
__global__ void __launch_bounds__(256) k_scatter(int n_hi, const int* __restrict__ n_dev, const uint32_t* __restrict__ cell, const int* __restrict__ ids,
                                                 int* __restrict__ cursor, uint2* __restrict__ tmp, uint32_t* __restrict__ cell_sorted) {
    const int n = n_dev ? __ldg(n_dev) : n_hi;  // exact count from the device in slab mode, else the launch bound
    const unsigned lane = threadIdx.x & 31;
    const int i0 = blockIdx.x * (256 * SCATTER_ROUNDS) + threadIdx.x;
    uint32_t c[SCATTER_ROUNDS];
    int id[SCATTER_ROUNDS], hl[SCATTER_ROUNDS], base[SCATTER_ROUNDS];
#pragma unroll
    for (int r = 0; r < SCATTER_ROUNDS; r++) {
        const int i = i0 + r * 256;
        c[r] = (i < n) ? cell[i] : 0xffffffffu;
        id[r] = (i < n) ? ids[i] : 0;
    }
#pragma unroll
    for (int r = 0; r < SCATTER_ROUNDS; r++) {
        uint32_t prev = __shfl_up_sync(SPHE_FULL, c[r], 1);
        bool head = (lane == 0) || (c[r] != prev);
        unsigned heads = __ballot_sync(SPHE_FULL, head);
        // lane of my run's head = highest head bit at or below my lane
        unsigned below = heads & ((2u << lane) - 1u);
        if (lane == 31) below = heads;
        hl[r] = 31 - __clz(below);
        base[r] = 0;
        if (head && c[r] != 0xffffffffu) {
            unsigned above = (lane == 31) ? 0u : (heads & ~((2u << lane) - 1u));
            int next = above ? (__ffs(above) - 1) : 32;
            base[r] = atomicAdd(&cursor[c[r]], next - (int)lane);
        }
    }
#pragma unroll
    for (int r = 0; r < SCATTER_ROUNDS; r++) {
        const int i = i0 + r * 256;
        const int b = __shfl_sync(SPHE_FULL, base[r], hl[r]);
        // every slot of a cell's range holds that cell's id, whatever the final in-cell order: the sorted cell ids are final here
        if (i < n && c[r] != 0xffffffffu) {
            const int slot = b + ((int)lane - hl[r]);
            tmp[slot] = make_uint2((uint32_t)id[r], (uint32_t)i);
            cell_sorted[slot] = c[r];
        }
    }
}

// ---------------------------------------------------------------- rank inside the cell + reorder
// One thread per scattered slot.  rank = number of particles of the same cell with a smaller id;
// destination = cell_start + rank.  Then gather the particle's state from its old storage slot.
__global__ void __launch_bounds__(256) k_rank_reorder(int n_hi, const int* __restrict__ n_dev, const uint2* __restrict__ tmp, const uint32_t* __restrict__ cell_sorted,
                                                      const int* __restrict__ cell_start,
                                                      const float4* __restrict__ posq_in, const float4* __restrict__ velv_in,
                                                      const float* __restrict__ sed_in,
                                                      float4* __restrict__ posq_out, float4* __restrict__ velv_out,
                                                      float* __restrict__ sed_out, int* __restrict__ ids_out,
                                                      int* __restrict__ src_of_slot) {
    const int n = n_dev ? __ldg(n_dev) : n_hi;  // exact count from the device in slab mode, else the launch bound
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    // The chain of dependent loads is what this kernel waits on (ncu long_scoreboard 29 warps per issue): the cell id comes
    // from the scatter (coalesced, no gather through the old slot) and the state gathers are issued before the ranking loop.
    const uint2 me = tmp[s];
    const uint32_t c = cell_sorted[s];
    const float4 p = posq_in[me.y];
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (!src_of_slot) v = velv_in[me.y];
    float sd = 0.f;
    if (sed_in) sd = sed_in[me.y];
    const int a = cell_start[c], b = cell_start[c + 1];
    int rank = 0;
    for (int k = a; k < b; k++) rank += ((tmp[k].x & SPHE_ID_MASK) < (me.x & SPHE_ID_MASK)) ? 1 : 0;
    const int dst = a + rank;
    posq_out[dst] = p;
    // src_of_slot != NULL (sphe_step_host): the velocities are still on the PCIe bus; only the permutation is recorded
    // and k_gather_vel moves them after the density pass, which does not read them
    if (src_of_slot) src_of_slot[dst] = (int)me.y;
    else velv_out[dst] = v;
    if (sed_in) sed_out[dst] = sd;
    ids_out[dst] = (int)me.x;
}

// ---------------------------------------------------------------- launch wrappers
void launch_hash(cudaStream_t st, int n, const int* n_dev, const float4* posq, const int* ids, const GridP& G, uint32_t* cell, int* count) {
    if (n <= 0) return;
    k_hash<<<(n + 255) / 256, 256, 0, st>>>(n, n_dev, posq, ids, G, cell, count);
}

int scan_tiles_for(long long ncells) { return (int)((ncells + SCAN_TILE - 1) / SCAN_TILE); }

// state: ntiles 64-bit status words (any contents with an older epoch); ticket: one word, monotonically increasing
void launch_scan_onepass(cudaStream_t st, long long ncells, int* count, int* cell_start, int* cursor, unsigned long long* state,
                         unsigned* ticket, unsigned ticket_base, unsigned epoch) {
    int ntiles = scan_tiles_for(ncells);
    k_scan_onepass<<<ntiles, SCAN_THREADS, 0, st>>>(ncells, count, cell_start, cursor, state, ticket, ticket_base, epoch);
}

void launch_scan(cudaStream_t st, long long ncells, int* count, int* tile_sum, int* cell_start, int* cursor) {
    int ntiles = scan_tiles_for(ncells);
    k_scan_reduce<<<ntiles, SCAN_THREADS, 0, st>>>(ncells, count, tile_sum);
    k_scan_final<<<ntiles, SCAN_THREADS, 0, st>>>(ncells, count, tile_sum, cell_start, cursor);
}

void launch_scatter(cudaStream_t st, int n, const int* n_dev, const uint32_t* cell, const int* ids, int* cursor, uint2* tmp, uint32_t* cell_sorted) {
    if (n <= 0) return;
    k_scatter<<<(n + 256 * SCATTER_ROUNDS - 1) / (256 * SCATTER_ROUNDS), 256, 0, st>>>(n, n_dev, cell, ids, cursor, tmp, cell_sorted);
}

void launch_rank_reorder(cudaStream_t st, int n, const int* n_dev, const uint2* tmp, const uint32_t* cell_sorted, const int* cell_start,
                         const float4* posq_in, const float4* velv_in, const float* sed_in,
                         float4* posq_out, float4* velv_out, float* sed_out, int* ids_out, int* src_of_slot) {
    if (n <= 0) return;
    k_rank_reorder<<<(n + 255) / 256, 256, 0, st>>>(n, n_dev, tmp, cell_sorted, cell_start, posq_in, velv_in, sed_in,
                                                    posq_out, velv_out, sed_out, ids_out, src_of_slot);
}

// velocities into sorted order after the fact (see k_rank_reorder); .w = m/rho was written by the density pass and stays
__global__ void __launch_bounds__(256) k_gather_vel(int n, const int* __restrict__ src_of_slot, const float4* __restrict__ velv_in,
                                                    float4* __restrict__ velv_out) {
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    const float4 v = velv_in[src_of_slot[s]];
    float* o = reinterpret_cast<float*>(velv_out + s);
    o[0] = v.x; o[1] = v.y; o[2] = v.z;
}
void launch_gather_vel(cudaStream_t st, int n, const int* src_of_slot, const float4* velv_in, float4* velv_out) {
    if (n > 0) k_gather_vel<<<(n + 255) / 256, 256, 0, st>>>(n, src_of_slot, velv_in, velv_out);
}

}  // namespace sphe
