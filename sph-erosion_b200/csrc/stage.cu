// Cell-cooperative neighbour passes with the candidates STAGED IN SHARED MEMORY (the default since round 2).
//
//   k_density_stage : pass 1 of FluidSystemSPH::Run (Erosion/fluid_system.h:108-124, kernDefault :410-416) + the
//                     neighbour test for both passes, recorded as ONE BIT per candidate
//   k_force_stage   : passes 2+3 (:128-178) + advance() (:306-353) + collisionS (:355-407) over the set bits only
//
// Why (profiles/r01c_ncu_k_density_list_c2.txt, r01c_ncu_k_force_list_c2.txt): the list kernels of round 1 sat on the
// L1 data pipe (lsu wavefronts 81 % / 64 % of peak, issue active 58 % / 45 %): every lane gathers its own candidate
// with an LDG.128 that touches ~10 distinct 128-byte lines per warp request, and the force pass chases
// index -> gather chains through L1/L2 (long scoreboard).  Here a CTA owns 2*THREADS consecutive sorted particles
// (targets).  They lie in a few z-runs of cells of one (cx, cy) column -- a "segment" -- and everything a segment's
// targets can interact with is the same z-run (+-1 cell) of the 9 columns around it.  The CTA copies that block
// into shared memory once: each of the 9 columns contributes ONE contiguous range of the sorted arrays (z is the least
// significant digit of the cell id), moved by a TMA bulk copy (cp.async.bulk -> UBLKCP, completion on an mbarrier): no
// per-particle instructions, no L1 wavefronts.  The candidates of a target in cell z are 9 sub-ranges of the stage
// (cells z-1 .. z+1 of every run); a lane walks them with LDS.128 at immediate offsets: no gathers, 29-cycle loads,
// and lanes that sit in the same cell read the same address (broadcast).
//
// Two targets per lane with packed fp32 math (FADD2/FMUL2/FFMA2, as in round 1): a candidate is loaded once for both.
//
// Neighbour "lists" are BIT MASKS over a lane's candidate range: the density pass sets bit i when candidate i is
// within h (+ a few ulp, see ST_EPS) of either target, 32 candidates per word, words stored per warp tile
// [warp][row][lane] (coalesced, ~36 B/particle/step instead of the 150 B/particle of the index lists).  The force
// pass stages the same block (plus velocities), recomputes the same ranges and visits the set bits.  There is no
// list capacity, no overflow, no spill: a range is at most the stage, and the rows are allocated for that.
//
// A cell neighbourhood that does not fit the stage (thousands of particles in one cell: tests/test_gpu_edges.py
// pile-up) takes the FALLBACK walk straight from global memory, in the same plane-major order with the same
// arithmetic, so results never depend on which path a particle took nor on how particles are grouped into CTAs or
// pairs (extra candidates contribute an exact +0): K slabs stay bit-equal to one GPU.
#include "common.cuh"
#include "sim.h"
#include "sph_device.cuh"

namespace sphe {

constexpr int ST_ZT = 30;               // target z-cells per segment at most
constexpr int ST_MAXP = ST_ZT + 2;      // planes staged per segment (one lane of warp 0 per plane in the fit test)
constexpr int ST_G = 4;                 // candidates per group of the density walk
constexpr int ST_PADG = ST_G - 1;       // sentinel slots behind every staged run (a group may overrun its range)
constexpr float ST_FAR = 1.0e18f;       // inactive target of a pair: every weight clamps to exactly 0
// A candidate is recorded when hh - d2 >= -ST_EPS*hh with the FMA-contracted d2.  The reference predicate is
// sqrt(d2_exact) <= h  <=>  d2_exact <= T, T within 1 ulp of hh; the contracted d2 differs from the exact one by
// < 3 ulp, so 2^-20 (8 ulp) makes the recorded set a SUPERSET of the exact neighbour set.  Every extra entry has
// clamped weights max(hh - d2, 0), max(h - r, 0) that are 0 or a few ulp: it adds (next to) nothing to any sum
// (tests/test_gpu_parity.py::test_pair_masks_cover_the_exact_neighbour_sets).
constexpr float ST_EPS = 9.5367431640625e-07f;

// Stage layout (RUN-MAJOR): run r = column (cx + r/3 - 1, cy + r%3 - 1), cells z = zlo-1 .. zhi+1, is ONE contiguous range
// of the sorted arrays (z is the least significant digit of the cell id), so a segment is staged with 9 bulk copies
// (cp.async.bulk = TMA, completion on an mbarrier) -- no per-particle instructions, no L1 wavefronts, no registers.
// The candidates of a target in cell z are 9 sub-ranges, one per run: stage slot of sorted index k = rb[r] + k.
template <int THREADS, int CAPC, bool VEL>
struct StageShared {
    float4 pos[CAPC + 9 * ST_PADG + 1];               // density: (x, y, z, -); force: (x, y, z, P/rho^2)
    float4 vel[VEL ? CAPC + 9 * ST_PADG + 1 : 1];     // force: (vx, vy, vz, m/rho)
    int raw[9][ST_MAXP + 2];                          // cell_start of column r at z = zlo-1+p, p = 0..np (clamped to the grid)
    int rb[9];                                        // stage slot of sorted index k of run r = rb[r] + k
    uint32_t tcell[2 * THREADS];                      // cell ids of the CTA's targets
    unsigned long long mbar;
    int seg[2];
};

__host__ __device__ constexpr int stage_mask_rows(int capc) { return 2 * (9 + (capc + 31) / 32); }

__device__ __forceinline__ uint32_t st_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void st_mbar_init(unsigned long long* mbar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(st_smem_u32(mbar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void st_mbar_expect_tx(unsigned long long* mbar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(st_smem_u32(mbar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void st_mbar_wait(unsigned long long* mbar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(ok) : "r"(st_smem_u32(mbar)), "r"(parity) : "memory");
    } while (!ok);
}
// global -> shared bulk copy (TMA unit, SASS UBLKCP); bytes is a multiple of 16, both addresses 16-byte aligned
__device__ __forceinline__ void st_bulk_g2s(void* dst, const void* src, uint32_t bytes, unsigned long long* mbar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(st_smem_u32(dst)), "l"(src), "r"(bytes), "r"(st_smem_u32(mbar)) : "memory");
}
__device__ __forceinline__ void st_fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// The segment loop shared by the density, force and decode kernels (it must be the SAME code: the force pass
// re-derives the candidate ranges the density pass recorded its bits against).
//   pass(useA, useB, pf, pl): targets A and/or B against planes [pf, pl) of the 9 staged runs:
//                             run r, stage slots [S.rb[r] + S.raw[r][pf], S.rb[r] + S.raw[r][pl])
//   fall(which, col): the neighbourhood of target A (0) / B (1) does not fit the stage -> walk global memory
// S.tcell must hold the CTA's target cells and S.mbar must be initialised (stage_prologue).
template <int THREADS, int CAPC, bool VEL, class PassF, class FallF>
__device__ __forceinline__ void stage_run(StageShared<THREADS, CAPC, VEL>& S, const int t0, const int t1, const float4* __restrict__ posq,
                                          const float4* __restrict__ velv, const int* __restrict__ cell_start, const GridP& G,
                                          const int a, const int cza, const int czb, const bool liveA, const bool liveB,
                                          PassF&& pass, FallF&& fall) {
    const int tid = threadIdx.x, lane = tid & 31;
    const int b = a + 1;
    const int ntar = t1 - t0;
    int seg_begin = t0;
    uint32_t parity = 0;
#pragma unroll 1
    while (seg_begin < t1) {
        __syncthreads();   // (A) everybody is done with the previous segment's stage and tables
        // segment = the targets from seg_begin on that share its column, at most ST_ZT z-cells (all threads compute this
        // redundantly from the target cells in shared memory: broadcast reads, no global latency, no barrier)
        const uint32_t c0 = S.tcell[seg_begin - t0];
        const int col = (int)(c0 / (uint32_t)G.nz), zlo = (int)(c0 - (uint32_t)col * (uint32_t)G.nz);
        int colend;
        {
            const uint32_t lim = (uint32_t)(col + 1) * (uint32_t)G.nz;   // first cell of the next column
            int lo = seg_begin - t0 + 1, hi = ntar;                     // first target index with cell >= lim
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (S.tcell[mid] < lim) lo = mid + 1; else hi = mid;
            }
            colend = t0 + lo;
        }
        const int zmax = (int)(S.tcell[colend - 1 - t0] - (uint32_t)col * (uint32_t)G.nz);
        const int zcap = min(zmax, zlo + ST_ZT - 1);
        const int np = zcap - zlo + 3;   // planes zlo-1 .. zcap+1
        const int cy = col % G.ny, cx = col / G.ny;
        for (int i = tid; i < 9 * (np + 1); i += THREADS) {
            const int r = i / (np + 1), p = i - r * (np + 1);
            const int x = cx + r / 3 - 1, y = cy + r % 3 - 1;
            int v = 0;   // a column outside the grid is empty
            if (x >= 0 && x < G.nx && y >= 0 && y < G.ny) {
                const int z = min(max(zlo - 1 + p, 0), G.nz);   // planes below / above the grid are empty
                v = __ldg(&cell_start[(x * G.ny + y) * G.nz + z]);
            }
            S.raw[r][p] = v;
        }
        __syncthreads();   // (B)
        if (tid < 32) {
            // lane p: candidates of planes 0..p over the 9 runs; the planes that fit the stage are a prefix
            int tot = 0x3fffffff;
            if (lane < np) {
                tot = 0;
#pragma unroll
                for (int r = 0; r < 9; r++) tot += S.raw[r][lane + 1] - S.raw[r][0];
            }
            const int npf = __popc(__ballot_sync(SPHE_FULL, tot <= CAPC));
            if (npf >= 3) {
                // lane r < 9: run r = sorted indices [raw[r][0], raw[r][npf]) -> stage slots [rbase, rbase + len) + sentinels
                const int first = lane < 9 ? S.raw[lane][0] : 0;
                const int len = lane < 9 ? S.raw[lane][npf] - first : 0;
                int incl = lane < 9 ? len + ST_PADG : 0;
#pragma unroll
                for (int o = 1; o < 16; o <<= 1) {
                    const int t = __shfl_up_sync(SPHE_FULL, incl, o);
                    if (lane >= o) incl += t;
                }
                const int rbase = incl - (len + ST_PADG);
                const int total = __reduce_add_sync(SPHE_FULL, len);
                if (lane == 0) {
                    st_fence_proxy_async();   // the stage was last read through the generic proxy
                    st_mbar_expect_tx(&S.mbar, (uint32_t)total * (VEL ? 32u : 16u));
                }
                __syncwarp();
                if (lane < 9) {
                    S.rb[lane] = rbase - first;
                    if (len > 0) {
                        st_bulk_g2s(&S.pos[rbase], posq + first, (uint32_t)len * 16u, &S.mbar);
                        if (VEL) st_bulk_g2s(&S.vel[rbase], velv + first, (uint32_t)len * 16u, &S.mbar);
                    }
#pragma unroll
                    for (int k = 0; k < ST_PADG; k++) {
                        S.pos[rbase + len + k] = make_float4(-ST_FAR, -ST_FAR, -ST_FAR, 0.f);
                        if (VEL) S.vel[rbase + len + k] = make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                }
            }
            if (lane == 0) S.seg[0] = npf;
        }
        __syncthreads();   // (C)
        const int npf = S.seg[0];          // planes that fit the stage
        const bool staged = npf >= 3;      // at least the neighbourhood of the first target's cell
        const int zhi = zlo + npf - 3;     // last target z-cell of this segment
        const int seg_end = min(colend, staged ? S.raw[4][zhi - zlo + 2] : S.raw[4][2]);
        const bool inA = liveA && a >= seg_begin && a < seg_end;
        const bool inB = liveB && b >= seg_begin && b < seg_end;
        if (staged) {
            st_mbar_wait(&S.mbar, parity);   // the 9 runs have landed
            parity ^= 1u;
            const bool merged = inA && inB && (czb - cza <= 1);
            // pass 0: A alone or A+B merged; pass 1: B alone
#pragma unroll 1
            for (int p = 0; p < 2; p++) {
                const bool useA = (p == 0) && inA;
                const bool useB = (p == 0) ? merged : (inB && !merged);
                if (!(useA || useB)) continue;
                const int zf = useA ? cza : czb, zl = useB ? czb : cza;
                pass(useA, useB, zf - zlo, zl - zlo + 3);
            }
        } else {
            if (inA) fall(0, col);
            if (inB) fall(1, col);
        }
        seg_begin = seg_end;
    }
}

// loads the CTA's target cells into shared memory and initialises the staging barrier
template <int THREADS, int CAPC, bool VEL>
__device__ __forceinline__ void stage_prologue(StageShared<THREADS, CAPC, VEL>& S, const uint32_t* __restrict__ cell_sorted, int t0, int t1,
                                               uint32_t& ca, uint32_t& cb) {
    const int tid = threadIdx.x;
    const int ia = min(t0 + 2 * tid, t1 - 1), ib = min(t0 + 2 * tid + 1, t1 - 1);
    ca = cell_sorted[ia]; cb = cell_sorted[ib];
    S.tcell[2 * tid] = ca; S.tcell[2 * tid + 1] = cb;
    if (tid == 0) st_mbar_init(&S.mbar, 1);
    // the first barrier of stage_run orders both before any use
}

// Plane-major... no: RUN-major walk of the 27 cells around cell (cx, cy, cz) straight from global memory: f(k) for
// every candidate, in exactly the order a staged pass visits them (run r = 0..8, then z-1 .. z+1 inside the run).
template <class F>
__device__ __forceinline__ void walk27_runs(const GridP& G, const int* __restrict__ cell_start, int cx, int cy, int cz, F&& f) {
    const int z0 = cz > 0 ? cz - 1 : 0, z1 = cz < G.nz - 1 ? cz + 1 : cz;
#pragma unroll 1
    for (int r = 0; r < 9; r++) {
        const int x = cx + r / 3 - 1, y = cy + r % 3 - 1;
        if (x < 0 || x >= G.nx || y < 0 || y >= G.ny) continue;
        const int base = (x * G.ny + y) * G.nz;
        const int s = __ldg(&cell_start[base + z0]), e = __ldg(&cell_start[base + z1 + 1]);
#pragma unroll 1
        for (int k = s; k < e; k++) f(k);
    }
}

// ------------------------------------------------------------------ pass 1: density + pressure + neighbour masks
// Writes rho and packs what pass 2 needs per neighbour into the arrays it stages anyway:
//   posq_q[i] = (x, y, z, P_i / rho_i^2)      velv[i].w = mass / rho_i
template <int THREADS, int CAPC, int MINB>
__global__ void __launch_bounds__(THREADS, MINB)
k_density_stage(int n_hi, const int* __restrict__ n_dev, const float4* __restrict__ posq, float4* __restrict__ posq_q,
                float4* __restrict__ velv, const uint32_t* __restrict__ cell_sorted, const int* __restrict__ cell_start,
                GridP G, StepC C, float* __restrict__ rho, unsigned* __restrict__ nmask) {
    using SS = StageShared<THREADS, CAPC, false>;
    extern __shared__ __align__(16) unsigned char st_smem[];
    SS& S = *reinterpret_cast<SS*>(st_smem);
    constexpr int MR = stage_mask_rows(CAPC);
    const int n = n_dev ? __ldg(n_dev) : n_hi;   // exact count from the device in slab mode, else the launch bound
    const int t0 = blockIdx.x * 2 * THREADS;
    if (t0 >= n) return;
    const int t1 = min(n, t0 + 2 * THREADS);
    const int tid = threadIdx.x;
    const int t = blockIdx.x * THREADS + tid;    // pair index
    const int a = 2 * t;
    const bool liveA = a < n, liveB = a + 1 < n;
    uint32_t ca, cb;
    stage_prologue(S, cell_sorted, t0, t1, ca, cb);
    const float4 pa = posq[min(a, t1 - 1)], pb = posq[min(a + 1, t1 - 1)];
    const int cza = (int)(ca % (uint32_t)G.nz), czb = (int)(cb % (uint32_t)G.nz);
    unsigned* const mrow = nmask + ((size_t)(t >> 5) * MR) * 32 + (tid & 31);
    int wrow = 0;
    float2 acc = make_float2(0.f, 0.f);
    const float2 HH = make_float2(C.hh, C.hh);
    const float neg_eps = -ST_EPS * C.hh;

    auto pass = [&](const bool useA, const bool useB, const int pf, const int pl) {
        const float2 X = make_float2(useA ? pa.x : ST_FAR, useB ? pb.x : ST_FAR);
        const float2 Y = make_float2(useA ? pa.y : ST_FAR, useB ? pb.y : ST_FAR);
        const float2 Z = make_float2(useA ? pa.z : ST_FAR, useB ? pb.z : ST_FAR);
#pragma unroll 1
        for (int r = 0; r < 9; r++) {
            const int rb = S.rb[r];
            const int s = rb + S.raw[r][pf], e = rb + S.raw[r][pl];
            const float4* cp = S.pos + s;
            const int ngroups = (e - s + ST_G - 1) / ST_G;
            unsigned m = 0;
#pragma unroll 1
            for (int g = 0; g < ngroups; g++, cp += ST_G) {
                unsigned gm = 0;
#pragma unroll
                for (int j = 0; j < ST_G; j++) {
                    const float4 pj = cp[j];
                    const float2 dx = __fadd2_rn(X, make_float2(-pj.x, -pj.x));
                    const float2 dy = __fadd2_rn(Y, make_float2(-pj.y, -pj.y));
                    const float2 dz = __fadd2_rn(Z, make_float2(-pj.z, -pj.z));
                    float2 d2 = __fmul2_rn(dx, dx);
                    d2 = __ffma2_rn(dy, dy, d2);
                    d2 = __ffma2_rn(dz, dz, d2);
                    float2 w = __fadd2_rn(HH, make_float2(-d2.x, -d2.y));
                    // the slots behind a range (two cells away in z, or the sentinels) are > h away: they fail on their own
                    if (fmaxf(w.x, w.y) >= neg_eps) gm |= 1u << j;
                    w.x = fmaxf(w.x, 0.f);
                    w.y = fmaxf(w.y, 0.f);
                    acc = __ffma2_rn(__fmul2_rn(w, w), w, acc);
                }
                m |= gm << ((g & 7) * ST_G);
                if ((g & 7) == 7) { mrow[(size_t)wrow * 32] = m; wrow++; m = 0; }
            }
            if (ngroups & 7) { mrow[(size_t)wrow * 32] = m; wrow++; }
        }
    };
    auto fall = [&](const int which, const int col) {
        const float4 q = which ? pb : pa;
        float s = 0.f;
        walk27_runs(G, cell_start, col / G.ny, col % G.ny, which ? czb : cza, [&](const int k) {
            const float4 pj = __ldg(&posq[k]);
            const float dx = __fsub_rn(q.x, pj.x), dy = __fsub_rn(q.y, pj.y), dz = __fsub_rn(q.z, pj.z);
            const float d2 = __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx)));
            const float w = fmaxf(__fsub_rn(C.hh, d2), 0.f);
            s = __fmaf_rn(__fmul_rn(w, w), w, s);
        });
        if (which) acc.y = s; else acc.x = s;
    };
    stage_run<THREADS, CAPC, false>(S, t0, t1, posq, nullptr, cell_start, G, a, cza, czb, liveA, liveB, pass, fall);

    if (!liveA) return;
    const float ra = acc.x * C.densK;
    const float Pa = C.k * (ra - C.p0);
    rho[a] = ra;
    posq_q[a] = make_float4(pa.x, pa.y, pa.z, Pa / (ra * ra));
    velv[a].w = C.mass / ra;
    if (liveB) {
        const float rb = acc.y * C.densK;
        const float Pb = C.k * (rb - C.p0);
        rho[a + 1] = rb;
        posq_q[a + 1] = make_float4(pb.x, pb.y, pb.z, Pb / (rb * rb));
        velv[a + 1].w = C.mass / rb;
    }
}

// ------------------------------------------------------------------ passes 2+3 + integrate + collide over the masks
template <int THREADS, int CAPC, int MINB, bool DIAG>
__global__ void __launch_bounds__(THREADS, MINB)
k_force_stage(int n_hi, const int* __restrict__ n_dev, const float4* __restrict__ posq_q, const float4* __restrict__ velv,
              const float* __restrict__ rho, const int* __restrict__ ids, const uint32_t* __restrict__ cell_sorted,
              const int* __restrict__ cell_start, GridP G, StepC C, const unsigned* __restrict__ nmask,
              float4* __restrict__ posq_out, float4* __restrict__ velv_out, DiagOut D) {
    using SS = StageShared<THREADS, CAPC, true>;
    extern __shared__ __align__(16) unsigned char st_smem[];
    SS& S = *reinterpret_cast<SS*>(st_smem);
    constexpr int MR = stage_mask_rows(CAPC);
    const int n = n_dev ? __ldg(n_dev) : n_hi;
    const int t0 = blockIdx.x * 2 * THREADS;
    if (t0 >= n) return;
    const int t1 = min(n, t0 + 2 * THREADS);
    const int tid = threadIdx.x;
    const int t = blockIdx.x * THREADS + tid;
    const int a = 2 * t;
    const bool liveA = a < n, liveB = a + 1 < n;
    uint32_t ca, cb;
    stage_prologue(S, cell_sorted, t0, t1, ca, cb);
    const int ia = min(a, t1 - 1), ib = min(a + 1, t1 - 1);
    const float4 pa = posq_q[ia], pb = posq_q[ib];
    const float4 va = velv[ia], vb = velv[ib];
    const int cza = (int)(ca % (uint32_t)G.nz), czb = (int)(cb % (uint32_t)G.nz);
    const unsigned* const mrow = nmask + ((size_t)(t >> 5) * MR) * 32 + (tid & 31);
    int wrow = 0;
    const float inv_sqrt3 = 0.57735026f;   // glm::normalize(vec3(1)) = 1 * (1/sqrt(3)), fluid_system.h:439

    float2 A_x = {0.f, 0.f}, A_y = {0.f, 0.f}, A_z = {0.f, 0.f};   // sum (q_i+q_j) (h-r)^2 dir
    float2 F_x = {0.f, 0.f}, F_y = {0.f, 0.f}, F_z = {0.f, 0.f};   // sum (v_j-v_i) vol_j (h-r)
    float2 N_x = {0.f, 0.f}, N_y = {0.f, 0.f}, N_z = {0.f, 0.f};   // sum vol_j (h^2-r^2)^2 d
    float2 CF = {0.f, 0.f};                                         // sum vol_j (h^2-r^2)(3h^2-7r^2)   (self included)
    const float2 Q = make_float2(pa.w, pb.w);
    const float2 VX = make_float2(va.x, vb.x), VY = make_float2(va.y, vb.y), VZ = make_float2(va.z, vb.z);
    const float2 HH = make_float2(C.hh, C.hh), H1 = make_float2(C.h, C.h), HH3 = make_float2(C.hh3, C.hh3);
    float2 X, Y, Z;

    // one neighbour against both targets (packed); selfA / selfB: the pressure sum excludes j == i (fluid_system.h:142)
    auto body = [&](const bool selfA, const bool selfB, const float4 pj, const float4 vj) {
        const float2 dx = __fadd2_rn(X, make_float2(-pj.x, -pj.x));
        const float2 dy = __fadd2_rn(Y, make_float2(-pj.y, -pj.y));
        const float2 dz = __fadd2_rn(Z, make_float2(-pj.z, -pj.z));
        float2 d2 = __fmul2_rn(dx, dx);
        d2 = __ffma2_rn(dy, dy, d2);
        d2 = __ffma2_rn(dz, dz, d2);
        float2 w = __fadd2_rn(HH, make_float2(-d2.x, -d2.y));
        w.x = fmaxf(w.x, 0.f); w.y = fmaxf(w.y, 0.f);
        const float2 vol = make_float2(vj.w, vj.w);
        const float2 vw = __fmul2_rn(w, vol);
        const float2 t7 = __ffma2_rn(d2, make_float2(-7.0f, -7.0f), HH3);
        CF = __ffma2_rn(vw, t7, CF);
        const float2 vww = __fmul2_rn(vw, w);
        N_x = __ffma2_rn(vww, dx, N_x); N_y = __ffma2_rn(vww, dy, N_y); N_z = __ffma2_rn(vww, dz, N_z);
        // d2 is clamped away from 0, so the flush-to-zero approximation never sees a denormal
        const float2 rinv = make_float2(rsqrt_ftz(fmaxf(d2.x, 1e-30f)), rsqrt_ftz(fmaxf(d2.y, 1e-30f)));
        const float2 r = __fmul2_rn(d2, rinv);
        float2 hm = __fadd2_rn(H1, make_float2(-r.x, -r.y));
        hm.x = fmaxf(hm.x, 0.f); hm.y = fmaxf(hm.y, 0.f);
        const float2 tv = __fmul2_rn(hm, vol);
        const float2 dvx = __fadd2_rn(make_float2(vj.x, vj.x), make_float2(-VX.x, -VX.y));
        const float2 dvy = __fadd2_rn(make_float2(vj.y, vj.y), make_float2(-VY.x, -VY.y));
        const float2 dvz = __fadd2_rn(make_float2(vj.z, vj.z), make_float2(-VZ.x, -VZ.y));
        F_x = __ffma2_rn(tv, dvx, F_x); F_y = __ffma2_rn(tv, dvy, F_y); F_z = __ffma2_rn(tv, dvz, F_z);
        const float2 sq = __fadd2_rn(Q, make_float2(pj.w, pj.w));
        float2 sc = __fmul2_rn(__fmul2_rn(sq, hm), hm);
        if (selfA) sc.x = 0.f;
        if (selfB) sc.y = 0.f;
        float2 ux = __fmul2_rn(dx, rinv), uy = __fmul2_rn(dy, rinv), uz = __fmul2_rn(dz, rinv);
        if (fminf(r.x, r.y) <= 1e-4f) {   // coincident pair: direction (1,1,1)/sqrt(3) (fluid_system.h:438-440); also the self entry
            if (r.x <= 1e-4f) { ux.x = inv_sqrt3; uy.x = inv_sqrt3; uz.x = inv_sqrt3; }
            if (r.y <= 1e-4f) { ux.y = inv_sqrt3; uy.y = inv_sqrt3; uz.y = inv_sqrt3; }
        }
        A_x = __ffma2_rn(sc, ux, A_x); A_y = __ffma2_rn(sc, uy, A_y); A_z = __ffma2_rn(sc, uz, A_z);
    };

    auto pass = [&](const bool useA, const bool useB, const int pf, const int pl) {
        // mask words of this pass: per run ceil(len / 32), in run order
        int nw = 0;
#pragma unroll
        for (int r = 0; r < 9; r++) nw += (S.raw[r][pl] - S.raw[r][pf] + 31) >> 5;
        if (nw == 0) return;
        X = make_float2(useA ? pa.x : ST_FAR, useB ? pb.x : ST_FAR);
        Y = make_float2(useA ? pa.y : ST_FAR, useB ? pb.y : ST_FAR);
        Z = make_float2(useA ? pa.z : ST_FAR, useB ? pb.z : ST_FAR);
        const int sa = useA ? S.rb[4] + a : -1, sb = useB ? S.rb[4] + a + 1 : -1;   // the targets' own stage slots
        const unsigned* mp = mrow + (size_t)wrow * 32;
        wrow += nw;
        unsigned m = 0u, mnext = mp[0];    // the next word is always in flight
        int wi = -1, r = -1, wleft = 0, base = 0;
        // next set bit of the lane's masks -> stage slot, -1 when the pass is exhausted
        auto next_slot = [&]() -> int {
            while (m == 0u) {
                if (++wi >= nw) return -1;
                m = mnext;
                mnext = (wi + 1 < nw) ? mp[(size_t)(wi + 1) * 32] : 0u;
                if (--wleft > 0) base += 32;
                else {
                    do {
                        r++;
                        base = S.rb[r] + S.raw[r][pf];
                        wleft = (S.rb[r] + S.raw[r][pl] - base + 31) >> 5;
                    } while (wleft == 0);
                }
            }
            const int bit = __ffs((int)m) - 1;
            m &= m - 1u;
            return base + bit;
        };
        int slot = next_slot();
        if (slot < 0) return;
        float4 pj = S.pos[slot], vj = S.vel[slot];
#pragma unroll 1
        for (;;) {
            const int nslot = next_slot();
            const int ls = nslot < 0 ? slot : nslot;
            const float4 pjn = S.pos[ls], vjn = S.vel[ls];   // the next neighbour's record is in flight during the body
            body(slot == sa, slot == sb, pj, vj);
            if (nslot < 0) break;
            slot = nslot; pj = pjn; vj = vjn;
        }
    };
    auto fall = [&](const int which, const int col) {
        X = make_float2(which ? ST_FAR : pa.x, which ? pb.x : ST_FAR);
        Y = make_float2(which ? ST_FAR : pa.y, which ? pb.y : ST_FAR);
        Z = make_float2(which ? ST_FAR : pa.z, which ? pb.z : ST_FAR);
        const float4 q = which ? pb : pa;
        const int self = which ? a + 1 : a;
        const float neg_eps = -ST_EPS * C.hh;
        walk27_runs(G, cell_start, col / G.ny, col % G.ny, which ? czb : cza, [&](const int k) {
            const float4 pj = __ldg(&posq_q[k]);
            const float dx = __fsub_rn(q.x, pj.x), dy = __fsub_rn(q.y, pj.y), dz = __fsub_rn(q.z, pj.z);
            const float d2 = __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx)));
            if (__fsub_rn(C.hh, d2) >= neg_eps) body(!which && k == self, which && k == self, pj, __ldg(&velv[k]));
        });
    };
    stage_run<THREADS, CAPC, true>(S, t0, t1, posq_q, velv, cell_start, G, a, cza, czb, liveA, liveB, pass, fall);

    if (!liveA) return;
    // DIAG: the reference's per-particle debug fields, by id; NeighbId comes from k_neighb_id (sph.cu)
    force_epilogue<DIAG>(a, pa, va, rho[a], A_x.x, A_y.x, A_z.x, F_x.x, F_y.x, F_z.x, N_x.x, N_y.x, N_z.x, CF.x, -1, C, ids, posq_out, velv_out, D);
    if (liveB)
        force_epilogue<DIAG>(a + 1, pb, vb, rho[a + 1], A_x.y, A_y.y, A_z.y, F_x.y, F_y.y, F_z.y, N_x.y, N_y.y, N_z.y, CF.y, -1, C, ids, posq_out, velv_out, D);
}

// ------------------------------------------------------------------ test hook: decode the production masks
// For every target: the sorted indices k of the candidates whose bit is set in a pass that served the target
// (out[target * cap + i], counts[target]; counts = -1 for a target that took the fallback walk).  Same segment loop,
// same ranges, same rows as k_force_stage.
template <int THREADS, int CAPC>
__global__ void __launch_bounds__(THREADS)
k_stage_decode(int n, const float4* __restrict__ posq, const uint32_t* __restrict__ cell_sorted, const int* __restrict__ cell_start,
               GridP G, const unsigned* __restrict__ nmask, int cap, int* __restrict__ counts, int* __restrict__ out) {
    using SS = StageShared<THREADS, CAPC, false>;
    extern __shared__ __align__(16) unsigned char st_smem[];
    SS& S = *reinterpret_cast<SS*>(st_smem);
    constexpr int MR = stage_mask_rows(CAPC);
    const int t0 = blockIdx.x * 2 * THREADS;
    if (t0 >= n) return;
    const int t1 = min(n, t0 + 2 * THREADS);
    const int tid = threadIdx.x;
    const int t = blockIdx.x * THREADS + tid;
    const int a = 2 * t;
    const bool liveA = a < n, liveB = a + 1 < n;
    uint32_t ca, cb;
    stage_prologue(S, cell_sorted, t0, t1, ca, cb);
    const int cza = (int)(ca % (uint32_t)G.nz), czb = (int)(cb % (uint32_t)G.nz);
    const unsigned* const mrow = nmask + ((size_t)(t >> 5) * MR) * 32 + (tid & 31);
    int wrow = 0, na = 0, nb = 0;
    auto pass = [&](const bool useA, const bool useB, const int pf, const int pl) {
        for (int r = 0; r < 9; r++) {
            const int k0 = S.raw[r][pf], len = S.raw[r][pl] - k0;
            const int nw = (len + 31) >> 5;
            for (int w = 0; w < nw; w++) {
                unsigned m = mrow[(size_t)(wrow + w) * 32];
                while (m) {
                    const int k = k0 + 32 * w + __ffs((int)m) - 1;   // sorted index of the candidate
                    m &= m - 1u;
                    if (useA) { if (na < cap) out[(size_t)a * cap + na] = k; na++; }
                    if (useB) { if (nb < cap) out[(size_t)(a + 1) * cap + nb] = k; nb++; }
                }
            }
            wrow += nw;
        }
    };
    auto fall = [&](const int which, const int) { if (which) nb = -1; else na = -1; };
    stage_run<THREADS, CAPC, false>(S, t0, t1, posq, nullptr, cell_start, G, a, cza, czb, liveA, liveB, pass, fall);
    if (liveA) counts[a] = na;
    if (liveB) counts[a + 1] = nb;
}

// ------------------------------------------------------------------ launch wrappers
constexpr int STG_THREADS = 64;
constexpr int STG_CAPC = 1536;

int stage_pairs_pad(int n) { return (((n + 1) / 2) + STG_THREADS - 1) / STG_THREADS * STG_THREADS; }
size_t stage_mask_words(int n) { return (size_t)stage_pairs_pad(n) * stage_mask_rows(STG_CAPC); }

void launch_density_stage(cudaStream_t st, int n, const int* n_dev, const float4* posq, float4* posq_q, float4* velv,
                          const uint32_t* cell_sorted, const int* cell_start, const GridP& G, const StepC& C, float* rho, unsigned* nmask) {
    if (n <= 0) return;
    auto kern = k_density_stage<STG_THREADS, STG_CAPC, 8>;
    const int smem = (int)sizeof(StageShared<STG_THREADS, STG_CAPC, false>);
    if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    kern<<<stage_pairs_pad(n) / STG_THREADS, STG_THREADS, smem, st>>>(n, n_dev, posq, posq_q, velv, cell_sorted, cell_start, G, C, rho, nmask);
}

void launch_force_stage(cudaStream_t st, int n, const int* n_dev, const float4* posq_q, const float4* velv, const float* rho,
                        const int* ids, const uint32_t* cell_sorted, const int* cell_start, const GridP& G, const StepC& C,
                        const unsigned* nmask, float4* posq_out, float4* velv_out, const DiagOut* diag) {
    if (n <= 0) return;
    const int smem = (int)sizeof(StageShared<STG_THREADS, STG_CAPC, true>);
    const dim3 g(stage_pairs_pad(n) / STG_THREADS), b(STG_THREADS);
    // the opt-in is per device (one process may drive several GPUs): set it before every launch, it is cheap
    if (diag) {
        auto kern = k_force_stage<STG_THREADS, STG_CAPC, 4, true>;
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        kern<<<g, b, smem, st>>>(n, n_dev, posq_q, velv, rho, ids, cell_sorted, cell_start, G, C, nmask, posq_out, velv_out, *diag);
        launch_neighb_id(st, n, n_dev, posq_q, ids, cell_sorted, cell_start, G, C, diag->neighb);
    } else {
        auto kern = k_force_stage<STG_THREADS, STG_CAPC, 4, false>;
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        kern<<<g, b, smem, st>>>(n, n_dev, posq_q, velv, rho, ids, cell_sorted, cell_start, G, C, nmask, posq_out, velv_out, DiagOut{});
    }
}

void launch_stage_decode(cudaStream_t st, int n, const float4* posq, const uint32_t* cell_sorted, const int* cell_start,
                         const GridP& G, const unsigned* nmask, int cap, int* counts, int* out) {
    if (n <= 0) return;
    auto kern = k_stage_decode<STG_THREADS, STG_CAPC>;
    const int smem = (int)sizeof(StageShared<STG_THREADS, STG_CAPC, false>);
    kern<<<stage_pairs_pad(n) / STG_THREADS, STG_THREADS, smem, st>>>(n, posq, cell_sorted, cell_start, G, nmask, cap, counts, out);
}

}  // namespace sphe
