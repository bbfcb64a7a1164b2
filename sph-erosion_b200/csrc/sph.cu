// SPH neighbour passes over the sorted SoA arrays.
//
//   k_density : pass 1 of FluidSystemSPH::Run (Erosion/fluid_system.h:108-124, kernDefault :410-416)
//   k_force   : passes 2+3 fused (:128-178; gradPressure :430-447, laplVisc :449-453,
//               gradDefault :418-422, laplDefault :424-428) + advance() (:306-353) + collisionS (:355-407)
//
// Pass 3 only needs particle i's own normal from pass 2, so 2 and 3 share one neighbour walk
// (SURVEY.md section 3.1).  advance() overwrites what neighbours still read, so the new state goes to
// a second buffer.
//
// Arithmetic: the neighbour PREDICATE is bit-exact (dist2_exact + threshold T, see common.cuh).  The
// sums use FMA and factor the kernel constants out of the loop; they match the reference within the
// fp32 tolerance stated in tests/test_gpu_parity.py (every SPH kernel vanishes at r = h, so the sums
// are continuous in the predicate).
#include "common.cuh"
#include "sim.h"
#include "sph_device.cuh"

namespace sphe {

// Visit the candidates of a particle in cell c in canonical grid-walk order: dx = -1..1, dy = -1..1,
// then the contiguous sorted range covering cells cz-1..cz+1 of column (cx+dx, cy+dy).
// Identical to WALK_BEGIN/WALK_END in oracle/sph_oracle.c.
template <class F>
__device__ __forceinline__ void walk27(const GridP& G, const int* __restrict__ cell_start, uint32_t c, F&& f) {
    int cz = (int)(c % (uint32_t)G.nz);
    uint32_t t = c / (uint32_t)G.nz;
    int cy = (int)(t % (uint32_t)G.ny);
    int cx = (int)(t / (uint32_t)G.ny);
    int z0 = cz > 0 ? cz - 1 : 0;
    int z1 = cz < G.nz - 1 ? cz + 1 : cz;
#pragma unroll 1
    for (int dx = -1; dx <= 1; dx++) {
        int x = cx + dx;
        if (x < 0 || x >= G.nx) continue;
#pragma unroll 1
        for (int dy = -1; dy <= 1; dy++) {
            int y = cy + dy;
            if (y < 0 || y >= G.ny) continue;
            int base = (x * G.ny + y) * G.nz;
            int s = __ldg(&cell_start[base + z0]);
            int e = __ldg(&cell_start[base + z1 + 1]);
            for (int k = s; k < e; k++) f(k);
        }
    }
}

// ------------------------------------------------------------------ pass 1: density + pressure
// Writes rho, and packs what pass 2 needs per neighbour into the arrays it will stream anyway:
//   posq_q[i] = (x, y, z, P_i / rho_i^2)      velv[i].w = mass / rho_i
__global__ void __launch_bounds__(128) k_density_tpp(int n_hi, const int* __restrict__ n_dev, const float4* __restrict__ posq, float4* __restrict__ posq_q,
                                                     float4* __restrict__ velv, const uint32_t* __restrict__ cell_sorted,
                                                     const int* __restrict__ cell_start, GridP G, StepC C,
                                                     float* __restrict__ rho) {
    const int n = n_dev ? __ldg(n_dev) : n_hi;  // exact count from the device in slab mode, else the launch bound
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 pi = posq[i];
    float sum = 0.0f;
    walk27(G, cell_start, cell_sorted[i], [&](int k) {
        float4 pj = __ldg(&posq[k]);
        float d2 = dist2_exact(pi.x - pj.x, pi.y - pj.y, pi.z - pj.z);
        if (d2 <= C.T) {
            float w = C.hh - d2;
            sum = fmaf(w * w, w, sum);
        }
    });
    float r = C.densK * sum;
    float P = C.k * (r - C.p0);
    rho[i] = r;
    posq_q[i] = make_float4(pi.x, pi.y, pi.z, P / (r * r));
    velv[i].w = C.mass / r;
}

// ------------------------------------------------------------------ passes 2+3 + integrate + collide
template <bool DIAG>
__global__ void __launch_bounds__(128) k_force_tpp(int n_hi, const int* __restrict__ n_dev, const float4* __restrict__ posq_q, const float4* __restrict__ velv,
                                                   const float* __restrict__ rho, const int* __restrict__ ids,
                                                   const uint32_t* __restrict__ cell_sorted,
                                                   const int* __restrict__ cell_start, GridP G, StepC C,
                                                   float4* __restrict__ posq_out, float4* __restrict__ velv_out, DiagOut D) {
    const int n = n_dev ? __ldg(n_dev) : n_hi;  // exact count from the device in slab mode, else the launch bound
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 pi = posq_q[i];
    const float4 vi = velv[i];
    const float rho_i = rho[i];
    const float inv_sqrt3 = 0.57735026f;  // glm::normalize(vec3(1)) = 1 * (1/sqrt(3)), fluid_system.h:439

    float ax = 0.f, ay = 0.f, az = 0.f;     // sum (q_i+q_j) (h-r)^2 dir
    float fx = 0.f, fy = 0.f, fz = 0.f;     // sum (v_j-v_i) vol_j (h-r)
    float nx = 0.f, ny = 0.f, nz = 0.f;     // sum vol_j (h^2-r^2)^2 d
    float cf = 0.f;                         // sum vol_j (h^2-r^2)(3h^2-7r^2)   (self included)
    int maxid = -1;

    walk27(G, cell_start, cell_sorted[i], [&](int k) {
        float4 pj = __ldg(&posq_q[k]);
        float dx = pi.x - pj.x, dy = pi.y - pj.y, dz = pi.z - pj.z;
        float d2 = dist2_exact(dx, dy, dz);
        if (d2 <= C.T) {
            float4 vj = __ldg(&velv[k]);
            float w = C.hh - d2;
            float vw = vj.w * w;
            cf = fmaf(vw, fmaf(-7.0f, d2, C.hh3), cf);
            float vww = vw * w;
            nx = fmaf(vww, dx, nx); ny = fmaf(vww, dy, ny); nz = fmaf(vww, dz, nz);
            float rinv = rsqrtf(fmaxf(d2, 1e-30f));
            float r = d2 * rinv;
            float hm = C.h - r;
            float tv = vj.w * hm;
            fx = fmaf(tv, vj.x - vi.x, fx); fy = fmaf(tv, vj.y - vi.y, fy); fz = fmaf(tv, vj.z - vi.z, fz);
            // pressure: j != i only (fluid_system.h:142); coincident pairs use direction (1,1,1)/sqrt(3)
            float s = (k != i) ? (pi.w + pj.w) * hm * hm : 0.0f;
            bool tiny = r <= 1e-4f;  // == (double)r < 10e-5, fluid_system.h:438
            float ux = tiny ? inv_sqrt3 : dx * rinv;
            float uy = tiny ? inv_sqrt3 : dy * rinv;
            float uz = tiny ? inv_sqrt3 : dz * rinv;
            ax = fmaf(s, ux, ax); ay = fmaf(s, uy, ay); az = fmaf(s, uz, az);
            if (DIAG) { if (k != i) maxid = max(maxid, __ldg(&ids[k])); }
        }
    });

    force_epilogue<DIAG>(i, pi, vi, rho_i, ax, ay, az, fx, fy, fz, nx, ny, nz, cf, maxid, C, ids, posq_out, velv_out, D);
}

// ------------------------------------------------------------------ the default passes: test ONCE, pair lists in HBM
//   k_density_list : pass 1 + the candidate test for BOTH passes.  Two targets per thread (consecutive particles of the sorted
//                    order, i.e. the same or adjacent cells), packed math (FADD2 / FMUL2 / FFMA2: one instruction serves both
//                    targets), branch-free append of every candidate within h of either target to a per-thread list staged in
//                    shared memory, flushed to HBM entry-major ([entry][pair]: a row is contiguous, fully coalesced).
//   k_force_list   : passes 2+3 + integrate + box (+ the terrain cull), walking the stored list only -- no candidate test.
// Cost: one 4-byte write + read per stored entry (about 50 per pair at 25 neighbours per particle).
// (History -- thread per particle, pair kernels without lists, 4 targets per thread, lane-strided lists: csrc/experiments,
// DESIGN.md section 3.)
constexpr int NLIST_CAP = 64;       // nominal entries per pair at the base level (sizing levels: 64 / 128 / 256)
// A candidate enters a pair list when hh - d2 >= -2^-20 hh with the FMA-contracted d2.  The reference predicate is
// sqrt(d2_exact) <= h  <=>  d2_exact <= T with T within 1 ulp of hh, and the contracted d2 is within 3 ulp of the exact one,
// so the lists are a SUPERSET of the exact neighbour sets; the extra entries have clamped weights of 0 or a few ulp
// (tests/test_gpu_parity.py::test_pair_masks_cover_the_exact_neighbour_sets reads the production lists back).
constexpr float LIST_NEG_EPS = -9.5367431640625e-07f;

// Template parameters of k_density_list:
// REC  = true: instead of (posC, vel.w) the pass writes ONE interleaved 32-byte record per particle for the force pass
//        (one 256-bit gather instead of two 128-bit ones).  Measured neutral (the gathers are latency bound); not instantiated
//        by the default dispatch.
// PF   = true (density variant 6, the default): software-pipelined candidate stream -- the next four candidates and the next
//        run's bounds are in flight while the current four are processed (ncu source view of the plain kernel: 45 % of the
//        stall samples on the first use of the gathered positions, 9 % on the run bounds).
// CAP / THREADS: entries per pair staged in shared memory and CTA size; the list is (CAP + 1) * THREADS ints.  Longer lists
//        spill to their HBM rows directly (see `spill`), lists beyond the allocated rows fall back to a direct walk in the
//        force pass; the host raises the level (128 / 256 staged entries: k_density_list16) when the spill counters say so
//        (api.cu step_device).
// Base level: 56 staged entries per pair, 64-thread CTAs, 14 CTAs/SM = 28 warps at 72 registers and 14.6 KB of shared memory
// per CTA.  With the 64-entry stage of round 1 (33 KB per 128-thread CTA, 80 registers) 24 warps fit -- c3 density pass
// 0.644 -> 0.616 ms; 48 entries at 16 CTAs/SM spill too often (0.657 ms).  A/B: SPHE_NVCC_EXTRA="-DDL_CAP=.. -DDL_THREADS=.. -DDL_MINB=.."
#ifndef DL_LD256
#define DL_LD256 1
#endif
struct __align__(32) F8 { float4 a, b; };
// ld.global.nc.v8.f32 (LDG.E.256.CONSTANT on sm_100a): p must be 32-byte aligned
__device__ __forceinline__ F8 ldg256(const float4* p) {
    F8 r;
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(r.a.x), "=f"(r.a.y), "=f"(r.a.z), "=f"(r.a.w), "=f"(r.b.x), "=f"(r.b.y), "=f"(r.b.z), "=f"(r.b.w) : "l"(p));
    return r;
}
#ifndef DL_CAP
#define DL_CAP 56
#define DL_THREADS 64
#define DL_MINB 14
#endif
template <bool REC, bool PF, int CAP, int THREADS, int UNROLL = 4>
__global__ void __launch_bounds__(THREADS, CAP == DL_CAP ? DL_MINB : (CAP >= 256 ? 3 : 1)) k_density_list(int n_hi, const int* __restrict__ n_dev, int npairs_pad, const float4* __restrict__ posq,
                                                          float4* __restrict__ posq_q, float4* __restrict__ velv,
                                                          const uint32_t* __restrict__ cell_sorted,
                                                          const int* __restrict__ cell_start, GridP G, StepC C,
                                                          float* __restrict__ rho, int* __restrict__ nlist,
                                                          int2* __restrict__ ncount, int* __restrict__ overflow, int rows) {
    constexpr int NLIST_CAP = CAP, NLIST_THREADS = THREADS;   // shadow the file-scope defaults
    const int n = n_dev ? __ldg(n_dev) : n_hi;  // exact count from the device in slab mode, else the launch bound
    // (CAP + 1) * THREADS ints, +1: trash slot for saturated appends.  Static when it fits the 48 KB static limit:
    // with a dynamic array ptxas does not know the CTAs/SM bound and settles for 48 registers (serialised gathers,
    // measured 0.23 ms instead of 0.15 ms at 1M particles)
    constexpr bool STATIC_LIST = (CAP + 1) * THREADS * sizeof(int) <= 48 * 1024;
    __shared__ int list_static[STATIC_LIST ? (CAP + 1) * THREADS : 1];
    extern __shared__ int list_dynamic[];
    int* const list = STATIC_LIST ? list_static : list_dynamic;
    const int tid = threadIdx.x;
    const int t = blockIdx.x * blockDim.x + tid;
    int a = 2 * t;
    const bool live = a < n;
    if (!live) a = 0;
    const int b = (a + 1 < n) ? a + 1 : a;
    const float4 pa = posq[a], pb = posq[b];
    const uint32_t ca = cell_sorted[a], cb = cell_sorted[b];
    const uint32_t cola = ca / (uint32_t)G.nz, colb = cb / (uint32_t)G.nz;
    const int cza = (int)(ca - cola * (uint32_t)G.nz), czb = (int)(cb - colb * (uint32_t)G.nz);
    const bool merged = (b != a) && (cola == colb) && (czb - cza <= 3);
    const float FAR = 1.0e18f;
    const int npass = !live ? 0 : ((merged || b == a) ? 1 : 2);

    float2 acc = make_float2(0.f, 0.f);
    int* const lbase = list + tid;
    // The append cursor is kept as a SHARED-MEMORY BYTE ADDRESS (wp): per candidate the append is then VIMNMX + STS +
    // one predicated add instead of min / add / shift-add / store / add / select (SASS: 27 -> 22 instructions per
    // candidate for the two targets).  `off` (slot * THREADS, as everywhere else) is derived where it is needed.
    const unsigned lbase_sa = (unsigned)__cvta_generic_to_shared(lbase);
    const unsigned cap_sa = lbase_sa + 4u * NLIST_CAP * NLIST_THREADS;
    unsigned wp = lbase_sa;
    auto OFF = [&]() { return (int)((wp - lbase_sa) >> 2); };
    int off0 = 0;  // slot * NLIST_THREADS
#pragma unroll 1
    for (int p = 0; p < npass; p++) {
        const bool useA = merged || p == 0, useB = merged || p == 1;
        const float2 X = make_float2(useA ? pa.x : FAR, useB ? pb.x : FAR);
        const float2 Y = make_float2(useA ? pa.y : FAR, useB ? pb.y : FAR);
        const float2 Z = make_float2(useA ? pa.z : FAR, useB ? pb.z : FAR);
        const uint32_t col = p ? colb : cola;
        const int czlo = p ? czb : cza, czhi = merged ? czb : czlo;
        const int cy = (int)(col % (uint32_t)G.ny), cx = (int)(col / (uint32_t)G.ny);
        const int z0 = czlo > 0 ? czlo - 1 : 0, z1 = czhi < G.nz - 1 ? czhi + 1 : czhi;
        if (p == 1) off0 = OFF();
        auto test = [&](const int k, const float4 pj) {
            float2 dx = __fadd2_rn(X, make_float2(-pj.x, -pj.x));
            float2 dy = __fadd2_rn(Y, make_float2(-pj.y, -pj.y));
            float2 dz = __fadd2_rn(Z, make_float2(-pj.z, -pj.z));
            float2 d2 = __fmul2_rn(dx, dx);
            d2 = __ffma2_rn(dy, dy, d2);
            d2 = __ffma2_rn(dz, dz, d2);
            float2 w = __fadd2_rn(make_float2(C.hh, C.hh), make_float2(-d2.x, -d2.y));
            // branch-free append: store at the current slot (the trash slot once saturated), advance only on a pass
            asm volatile("st.shared.u32 [%0], %1;" ::"r"(min(wp, cap_sa)), "r"(k) : "memory");
            asm("{ .reg .pred q; setp.ge.f32 q, %1, %3; @q add.u32 %0, %0, %2; }" : "+r"(wp) : "f"(fmaxf(w.x, w.y)), "n"(4 * NLIST_THREADS), "f"(LIST_NEG_EPS * C.hh));
            // clamp without the two FMNMX: w + |w| = 2 max(w, 0) EXACTLY, in one packed add (FADD2 takes |.| on an operand);
            // the sum then carries a factor 2^3 that the final scale removes exactly -- bit-identical to clamping
            w = __fadd2_rn(w, make_float2(fabsf(w.x), fabsf(w.y)));
            acc = __ffma2_rn(__fmul2_rn(w, w), w, acc);
        };
        // SPILL: the shared-memory list holds CAP entries; when a run saturates it, that run is walked again (rare,
        // outside the hot loop) and the entries with index >= CAP go straight to their rows in HBM, up to the `rows`
        // the host allocated.  Same candidates, same order, same predicate: the list the force pass reads is exactly
        // the one an unbounded list would have held, so results do not depend on the capacity.
        auto spill = [&](const int s, const int e, int idx) {
#pragma unroll 4
            for (int k = s; k < e; k++) {
                const float4 pj = __ldg(&posq[k]);
                float2 dx = __fadd2_rn(X, make_float2(-pj.x, -pj.x));
                float2 dy = __fadd2_rn(Y, make_float2(-pj.y, -pj.y));
                float2 dz = __fadd2_rn(Z, make_float2(-pj.z, -pj.z));
                float2 d2 = __fmul2_rn(dx, dx);
                d2 = __ffma2_rn(dy, dy, d2);
                d2 = __ffma2_rn(dz, dz, d2);
                float2 w = __fadd2_rn(make_float2(C.hh, C.hh), make_float2(-d2.x, -d2.y));
                if (fmaxf(w.x, w.y) >= LIST_NEG_EPS * C.hh) {
                    if (idx >= NLIST_CAP && idx < rows) nlist[(size_t)idx * npairs_pad + t] = k;
                    idx++;
                }
            }
        };
        auto bounds = [&](const int r, int& s, int& e) {
            int x = cx + r / 3 - 1, y = cy + r % 3 - 1;
            bool ok = r < 9 && x >= 0 && x < G.nx && y >= 0 && y < G.ny;
            int base = ok ? (x * G.ny + y) * G.nz : 0;
            s = __ldg(&cell_start[base + z0]);
            e = ok ? __ldg(&cell_start[base + z1 + 1]) : s;   // empty range for a column outside the grid
        };
        if (!PF) {
#pragma unroll 1
            for (int r = 0; r < 9; r++) {
                int s, e;
                bounds(r, s, e);
                const int off_run = OFF();
#pragma unroll UNROLL
                for (int k = s; k < e; k++) test(k, __ldg(&posq[k]));
                if (wp > cap_sa) spill(s, e, off_run / NLIST_THREADS);
            }
        } else {
            int s, e, sn, en;
            bounds(0, s, e);
#pragma unroll 1
            for (int r = 0; r < 9; r++) {
                bounds(r + 1, sn, en);   // next run's bounds in flight (r + 1 == 9 yields an empty range)
                const int off_run = OFF();
                int k = s;
#if DL_LD256
                // 256-bit candidate loads: two consecutive candidates (32 B, k even) per request.  The pass keeps the L1 data
                // pipe 75 % busy with ~5.6 wavefronts per 128-bit request (the lanes of a warp walk runs staggered by their
                // cells); a request costs its wavefronts whatever its width, so half the requests is half the wavefronts.
                // Same candidates in the same order: lists and sums are bit-identical to the 128-bit path.
                if ((k & 1) && k < e) { test(k, __ldg(&posq[k])); k++; }
                if (k + 4 <= e) {
                    F8 q01 = ldg256(&posq[k]), q23 = ldg256(&posq[k + 2]);
#pragma unroll 1
                    for (; k + 8 <= e; k += 4) {
                        const F8 n01 = ldg256(&posq[k + 4]), n23 = ldg256(&posq[k + 6]);
                        test(k, q01.a); test(k + 1, q01.b); test(k + 2, q23.a); test(k + 3, q23.b);
                        q01 = n01; q23 = n23;
                    }
                    test(k, q01.a); test(k + 1, q01.b); test(k + 2, q23.a); test(k + 3, q23.b);
                    k += 4;
                }
                if (k + 2 <= e) {
                    const F8 q = ldg256(&posq[k]);
                    test(k, q.a); test(k + 1, q.b);
                    k += 2;
                }
#else
                if (k + 4 <= e) {
                    // (a ping-pong version without the register rotation needs 88 registers -> 5 CTAs/SM: no faster, measured)
                    float4 q0 = __ldg(&posq[k]), q1 = __ldg(&posq[k + 1]), q2 = __ldg(&posq[k + 2]), q3 = __ldg(&posq[k + 3]);
#pragma unroll 1
                    for (; k + 8 <= e; k += 4) {
                        const float4 n0 = __ldg(&posq[k + 4]), n1 = __ldg(&posq[k + 5]), n2 = __ldg(&posq[k + 6]), n3 = __ldg(&posq[k + 7]);
                        test(k, q0); test(k + 1, q1); test(k + 2, q2); test(k + 3, q3);
                        q0 = n0; q1 = n1; q2 = n2; q3 = n3;
                    }
                    test(k, q0); test(k + 1, q1); test(k + 2, q2); test(k + 3, q3);
                    k += 4;
                }
#endif
#pragma unroll 1
                for (; k < e; k++) test(k, __ldg(&posq[k]));
                if (wp > cap_sa) spill(s, e, off_run / NLIST_THREADS);
                s = sn; e = en;
            }
        }
    }
    const int off = OFF();
    if (npass == 1) off0 = off;
    // coalesced flush of the lists: entry e of all threads of the block is one contiguous row
    const int cnt = off / NLIST_THREADS;
    const bool fits = cnt <= rows;
    if (live) ncount[t] = fits ? make_int2(off0 / NLIST_THREADS, cnt) : make_int2(-1, -1);
    if (live && !fits) atomicAdd(overflow, 1);                          // beyond the allocated rows: direct walk in the force pass
    if (live && cnt > NLIST_CAP) atomicAdd(overflow + 1, 1);            // spilled (statistics for the host's choice of CAP)
    if (live && cnt > NLIST_CAP / 2) atomicAdd(overflow + 2, 1);        // would spill at the next smaller CAP
    if (t == 0) { overflow[3] = (NLIST_CAP == DL_CAP) ? 64 : NLIST_CAP; overflow[4] = rows; }   // which sizing LEVEL (64 / 128 / 256) these counts belong to (the host reads them late)
    if (fits) {
        int* dst = nlist + t;
        const int stop = min(off, NLIST_CAP * NLIST_THREADS);           // the rest is already in place
#pragma unroll 4
        for (int o = 0, e = 0; o < stop; o += NLIST_THREADS, e++) dst[(size_t)e * npairs_pad] = lbase[o];
    }
    if (!live) return;
    const float dk = C.densK * 0.125f;   // exact: the accumulators hold 8 x sum (h^2 - r^2)^3 (see `test`)
    float ra = acc.x * dk, rb = acc.y * dk;
    float Pa = C.k * (ra - C.p0), Pb = C.k * (rb - C.p0);
    rho[a] = ra;
    if (REC) {
        // posq_q doubles as the record array (2n float4)
        const float4 va = velv[a];
        posq_q[2 * a] = make_float4(pa.x, pa.y, pa.z, Pa / (ra * ra));
        posq_q[2 * a + 1] = make_float4(va.x, va.y, va.z, C.mass / ra);
    } else {
        posq_q[a] = make_float4(pa.x, pa.y, pa.z, Pa / (ra * ra));
        velv[a].w = C.mass / ra;
    }
    if (b != a) {
        rho[b] = rb;
        if (REC) {
            const float4 vb = velv[b];
            posq_q[2 * b] = make_float4(pb.x, pb.y, pb.z, Pb / (rb * rb));
            posq_q[2 * b + 1] = make_float4(vb.x, vb.y, vb.z, C.mass / rb);
        } else {
            posq_q[b] = make_float4(pb.x, pb.y, pb.z, Pb / (rb * rb));
            velv[b].w = C.mass / rb;
        }
    }
}

// ------------------------------------------------------------------ variant 10: 16-bit list entries, 128 per pair
// Along the dam-break transient the fluid compresses to ~2x rest density (the reference's equation of state is soft,
// k = 3): a third of the pairs then overflow 64-entry lists and the direct-walk fallback of the force pass costs
// 6x (scripts/transient.py).  Doubling the capacity with 32-bit entries doubles the shared memory and halves the
// occupancy of this pass (+30 %).  Here an entry is 16 bits -- run index (4 bits) | offset inside the run (12 bits) --
// so 128 entries per pair fit the SAME 33 KB per CTA; the flush expands them to particle indices with the run starts
// kept in shared memory (9 ints per thread).  A run longer than 4095 particles marks the pair as overflowed.
constexpr int L16_CAP = 128, L16_THREADS = 128;   // the 128-entry level; the 256-entry level runs <256, 64>: 35 KB per CTA, 6 CTAs/SM
// (32-bit entries need 66 KB for 256 x 64: 3 CTAs/SM = 6 warps, the density pass at 115 neighbours was occupancy starved)

template <bool PF, int L16_CAP = sphe::L16_CAP, int L16_THREADS = sphe::L16_THREADS>
__global__ void __launch_bounds__(L16_THREADS) k_density_list16(int n_hi, const int* __restrict__ n_dev, int npairs_pad, const float4* __restrict__ posq,
                                                                float4* __restrict__ posq_q, float4* __restrict__ velv,
                                                                const uint32_t* __restrict__ cell_sorted,
                                                                const int* __restrict__ cell_start, GridP G, StepC C,
                                                                float* __restrict__ rho, int* __restrict__ nlist,
                                                                int2* __restrict__ ncount, int* __restrict__ overflow, int rows) {
    constexpr int T = L16_THREADS;
    const int n = n_dev ? __ldg(n_dev) : n_hi;
    __shared__ unsigned short list16[(L16_CAP + 1) * T];   // +1: trash slot for saturated appends
    __shared__ int sbase[9 * T];                           // run starts of pass 0
    const int tid = threadIdx.x;
    const int t = blockIdx.x * blockDim.x + tid;
    int a = 2 * t;
    const bool live = a < n;
    if (!live) a = 0;
    const int b = (a + 1 < n) ? a + 1 : a;
    const float4 pa = posq[a], pb = posq[b];
    const uint32_t ca = cell_sorted[a], cb = cell_sorted[b];
    const uint32_t cola = ca / (uint32_t)G.nz, colb = cb / (uint32_t)G.nz;
    const int cza = (int)(ca - cola * (uint32_t)G.nz), czb = (int)(cb - colb * (uint32_t)G.nz);
    const bool merged = (b != a) && (cola == colb) && (czb - cza <= 3);
    const float FAR = 1.0e18f;
    const int npass = !live ? 0 : ((merged || b == a) ? 1 : 2);

    float2 acc = make_float2(0.f, 0.f);
    unsigned short* const lbase = list16 + tid;
    int* const sb = sbase + tid;
    const unsigned lbase_sa = (unsigned)__cvta_generic_to_shared(lbase);
    const unsigned cap_sa = lbase_sa + 2u * L16_CAP * T;
    unsigned wp = lbase_sa;                                    // append cursor: shared-memory byte address
    auto OFF = [&]() { return (int)((wp - lbase_sa) >> 1); };  // slot * T
    int off0 = 0;
    bool too_long = false;
#pragma unroll 1
    for (int p = 0; p < npass; p++) {
        const bool useA = merged || p == 0, useB = merged || p == 1;
        const float2 X = make_float2(useA ? pa.x : FAR, useB ? pb.x : FAR);
        const float2 Y = make_float2(useA ? pa.y : FAR, useB ? pb.y : FAR);
        const float2 Z = make_float2(useA ? pa.z : FAR, useB ? pb.z : FAR);
        const uint32_t col = p ? colb : cola;
        const int czlo = p ? czb : cza, czhi = merged ? czb : czlo;
        const int cy = (int)(col % (uint32_t)G.ny), cx = (int)(col / (uint32_t)G.ny);
        const int z0 = czlo > 0 ? czlo - 1 : 0, z1 = czhi < G.nz - 1 ? czhi + 1 : czhi;
        if (p == 1) off0 = OFF();
        auto test = [&](const unsigned code, const float4 pj) {
            float2 dx = __fadd2_rn(X, make_float2(-pj.x, -pj.x));
            float2 dy = __fadd2_rn(Y, make_float2(-pj.y, -pj.y));
            float2 dz = __fadd2_rn(Z, make_float2(-pj.z, -pj.z));
            float2 d2 = __fmul2_rn(dx, dx);
            d2 = __ffma2_rn(dy, dy, d2);
            d2 = __ffma2_rn(dz, dz, d2);
            float2 w = __fadd2_rn(make_float2(C.hh, C.hh), make_float2(-d2.x, -d2.y));
            // branch-free append through a shared-memory BYTE-address cursor and the w + |w| clamp, as in k_density_list
            asm volatile("st.shared.u16 [%0], %1;" ::"r"(min(wp, cap_sa)), "h"((unsigned short)code) : "memory");
            asm("{ .reg .pred q; setp.ge.f32 q, %1, %3; @q add.u32 %0, %0, %2; }" : "+r"(wp) : "f"(fmaxf(w.x, w.y)), "n"(2 * L16_THREADS), "f"(LIST_NEG_EPS * C.hh));
            w = __fadd2_rn(w, make_float2(fabsf(w.x), fabsf(w.y)));
            acc = __ffma2_rn(__fmul2_rn(w, w), w, acc);
        };
        // entries beyond the shared-memory capacity go straight to their rows in HBM (see k_density_list)
        auto spill = [&](const int s, const int e, int idx) {
#pragma unroll 4
            for (int k = s; k < e; k++) {
                const float4 pj = __ldg(&posq[k]);
                float2 dx = __fadd2_rn(X, make_float2(-pj.x, -pj.x));
                float2 dy = __fadd2_rn(Y, make_float2(-pj.y, -pj.y));
                float2 dz = __fadd2_rn(Z, make_float2(-pj.z, -pj.z));
                float2 d2 = __fmul2_rn(dx, dx);
                d2 = __ffma2_rn(dy, dy, d2);
                d2 = __ffma2_rn(dz, dz, d2);
                float2 w = __fadd2_rn(make_float2(C.hh, C.hh), make_float2(-d2.x, -d2.y));
                if (fmaxf(w.x, w.y) >= LIST_NEG_EPS * C.hh) {
                    if (idx >= L16_CAP && idx < rows) nlist[(size_t)idx * npairs_pad + t] = k;
                    idx++;
                }
            }
        };
        auto bounds = [&](const int r, int& s, int& e) {
            int x = cx + r / 3 - 1, y = cy + r % 3 - 1;
            bool ok = r < 9 && x >= 0 && x < G.nx && y >= 0 && y < G.ny;
            int base = ok ? (x * G.ny + y) * G.nz : 0;
            s = __ldg(&cell_start[base + z0]);
            e = ok ? __ldg(&cell_start[base + z1 + 1]) : s;
        };
        if (!PF) {
#pragma unroll 1
            for (int r = 0; r < 9; r++) {
                int s, e;
                bounds(r, s, e);
                if (p == 0) sb[r * T] = s;
                too_long |= (e - s) > 4095;
                const unsigned rb = (unsigned)r << 12;
                const int off_run = OFF();
#pragma unroll 4
                for (int k = s; k < e; k++) test(rb + (unsigned)(k - s), __ldg(&posq[k]));
                if (wp > cap_sa) spill(s, e, off_run / T);
            }
        } else {
            int s, e, sn, en;
            bounds(0, s, e);
#pragma unroll 1
            for (int r = 0; r < 9; r++) {
                bounds(r + 1, sn, en);   // next run's bounds in flight (r + 1 == 9 yields an empty range)
                if (p == 0) sb[r * T] = s;
                too_long |= (e - s) > 4095;
                unsigned code = (unsigned)r << 12;
                const int off_run = OFF();
                int k = s;
#if DL_LD256
                // 256-bit candidate loads (see k_density_list)
                if ((k & 1) && k < e) { test(code, __ldg(&posq[k])); k++; code++; }
                if (k + 4 <= e) {
                    F8 q01 = ldg256(&posq[k]), q23 = ldg256(&posq[k + 2]);
#pragma unroll 1
                    for (; k + 8 <= e; k += 4, code += 4) {
                        const F8 n01 = ldg256(&posq[k + 4]), n23 = ldg256(&posq[k + 6]);
                        test(code, q01.a); test(code + 1, q01.b); test(code + 2, q23.a); test(code + 3, q23.b);
                        q01 = n01; q23 = n23;
                    }
                    test(code, q01.a); test(code + 1, q01.b); test(code + 2, q23.a); test(code + 3, q23.b);
                    k += 4; code += 4;
                }
                if (k + 2 <= e) {
                    const F8 q = ldg256(&posq[k]);
                    test(code, q.a); test(code + 1, q.b);
                    k += 2; code += 2;
                }
#else
                if (k + 4 <= e) {
                    float4 q0 = __ldg(&posq[k]), q1 = __ldg(&posq[k + 1]), q2 = __ldg(&posq[k + 2]), q3 = __ldg(&posq[k + 3]);
#pragma unroll 1
                    for (; k + 8 <= e; k += 4, code += 4) {
                        const float4 n0 = __ldg(&posq[k + 4]), n1 = __ldg(&posq[k + 5]), n2 = __ldg(&posq[k + 6]), n3 = __ldg(&posq[k + 7]);
                        test(code, q0); test(code + 1, q1); test(code + 2, q2); test(code + 3, q3);
                        q0 = n0; q1 = n1; q2 = n2; q3 = n3;
                    }
                    test(code, q0); test(code + 1, q1); test(code + 2, q2); test(code + 3, q3);
                    k += 4; code += 4;
                }
#endif
#pragma unroll 1
                for (; k < e; k++, code++) test(code, __ldg(&posq[k]));
                if (wp > cap_sa) spill(s, e, off_run / T);
                s = sn; e = en;
            }
        }
    }
    const int off = OFF();
    if (npass == 1) off0 = off;
    const int cnt = off / T;
    const bool fits = cnt <= max(rows, L16_CAP) && !too_long;
    if (live) ncount[t] = fits ? make_int2(off0 / T, cnt) : make_int2(-1, -1);
    if (live && !fits) atomicAdd(overflow, 1);
    if (live && cnt > L16_CAP) atomicAdd(overflow + 1, 1);
    if (live && cnt > L16_CAP / 2) atomicAdd(overflow + 2, 1);
    if (t == 0) { overflow[3] = L16_CAP; overflow[4] = max(rows, L16_CAP); }
    if (fits) {
        // coalesced flush: entry e of all threads of the block is one contiguous row; decode run | offset -> particle index
        int* dst = nlist + t;
        const int seg = (npass == 2) ? off0 : off;          // entries from pass 0 use the stored run starts
        const int cyb = (int)(colb % (uint32_t)G.ny), cxb = (int)(colb / (uint32_t)G.ny), z0b = czb > 0 ? czb - 1 : 0;
        const int stop = min(off, L16_CAP * T);             // longer lists: the rest is already in place
#pragma unroll 2
        for (int o = 0, e = 0; o < stop; o += T, e++) {
            const unsigned v = lbase[o];
            const int r = (int)(v >> 12);
            int base;
            if (o < seg) base = sb[r * T];
            else base = __ldg(&cell_start[((cxb + r / 3 - 1) * G.ny + (cyb + r % 3 - 1)) * G.nz + z0b]);   // split pair, second pass (rare)
            dst[(size_t)e * npairs_pad] = base + (int)(v & 4095u);
        }
    }
    if (!live) return;
    const float dk = C.densK * 0.125f;   // exact: the accumulators hold 8 x sum (h^2 - r^2)^3 (see `test`)
    float ra = acc.x * dk, rb = acc.y * dk;
    float Pa = C.k * (ra - C.p0), Pb = C.k * (rb - C.p0);
    rho[a] = ra;
    posq_q[a] = make_float4(pa.x, pa.y, pa.z, Pa / (ra * ra));
    velv[a].w = C.mass / ra;
    if (b != a) {
        rho[b] = rb;
        posq_q[b] = make_float4(pb.x, pb.y, pb.z, Pb / (rb * rb));
        velv[b].w = C.mass / rb;
    }
}

// one 256-bit read-only gather of an interleaved particle record
__device__ __forceinline__ void ldg_rec(const float4* __restrict__ rec, int k, float4& p, float4& v) {
    asm("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=f"(p.x), "=f"(p.y), "=f"(p.z), "=f"(p.w), "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
        : "l"(rec + 2 * (size_t)k));
}

#ifndef FL_MINB
#define FL_MINB 16
#endif
#ifndef FL_THREADS
#define FL_THREADS 64    // CTA size of the force pass (independent of the list layout).  The pass is latency bound (long scoreboard: gathers
// that miss L1): 32 resident warps at 64 registers (a few spilled words) beat 28 at 72 -- c3 force pass 0.541 -> 0.517 ms;
// 24 warps: 0.62 ms, 36 warps at 56 registers: 0.58 ms (profiles/r02/ab_force_occupancy.txt)
#endif
// PF = true (variant 6): the list index is fetched TWO entries ahead.  ncu source view of the plain kernel: 38 % of
// the stall samples sit on the address computation of the next gather, i.e. on the index load it depends on
// (index -> gather is a dependent chain; the lists stream from HBM).  One more register hides it.
template <bool DIAG, bool REC, bool PF>
__global__ void __launch_bounds__(FL_THREADS, FL_MINB) k_force_list(int n_hi, const int* __restrict__ n_dev, int npairs_pad, const float4* __restrict__ posq_q,
                                                              const float4* __restrict__ velv, const float* __restrict__ rho,
                                                              const int* __restrict__ ids, const uint32_t* __restrict__ cell_sorted,
                                                              const int* __restrict__ cell_start, GridP G, StepC C,
                                                              const int* __restrict__ nlist, const int2* __restrict__ ncount,
                                                              float4* __restrict__ posq_out, float4* __restrict__ velv_out, DiagOut D) {
    const int n = n_dev ? __ldg(n_dev) : n_hi;  // exact count from the device in slab mode, else the launch bound
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int a = 2 * t;
    if (a >= n) return;
    const int b = (a + 1 < n) ? a + 1 : a;
    float4 pa, pb, va, vb;
    if (REC) { ldg_rec(posq_q, a, pa, va); ldg_rec(posq_q, b, pb, vb); }
    else { pa = posq_q[a]; pb = posq_q[b]; va = velv[a]; vb = velv[b]; }
    const int2 cn = ncount[t];
    const float inv_sqrt3 = 0.57735026f;
    const float FAR = 1.0e18f;

    float2 A_x = {0.f, 0.f}, A_y = {0.f, 0.f}, A_z = {0.f, 0.f};
    float2 F_x = {0.f, 0.f}, F_y = {0.f, 0.f}, F_z = {0.f, 0.f};
    float2 N_x = {0.f, 0.f}, N_y = {0.f, 0.f}, N_z = {0.f, 0.f};
    float2 CF = {0.f, 0.f};
    int maxa = -1, maxb = -1;
    const float2 Q = make_float2(pa.w, pb.w);
    const float2 VX = make_float2(va.x, vb.x), VY = make_float2(va.y, vb.y), VZ = make_float2(va.z, vb.z);

    // segment 0 = entries [0, cn.x): targets (a, b) when the pair was merged (cn.x == cn.y), else (a, FAR);
    // segment 1 = entries [cn.x, cn.y): targets (FAR, b)
    const bool merged = (cn.x == cn.y);
    float2 X, Y, Z;
    int ia, ib;
    auto body = [&](const int k, const float4 pj, const float4 vj) {
        float2 dx = __fadd2_rn(X, make_float2(-pj.x, -pj.x));
        float2 dy = __fadd2_rn(Y, make_float2(-pj.y, -pj.y));
        float2 dz = __fadd2_rn(Z, make_float2(-pj.z, -pj.z));
        float2 d2 = __fmul2_rn(dx, dx);
        d2 = __ffma2_rn(dy, dy, d2);
        d2 = __ffma2_rn(dz, dz, d2);
        float2 w = __fadd2_rn(make_float2(C.hh, C.hh), make_float2(-d2.x, -d2.y));
        // clamps as w + |w| = 2 max(w, 0) (one FADD2 instead of two FMNMX, exact): CF and F then carry a factor 2, N and A a
        // factor 4, removed exactly in the epilogue call below -- bit-identical to clamping
        w = __fadd2_rn(w, make_float2(fabsf(w.x), fabsf(w.y)));
        float2 vw = __fmul2_rn(w, make_float2(vj.w, vj.w));
        float2 t7 = __ffma2_rn(d2, make_float2(-7.0f, -7.0f), make_float2(C.hh3, C.hh3));
        CF = __ffma2_rn(vw, t7, CF);
        float2 vww = __fmul2_rn(vw, w);
        N_x = __ffma2_rn(vww, dx, N_x); N_y = __ffma2_rn(vww, dy, N_y); N_z = __ffma2_rn(vww, dz, N_z);
        float2 rinv = make_float2(rsqrt_ftz(fmaxf(d2.x, 1e-30f)), rsqrt_ftz(fmaxf(d2.y, 1e-30f)));
        float2 r = __fmul2_rn(d2, rinv);
        float2 hm = __fadd2_rn(make_float2(C.h, C.h), make_float2(-r.x, -r.y));
        hm = __fadd2_rn(hm, make_float2(fabsf(hm.x), fabsf(hm.y)));   // 2 max(h - r, 0)
        float2 tv = __fmul2_rn(hm, make_float2(vj.w, vj.w));
        float2 dvx = __fadd2_rn(make_float2(vj.x, vj.x), make_float2(-VX.x, -VX.y));
        float2 dvy = __fadd2_rn(make_float2(vj.y, vj.y), make_float2(-VY.x, -VY.y));
        float2 dvz = __fadd2_rn(make_float2(vj.z, vj.z), make_float2(-VZ.x, -VZ.y));
        F_x = __ffma2_rn(tv, dvx, F_x); F_y = __ffma2_rn(tv, dvy, F_y); F_z = __ffma2_rn(tv, dvz, F_z);
        float2 sq = __fadd2_rn(Q, make_float2(pj.w, pj.w));
        float2 sc = __fmul2_rn(__fmul2_rn(sq, hm), hm);
        if (k == ia) sc.x = 0.f;  // pressure excludes j == i (fluid_system.h:142)
        if (k == ib) sc.y = 0.f;
        float2 ux = __fmul2_rn(dx, rinv), uy = __fmul2_rn(dy, rinv), uz = __fmul2_rn(dz, rinv);
        if (fminf(r.x, r.y) <= 1e-4f) {  // coincident pair (fluid_system.h:438-440) or the self entry
            if (r.x <= 1e-4f) { ux.x = inv_sqrt3; uy.x = inv_sqrt3; uz.x = inv_sqrt3; }
            if (r.y <= 1e-4f) { ux.y = inv_sqrt3; uy.y = inv_sqrt3; uz.y = inv_sqrt3; }
        }
        A_x = __ffma2_rn(sc, ux, A_x); A_y = __ffma2_rn(sc, uy, A_y); A_z = __ffma2_rn(sc, uz, A_z);
        if (DIAG) {
            int idk = __ldg(&ids[k]);
            if (k != ia && ia >= 0 && dist2_exact(pa.x - pj.x, pa.y - pj.y, pa.z - pj.z) <= C.T) maxa = max(maxa, idk);
            if (k != ib && ib >= 0 && dist2_exact(pb.x - pj.x, pb.y - pj.y, pb.z - pj.z) <= C.T) maxb = max(maxb, idk);
        }
    };

    if (cn.y >= 0) {
        const int* src = nlist + t;
#pragma unroll 1
        for (int seg = 0; seg < 2; seg++) {
            const int e0 = seg ? cn.x : 0, e1 = seg ? cn.y : cn.x;
            if (e0 >= e1) continue;
            const bool useA = seg == 0, useB = merged || seg == 1;
            X = make_float2(useA ? pa.x : FAR, useB ? pb.x : FAR);
            Y = make_float2(useA ? pa.y : FAR, useB ? pb.y : FAR);
            Z = make_float2(useA ? pa.z : FAR, useB ? pb.z : FAR);
            ia = useA ? a : -1; ib = (useB && b != a) ? b : -1;
            // software pipeline: the next entry's index and gathers are in flight while this one is processed
            // (a two-stage version raised the register count to 94 and lost a CTA/SM: slower, measured)
            int k = __ldg(&src[(size_t)e0 * npairs_pad]);
            float4 pj, vj;
            if (!PF) {
                if (REC) ldg_rec(posq_q, k, pj, vj); else { pj = __ldg(&posq_q[k]); vj = __ldg(&velv[k]); }
#pragma unroll 1
                for (int e = e0; e < e1; e++) {
                    const int kn = (e + 1 < e1) ? __ldg(&src[(size_t)(e + 1) * npairs_pad]) : k;
                    float4 pjn, vjn;
                    if (REC) ldg_rec(posq_q, kn, pjn, vjn); else { pjn = __ldg(&posq_q[kn]); vjn = __ldg(&velv[kn]); }
                    body(k, pj, vj);
                    k = kn; pj = pjn; vj = vjn;
                }
            } else {
                int k1 = (e0 + 1 < e1) ? __ldg(&src[(size_t)(e0 + 1) * npairs_pad]) : k;
                if (REC) ldg_rec(posq_q, k, pj, vj); else { pj = __ldg(&posq_q[k]); vj = __ldg(&velv[k]); }
#pragma unroll 1
                for (int e = e0; e < e1; e++) {
                    const int k2 = (e + 2 < e1) ? __ldg(&src[(size_t)(e + 2) * npairs_pad]) : k1;   // index two ahead
                    float4 pjn, vjn;                                                                // gathers one ahead
                    if (REC) ldg_rec(posq_q, k1, pjn, vjn); else { pjn = __ldg(&posq_q[k1]); vjn = __ldg(&velv[k1]); }
                    body(k, pj, vj);
                    k = k1; k1 = k2; pj = pjn; vj = vjn;
                }
            }
        }
    } else {
        // list overflowed (strong compression): direct walks with the body under the predicate
        const uint32_t ca = cell_sorted[a], cb = cell_sorted[b];
        const uint32_t cola = ca / (uint32_t)G.nz, colb = cb / (uint32_t)G.nz;
        const int cza = (int)(ca - cola * (uint32_t)G.nz), czb = (int)(cb - colb * (uint32_t)G.nz);
        const bool mg = (b != a) && (cola == colb) && (czb - cza <= 3);
        const int npass = (mg || b == a) ? 1 : 2;
#pragma unroll 1
        for (int p = 0; p < npass; p++) {
            const bool useA = mg || p == 0, useB = mg || p == 1;
            X = make_float2(useA ? pa.x : FAR, useB ? pb.x : FAR);
            Y = make_float2(useA ? pa.y : FAR, useB ? pb.y : FAR);
            Z = make_float2(useA ? pa.z : FAR, useB ? pb.z : FAR);
            ia = useA ? a : -1; ib = (useB && b != a) ? b : -1;
            const uint32_t col = p ? colb : cola;
            const int czlo = p ? czb : cza, czhi = mg ? czb : czlo;
            const int cy = (int)(col % (uint32_t)G.ny), cx = (int)(col / (uint32_t)G.ny);
            const int z0 = czlo > 0 ? czlo - 1 : 0, z1 = czhi < G.nz - 1 ? czhi + 1 : czhi;
#pragma unroll 1
            for (int r = 0; r < 9; r++) {
                int x = cx + r / 3 - 1, y = cy + r % 3 - 1;
                if (x < 0 || x >= G.nx || y < 0 || y >= G.ny) continue;
                int base = (x * G.ny + y) * G.nz;
                int s = __ldg(&cell_start[base + z0]);
                int e = __ldg(&cell_start[base + z1 + 1]);
#pragma unroll 1
                for (int k = s; k < e; k++) {
                    float4 pj, vj;
                    if (REC) ldg_rec(posq_q, k, pj, vj); else { pj = __ldg(&posq_q[k]); vj = __ldg(&velv[k]); }
                    float ex = X.x - pj.x, ey = Y.x - pj.y, ez = Z.x - pj.z;
                    float gx = X.y - pj.x, gy = Y.y - pj.y, gz = Z.y - pj.z;
                    float da = fmaf(ez, ez, fmaf(ey, ey, ex * ex)), db = fmaf(gz, gz, fmaf(gy, gy, gx * gx));
                    if (fminf(da, db) <= C.hh) body(k, pj, vj);
                }
            }
        }
    }

#pragma unroll 1
    for (int p = 0; p < 2; p++) {
        if (p == 1 && b == a) break;
        const int i = p ? b : a;
        const float4 pi = p ? pb : pa;
        const float4 vi = p ? vb : va;
        const float ax = p ? A_x.y : A_x.x, ay = p ? A_y.y : A_y.x, az = p ? A_z.y : A_z.x;
        const float fx = p ? F_x.y : F_x.x, fy = p ? F_y.y : F_y.x, fz = p ? F_z.y : F_z.x;
        const float nx = p ? N_x.y : N_x.x, ny = p ? N_y.y : N_y.x, nz = p ? N_z.y : N_z.x;
        const float cf = p ? CF.y : CF.x;
        // exact powers of two: the clamps in `body` were 2 max(., 0) (A and N carry 4x, F and CF 2x)
        force_epilogue<DIAG>(i, pi, vi, rho[i], 0.25f * ax, 0.25f * ay, 0.25f * az, 0.5f * fx, 0.5f * fy, 0.5f * fz, 0.25f * nx, 0.25f * ny, 0.25f * nz,
                             0.5f * cf, p ? maxb : maxa, C, ids, posq_out, velv_out, D);
    }
}

#ifdef SPHE_WITH_EXPERIMENTS
#include "experiments/sph_variants_r1.cuh"
#endif

// ------------------------------------------------------------------ neighbour-list test hooks
__global__ void __launch_bounds__(128) k_nbr_count(int n, const float4* __restrict__ posq, const uint32_t* __restrict__ cell_sorted,
                                                   const int* __restrict__ cell_start, GridP G, StepC C, int* __restrict__ counts) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 pi = posq[i];
    int c = 0;
    walk27(G, cell_start, cell_sorted[i], [&](int k) {
        float4 pj = __ldg(&posq[k]);
        if (dist2_exact(pi.x - pj.x, pi.y - pj.y, pi.z - pj.z) <= C.T) c++;
    });
    counts[i] = c;
}

__global__ void __launch_bounds__(128) k_nbr_fill(int n, const float4* __restrict__ posq, const int* __restrict__ ids,
                                                  const uint32_t* __restrict__ cell_sorted, const int* __restrict__ cell_start,
                                                  GridP G, StepC C, const long long* __restrict__ nbr_start, int* __restrict__ nbr) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 pi = posq[i];
    long long o = nbr_start[i];
    walk27(G, cell_start, cell_sorted[i], [&](int k) {
        float4 pj = __ldg(&posq[k]);
        if (dist2_exact(pi.x - pj.x, pi.y - pj.y, pi.z - pj.z) <= C.T) nbr[o++] = ids[k];
    });
}

// Test hook: the PRODUCTION pair lists of the index-list kernels (what k_force_list walks), per target:
// out[target * cap + i] = sorted slot of the i-th recorded candidate, counts[target] (-1: the pair's list overflowed its
// rows and the force pass walked the cells directly).  Segment 0 of a pair list serves (a, b) when the pair was merged,
// else a; segment 1 serves b.
__global__ void k_list_decode(int n, int npairs_pad, const int* __restrict__ nlist, const int2* __restrict__ ncount, int cap,
                              int* __restrict__ counts, int* __restrict__ out) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int a = 2 * t;
    if (a >= n) return;
    const bool hasb = a + 1 < n;
    const int2 cn = ncount[t];
    if (cn.y < 0) { counts[a] = -1; if (hasb) counts[a + 1] = -1; return; }
    const bool merged = cn.x == cn.y;
    int na = 0, nb = 0;
    for (int e = 0; e < cn.y; e++) {
        const int k = nlist[(size_t)e * npairs_pad + t];
        const bool forA = e < cn.x, forB = hasb && (merged || e >= cn.x);
        if (forA) { if (na < cap) out[(size_t)a * cap + na] = k; na++; }
        if (forB) { if (nb < cap) out[(size_t)(a + 1) * cap + nb] = k; nb++; }
    }
    counts[a] = na;
    if (hasb) counts[a + 1] = nb;
}

// NeighbId (fluid_system.h:144, debug only): the last neighbour in ascending-id order = the largest id among the exact
// neighbours j != i.  The staged force kernel has no ids at hand, so the diagnostics get them from this walk.
__global__ void __launch_bounds__(128) k_neighb_id(int n_hi, const int* __restrict__ n_dev, const float4* __restrict__ posq, const int* __restrict__ ids,
                                                   const uint32_t* __restrict__ cell_sorted, const int* __restrict__ cell_start,
                                                   GridP G, StepC C, int* __restrict__ neighb) {
    const int n = n_dev ? __ldg(n_dev) : n_hi;
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 pi = posq[i];
    int maxid = -1;
    walk27(G, cell_start, cell_sorted[i], [&](int k) {
        float4 pj = __ldg(&posq[k]);
        if (k != i && dist2_exact(pi.x - pj.x, pi.y - pj.y, pi.z - pj.z) <= C.T) maxid = max(maxid, __ldg(&ids[k]));
    });
    if (maxid >= 0) neighb[ids[i]] = maxid;
}

// ------------------------------------------------------------------ id-order gathers / packing
__global__ void k_unsort_f4(int n, const float4* __restrict__ src, const int* __restrict__ ids, float* __restrict__ dst) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 v = src[i];
    size_t o = 3 * (size_t)ids[i];
    dst[o] = v.x; dst[o + 1] = v.y; dst[o + 2] = v.z;
}
__global__ void k_unsort_f1(int n, const float* __restrict__ src, const int* __restrict__ ids, float* __restrict__ dst) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[ids[i]] = src[i];
}
__global__ void k_unsort_u32(int n, const uint32_t* __restrict__ src, const int* __restrict__ ids, int* __restrict__ dst) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[ids[i]] = (int)src[i];
}
// pos / vel may be NULL: the end-to-end path packs the two halves on different streams (velocities are still on the
// PCIe bus while the positions are already being binned)
__global__ void k_pack_state(int n, const float* __restrict__ pos, const float* __restrict__ vel, float4* __restrict__ posq,
                             float4* __restrict__ velv, int* __restrict__ ids, float* __restrict__ sed) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (pos) {
        posq[i] = make_float4(pos[3 * (size_t)i], pos[3 * (size_t)i + 1], pos[3 * (size_t)i + 2], 0.f);
        ids[i] = i;
        if (sed) sed[i] = 0.f;
    }
    if (vel) velv[i] = make_float4(vel[3 * (size_t)i], vel[3 * (size_t)i + 1], vel[3 * (size_t)i + 2], 0.f);
}
__global__ void k_slot_of_id(int n, const int* __restrict__ ids, int* __restrict__ slot) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) slot[ids[i]] = i;
}

// ------------------------------------------------------------------ launch wrappers
static inline int nblk(int n, int b) { return (n + b - 1) / b; }

int nlist_cap() { return NLIST_CAP; }

// which kernel families this build carries (sphe_set_variant)
bool variant_supported(int density, int force) {
    auto base = [](int v) { return v == 0 || v == 3 || v == 6 || v == 10 || v == 11; };
#ifdef SPHE_WITH_EXPERIMENTS
    auto ex = [](int v) { return v == 1 || v == 2 || v == 4 || v == 8 || v == 16 || v == 7 || v == 9 || v == 52 || v == 54 || v == 58; };
#else
    auto ex = [](int) { return false; };
#endif
    return (base(density) || density == 20 || ex(density)) && (base(force) || force == 20 || ex(force));
}
int nlist_pairs_pad(int n) { return (((n + 1) / 2) + 127) & ~127; }

template <bool REC, bool PF, int CAP, int THREADS, int UNROLL = 4>
static void launch_density_list(cudaStream_t st, int n, const int* n_dev, int pp, const float4* posq, float4* posq_q, float4* velv,
                                const uint32_t* cell_sorted, const int* cell_start, const GridP& G, const StepC& C, float* rho,
                                int* nlist, int2* ncount, int* overflow, int rows) {
    int smem = (CAP + 1) * THREADS * (int)sizeof(int);
    if (smem <= 48 * 1024) smem = 0;   // static in the kernel
    auto kern = k_density_list<REC, PF, CAP, THREADS, UNROLL>;
    // the opt-in is per DEVICE (one process may drive several GPUs): set it before every launch that needs it, it is cheap
    if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    kern<<<pp / THREADS, THREADS, smem, st>>>(n, n_dev, pp, posq, posq_q, velv, cell_sorted, cell_start, G, C, rho, nlist, ncount, overflow, rows < CAP ? CAP : rows);
}

void launch_density(cudaStream_t st, int variant, int n, const int* n_dev, const float4* posq, float4* posq_q, float4* velv,
                    const uint32_t* cell_sorted, const int* cell_start, const GridP& G, const StepC& C, float* rho,
                    int* nlist, int2* ncount, int cap, int* overflow, int smem_cap) {
    if (n <= 0) return;
    [[maybe_unused]] int pairs = (n + 1) / 2;
#ifdef SPHE_WITH_EXPERIMENTS
    if (variant == 52) return launch_density_s<2>(st, n, n_dev, posq, posq_q, velv, cell_sorted, cell_start, G, C, rho, nlist, ncount);
    if (variant == 54) return launch_density_s<4>(st, n, n_dev, posq, posq_q, velv, cell_sorted, cell_start, G, C, rho, nlist, ncount);
    if (variant == 58) return launch_density_s<8>(st, n, n_dev, posq, posq_q, velv, cell_sorted, cell_start, G, C, rho, nlist, ncount);
    if (variant == 7 || variant == 9) {
        int pp = nlist_pairs_pad(n);
        if (variant == 9) k_density_quad<true><<<pp / 2 / QUAD_THREADS, QUAD_THREADS, 0, st>>>(n, n_dev, pp, posq, posq_q, velv, cell_sorted, cell_start, G, C, rho, nlist, ncount);
        else k_density_quad<false><<<pp / 2 / QUAD_THREADS, QUAD_THREADS, 0, st>>>(n, n_dev, pp, posq, posq_q, velv, cell_sorted, cell_start, G, C, rho, nlist, ncount);
        return;
    }
#endif
    if (variant == 10 || variant == 11) {
        int pp = nlist_pairs_pad(n);
        if (variant == 10) k_density_list16<true><<<pp / L16_THREADS, L16_THREADS, 0, st>>>(n, n_dev, pp, posq, posq_q, velv, cell_sorted, cell_start, G, C, rho, nlist, ncount, overflow, cap);
        else k_density_list16<false><<<pp / L16_THREADS, L16_THREADS, 0, st>>>(n, n_dev, pp, posq, posq_q, velv, cell_sorted, cell_start, G, C, rho, nlist, ncount, overflow, cap);
        return;
    }
    if (variant == 3 || variant == 6) {
        int pp = nlist_pairs_pad(n);
#define SPHE_DL(RC, PFv, CP, TH) launch_density_list<RC, PFv, CP, TH>(st, n, n_dev, pp, posq, posq_q, velv, cell_sorted, cell_start, G, C, rho, nlist, ncount, overflow, cap)
        // cap = rows allocated per pair.  The shared-memory part stays at 64 entries (occupancy), longer lists spill; only
        // when MOST pairs spill (smem_cap, chosen by the host from the spill statistics) the wide-shared-memory
        // instantiations take over (explicit prefetch: ptxas serialises the gathers of those otherwise, 48 registers)
#ifdef SPHE_DL256_32BIT
        if (smem_cap > 128) SPHE_DL(false, true, 256, 64);
#else
        if (smem_cap > 128) k_density_list16<true, 256, 64><<<pp / 64, 64, 0, st>>>(n, n_dev, pp, posq, posq_q, velv, cell_sorted, cell_start, G, C, rho, nlist, ncount, overflow, cap);
#endif
        else if (smem_cap > 64) {   // 128 staged entries: the 16-bit kernel holds them in the same 33 KB (6 CTAs/SM)
            if (variant == 6) k_density_list16<true><<<pp / L16_THREADS, L16_THREADS, 0, st>>>(n, n_dev, pp, posq, posq_q, velv, cell_sorted, cell_start, G, C, rho, nlist, ncount, overflow, cap);
            else k_density_list16<false><<<pp / L16_THREADS, L16_THREADS, 0, st>>>(n, n_dev, pp, posq, posq_q, velv, cell_sorted, cell_start, G, C, rho, nlist, ncount, overflow, cap);
        }
        else if (variant == 6) SPHE_DL(false, true, DL_CAP, DL_THREADS);
        else SPHE_DL(false, false, DL_CAP, DL_THREADS);
#undef SPHE_DL
        return;
    }
#ifdef SPHE_WITH_EXPERIMENTS
    if (variant == 1) k_density_pair<1><<<nblk(pairs, 128), 128, 0, st>>>(n, n_dev, posq, posq_q, velv, cell_sorted, cell_start, G, C, rho);
    else if (variant == 2) k_density_pair<2><<<nblk(pairs * 2, 128), 128, 0, st>>>(n, n_dev, posq, posq_q, velv, cell_sorted, cell_start, G, C, rho);
    else if (variant == 4) k_density_pair<4><<<nblk(pairs * 4, 128), 128, 0, st>>>(n, n_dev, posq, posq_q, velv, cell_sorted, cell_start, G, C, rho);
    else if (variant == 8) k_density_pair<8><<<nblk(pairs * 8, 128), 128, 0, st>>>(n, n_dev, posq, posq_q, velv, cell_sorted, cell_start, G, C, rho);
    else if (variant == 16) k_density_pair<16><<<nblk(pairs * 16, 128), 128, 0, st>>>(n, n_dev, posq, posq_q, velv, cell_sorted, cell_start, G, C, rho);
    if (variant == 1 || variant == 2 || variant == 4 || variant == 8 || variant == 16) return;
#endif
    k_density_tpp<<<nblk(n, 128), 128, 0, st>>>(n, n_dev, posq, posq_q, velv, cell_sorted, cell_start, G, C, rho);
}

void launch_force(cudaStream_t st, int variant, int n, const int* n_dev, const float4* posq_q, const float4* velv, const float* rho,
                  const int* ids, const uint32_t* cell_sorted, const int* cell_start, const GridP& G, const StepC& C,
                  float4* posq_out, float4* velv_out, const DiagOut* diag, const int* nlist, const int2* ncount) {
    if (n <= 0) return;
#ifdef SPHE_WITH_EXPERIMENTS
    if (variant == 52) return launch_force_s<2>(st, n, n_dev, posq_q, velv, rho, ids, cell_sorted, cell_start, G, C, posq_out, velv_out, diag, nlist, ncount);
    if (variant == 54) return launch_force_s<4>(st, n, n_dev, posq_q, velv, rho, ids, cell_sorted, cell_start, G, C, posq_out, velv_out, diag, nlist, ncount);
    if (variant == 58) return launch_force_s<8>(st, n, n_dev, posq_q, velv, rho, ids, cell_sorted, cell_start, G, C, posq_out, velv_out, diag, nlist, ncount);
#endif
    if (variant == 7 || variant == 9 || variant == 10 || variant == 11) variant = 3;   // these density passes write the plain pair lists
    if (variant == 3 || variant == 6) {
        int pp = nlist_pairs_pad(n);
        dim3 g((pp + FL_THREADS - 1) / FL_THREADS), b(FL_THREADS);
#define SPHE_FL(DG, RC, PFv, dg) k_force_list<DG, RC, PFv><<<g, b, 0, st>>>(n, n_dev, pp, posq_q, velv, rho, ids, cell_sorted, cell_start, G, C, nlist, ncount, posq_out, velv_out, dg)
        if (variant == 6) { if (diag) SPHE_FL(true, false, true, *diag); else SPHE_FL(false, false, true, DiagOut{}); }
        else { if (diag) SPHE_FL(true, false, false, *diag); else SPHE_FL(false, false, false, DiagOut{}); }
#undef SPHE_FL
        return;
    }
#ifdef SPHE_WITH_EXPERIMENTS
    if (variant == 1) {
        int nb = nblk((n + 1) / 2, FORCE_THREADS);
        if (diag) k_force_pair<true><<<nb, FORCE_THREADS, 0, st>>>(n, n_dev, posq_q, velv, rho, ids, cell_sorted, cell_start, G, C, posq_out, velv_out, *diag);
        else k_force_pair<false><<<nb, FORCE_THREADS, 0, st>>>(n, n_dev, posq_q, velv, rho, ids, cell_sorted, cell_start, G, C, posq_out, velv_out, DiagOut{});
        return;
    }
#endif
    if (diag) k_force_tpp<true><<<nblk(n, 128), 128, 0, st>>>(n, n_dev, posq_q, velv, rho, ids, cell_sorted, cell_start, G, C, posq_out, velv_out, *diag);
    else k_force_tpp<false><<<nblk(n, 128), 128, 0, st>>>(n, n_dev, posq_q, velv, rho, ids, cell_sorted, cell_start, G, C, posq_out, velv_out, DiagOut{});
}

void launch_list_decode(cudaStream_t st, int n, const int* nlist, const int2* ncount, int cap, int* counts, int* out) {
    if (n > 0) k_list_decode<<<nblk((n + 1) / 2, 128), 128, 0, st>>>(n, nlist_pairs_pad(n), nlist, ncount, cap, counts, out);
}
void launch_neighb_id(cudaStream_t st, int n, const int* n_dev, const float4* posq, const int* ids, const uint32_t* cell_sorted,
                      const int* cell_start, const GridP& G, const StepC& C, int* neighb_by_id) {
    if (n > 0) k_neighb_id<<<nblk(n, 128), 128, 0, st>>>(n, n_dev, posq, ids, cell_sorted, cell_start, G, C, neighb_by_id);
}
void launch_neighbour_count(cudaStream_t st, int n, const float4* posq, const uint32_t* cell_sorted, const int* cell_start,
                            const GridP& G, const StepC& C, int* counts) {
    if (n > 0) k_nbr_count<<<nblk(n, 128), 128, 0, st>>>(n, posq, cell_sorted, cell_start, G, C, counts);
}
void launch_neighbour_fill(cudaStream_t st, int n, const float4* posq, const int* ids, const uint32_t* cell_sorted,
                           const int* cell_start, const GridP& G, const StepC& C, const long long* nbr_start, int* nbr) {
    if (n > 0) k_nbr_fill<<<nblk(n, 128), 128, 0, st>>>(n, posq, ids, cell_sorted, cell_start, G, C, nbr_start, nbr);
}
void launch_unsort_f4(cudaStream_t st, int n, const float4* src, const int* ids, float* dst_xyz) {
    if (n > 0) k_unsort_f4<<<nblk(n, 256), 256, 0, st>>>(n, src, ids, dst_xyz);
}
void launch_unsort_f1(cudaStream_t st, int n, const float* src, const int* ids, float* dst) {
    if (n > 0) k_unsort_f1<<<nblk(n, 256), 256, 0, st>>>(n, src, ids, dst);
}
void launch_unsort_u32(cudaStream_t st, int n, const uint32_t* src, const int* ids, int* dst) {
    if (n > 0) k_unsort_u32<<<nblk(n, 256), 256, 0, st>>>(n, src, ids, dst);
}
void launch_pack_state(cudaStream_t st, int n, const float* pos_xyz, const float* vel_xyz, float4* posq, float4* velv,
                       int* ids, float* sed) {
    if (n > 0) k_pack_state<<<nblk(n, 256), 256, 0, st>>>(n, pos_xyz, vel_xyz, posq, velv, ids, sed);
}
void launch_slot_of_id(cudaStream_t st, int n, const int* ids, int* slot_of_id) {
    if (n > 0) k_slot_of_id<<<nblk(n, 256), 256, 0, st>>>(n, ids, slot_of_id);
}

}  // namespace sphe
