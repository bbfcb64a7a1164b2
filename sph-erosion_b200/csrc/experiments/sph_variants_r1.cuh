// Round-1 kernel experiments, measured and rejected (numbers in the comments and in DESIGN.md section 3): packed pair without
// lists (variants 1, 2, 4, 8, 16), four targets per thread (7, 9), S lanes per pair with sub-lists (52, 54, 58).
// NOT part of the default build: sph.cu includes this file only with -DSPHE_WITH_EXPERIMENTS
// (SPHE_WITH_EXPERIMENTS=1 python sph-erosion_b200/build.py); sphe_set_variant refuses these numbers otherwise.
#pragma once

// ------------------------------------------------------------------ pass 1, variant 1: two targets per thread
// B200-specific: sm_100 has packed fp32 math (FADD2 / FMUL2 / FFMA2, PTX add/mul/fma.f32x2) whose
// second operand can be a scalar broadcast.  A thread owns two consecutive sorted particles (a, b),
// keeps their coordinates packed as (xa,xb),(ya,yb),(za,zb) and streams each candidate ONCE for both:
// one LDG.128 + 9 packed ops + 2 FMNMX per candidate instead of 2 x (LDG + 12 scalar ops).  That
// halves both the issue slots and the L1 wavefronts per (target,candidate) pair -- the two limits ncu
// showed for the thread-per-particle kernel (profiles/r01_ncu_density_force_tpp.txt).
// a and b are usually in the same cell; if they are in the same column and at most 3 cells apart the
// walk covers the union z-range (extra candidates fail the distance test); otherwise two walks.
// The weight is clamped, max(h^2 - d2, 0)^3, instead of predicated: contributions vanish continuously
// at r = h so the FMA-contracted d2 is within tolerance (the exact predicate is only needed for the
// neighbour LISTS, see k_nbr_fill).
template <int S>
__device__ __forceinline__ float2 density_walk_pair(const GridP& G, const int* __restrict__ cell_start,
                                                    const float4* __restrict__ posq, int cx, int cy, int z0, int z1,
                                                    float2 X, float2 Y, float2 Z, float hh, int slice) {
    float2 acc = make_float2(0.f, 0.f);
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
#pragma unroll 4
            for (int k = s + slice; k < e; k += S) {
                float4 p = __ldg(&posq[k]);
                float2 ddx = __fadd2_rn(X, make_float2(-p.x, -p.x));
                float2 ddy = __fadd2_rn(Y, make_float2(-p.y, -p.y));
                float2 ddz = __fadd2_rn(Z, make_float2(-p.z, -p.z));
                float2 d2 = __fmul2_rn(ddx, ddx);
                d2 = __ffma2_rn(ddy, ddy, d2);
                d2 = __ffma2_rn(ddz, ddz, d2);
                float2 w = __fadd2_rn(make_float2(hh, hh), make_float2(-d2.x, -d2.y));
                w.x = fmaxf(w.x, 0.f);
                w.y = fmaxf(w.y, 0.f);
                acc = __ffma2_rn(__fmul2_rn(w, w), w, acc);
            }
        }
    }
    return acc;
}

// S lanes share one target pair and stride the candidate ranges (lane%S, step S): the S lanes read S
// consecutive float4 (one or two 128-byte lines) instead of S unrelated ranges, which cuts the L1
// wavefronts per request -- the limiter ncu reports for S = 1 (l1tex lsu wavefronts 91 % of peak) -- and
// leaves fewer distinct cells per warp (less trip-count divergence).  Partial sums are combined with
// __shfl_xor.
template <int S>
__global__ void __launch_bounds__(128) k_density_pair(int n_hi, const int* __restrict__ n_dev, const float4* __restrict__ posq, float4* __restrict__ posq_q,
                                                      float4* __restrict__ velv, const uint32_t* __restrict__ cell_sorted,
                                                      const int* __restrict__ cell_start, GridP G, StepC C,
                                                      float* __restrict__ rho) {
    const int n = n_dev ? __ldg(n_dev) : n_hi;  // exact count from the device in slab mode, else the launch bound
    const int gt = blockIdx.x * blockDim.x + threadIdx.x;
    const int slice = gt % S;
    int a = 2 * (gt / S);
    if (a >= n) a = (n - 1) & ~1;  // keep whole warps alive for the shuffles; duplicates write identical values
    int b = (a + 1 < n) ? a + 1 : a;
    float4 pa = posq[a], pb = posq[b];
    uint32_t ca = cell_sorted[a], cb = cell_sorted[b];
    uint32_t cola = ca / (uint32_t)G.nz, colb = cb / (uint32_t)G.nz;
    int cza = (int)(ca - cola * (uint32_t)G.nz), czb = (int)(cb - colb * (uint32_t)G.nz);
    bool merged = (cola == colb) && (czb - cza <= 3);
    // one walk over the union z-range when merged, else one walk per particle (single code copy so
    // merged and split lanes of a warp stay converged inside the walk)
    float ra = 0.f, rb = 0.f;
    int npass = merged ? 1 : 2;
#pragma unroll 1
    for (int p = 0; p < npass; p++) {
        float4 q0 = p ? pb : pa;
        float4 q1 = merged ? pb : q0;
        uint32_t col = p ? colb : cola;
        int czlo = p ? czb : cza;
        int czhi = merged ? czb : czlo;
        int cy = (int)(col % (uint32_t)G.ny), cx = (int)(col / (uint32_t)G.ny);
        int z0 = czlo > 0 ? czlo - 1 : 0, z1 = czhi < G.nz - 1 ? czhi + 1 : czhi;
        float2 r = density_walk_pair<S>(G, cell_start, posq, cx, cy, z0, z1, make_float2(q0.x, q1.x), make_float2(q0.y, q1.y),
                                        make_float2(q0.z, q1.z), C.hh, slice);
        if (p == 0) { ra = r.x; rb = r.y; } else { rb = r.y; }
    }
#pragma unroll
    for (int o = 1; o < S; o <<= 1) {
        ra += __shfl_xor_sync(SPHE_FULL, ra, o);
        rb += __shfl_xor_sync(SPHE_FULL, rb, o);
    }
    if (slice != 0) return;
    ra *= C.densK; rb *= C.densK;
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

// ------------------------------------------------------------------ passes 2+3, variant 1: pair + compaction
// ncu on k_force_tpp (profiles/r01_ncu_density_force_tpp.txt): 15.7 of 32 lanes active per instruction --
// the ~45-instruction neighbour body ran on every candidate iteration with ~15 % of the lanes.  Here:
//   phase 1 (test):    two targets per thread, packed FADD2/FMUL2/FFMA2 distance test per candidate, and the
//                      indices of candidates that are a neighbour of EITHER target are appended to a
//                      per-thread list in shared memory ([entry][thread] layout, conflict free);
//   phase 2 (process): the thread walks its own dense list with the packed neighbour body for both targets.
// Weights are clamped (max(h^2-d2,0), max(h-r,0)) so a candidate that is a neighbour of only one of the
// two targets contributes exactly 0 to the other.  A full list is flushed in place (rare).
constexpr int FORCE_LIST_CAP = 48;
constexpr int FORCE_THREADS = 128;

template <bool DIAG>
__global__ void __launch_bounds__(FORCE_THREADS) k_force_pair(int n_hi, const int* __restrict__ n_dev, const float4* __restrict__ posq_q, const float4* __restrict__ velv,
                                                    const float* __restrict__ rho, const int* __restrict__ ids,
                                                    const uint32_t* __restrict__ cell_sorted,
                                                    const int* __restrict__ cell_start, GridP G, StepC C,
                                                    float4* __restrict__ posq_out, float4* __restrict__ velv_out, DiagOut D) {
    const int n = n_dev ? __ldg(n_dev) : n_hi;  // exact count from the device in slab mode, else the launch bound
    __shared__ int list[(FORCE_LIST_CAP + 1) * FORCE_THREADS];  // +1: trash slot for saturated appends
    const int tid = threadIdx.x;
    int a = 2 * (blockIdx.x * blockDim.x + tid);
    if (a >= n) return;
    int b = (a + 1 < n) ? a + 1 : a;
    const float4 pa = posq_q[a], pb = posq_q[b];
    const float4 va = velv[a], vb = velv[b];
    const uint32_t ca = cell_sorted[a], cb = cell_sorted[b];
    const uint32_t cola = ca / (uint32_t)G.nz, colb = cb / (uint32_t)G.nz;
    const int cza = (int)(ca - cola * (uint32_t)G.nz), czb = (int)(cb - colb * (uint32_t)G.nz);
    const bool merged = (b != a) && (cola == colb) && (czb - cza <= 3);
    const float inv_sqrt3 = 0.57735026f;
    const float FAR = 1.0e18f;  // dummy target: every weight clamps to exactly 0

    float2 A_x = {0.f, 0.f}, A_y = {0.f, 0.f}, A_z = {0.f, 0.f};  // sum (q_i+q_j)(h-r)^2 dir
    float2 F_x = {0.f, 0.f}, F_y = {0.f, 0.f}, F_z = {0.f, 0.f};  // sum (v_j-v_i) vol_j (h-r)
    float2 N_x = {0.f, 0.f}, N_y = {0.f, 0.f}, N_z = {0.f, 0.f};  // sum vol_j (h^2-r^2)^2 d
    float2 CF = {0.f, 0.f};                                        // sum vol_j (h^2-r^2)(3h^2-7r^2)
    int maxa = -1, maxb = -1;

    const int npass = (merged || b == a) ? 1 : 2;
#pragma unroll 1
    for (int p = 0; p < npass; p++) {
        // targets of this pass: (a,b) merged, else (a,FAR) then (FAR,b)
        const bool useA = merged || p == 0, useB = merged || p == 1;
        const float2 X = make_float2(useA ? pa.x : FAR, useB ? pb.x : FAR);
        const float2 Y = make_float2(useA ? pa.y : FAR, useB ? pb.y : FAR);
        const float2 Z = make_float2(useA ? pa.z : FAR, useB ? pb.z : FAR);
        const float2 Q = make_float2(pa.w, pb.w);
        const float2 VX = make_float2(va.x, vb.x), VY = make_float2(va.y, vb.y), VZ = make_float2(va.z, vb.z);
        const int ia = useA ? a : -1, ib = (useB && b != a) ? b : -1;
        const uint32_t col = p ? colb : cola;
        const int czlo = p ? czb : cza, czhi = merged ? czb : czlo;
        const int cy = (int)(col % (uint32_t)G.ny), cx = (int)(col / (uint32_t)G.ny);
        const int z0 = czlo > 0 ? czlo - 1 : 0, z1 = czhi < G.nz - 1 ? czhi + 1 : czhi;

        // neighbour body for candidate k against both targets (packed)
        auto body = [&](const int k) {
            const float4 pj = __ldg(&posq_q[k]);
            const float4 vj = __ldg(&velv[k]);
            float2 dx = __fadd2_rn(X, make_float2(-pj.x, -pj.x));
            float2 dy = __fadd2_rn(Y, make_float2(-pj.y, -pj.y));
            float2 dz = __fadd2_rn(Z, make_float2(-pj.z, -pj.z));
            float2 d2 = __fmul2_rn(dx, dx);
            d2 = __ffma2_rn(dy, dy, d2);
            d2 = __ffma2_rn(dz, dz, d2);
            float2 w = __fadd2_rn(make_float2(C.hh, C.hh), make_float2(-d2.x, -d2.y));
            w.x = fmaxf(w.x, 0.f); w.y = fmaxf(w.y, 0.f);
            float2 vw = __fmul2_rn(w, make_float2(vj.w, vj.w));
            float2 t7 = __ffma2_rn(d2, make_float2(-7.0f, -7.0f), make_float2(C.hh3, C.hh3));
            CF = __ffma2_rn(vw, t7, CF);
            float2 vww = __fmul2_rn(vw, w);
            N_x = __ffma2_rn(vww, dx, N_x); N_y = __ffma2_rn(vww, dy, N_y); N_z = __ffma2_rn(vww, dz, N_z);
            // d2 is clamped away from 0, so the flush-to-zero approximation never sees a denormal
            float2 rinv = make_float2(rsqrt_ftz(fmaxf(d2.x, 1e-30f)), rsqrt_ftz(fmaxf(d2.y, 1e-30f)));
            float2 r = __fmul2_rn(d2, rinv);
            float2 hm = __fadd2_rn(make_float2(C.h, C.h), make_float2(-r.x, -r.y));
            hm.x = fmaxf(hm.x, 0.f); hm.y = fmaxf(hm.y, 0.f);
            float2 tv = __fmul2_rn(hm, make_float2(vj.w, vj.w));
            float2 dvx = __fadd2_rn(make_float2(vj.x, vj.x), make_float2(-VX.x, -VX.y));
            float2 dvy = __fadd2_rn(make_float2(vj.y, vj.y), make_float2(-VY.x, -VY.y));
            float2 dvz = __fadd2_rn(make_float2(vj.z, vj.z), make_float2(-VZ.x, -VZ.y));
            F_x = __ffma2_rn(tv, dvx, F_x); F_y = __ffma2_rn(tv, dvy, F_y); F_z = __ffma2_rn(tv, dvz, F_z);
            float2 sq = __fadd2_rn(Q, make_float2(pj.w, pj.w));
            float2 sc = __fmul2_rn(__fmul2_rn(sq, hm), hm);
            // pressure excludes j == i (fluid_system.h:142)
            if (k == ia) sc.x = 0.f;
            if (k == ib) sc.y = 0.f;
            float2 ux = __fmul2_rn(dx, rinv), uy = __fmul2_rn(dy, rinv), uz = __fmul2_rn(dz, rinv);
            if (fminf(r.x, r.y) <= 1e-4f) {
                // coincident pair: direction (1,1,1)/sqrt(3) (fluid_system.h:438-440); also taken by the
                // self entry, whose pressure weight is already 0
                if (r.x <= 1e-4f) { ux.x = inv_sqrt3; uy.x = inv_sqrt3; uz.x = inv_sqrt3; }
                if (r.y <= 1e-4f) { ux.y = inv_sqrt3; uy.y = inv_sqrt3; uz.y = inv_sqrt3; }
            }
            A_x = __ffma2_rn(sc, ux, A_x); A_y = __ffma2_rn(sc, uy, A_y); A_z = __ffma2_rn(sc, uz, A_z);
            if (DIAG) {
                // NeighbId needs the reference's exact predicate (bit-exact neighbour set)
                int idk = __ldg(&ids[k]);
                if (k != ia && ia >= 0 && dist2_exact(pa.x - pj.x, pa.y - pj.y, pa.z - pj.z) <= C.T) maxa = max(maxa, idk);
                if (k != ib && ib >= 0 && dist2_exact(pb.x - pj.x, pb.y - pj.y, pb.z - pj.z) <= C.T) maxb = max(maxb, idk);
            }
        };

        // Phase 1: fill.  Same nested loops for every lane (warp stays converged) and a BRANCH-FREE append:
        // every candidate is stored at the lane's current slot and the slot only advances on a pass, so a
        // failing candidate is overwritten by the next one.  Slots saturate at CAP; overflow is detected
        // afterwards and those (rare, strongly compressed) lanes redo the walk without a list.
        int* const lbase = list + tid;
        int off = 0;  // slot * FORCE_THREADS
#pragma unroll 1
        for (int r = 0; r < 9; r++) {
            int x = cx + r / 3 - 1, y = cy + r % 3 - 1;
            if (x < 0 || x >= G.nx || y < 0 || y >= G.ny) continue;
            int base = (x * G.ny + y) * G.nz;
            int s = __ldg(&cell_start[base + z0]);
            int e = __ldg(&cell_start[base + z1 + 1]);
#pragma unroll 4
            for (int k = s; k < e; k++) {
                float4 pj = __ldg(&posq_q[k]);
                float2 dx = __fadd2_rn(X, make_float2(-pj.x, -pj.x));
                float2 dy = __fadd2_rn(Y, make_float2(-pj.y, -pj.y));
                float2 dz = __fadd2_rn(Z, make_float2(-pj.z, -pj.z));
                float2 d2 = __fmul2_rn(dx, dx);
                d2 = __ffma2_rn(dy, dy, d2);
                d2 = __ffma2_rn(dz, dz, d2);
                lbase[min(off, FORCE_LIST_CAP * FORCE_THREADS)] = k;
                off += (fminf(d2.x, d2.y) <= C.hh) ? FORCE_THREADS : 0;
            }
        }
        const int cnt = off / FORCE_THREADS;
        if (cnt <= FORCE_LIST_CAP) {
            // Phase 2: dense walk over the lane's own list
#pragma unroll 1
            for (int o = 0; o < off; o += FORCE_THREADS) body(lbase[o]);
        } else {
            // overflow: direct walk, body executed under the (divergent) predicate
#pragma unroll 1
            for (int r = 0; r < 9; r++) {
                int x = cx + r / 3 - 1, y = cy + r % 3 - 1;
                if (x < 0 || x >= G.nx || y < 0 || y >= G.ny) continue;
                int base = (x * G.ny + y) * G.nz;
                int s = __ldg(&cell_start[base + z0]);
                int e = __ldg(&cell_start[base + z1 + 1]);
#pragma unroll 1
                for (int k = s; k < e; k++) {
                    float4 pj = __ldg(&posq_q[k]);
                    float ex = X.x - pj.x, ey = Y.x - pj.y, ez = Z.x - pj.z;
                    float gx = X.y - pj.x, gy = Y.y - pj.y, gz = Z.y - pj.z;
                    float da = fmaf(ez, ez, fmaf(ey, ey, ex * ex)), db = fmaf(gz, gz, fmaf(gy, gy, gx * gx));
                    if (fminf(da, db) <= C.hh) body(k);
                }
            }
        }
    }

    // epilogue per target
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
        force_epilogue<DIAG>(i, pi, vi, rho[i], ax, ay, az, fx, fy, fz, nx, ny, nz, cf, p ? maxb : maxa, C, ids, posq_out, velv_out, D);
    }
}

// ------------------------------------------------------------------ variant 7: FOUR targets per thread
// ncu on k_density_list (profiles/r01_ncu_density_list_c2.txt): the L1 data pipe is the binding resource (lsu
// wavefronts 80 % of peak, issue active 55 %): every candidate costs one divergent LDG.128 (~8 distinct lines per
// warp request) that serves only two targets.  Here one candidate load serves FOUR consecutive particles (two
// packed pairs): the union z-range of 4 neighbours in the sorted order is almost always the same 3-4 cells as for
// 2, so the gathers per target halve while the packed math per target stays the same.  Each thread keeps TWO
// pair lists (the format k_force_list reads), 64 threads per CTA = the same 33 KB of shared memory per CTA.
constexpr int QUAD_THREADS = 64;

template <bool PF>
__global__ void __launch_bounds__(QUAD_THREADS) k_density_quad(int n_hi, const int* __restrict__ n_dev, int npairs_pad, const float4* __restrict__ posq,
                                                               float4* __restrict__ posq_q, float4* __restrict__ velv,
                                                               const uint32_t* __restrict__ cell_sorted,
                                                               const int* __restrict__ cell_start, GridP G, StepC C,
                                                               float* __restrict__ rho, int* __restrict__ nlist,
                                                               int2* __restrict__ ncount) {
    const int n = n_dev ? __ldg(n_dev) : n_hi;
    __shared__ int list[2 * (NLIST_CAP + 1) * QUAD_THREADS];   // [pair of the quad][entry][thread], +1: trash slot
    const int tid = threadIdx.x;
    const int q = blockIdx.x * blockDim.x + tid;
    const float FAR = 1.0e18f;
    // targets: pair 0 = (i[0], i[1]), pair 1 = (i[2], i[3]); a missing partner repeats its pair's first target
    int i[4];
    const bool live0 = 4 * q < n, live1 = 4 * q + 2 < n;
    i[0] = live0 ? 4 * q : 0;
    i[1] = (i[0] + 1 < n) ? i[0] + 1 : i[0];
    i[2] = live1 ? 4 * q + 2 : i[0];
    i[3] = (live1 && 4 * q + 3 < n) ? 4 * q + 3 : i[2];
    float4 p[4];
    uint32_t col[4];
    int cz[4];
#pragma unroll
    for (int j = 0; j < 4; j++) {
        p[j] = posq[i[j]];
        const uint32_t c = cell_sorted[i[j]];
        col[j] = c / (uint32_t)G.nz;
        cz[j] = (int)(c - col[j] * (uint32_t)G.nz);
    }
    // the sorted order makes cz non-decreasing inside a column
    const bool quad = live1 && col[0] == col[3] && col[0] == col[1] && col[0] == col[2] && (cz[3] - cz[0] <= 3);
    const bool m0 = (i[1] != i[0]) && col[0] == col[1] && (cz[1] - cz[0] <= 3);
    const bool m1 = (i[3] != i[2]) && col[2] == col[3] && (cz[3] - cz[2] <= 3);
    const int np0 = !live0 ? 0 : ((m0 || i[1] == i[0]) ? 1 : 2);
    const int np1 = !live1 ? 0 : ((m1 || i[3] == i[2]) ? 1 : 2);
    const int nwalk = quad ? 1 : np0 + np1;

    float2 acc0 = make_float2(0.f, 0.f), acc1 = make_float2(0.f, 0.f);
    int* const l0 = list + tid;
    int* const l1 = list + (NLIST_CAP + 1) * QUAD_THREADS + tid;
    int off0 = 0, off1 = 0;      // slot * QUAD_THREADS
    int seg0 = -1, seg1 = -1;    // start of the second segment of a split pair
#pragma unroll 1
    for (int w = 0; w < nwalk; w++) {
        // which targets take part in this walk
        bool u[4];
        int lo, hi;              // targets whose cells bound the z-range
        if (quad) { u[0] = u[1] = u[2] = u[3] = true; lo = 0; hi = 3; }
        else if (w < np0) {
            u[0] = m0 || w == 0; u[1] = m0 || w == 1; u[2] = u[3] = false;
            lo = (w == 0) ? 0 : 1; hi = m0 ? 1 : lo;
            if (w == 1) seg0 = off0;
        } else {
            const int v = w - np0;
            u[0] = u[1] = false; u[2] = m1 || v == 0; u[3] = m1 || v == 1;
            lo = (v == 0) ? 2 : 3; hi = m1 ? 3 : lo;
            if (v == 1) seg1 = off1;
        }
        const float2 X0 = make_float2(u[0] ? p[0].x : FAR, u[1] ? p[1].x : FAR), X1 = make_float2(u[2] ? p[2].x : FAR, u[3] ? p[3].x : FAR);
        const float2 Y0 = make_float2(u[0] ? p[0].y : FAR, u[1] ? p[1].y : FAR), Y1 = make_float2(u[2] ? p[2].y : FAR, u[3] ? p[3].y : FAR);
        const float2 Z0 = make_float2(u[0] ? p[0].z : FAR, u[1] ? p[1].z : FAR), Z1 = make_float2(u[2] ? p[2].z : FAR, u[3] ? p[3].z : FAR);
        const uint32_t cc = lo == 0 ? col[0] : (lo == 1 ? col[1] : (lo == 2 ? col[2] : col[3]));
        const int czlo = lo == 0 ? cz[0] : (lo == 1 ? cz[1] : (lo == 2 ? cz[2] : cz[3]));
        const int czhi = hi == 0 ? cz[0] : (hi == 1 ? cz[1] : (hi == 2 ? cz[2] : cz[3]));
        const int cy = (int)(cc % (uint32_t)G.ny), cx = (int)(cc / (uint32_t)G.ny);
        const int z0 = czlo > 0 ? czlo - 1 : 0, z1 = czhi < G.nz - 1 ? czhi + 1 : czhi;
        auto test = [&](const int k, const float4 pj) {
            const float2 nx = make_float2(-pj.x, -pj.x), ny = make_float2(-pj.y, -pj.y), nz = make_float2(-pj.z, -pj.z);
            float2 dx = __fadd2_rn(X0, nx), dy = __fadd2_rn(Y0, ny), dz = __fadd2_rn(Z0, nz);
            float2 ex = __fadd2_rn(X1, nx), ey = __fadd2_rn(Y1, ny), ez = __fadd2_rn(Z1, nz);
            float2 d2 = __fmul2_rn(dx, dx), e2 = __fmul2_rn(ex, ex);
            d2 = __ffma2_rn(dy, dy, d2); e2 = __ffma2_rn(ey, ey, e2);
            d2 = __ffma2_rn(dz, dz, d2); e2 = __ffma2_rn(ez, ez, e2);
            float2 wa = __fadd2_rn(make_float2(C.hh, C.hh), make_float2(-d2.x, -d2.y));
            float2 wb = __fadd2_rn(make_float2(C.hh, C.hh), make_float2(-e2.x, -e2.y));
            // branch-free append to both pair lists: store at the current slot, advance only on a pass
            l0[min(off0, NLIST_CAP * QUAD_THREADS)] = k;
            l1[min(off1, NLIST_CAP * QUAD_THREADS)] = k;
            off0 += (fmaxf(wa.x, wa.y) >= 0.f) ? QUAD_THREADS : 0;
            off1 += (fmaxf(wb.x, wb.y) >= 0.f) ? QUAD_THREADS : 0;
            wa.x = fmaxf(wa.x, 0.f); wa.y = fmaxf(wa.y, 0.f);
            wb.x = fmaxf(wb.x, 0.f); wb.y = fmaxf(wb.y, 0.f);
            acc0 = __ffma2_rn(__fmul2_rn(wa, wa), wa, acc0);
            acc1 = __ffma2_rn(__fmul2_rn(wb, wb), wb, acc1);
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
#pragma unroll 4
                for (int k = s; k < e; k++) test(k, __ldg(&posq[k]));
            }
        } else {
            int s, e, sn, en;
            bounds(0, s, e);
#pragma unroll 1
            for (int r = 0; r < 9; r++) {
                bounds(r + 1, sn, en);
                int k = s;
                if (k + 4 <= e) {
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
#pragma unroll 1
                for (; k < e; k++) test(k, __ldg(&posq[k]));
                s = sn; e = en;
            }
        }
    }
    // lists: pair 2q and 2q+1 sit next to each other in every entry row -> one 8-byte store per entry per thread
    const int cnt0 = off0 / QUAD_THREADS, cnt1 = off1 / QUAD_THREADS;
    const bool fit0 = cnt0 <= NLIST_CAP, fit1 = cnt1 <= NLIST_CAP;
    if (seg0 < 0) seg0 = off0;
    if (seg1 < 0) seg1 = off1;
    const int2 c0 = fit0 ? make_int2(seg0 / QUAD_THREADS, cnt0) : make_int2(-1, -1);
    const int2 c1 = fit1 ? make_int2(seg1 / QUAD_THREADS, cnt1) : make_int2(-1, -1);
    if (live0) ncount[2 * q] = c0;
    if (live1) ncount[2 * q + 1] = c1;
    {
        const int rows = max(fit0 ? cnt0 : 0, fit1 ? cnt1 : 0);
        int2* dst = reinterpret_cast<int2*>(nlist + 2 * q);
        const size_t stride = (size_t)npairs_pad / 2;   // in int2
#pragma unroll 4
        for (int e = 0; e < rows; e++) dst[e * stride] = make_int2(l0[min(e, NLIST_CAP) * QUAD_THREADS], l1[min(e, NLIST_CAP) * QUAD_THREADS]);
    }
    const float racc[4] = {acc0.x, acc0.y, acc1.x, acc1.y};
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const bool first = (j == 0 && live0) || (j == 1 && live0 && i[1] != i[0]) || (j == 2 && live1) || (j == 3 && live1 && i[3] != i[2]);
        if (!first) continue;
        const float r = racc[j] * C.densK;
        const float P = C.k * (r - C.p0);
        rho[i[j]] = r;
        posq_q[i[j]] = make_float4(p[j].x, p[j].y, p[j].z, P / (r * r));
        velv[i[j]].w = C.mass / r;
    }
}

// ------------------------------------------------------------------ variant 5x: S lanes per target pair
// ncu on the list kernels (profiles/r01_ncu_density_list_c2.txt, r01_ncu_force_list_c2.txt): both sit on the
// L1 data pipe (lsu wavefronts 80 % / 65 % of peak, issue active 55 % / 45 %).  The pipe delivers one 128-byte
// line per cycle per SM, and with one pair per lane a warp spans ~10 cells, so every gather request touches
// ~10-20 distinct lines.  Here S consecutive lanes share one target pair and stride the candidates
// (lane % S, step S): a warp spans 32/S pairs (~3 cells at S = 4), the S lanes of a pair read S consecutive
// candidates (one line), pairs of the same cell read the same addresses -> 3-4x fewer lines per request at
// the same number of lane-candidates per request.  Each lane keeps its own sub-list (entry-major in HBM as
// before); the force pass walks the sub-list it is given and the S partial sums are combined with
// __shfl_xor (20 values, log2 S levels: ~3 % of the body work).  Shared memory per CTA drops with S
// (shorter sub-lists), which also lifts the occupancy limit of the S = 1 kernel.
template <int S> struct SList { static constexpr int CAP = S == 1 ? 64 : (S == 2 ? 36 : (S == 4 ? 20 : 12)); };

template <int S>
__global__ void __launch_bounds__(NLIST_THREADS) k_density_slist(int n_hi, const int* __restrict__ n_dev, int stride,
                                                                 const float4* __restrict__ posq, float4* __restrict__ posq_q,
                                                                 float4* __restrict__ velv, const uint32_t* __restrict__ cell_sorted,
                                                                 const int* __restrict__ cell_start, GridP G, StepC C,
                                                                 float* __restrict__ rho, int* __restrict__ nlist,
                                                                 int2* __restrict__ ncount) {
    constexpr int CAP = SList<S>::CAP;
    const int n = n_dev ? __ldg(n_dev) : n_hi;
    __shared__ int list[(CAP + 1) * NLIST_THREADS];  // +1: trash slot for saturated appends
    const int tid = threadIdx.x;
    const int gt = blockIdx.x * blockDim.x + tid;
    const int slice = gt % S;
    int a = 2 * (gt / S);
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
    int off = 0, off0 = 0;  // slot * NLIST_THREADS
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
        if (p == 1) off0 = off;
#pragma unroll 1
        for (int r = 0; r < 9; r++) {
            int x = cx + r / 3 - 1, y = cy + r % 3 - 1;
            if (x < 0 || x >= G.nx || y < 0 || y >= G.ny) continue;
            int base = (x * G.ny + y) * G.nz;
            int s = __ldg(&cell_start[base + z0]);
            int e = __ldg(&cell_start[base + z1 + 1]);
#pragma unroll 4
            for (int k = s + slice; k < e; k += S) {
                float4 pj = __ldg(&posq[k]);
                float2 dx = __fadd2_rn(X, make_float2(-pj.x, -pj.x));
                float2 dy = __fadd2_rn(Y, make_float2(-pj.y, -pj.y));
                float2 dz = __fadd2_rn(Z, make_float2(-pj.z, -pj.z));
                float2 d2 = __fmul2_rn(dx, dx);
                d2 = __ffma2_rn(dy, dy, d2);
                d2 = __ffma2_rn(dz, dz, d2);
                float2 w = __fadd2_rn(make_float2(C.hh, C.hh), make_float2(-d2.x, -d2.y));
                lbase[min(off, CAP * NLIST_THREADS)] = k;  // branch-free append
                off += (fmaxf(w.x, w.y) >= 0.f) ? NLIST_THREADS : 0;
                w.x = fmaxf(w.x, 0.f);
                w.y = fmaxf(w.y, 0.f);
                acc = __ffma2_rn(__fmul2_rn(w, w), w, acc);
            }
        }
    }
    if (npass == 1) off0 = off;
    const int cnt = off / NLIST_THREADS;
    // the S sub-lists of a pair are used together: if one overflowed, all of them fall back
    int fits = cnt <= CAP;
#pragma unroll
    for (int o = 1; o < S; o <<= 1) {
        fits &= __shfl_xor_sync(SPHE_FULL, fits, o);
        acc.x += __shfl_xor_sync(SPHE_FULL, acc.x, o);
        acc.y += __shfl_xor_sync(SPHE_FULL, acc.y, o);
    }
    if (live) ncount[gt] = fits ? make_int2(off0 / NLIST_THREADS, cnt) : make_int2(-1, -1);
    if (fits) {
        int* dst = nlist + gt;
#pragma unroll 4
        for (int o = 0, e = 0; o < off; o += NLIST_THREADS, e++) dst[(size_t)e * stride] = lbase[o];
    }
    if (!live || slice != 0) return;
    float ra = acc.x * C.densK, rb = acc.y * C.densK;
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

template <bool DIAG, int S>
__global__ void __launch_bounds__(NLIST_THREADS, FL_MINB) k_force_slist(int n_hi, const int* __restrict__ n_dev, int stride,
                                                               const float4* __restrict__ posq_q, const float4* __restrict__ velv,
                                                               const float* __restrict__ rho, const int* __restrict__ ids,
                                                               const uint32_t* __restrict__ cell_sorted,
                                                               const int* __restrict__ cell_start, GridP G, StepC C,
                                                               const int* __restrict__ nlist, const int2* __restrict__ ncount,
                                                               float4* __restrict__ posq_out, float4* __restrict__ velv_out, DiagOut D) {
    const int n = n_dev ? __ldg(n_dev) : n_hi;
    const int gt = blockIdx.x * blockDim.x + threadIdx.x;
    const int slice = gt % S;
    int a = 2 * (gt / S);
    const bool live = a < n;   // dead lanes stay for the shuffles (a whole S-group is dead or alive together)
    if (!live) a = 0;
    const int b = (a + 1 < n) ? a + 1 : a;
    const float4 pa = posq_q[a], pb = posq_q[b];
    const float4 va = velv[a], vb = velv[b];
    const int2 cn = live ? ncount[gt] : make_int2(0, 0);
    const float inv_sqrt3 = 0.57735026f;
    const float FAR = 1.0e18f;

    float2 A_x = {0.f, 0.f}, A_y = {0.f, 0.f}, A_z = {0.f, 0.f};
    float2 F_x = {0.f, 0.f}, F_y = {0.f, 0.f}, F_z = {0.f, 0.f};
    float2 N_x = {0.f, 0.f}, N_y = {0.f, 0.f}, N_z = {0.f, 0.f};
    float2 CF = {0.f, 0.f};
    int maxa = -1, maxb = -1;
    const float2 Q = make_float2(pa.w, pb.w);
    const float2 VX = make_float2(va.x, vb.x), VY = make_float2(va.y, vb.y), VZ = make_float2(va.z, vb.z);

    // the pair's own geometry decides merged/split (all S lanes agree; a sub-list may be empty)
    const uint32_t ca = cell_sorted[a], cb = cell_sorted[b];
    const uint32_t cola = ca / (uint32_t)G.nz, colb = cb / (uint32_t)G.nz;
    const int cza = (int)(ca - cola * (uint32_t)G.nz), czb = (int)(cb - colb * (uint32_t)G.nz);
    const bool merged = (b != a) && (cola == colb) && (czb - cza <= 3);
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
        w.x = fmaxf(w.x, 0.f); w.y = fmaxf(w.y, 0.f);
        float2 vw = __fmul2_rn(w, make_float2(vj.w, vj.w));
        float2 t7 = __ffma2_rn(d2, make_float2(-7.0f, -7.0f), make_float2(C.hh3, C.hh3));
        CF = __ffma2_rn(vw, t7, CF);
        float2 vww = __fmul2_rn(vw, w);
        N_x = __ffma2_rn(vww, dx, N_x); N_y = __ffma2_rn(vww, dy, N_y); N_z = __ffma2_rn(vww, dz, N_z);
        float2 rinv = make_float2(rsqrt_ftz(fmaxf(d2.x, 1e-30f)), rsqrt_ftz(fmaxf(d2.y, 1e-30f)));
        float2 r = __fmul2_rn(d2, rinv);
        float2 hm = __fadd2_rn(make_float2(C.h, C.h), make_float2(-r.x, -r.y));
        hm.x = fmaxf(hm.x, 0.f); hm.y = fmaxf(hm.y, 0.f);
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

    if (live && cn.y >= 0) {
        const int* src = nlist + gt;
#pragma unroll 1
        for (int seg = 0; seg < 2; seg++) {
            // segment 0 = entries [0, cn.x): targets (a, b) when merged, else (a, FAR); segment 1 = [cn.x, cn.y): (FAR, b)
            const int e0 = seg ? cn.x : 0, e1 = seg ? cn.y : cn.x;
            if (e0 >= e1) continue;
            const bool useA = seg == 0, useB = merged || seg == 1;
            X = make_float2(useA ? pa.x : FAR, useB ? pb.x : FAR);
            Y = make_float2(useA ? pa.y : FAR, useB ? pb.y : FAR);
            Z = make_float2(useA ? pa.z : FAR, useB ? pb.z : FAR);
            ia = useA ? a : -1; ib = (useB && b != a) ? b : -1;
            int k = __ldg(&src[(size_t)e0 * stride]);
            float4 pj = __ldg(&posq_q[k]);
            float4 vj = __ldg(&velv[k]);
#pragma unroll 1
            for (int e = e0; e < e1; e++) {
                const int kn = (e + 1 < e1) ? __ldg(&src[(size_t)(e + 1) * stride]) : k;
                const float4 pjn = __ldg(&posq_q[kn]);
                const float4 vjn = __ldg(&velv[kn]);
                body(k, pj, vj);
                k = kn; pj = pjn; vj = vjn;
            }
        }
    } else if (live) {
        // a sub-list overflowed (strong compression): direct walks, the S lanes stride the candidates
        const int npass = (merged || b == a) ? 1 : 2;
#pragma unroll 1
        for (int p = 0; p < npass; p++) {
            const bool useA = merged || p == 0, useB = merged || p == 1;
            X = make_float2(useA ? pa.x : FAR, useB ? pb.x : FAR);
            Y = make_float2(useA ? pa.y : FAR, useB ? pb.y : FAR);
            Z = make_float2(useA ? pa.z : FAR, useB ? pb.z : FAR);
            ia = useA ? a : -1; ib = (useB && b != a) ? b : -1;
            const uint32_t col = p ? colb : cola;
            const int czlo = p ? czb : cza, czhi = merged ? czb : czlo;
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
                for (int k = s + slice; k < e; k += S) {
                    float4 pj = __ldg(&posq_q[k]);
                    float ex = X.x - pj.x, ey = Y.x - pj.y, ez = Z.x - pj.z;
                    float gx = X.y - pj.x, gy = Y.y - pj.y, gz = Z.y - pj.z;
                    float da = fmaf(ez, ez, fmaf(ey, ey, ex * ex)), db = fmaf(gz, gz, fmaf(gy, gy, gx * gx));
                    if (fminf(da, db) <= C.hh) body(k, pj, __ldg(&velv[k]));
                }
            }
        }
    }

    // combine the S partial sums of the pair
#pragma unroll
    for (int o = 1; o < S; o <<= 1) {
#define SPHE_RED2(v) v.x += __shfl_xor_sync(SPHE_FULL, v.x, o); v.y += __shfl_xor_sync(SPHE_FULL, v.y, o);
        SPHE_RED2(A_x) SPHE_RED2(A_y) SPHE_RED2(A_z) SPHE_RED2(F_x) SPHE_RED2(F_y) SPHE_RED2(F_z)
        SPHE_RED2(N_x) SPHE_RED2(N_y) SPHE_RED2(N_z) SPHE_RED2(CF)
#undef SPHE_RED2
        if (DIAG) { maxa = max(maxa, __shfl_xor_sync(SPHE_FULL, maxa, o)); maxb = max(maxb, __shfl_xor_sync(SPHE_FULL, maxb, o)); }
    }
    if (!live) return;
    // lane 0 of the group finishes target a, lane 1 (if any) target b
#pragma unroll 1
    for (int p = 0; p < 2; p++) {
        if (p == 1 && b == a) break;
        if (S > 1 && slice != p) continue;
        if (S == 1 || slice == p) {
            const int i = p ? b : a;
            const float4 pi = p ? pb : pa;
            const float4 vi = p ? vb : va;
            const float ax = p ? A_x.y : A_x.x, ay = p ? A_y.y : A_y.x, az = p ? A_z.y : A_z.x;
            const float fx = p ? F_x.y : F_x.x, fy = p ? F_y.y : F_y.x, fz = p ? F_z.y : F_z.x;
            const float nx = p ? N_x.y : N_x.x, ny = p ? N_y.y : N_y.x, nz = p ? N_z.y : N_z.x;
            const float cf = p ? CF.y : CF.x;
            force_epilogue<DIAG>(i, pi, vi, rho[i], ax, ay, az, fx, fy, fz, nx, ny, nz, cf, p ? maxb : maxa, C, ids, posq_out, velv_out, D);
        }
    }
}

int slist_threads_pad(int n, int S) { return ((((n + 1) / 2) * S) + NLIST_THREADS - 1) / NLIST_THREADS * NLIST_THREADS; }
int slist_entries(int S) { return S == 1 ? 64 : (S == 2 ? 36 : (S == 4 ? 20 : 12)); }

template <int S>
static void launch_density_s(cudaStream_t st, int n, const int* n_dev, const float4* posq, float4* posq_q, float4* velv,
                             const uint32_t* cell_sorted, const int* cell_start, const GridP& G, const StepC& C, float* rho,
                             int* nlist, int2* ncount) {
    int tp = slist_threads_pad(n, S);
    k_density_slist<S><<<tp / NLIST_THREADS, NLIST_THREADS, 0, st>>>(n, n_dev, tp, posq, posq_q, velv, cell_sorted, cell_start, G, C, rho, nlist, ncount);
}
template <int S>
static void launch_force_s(cudaStream_t st, int n, const int* n_dev, const float4* posq_q, const float4* velv, const float* rho,
                           const int* ids, const uint32_t* cell_sorted, const int* cell_start, const GridP& G, const StepC& C,
                           float4* posq_out, float4* velv_out, const DiagOut* diag, const int* nlist, const int2* ncount) {
    int tp = slist_threads_pad(n, S);
    dim3 g(tp / NLIST_THREADS), b(NLIST_THREADS);
    if (diag) k_force_slist<true, S><<<g, b, 0, st>>>(n, n_dev, tp, posq_q, velv, rho, ids, cell_sorted, cell_start, G, C, nlist, ncount, posq_out, velv_out, *diag);
    else k_force_slist<false, S><<<g, b, 0, st>>>(n, n_dev, tp, posq_q, velv, rho, ids, cell_sorted, cell_start, G, C, nlist, ncount, posq_out, velv_out, DiagOut{});
}

