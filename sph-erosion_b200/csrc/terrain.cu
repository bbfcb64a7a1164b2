// Terrain half of the hot path.  THIS TRANSLATION UNIT IS COMPILED WITH -fmad=false (build.py): the
// contact search is a chain of float comparisons (ray/triangle tests, nearest-corner selection), so
// its decisions only agree with the reference if every multiply and add rounds separately like the
// reference's x86-64 build.  Divisions and square roots are IEEE (nvcc defaults).
//
//   collide()          Grid::collision + helpers, Erosion/grid.h:178-805: contact point and normal of a
//                      moving particle against the two-triangles-per-cell heightfield surface.
//   (cull)             in the force kernels' epilogue (common.cuh terrain_may_touch): only particles at or below the
//                      local maximum of the heightfield round their cells reach the contact search.
//   k_terrain_contact  the call the reference has commented out in advance() (fluid_system.h:335-340):
//                      contact response, then the box collision (:342-347), + this project's erosion
//                      requests (the reference has no erosion code; model in DESIGN.md).
//   k_terrain_grant    shares the material available above bedrock among the pick-up requests.
//   k_terrain_apply    heights += deposits - grants; clears the per-vertex accumulators.
//   k_terrain_lmax     per-cell maximum of the 4x4 vertex neighbourhood the contact search can touch (the cull bound).
//   k_terrain_surface  UpdateGrid (grid.h:138-176) vertices + normals;  k_terrain_indices  genIndices (:118-136).
//
// Heights are fixed point (int32, 1/4096 of a height unit): every per-vertex accumulation is an integer
// atomic, so terrain and sediment totals are conserved EXACTLY and the result does not depend on the
// order in which particles arrive.  Atomics are warp-aggregated: lanes that hit the same vertex are
// found with __match_any_sync and their amounts summed with __reduce_add_sync before one atomicAdd.
#include "common.cuh"
#include "sim.h"

namespace sphe {

namespace {

struct V3 { float x, y, z; };
struct Tri { V3 A, B, C, n; };

__device__ __forceinline__ V3 mk(float x, float y, float z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ V3 operator+(V3 a, V3 b) { return mk(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ V3 operator-(V3 a, V3 b) { return mk(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ V3 operator*(float s, V3 a) { return mk(s * a.x, s * a.y, s * a.z); }
__device__ __forceinline__ V3 operator-(V3 a) { return mk(-a.x, -a.y, -a.z); }
// glm 0.9.9.7 evaluation order (vendor/glm/glm/detail/func_geometric.inl:47-90)
__device__ __forceinline__ float dot(V3 a, V3 b) { float tx = a.x * b.x, ty = a.y * b.y, tz = a.z * b.z; return tx + ty + tz; }
__device__ __forceinline__ V3 cross(V3 x, V3 y) { return mk(x.y * y.z - y.y * x.z, x.z * y.x - y.z * x.x, x.x * y.y - y.x * x.y); }
__device__ __forceinline__ V3 unit(V3 a) { float inv = 1.0f / sqrtf(dot(a, a)); return mk(a.x * inv, a.y * inv, a.z * inv); }
__device__ __forceinline__ float norm3(V3 a) { return sqrtf(dot(a, a)); }
__device__ __forceinline__ float norm2(float x, float y) { return sqrtf(x * x + y * y); }

// GetHeightfieldAt (grid.h:104-107), index clamped (the reference reads out of bounds next to the border)
__device__ __forceinline__ float height(const TerrainDev& T, int x, int z) {
    x = min(max(x, 0), T.rows - 1);
    z = min(max(z, 0), T.cols - 1);
    return (float)__ldg(&T.hfx[(size_t)T.cols * x + z]) * (1.0f / 4096.0f);
}

// getCellTriangles (grid.h:195-207): which = 0 -> Triangle(C, B, A), 1 -> Triangle(AA, B, C)
__device__ Tri cell_tri(const TerrainDev& T, float cx, float cz, int which) {
    int ix = (int)cx, iz = (int)cz, ix1 = (int)(cx + 1.0f), iz1 = (int)(cz + 1.0f);
    V3 B = mk(cx + 1.0f, height(T, ix1, iz), cz);
    V3 C = mk(cx, height(T, ix, iz1), cz + 1.0f);
    Tri t;
    if (which == 0) { t.A = C; t.B = B; t.C = mk(cx, height(T, ix, iz), cz); }
    else { t.A = mk(cx + 1.0f, height(T, ix1, iz1), cz + 1.0f); t.B = B; t.C = C; }
    t.n = unit(cross(t.B - t.A, t.C - t.A));  // Triangle ctor, grid.h:14-18
    return t;
}

// rayIntersectsTriangle (grid.h:178-193)
__device__ bool ray_tri(V3 o, V3 d, const Tri& t, float& tt) {
    V3 E1 = t.B - t.A, E2 = t.C - t.A;
    V3 N = cross(E1, E2);
    float det = -dot(d, N);
    float inv = 1.0f / det;  // == (float)(1.0 / det): double division then rounding is innocuous for a quotient
    V3 AO = o - t.A;
    V3 DAO = cross(AO, d);
    float u = dot(E2, DAO) * inv;
    float v = -dot(E1, DAO) * inv;
    tt = dot(AO, N) * inv;
    return (double)fabsf(det) >= 1e-6 && tt >= 0.0f && u >= 0.0f && v >= 0.0f && (u + v) <= 1.0f;
}

__device__ float ray_dist(V3 o, V3 d, const Tri& t) {  // grid.h:486-499, :719-730
    float tt;
    if (!ray_tri(o, d, t, tt)) return INFINITY;
    return norm3((o + tt * d) - o);
}

__device__ __forceinline__ bool first_wins(float d0, float d1) {  // grid.h:504, :733
    return d0 < d1 || (fabsf(d0 - d1) < 1.1920928955078125e-7f && d0 != INFINITY);
}

__device__ bool project_on(const Tri& t, V3 p, V3 d, V3& cp, V3& n) {  // mappedOnTriangle, grid.h:271-283
    float tt;
    if (!ray_tri(p, d, t, tt)) return false;
    n = d; cp = p + tt * d;
    return true;
}

__device__ bool project_between(const Tri& a, const Tri& b, V3 p, V3& cp, V3& n) {  // mappedBetweenTriangles, :285-305
    V3 d = unit(a.n + b.n);
    return project_on(a, p, d, cp, n) || project_on(b, p, d, cp, n);
}

// findAdjacentCell (grid.h:210-269): one 2-D DDA step from the cell of `p` along d
__device__ bool next_cell(const TerrainDev& T, V3 p, V3 d, float& cx, float& cz) {
    float tx = INFINITY, tz = INFINITY;
    if (d.x < 0.0f) tx = (floorf(p.x) - p.x) / d.x; else if (d.x > 0.0f) tx = ((floorf(p.x) + 1.0f) - p.x) / d.x;
    if (d.z < 0.0f) tz = (floorf(p.z) - p.z) / d.z; else if (d.z > 0.0f) tz = ((floorf(p.z) + 1.0f) - p.z) / d.z;
    if (tx < tz) { if (d.x < 0.0f) cx -= 1.0f; else if (d.x > 0.0f) cx += 1.0f; }
    else { if (d.z < 0.0f) cz -= 1.0f; else if (d.z > 0.0f) cz += 1.0f; }
    return !(cx < 0.0f || cx >= (float)T.dimx || cz < 0.0f || cz >= (float)T.dimz);
}

// cornerCaseABC / cornerCaseAABC (grid.h:320-440).  The fan of six triangles round the grid vertex
// nearest to posCurr, in the reference's order (table FAN); `low` selects the ABC flavour (vertex candidates origin/right/down) or the AABC flavour
// (origin+1/up/left).
// rows: [flavour][nearest][entry] = {dx, dz, which}
__constant__ signed char FAN[2][3][6][3] = {
    { { {0,0,0}, {-1,0,0}, {-1,0,1}, {0,-1,0}, {0,-1,1}, {-1,-1,1} },      // ABC, nearest = origin
      { {0,0,0}, {0,0,1}, {1,0,0}, {0,-1,1}, {1,-1,0}, {1,-1,1} },         // ABC, nearest = right
      { {0,0,0}, {0,0,1}, {-1,0,1}, {0,1,0}, {-1,1,0}, {-1,1,1} } },       // ABC, nearest = down
    { { {0,0,1}, {1,0,0}, {1,0,1}, {0,1,0}, {0,1,1}, {1,1,0} },            // AABC, nearest = origin + (1,1)
      { {0,0,0}, {0,0,1}, {1,0,0}, {0,-1,1}, {1,-1,0}, {1,-1,1} },         // AABC, nearest = up
      { {0,0,0}, {0,0,1}, {-1,0,1}, {0,1,0}, {-1,1,0}, {-1,1,1} } } };     // AABC, nearest = left
__device__ __forceinline__ void fan_entry(bool low, int sel, int k, int& dx, int& dz, int& which) {
    const signed char* e = FAN[low ? 0 : 1][sel][k];
    dx = e[0]; dz = e[1]; which = e[2];
}

__device__ __noinline__ bool corner(const TerrainDev& T, bool low, V3 pc, V3 pn, float cx, float cz, V3& cp, V3& n) {
    float ox = floorf(pc.x) + (low ? 0.0f : 1.0f), oz = floorf(pc.z) + (low ? 0.0f : 1.0f);
    float s = low ? 1.0f : -1.0f;
    float d0 = norm2(ox - pc.x, oz - pc.z);
    float dX = norm2((ox + s) - pc.x, oz - pc.z);   // right (ABC) / left (AABC)
    float dZ = norm2(ox - pc.x, (oz + s) - pc.z);   // down (ABC) / up (AABC)
    // min3(dorigin, dright, ddown) resp. min3(dorigin, dup, dleft), grid.h:307-318
    float a = d0, b = low ? dX : dZ, c = low ? dZ : dX;
    float m = (a < b) ? ((a < c) ? a : c) : ((b < c) ? b : c);
    int sel = (m == a) ? 0 : ((m == b) ? 1 : 2);
    V3 sum = mk(0.f, 0.f, 0.f);
    for (int k = 0; k < 6; k++) {
        int dx, dz, w;
        fan_entry(low, sel, k, dx, dz, w);
        sum = sum + cell_tri(T, cx + (float)dx, cz + (float)dz, w).n;
    }
    V3 d = unit(sum);
    for (int k = 0; k < 6; k++) {
        int dx, dz, w;
        fan_entry(low, sel, k, dx, dz, w);
        if (project_on(cell_tri(T, cx + (float)dx, cz + (float)dz, w), pn, d, cp, n)) return true;
    }
    return false;
}

// Grid::collision (grid.h:462-805) in terrain coordinates
__device__ bool collide(const TerrainDev& T, V3 pc, V3 pn, V3 vel, V3& cp, V3& n) {
    V3 dir = unit(vel), back = -dir;
    float cx = floorf(pc.x), cz = floorf(pc.z), nx = floorf(pn.x), nz = floorf(pn.z);
    float mx = (float)(T.dimx - 1), mz = (float)(T.dimz - 1);
    if (cx < 0.0f || cx >= mx || cz < 0.0f || cz >= mz || nx < 0.0f || nx >= mx || nz < 0.0f || nz >= mz) return false;
    Tri t0 = cell_tri(T, cx, cz, 0), t1 = cell_tri(T, cx, cz, 1);
    float tt;
    if (cx == nx && cz == nz) {
        float d0 = ray_dist(pn, back, t0), d1 = ray_dist(pn, back, t1);
        int pick = first_wins(d0, d1) ? 0 : (d1 < d0 ? 1 : -1);
        if (pick < 0) return false;
        const Tri& me = pick ? t1 : t0;
        if (ray_tri(pn, me.n, me, tt)) { n = me.n; cp = pn + tt * n; return true; }
        float ax = cx, az = cz;
        if (!next_cell(T, pn, unit(dir + me.n), ax, az)) return false;
        float ex = ax - cx, ez = az - cz;
        bool hyp = pick == 0 ? (ex > 0.0f || ez > 0.0f) : (ex < 0.0f || ez < 0.0f);
        bool hit = hyp ? project_between(t0, t1, pn, cp, n) : project_between(me, cell_tri(T, ax, az, pick ? 0 : 1), pn, cp, n);
        return hit || corner(T, pick == 0, pc, pn, cx, cz, cp, n);
    }
    Tri u0 = cell_tri(T, nx, nz, 0), u1 = cell_tri(T, nx, nz, 1);
    float ex = nx - cx, ez = nz - cz;
    if (ex != 0.0f && ez != 0.0f) return corner(T, !(ex + ez == 2.0f), pc, pn, cx, cz, cp, n);
    if (ray_tri(pn, back, t0, tt)) {
        if ((ex == -1.0f || ez == -1.0f) && project_between(t0, u1, pn, cp, n)) return true;
        return corner(T, true, pc, pn, cx, cz, cp, n);
    }
    if (ray_tri(pn, back, t1, tt)) {
        if ((ex == 1.0f || ez == 1.0f) && project_between(t1, u0, pn, cp, n)) return true;
        return corner(T, false, pc, pn, cx, cz, cp, n);
    }
    float d0 = ray_dist(pc, dir, u0), d1 = ray_dist(pc, dir, u1);
    if (first_wins(d0, d1)) {
        if (ray_tri(pn, u0.n, u0, tt)) { n = u0.n; cp = pn + tt * n; return true; }
        if (ex == 1.0f || ez == 1.0f) return project_between(u0, t1, pn, cp, n) || corner(T, false, pc, pn, cx, cz, cp, n);
        return corner(T, true, pc, pn, cx, cz, cp, n);
    }
    if (d1 < d0) {
        if (ray_tri(pn, u1.n, u1, tt)) { n = u1.n; cp = pn + tt * n; return true; }
        if (ex == -1.0f || ez == -1.0f) return project_between(u1, t0, pn, cp, n) || corner(T, true, pc, pn, cx, cz, cp, n);
        return corner(T, false, pc, pn, cx, cz, cp, n);
    }
    return false;
}

// one atomicAdd per distinct vertex per warp: lanes with the same vertex sum their amounts first
__device__ __forceinline__ void vertex_add(unsigned participants, bool active, int* table, int vertex, int amount) {
    if (!active) return;
    unsigned peers = __match_any_sync(participants, vertex);
    int total = __reduce_add_sync(peers, amount);
    if ((int)(threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(&table[vertex], total);
}

// Survivors come in SPHE_SURV_CLASSES lists (surv + c * cap, counts[c]).  The stage kernels walk ONE index space in which
// every class starts at a multiple of 32, so a warp only ever holds survivors of one class: j -> (slot i, live).
struct SurvMap {
    int start[SPHE_SURV_CLASSES + 1];
    __device__ __forceinline__ SurvMap(const int* __restrict__ counts) {
        start[0] = 0;
#pragma unroll
        for (int c = 0; c < SPHE_SURV_CLASSES; c++) start[c + 1] = ((start[c] + __ldg(&counts[c]) + 31) & ~31);
        // start[c + 1] - start[c] is the padded size of class c; the live part is counts[c]
    }
    __device__ __forceinline__ int total() const { return start[SPHE_SURV_CLASSES]; }
    __device__ __forceinline__ int slot(const int* __restrict__ surv, const int* __restrict__ counts, int cap, int j) const {
        int c = 0;
#pragma unroll
        for (int k = 1; k < SPHE_SURV_CLASSES; k++) c += (j >= start[k]) ? 1 : 0;
        const int local = j - start[c];
        return (local < __ldg(&counts[c])) ? __ldg(&surv[(size_t)c * cap + local]) : -1;
    }
};

}  // namespace

// ------------------------------------------------------------------ contact response + erosion requests
// Runs over the SURVIVORS of the exact cull only (a few per cent of the particles): surv[j] is the sorted
// slot of survivor j, *surv_count their number (device word: the host never learns it; the grid is a fixed
// persistent one and every warp strides over the list).  pos_old: positions before the step (.xyz);
// posq / velv: the integrated, un-boxed state the force kernel wrote for survivors; updated in place,
// then the box collision.  req_vertex[j] = vertex of survivor j's pending pick-up request (or -1).
__global__ void __launch_bounds__(128) k_terrain_contact(const int* __restrict__ surv, const int* __restrict__ surv_count, int surv_cap,
                                                         const float4* __restrict__ pos_old, float4* __restrict__ posq,
                                                         float4* __restrict__ velv, int* __restrict__ sediment, StepC C,
                                                         TerrainDev T, int apply_box, int* __restrict__ req_vertex,
                                                         int* __restrict__ req_amount, int* __restrict__ hit_out) {
    const SurvMap M(surv_count);
    const int count = M.total();
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    for (int base = warp * 32; base < count; base += nwarps * 32) {
        const int j = base + lane;
        int i = M.slot(surv, surv_count, surv_cap, j);
        const bool live = i >= 0;
        float4 po = make_float4(0, 0, 0, 0), p4 = po, v4 = po;
        bool hit = false;
        V3 cp = mk(0, 0, 0), nn = mk(0, 0, 0);
        if (live) {
            po = pos_old[i]; p4 = posq[i]; v4 = velv[i];
            V3 pc = mk((po.x - T.ox) * T.inv_scale, (po.y - T.oy) * T.inv_scale, (po.z - T.oz) * T.inv_scale);
            V3 pn = mk((p4.x - T.ox) * T.inv_scale, (p4.y - T.oy) * T.inv_scale, (p4.z - T.oz) * T.inv_scale);
            V3 vn = mk(v4.x * T.inv_scale, v4.y * T.inv_scale, v4.z * T.inv_scale);
            hit = C.dt != 0.0f && collide(T, pc, pn, vn, cp, nn);
            // slab-local terrain: everything this contact read or will write must lie inside the maintained rows
            const float lo = fminf(pc.x, pn.x), hi = fmaxf(pc.x, pn.x);
            if (hit && ((T.win0 > 0 && lo < (float)(T.win0 + 2)) || (T.win1 < T.rows && hi > (float)(T.win1 - 3))))
                atomicAdd(T.violations, 1ull);
        }
        int dep_vertex = 0, dep_amount = 0, want_vertex = 0, want_amount = 0;
        bool dep = false, want = false;
        if (hit) {
            // fluid_system.h:337-339 (world units)
            V3 v = mk(v4.x, v4.y, v4.z);
            V3 cw = mk(cp.x * T.scale + T.ox, cp.y * T.scale + T.oy, cp.z * T.scale + T.oz);
            float d = norm3(mk(p4.x, p4.y, p4.z) - cw);
            float vn_ = dot(v, nn);
            float k = 1.0f + C.cR * (d / (C.dt * norm3(v)));
            float vtl = norm3(v - vn_ * nn);
            V3 v2 = v - (k * vn_) * nn;
            v4.x = v2.x; v4.y = v2.y; v4.z = v2.z;
            p4.x = cw.x; p4.y = cw.y; p4.z = cw.z;
            if (T.erosion) {
                int vx = min(max((int)floorf(cp.x + 0.5f), 0), T.rows - 1);
                int vz = min(max((int)floorf(cp.z + 0.5f), 0), T.cols - 1);
                int c = vx * T.cols + vz;
                int s_fx = sediment[i];
                float cap = T.Kc * vtl;
                float s = (float)s_fx * (1.0f / 4096.0f);
                if (s > cap) {
                    int q = min(__float2int_rn((s - cap) * T.Kd * 4096.0f), s_fx);
                    if (q > 0) { dep = true; dep_vertex = c; dep_amount = q; sediment[i] = s_fx - q; }
                } else if (s < cap) {
                    int q = min(__float2int_rn((cap - s) * T.Ke * 4096.0f), T.max_pickup_fx);
                    if (q > 0) { want = true; want_vertex = c; want_amount = q; }
                }
            }
        }
        unsigned m_hit = __ballot_sync(SPHE_FULL, hit);
        if (m_hit && lane == 0) atomicAdd(T.contacts, (unsigned long long)__popc(m_hit));
        unsigned m_dep = __ballot_sync(SPHE_FULL, dep), m_want = __ballot_sync(SPHE_FULL, want);
        vertex_add(m_dep, dep, T.delta, dep_vertex, dep_amount);
        vertex_add(m_want, want, T.want, want_vertex, want_amount);
        if (!live) continue;
        req_vertex[j] = want ? want_vertex : -1;
        req_amount[j] = want_amount;
        if (hit_out) hit_out[i] = hit ? 1 : 0;
        bool boxed = apply_box && box_collide(C, p4.x, p4.y, p4.z, v4.x, v4.y, v4.z);  // fluid_system.h:342-347
        if (hit || boxed) { posq[i] = p4; velv[i] = v4; }
    }
}

// ------------------------------------------------------------------ share what is above bedrock
__global__ void __launch_bounds__(128) k_terrain_grant(const int* __restrict__ surv, const int* __restrict__ surv_count, int surv_cap,
                                                       const int* __restrict__ req_vertex, const int* __restrict__ req_amount,
                                                       int* __restrict__ sediment, TerrainDev T) {
    const SurvMap M(surv_count);
    const int count = M.total();
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    for (int base = warp * 32; base < count; base += nwarps * 32) {
        const int j = base + lane;
        const int i = M.slot(surv, surv_count, surv_cap, j);
        int c = (i >= 0) ? req_vertex[j] : -1;
        bool act = c >= 0;
        int g = 0;
        if (act) {
            long long avail = (long long)T.hfx_rw[c] - T.hmin_fx;
            if (avail < 0) avail = 0;
            long long w = T.want[c];
            int q = req_amount[j];
            g = (w <= avail) ? q : (int)(((long long)q * avail) / w);
            sediment[i] += g;
        }
        unsigned m = __ballot_sync(SPHE_FULL, act);
        vertex_add(m, act, T.delta, c, -g);
    }
}

__global__ void __launch_bounds__(256) k_terrain_apply(int first, int cells, TerrainDev T) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= cells) return;
    c += first;
    int d = T.delta[c];
    if (d) { T.hfx_rw[c] += d; T.delta[c] = 0; }
    if (T.want[c]) T.want[c] = 0;
}

// every slot is a survivor, all in class 0 (the test hook sphe_terrain_stage_host has no force kernel in front of it)
__global__ void k_iota(int n, int* __restrict__ a, int* __restrict__ count) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) a[i] = i;
    if (i == 0) { count[0] = n; for (int c = 1; c < SPHE_SURV_CLASSES; c++) count[c] = 0; }
}

// lmax[x, z] = max height over the vertices [x-1, x+2] x [z-1, z+2]: everything collide() can touch from cell (x, z)
__global__ void __launch_bounds__(256) k_terrain_lmax(TerrainDev T, int* __restrict__ lmax) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= (T.win1 - T.win0) * T.cols) return;
    c += T.win0 * T.cols;
    int x = c / T.cols, z = c - x * T.cols;
    int m = -0x7fffffff;
#pragma unroll
    for (int dx = -1; dx <= 2; dx++)
#pragma unroll
        for (int dz = -1; dz <= 2; dz++) {
            int xx = min(max(x + dx, 0), T.rows - 1), zz = min(max(z + dz, 0), T.cols - 1);
            m = max(m, T.hfx_rw[xx * T.cols + zz]);
        }
    lmax[c] = m;
}

// ------------------------------------------------------------------ render mesh (UpdateGrid, grid.h:138-176)
__global__ void __launch_bounds__(256) k_terrain_surface(TerrainDev T, float* __restrict__ out) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= T.dimx * T.dimz) return;
    int z = k / T.dimx, x = k - z * T.dimx;
    int y = (int)height(T, x, z);
    if (T.dimy <= y) y = T.dimy - 1;
    V3 u = mk(0, 0, 0), d = u, r = u, l = u;
    if (x - 1 >= 0) l = mk(1.0f, (float)(y - (int)height(T, x - 1, z)), 0.0f);
    if (x + 1 < T.dimx) r = mk(1.0f, (float)((int)height(T, x + 1, z) - y), 0.0f);
    if (z - 1 >= 0) u = mk(0.0f, (float)(y - (int)height(T, x, z - 1)), 1.0f);
    if (z + 1 < T.dimy) d = mk(0.0f, (float)((int)height(T, x, z + 1) - y), 1.0f);  // sic: dim.y, grid.h:167
    V3 nn = unit(((cross(u, l) + cross(u, r)) + cross(d, l)) + cross(d, r));
    float* o = out + 6 * (size_t)k;
    o[0] = (float)x; o[1] = (float)y; o[2] = (float)z; o[3] = nn.x; o[4] = nn.y; o[5] = nn.z;
}

// genIndices (grid.h:118-136): iteration (z, x) emits two triangles iff z < dimz-1 and x < dimx-1
// (the second condition j-1 >= 0 && i-1 >= 0 with j = dimz-1-z, i = dimx-1-x is the same set).
__global__ void __launch_bounds__(256) k_terrain_indices(int dimx, int dimz, unsigned* __restrict__ out) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    int qx = dimx - 1, qz = dimz - 1;
    if (k >= qx * qz) return;
    int z = k / qx, x = k - z * qx;
    int j = dimz - 1 - z, i = dimx - 1 - x;
    unsigned* o = out + 6 * (size_t)k;
    o[0] = z * dimx + x; o[1] = z * dimx + (x + 1); o[2] = (z + 1) * dimx + x;
    o[3] = j * dimx + i; o[4] = j * dimx + (i - 1); o[5] = (j - 1) * dimx + i;
}

// batched Grid::collision for the Grid shim and the parity tests (terrain coordinates in and out)
__global__ void __launch_bounds__(128) k_terrain_collide(int n, const float* __restrict__ pc, const float* __restrict__ pn,
                                                         const float* __restrict__ vn, TerrainDev T, int* __restrict__ hit,
                                                         float* __restrict__ cp_out, float* __restrict__ n_out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    V3 cp = mk(0, 0, 0), nn = mk(0, 0, 0);
    bool h = collide(T, mk(pc[3 * i], pc[3 * i + 1], pc[3 * i + 2]), mk(pn[3 * i], pn[3 * i + 1], pn[3 * i + 2]),
                     mk(vn[3 * i], vn[3 * i + 1], vn[3 * i + 2]), cp, nn);
    hit[i] = h ? 1 : 0;
    cp_out[3 * i] = cp.x; cp_out[3 * i + 1] = cp.y; cp_out[3 * i + 2] = cp.z;
    n_out[3 * i] = nn.x; n_out[3 * i + 1] = nn.y; n_out[3 * i + 2] = nn.z;
}

__global__ void k_heights_from_u8(int cells, const unsigned char* __restrict__ img, int* __restrict__ hfx, int* __restrict__ hmax) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    int h = -0x7fffffff;
    if (c < cells) { h = (int)img[c] * 4096; hfx[c] = h; }
    h = __reduce_max_sync(SPHE_FULL, h);
    if ((threadIdx.x & 31) == 0 && h > -0x7fffffff) atomicMax(hmax, h);
}
__global__ void k_heights_from_f32(int cells, const float* __restrict__ src, int* __restrict__ hfx, int* __restrict__ hmax) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    int h = -0x7fffffff;
    if (c < cells) { h = __float2int_rn(src[c] * 4096.0f); hfx[c] = h; }
    h = __reduce_max_sync(SPHE_FULL, h);
    if ((threadIdx.x & 31) == 0 && h > -0x7fffffff) atomicMax(hmax, h);
}
__global__ void k_heights_to_f32(int cells, const int* __restrict__ hfx, float* __restrict__ dst) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < cells) dst[c] = (float)hfx[c] * (1.0f / 4096.0f);
}
// 64-bit sums for the conservation checks: out[0] += sum of a[0..n)
__global__ void __launch_bounds__(256) k_sum_i32(int n_hi, const int* __restrict__ n_dev, const int* __restrict__ a, const int* __restrict__ ghost_ids, long long* __restrict__ out) {
    const int n = n_dev ? __ldg(n_dev) : n_hi;
    long long s = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        if (!ghost_ids || !(ghost_ids[i] & SPHE_GHOST_BIT)) s += a[i];
#pragma unroll
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(SPHE_FULL, s, o);
    if ((threadIdx.x & 31) == 0 && s) atomicAdd((unsigned long long*)out, (unsigned long long)s);
}

// ------------------------------------------------------------------ launch wrappers
static inline int nb(int n, int b) { return (n + b - 1) / b; }

// persistent grid: (SMs of the current device) x 4 blocks of 128 threads stride over the survivor lists
static int stage_grid() {
    static int sms[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) dev = 0;
    if (!sms[dev]) {
        cudaDeviceProp p;
        sms[dev] = (cudaGetDeviceProperties(&p, dev) == cudaSuccess && p.multiProcessorCount > 0) ? p.multiProcessorCount : 148;
    }
    return sms[dev] * 4;
}
void launch_terrain_stage(cudaStream_t st, const int* surv, const int* surv_count, int surv_cap, const float4* pos_old, float4* posq, float4* velv,
                          int* sediment, const StepC& C, const TerrainDev& T, int apply_box, int* req_vertex, int* req_amount,
                          int* hit_out, int phases) {
    // Multi-GPU slabs run the phases separately: per-vertex `want` is summed over the ranks between contact and
    // grant, `delta` between grant and apply (integer sums: exact and order independent).
    if (phases & TERRAIN_CONTACT)
        k_terrain_contact<<<stage_grid(), 128, 0, st>>>(surv, surv_count, surv_cap, pos_old, posq, velv, sediment, C, T, apply_box, req_vertex, req_amount, hit_out);
    if (T.erosion && C.dt != 0.0f) {
        if (phases & TERRAIN_GRANT) k_terrain_grant<<<stage_grid(), 128, 0, st>>>(surv, surv_count, surv_cap, req_vertex, req_amount, sediment, T);
        if (phases & TERRAIN_APPLY) {
            const int cells = (T.win1 - T.win0) * T.cols;
            k_terrain_apply<<<nb(cells, 256), 256, 0, st>>>(T.win0 * T.cols, cells, T);
            k_terrain_lmax<<<nb(cells, 256), 256, 0, st>>>(T, T.lmax_rw);
        }
    }
}
int terrain_stage_launches(const StepC& C, const TerrainDev& T) { return (T.erosion && C.dt != 0.0f) ? 4 : 1; }
void launch_iota(cudaStream_t st, int n, int* a, int* count) {
    if (n > 0) k_iota<<<nb(n, 256), 256, 0, st>>>(n, a, count);
}
void launch_terrain_lmax(cudaStream_t st, const TerrainDev& T) {
    k_terrain_lmax<<<nb((T.win1 - T.win0) * T.cols, 256), 256, 0, st>>>(T, T.lmax_rw);
}

void launch_terrain_surface(cudaStream_t st, const TerrainDev& T, float* out) {
    int m = T.dimx * T.dimz;
    if (m > 0) k_terrain_surface<<<nb(m, 256), 256, 0, st>>>(T, out);
}
void launch_terrain_indices(cudaStream_t st, int dimx, int dimz, unsigned* out) {
    int m = (dimx - 1) * (dimz - 1);
    if (m > 0) k_terrain_indices<<<nb(m, 256), 256, 0, st>>>(dimx, dimz, out);
}
void launch_terrain_collide(cudaStream_t st, int n, const float* pc, const float* pn, const float* vn, const TerrainDev& T,
                            int* hit, float* cp, float* nrm) {
    if (n > 0) k_terrain_collide<<<nb(n, 128), 128, 0, st>>>(n, pc, pn, vn, T, hit, cp, nrm);
}
void launch_heights_from_u8(cudaStream_t st, int cells, const unsigned char* img, int* hfx, int* hmax) {
    k_heights_from_u8<<<nb(cells, 256), 256, 0, st>>>(cells, img, hfx, hmax);
}
void launch_heights_from_f32(cudaStream_t st, int cells, const float* src, int* hfx, int* hmax) {
    k_heights_from_f32<<<nb(cells, 256), 256, 0, st>>>(cells, src, hfx, hmax);
}
void launch_heights_to_f32(cudaStream_t st, int cells, const int* hfx, float* dst) {
    k_heights_to_f32<<<nb(cells, 256), 256, 0, st>>>(cells, hfx, dst);
}
void launch_sum_i32(cudaStream_t st, int n, const int* a, const int* ghost_ids, long long* out, const int* n_dev) {
    if (n > 0) k_sum_i32<<<min(nb(n, 256), 1184), 256, 0, st>>>(n, n_dev, a, ghost_ids, out);
}

}  // namespace sphe
