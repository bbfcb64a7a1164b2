/* TEST INFRASTRUCTURE ONLY -- CPU oracle for the terrain half of the hot path.
 *
 * (1) Grid::collision and helpers (Erosion/grid.h:178-805): restated as a decision procedure over a
 *     float heightfield.  PINNED: bit-identical hit/miss, contact point and normal against the compiled
 *     reference (oracle/_ref, tests/test_oracle_vs_ref.py) and the committed fixtures
 *     tests/golden/terrain_collision.npz.
 * (2) Grid::UpdateGrid surface vertices + normals and genIndices (grid.h:118-176).  PINNED the same way.
 * (3) Erosion / sediment transport: PARITY UNPINNED -- the reference contains no such code
 *     (SURVEY.md F2).  so_erode_* below is THIS PROJECT'S specification (DESIGN.md "erosion model"),
 *     written independently of the CUDA kernels; tests compare the two bit-exactly and check the
 *     invariants (exact conservation of sediment + terrain volume, monotonicity, no change at rest).
 *
 * Arithmetic: gcc -O2 -ffp-contract=off (no FMA), glm 0.9.9.7 evaluation order:
 *   dot(a,b) = (ax*bx + ay*by) + az*bz        vendor/glm/glm/detail/func_geometric.inl:47-55
 *   cross(x,y) = (x.y*y.z - y.y*x.z, x.z*y.x - y.z*x.x, x.x*y.y - y.x*x.y)             :68-78
 *   normalize(v) = v * (1/sqrt(dot(v,v)))     :82-90, func_exponential.inl:135-139
 */
#include "terrain_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

typedef struct { float x, y, z; } v3;
typedef struct { v3 A, B, C, n; } tri;

static inline v3 mk(float x, float y, float z) { v3 r = { x, y, z }; return r; }
static inline v3 add(v3 a, v3 b) { return mk(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline v3 sub(v3 a, v3 b) { return mk(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline v3 scl(float s, v3 a) { return mk(s * a.x, s * a.y, s * a.z); }
static inline v3 neg(v3 a) { return mk(-a.x, -a.y, -a.z); }
static inline float dot(v3 a, v3 b) { float tx = a.x * b.x, ty = a.y * b.y, tz = a.z * b.z; return tx + ty + tz; }
static inline v3 cross(v3 x, v3 y) { return mk(x.y * y.z - y.y * x.z, x.z * y.x - y.z * x.x, x.x * y.y - y.x * x.y); }
static inline v3 normalize(v3 a) { float inv = 1.0f / sqrtf(dot(a, a)); return mk(a.x * inv, a.y * inv, a.z * inv); }
static inline float length(v3 a) { return sqrtf(dot(a, a)); }
static inline float length2(float x, float y) { return sqrtf(x * x + y * y); }

/* GetHeightfieldAt (grid.h:104-107): map[dimY * x + y].  The reference indexes unchecked (the corner
 * cases read cells at index -1 / Dim next to the border: undefined behaviour there); this restatement
 * clamps, and the parity tests stay one cell away from the border. */
static inline float H(const so_terrain* T, int x, int z) {
    if (x < 0) x = 0; if (x >= T->rows) x = T->rows - 1;
    if (z < 0) z = 0; if (z >= T->cols) z = T->cols - 1;
    return T->h[(size_t)T->cols * x + z];
}

/* Triangle ctor, grid.h:14-18 */
static inline tri mktri(v3 a, v3 b, v3 c) { tri t; t.A = a; t.B = b; t.C = c; t.n = normalize(cross(sub(b, a), sub(c, a))); return t; }

/* getCellTriangles, grid.h:195-207: which = 0 -> (C,B,A) "ABC", 1 -> (AA,B,C) "AABC" */
static tri cell_tri(const so_terrain* T, float cx, float cz, int which) {
    int ix = (int)cx, iz = (int)cz, ix1 = (int)(cx + 1), iz1 = (int)(cz + 1);
    v3 A = mk(cx, H(T, ix, iz), cz);
    v3 AA = mk(cx + 1, H(T, ix1, iz1), cz + 1);
    v3 B = mk(cx + 1, H(T, ix1, iz), cz);
    v3 C = mk(cx, H(T, ix, iz1), cz + 1);
    return which == 0 ? mktri(C, B, A) : mktri(AA, B, C);
}

/* rayIntersectsTriangle, grid.h:178-193 */
static int ray_tri(v3 pos, v3 dir, const tri* t, float* tt) {
    v3 E1 = sub(t->B, t->A), E2 = sub(t->C, t->A);
    v3 N = cross(E1, E2);
    float det = -dot(dir, N);
    float invdet = (float)(1.0 / det);
    v3 AO = sub(pos, t->A);
    v3 DAO = cross(AO, dir);
    float u = dot(E2, DAO) * invdet;
    float v = -dot(E1, DAO) * invdet;
    *tt = dot(AO, N) * invdet;
    return (fabsf(det) >= 1e-6 && *tt >= 0.0 && u >= 0.0 && v >= 0.0 && (u + v) <= 1.0);
}

/* findAdjacentCell, grid.h:210-269 (unit cells).  Returns 0 when the step leaves the grid. */
static int adjacent_cell(const so_terrain* T, v3 pos, v3 dir, float* cx, float* cz) {
    float ox = pos.x, oz = pos.z, dx = dir.x, dz = dir.z, t_x, t_z;
    if (dx < 0) t_x = (floorf(ox / 1.0f) * 1.0f - ox) / dx;
    else if (dx > 0) t_x = ((floorf(ox / 1.0f) + 1) * 1.0f - ox) / dx;
    else t_x = INFINITY;
    if (dz < 0) t_z = (floorf(oz / 1.0f) * 1.0f - oz) / dz;
    else if (dz > 0) t_z = ((floorf(oz / 1.0f) + 1) * 1.0f - oz) / dz;
    else t_z = INFINITY;
    if (t_x < t_z) { if (dx < 0) (*cx)--; else if (dx > 0) (*cx)++; }
    else { if (dz < 0) (*cz)--; else if (dz > 0) (*cz)++; }
    if (*cx < 0 || *cx >= T->dimx || *cz < 0 || *cz >= T->dimz) return 0;
    return 1;
}

/* mappedOnTriangle, grid.h:271-283 */
static int map_on(const tri* t, v3 pos, v3 dir, v3* cp, v3* n) {
    float tt;
    if (ray_tri(pos, dir, t, &tt)) { *n = dir; *cp = add(pos, scl(tt, *n)); return 1; }
    return 0;
}

/* mappedBetweenTriangles, grid.h:285-305 */
static int map_between(const tri* t1, const tri* t2, v3 pos, v3* cp, v3* n) {
    v3 d = normalize(add(t1->n, t2->n));
    float tt;
    if (ray_tri(pos, d, t1, &tt)) { *n = d; *cp = add(pos, scl(tt, *n)); return 1; }
    if (ray_tri(pos, d, t2, &tt)) { *n = d; *cp = add(pos, scl(tt, *n)); return 1; }
    return 0;
}

static float min3(float a, float b, float c) { /* grid.h:307-318 */
    if (a < b) { if (a < c) return a; else return c; }
    else if (b < c) return b;
    else return c;
}

/* cornerCaseABC / cornerCaseAABC, grid.h:320-440: the fan of triangles round the nearest corner, in
 * the reference's push order, as (dx, dz, which) offsets from the current cell; -1 terminates. */
typedef struct { signed char dx, dz, which; } fan_e;
static const fan_e FAN_ABC[3][7] = {
    /* nearest = origin  */ { {0,0,0}, {-1,0,0}, {-1,0,1}, {0,-1,0}, {0,-1,1}, {-1,-1,1}, {0,0,-1} },
    /* nearest = right   */ { {0,0,0}, {0,0,1}, {1,0,0}, {0,-1,1}, {1,-1,0}, {1,-1,1}, {0,0,-1} },
    /* nearest = down    */ { {0,0,0}, {0,0,1}, {-1,0,1}, {0,1,0}, {-1,1,0}, {-1,1,1}, {0,0,-1} } };
static const fan_e FAN_AABC[3][7] = {
    /* nearest = origin  */ { {0,0,1}, {1,0,0}, {1,0,1}, {0,1,0}, {0,1,1}, {1,1,0}, {0,0,-1} },
    /* nearest = up      */ { {0,0,0}, {0,0,1}, {1,0,0}, {0,-1,1}, {1,-1,0}, {1,-1,1}, {0,0,-1} },
    /* nearest = left    */ { {0,0,0}, {0,0,1}, {-1,0,1}, {0,1,0}, {-1,1,0}, {-1,1,1}, {0,0,-1} } };

static int corner_fan(const so_terrain* T, const fan_e* fan, float cx, float cz, v3 posNext, v3* cp, v3* n) {
    tri ts[6];
    int m = 0;
    for (; fan[m].which >= 0; m++) ts[m] = cell_tri(T, cx + fan[m].dx, cz + fan[m].dz, fan[m].which);
    v3 s = mk(0, 0, 0);
    for (int i = 0; i < m; i++) s = add(s, ts[i].n);
    s = normalize(s);
    for (int i = 0; i < m; i++) if (map_on(&ts[i], posNext, s, cp, n)) return 1;
    return 0;
}

static int corner_abc(const so_terrain* T, v3 posCurr, v3 posNext, float cx, float cz, v3* cp, v3* n) {
    float ox = floorf(posCurr.x), oz = floorf(posCurr.z);
    float dorigin = length2(ox - posCurr.x, oz - posCurr.z);
    float dright = length2((ox + 1) - posCurr.x, oz - posCurr.z);
    float ddown = length2(ox - posCurr.x, (oz + 1) - posCurr.z);
    float dmin = min3(dorigin, dright, ddown);
    int sel = (dmin == dorigin) ? 0 : (dmin == dright) ? 1 : 2;
    return corner_fan(T, FAN_ABC[sel], cx, cz, posNext, cp, n);
}

static int corner_aabc(const so_terrain* T, v3 posCurr, v3 posNext, float cx, float cz, v3* cp, v3* n) {
    float ox = floorf(posCurr.x) + 1, oz = floorf(posCurr.z) + 1;
    float dorigin = length2(ox - posCurr.x, oz - posCurr.z);
    float dleft = length2((ox - 1) - posCurr.x, oz - posCurr.z);
    float dup = length2(ox - posCurr.x, (oz - 1) - posCurr.z);
    float dmin = min3(dorigin, dup, dleft);
    int sel = (dmin == dorigin) ? 0 : (dmin == dup) ? 1 : 2;
    return corner_fan(T, FAN_AABC[sel], cx, cz, posNext, cp, n);
}

/* distance along a ray to a triangle, INFINITY if missed (grid.h:486-499, 719-730) */
static float ray_dist(v3 o, v3 d, const tri* t) {
    float tt;
    if (!ray_tri(o, d, t, &tt)) return INFINITY;
    v3 p = add(o, scl(tt, d));
    return length(sub(p, o));
}

/* first branch of the dABC / dAABC comparison, grid.h:504 and :733 */
static inline int first_wins(float d0, float d1) { return d0 < d1 || (fabs(d0 - d1) < 1.1920928955078125e-7 && d0 != INFINITY); }

/* Grid::collision, grid.h:462-805 */
int so_terrain_collision(const so_terrain* T, const float pc[3], const float pn[3], const float vn[3], float cp_out[3], float n_out[3]) {
    v3 posCurr = mk(pc[0], pc[1], pc[2]), posNext = mk(pn[0], pn[1], pn[2]);
    v3 dir = normalize(mk(vn[0], vn[1], vn[2]));
    v3 back = neg(dir);
    v3 cp = mk(cp_out[0], cp_out[1], cp_out[2]), n = mk(n_out[0], n_out[1], n_out[2]);
    float cx = floorf(posCurr.x), cz = floorf(posCurr.z), nx = floorf(posNext.x), nz = floorf(posNext.z);
    if (cx < 0 || cx >= T->dimx - 1 || cz < 0 || cz >= T->dimz - 1 || nx < 0 || nx >= T->dimx - 1 || nz < 0 || nz >= T->dimz - 1)
        return 0;
    int hit = 0;
    tri t0 = cell_tri(T, cx, cz, 0), t1 = cell_tri(T, cx, cz, 1); /* ABC, AABC of the current cell */
    if (cx == nx && cz == nz) {
        /* same cell, :476-622 */
        float d0 = ray_dist(posNext, back, &t0), d1 = ray_dist(posNext, back, &t1);
        int pick = first_wins(d0, d1) ? 0 : (d1 < d0 ? 1 : -1);
        if (pick >= 0) {
            const tri* me = pick ? &t1 : &t0;
            float tt;
            if (ray_tri(posNext, me->n, me, &tt)) { n = me->n; cp = add(posNext, scl(tt, n)); hit = 1; }
            else {
                float ax = cx, az = cz;
                if (!adjacent_cell(T, posNext, normalize(add(dir, me->n)), &ax, &az)) return 0;
                float ddx = ax - cx, ddz = az - cz;
                /* towards the hypotenuse (shared with the cell's other triangle) or towards a cathetus
                 * (shared with the adjacent cell's opposite triangle) */
                int hyp = pick == 0 ? (ddx > 0 || ddz > 0) : (ddx < 0 || ddz < 0);
                if (hyp) hit = map_between(&t0, &t1, posNext, &cp, &n);
                else { tri adj = cell_tri(T, ax, az, pick ? 0 : 1); hit = map_between(me, &adj, posNext, &cp, &n); }
                if (!hit) hit = pick ? corner_aabc(T, posCurr, posNext, cx, cz, &cp, &n) : corner_abc(T, posCurr, posNext, cx, cz, &cp, &n);
            }
        }
    } else {
        /* different cells, :623-804 */
        tri u0 = cell_tri(T, nx, nz, 0), u1 = cell_tri(T, nx, nz, 1); /* ABC, AABC of the next cell */
        float ddx = nx - cx, ddz = nz - cz, tt;
        if (ddx != 0 && ddz != 0) {
            /* diagonal move, :636-662 */
            if (ddx + ddz == 2) hit = corner_aabc(T, posCurr, posNext, cx, cz, &cp, &n);
            else hit = corner_abc(T, posCurr, posNext, cx, cz, &cp, &n);
        } else if (ray_tri(posNext, back, &t0, &tt)) {
            if (ddx == -1 || ddz == -1) hit = map_between(&t0, &u1, posNext, &cp, &n);
            if (!hit) hit = corner_abc(T, posCurr, posNext, cx, cz, &cp, &n);
        } else if (ray_tri(posNext, back, &t1, &tt)) {
            if (ddx == 1 || ddz == 1) hit = map_between(&t1, &u0, posNext, &cp, &n);
            if (!hit) hit = corner_aabc(T, posCurr, posNext, cx, cz, &cp, &n);
        } else {
            float d0 = ray_dist(posCurr, dir, &u0), d1 = ray_dist(posCurr, dir, &u1);
            if (first_wins(d0, d1)) {
                if (ray_tri(posNext, u0.n, &u0, &tt)) { n = u0.n; cp = add(posNext, scl(tt, n)); hit = 1; }
                else if (ddx == 1 || ddz == 1) {
                    hit = map_between(&u0, &t1, posNext, &cp, &n);
                    if (!hit) hit = corner_aabc(T, posCurr, posNext, cx, cz, &cp, &n);
                } else hit = corner_abc(T, posCurr, posNext, cx, cz, &cp, &n);
            } else if (d1 < d0) {
                if (ray_tri(posNext, u1.n, &u1, &tt)) { n = u1.n; cp = add(posNext, scl(tt, n)); hit = 1; }
                else if (ddx == -1 || ddz == -1) {
                    hit = map_between(&u1, &t0, posNext, &cp, &n);
                    if (!hit) hit = corner_abc(T, posCurr, posNext, cx, cz, &cp, &n);
                } else hit = corner_aabc(T, posCurr, posNext, cx, cz, &cp, &n);
            }
        }
    }
    /* the reference writes cp / norm through references even on a failed mappedOnTriangle chain only
     * when a test succeeds, so on a miss the outputs keep their input values */
    if (hit) { cp_out[0] = cp.x; cp_out[1] = cp.y; cp_out[2] = cp.z; n_out[0] = n.x; n_out[1] = n.y; n_out[2] = n.z; }
    return hit;
}

/* UpdateGrid, grid.h:138-176: per (z, x): vertex (x, y, z) with y = min(H(x,z), dimy-1) as int, normal
 * from the four neighbour differences.  out = 6 floats per vertex, z-major. */
void so_terrain_surface(const so_terrain* T, float* out) {
    size_t k = 0;
    for (int z = 0; z < T->dimz; z++)
        for (int x = 0; x < T->dimx; x++) {
            int y = (int)H(T, x, z);
            if (T->dimy <= y) y = T->dimy - 1;
            v3 u = mk(0, 0, 0), d = mk(0, 0, 0), r = mk(0, 0, 0), l = mk(0, 0, 0);
            if (x - 1 >= 0) l = mk(x - (x - 1), y - (int)H(T, x - 1, z), z - z);
            if (x + 1 < T->dimx) r = mk((x + 1) - x, (int)H(T, x + 1, z) - y, z - z);
            if (z - 1 >= 0) u = mk(x - x, y - (int)H(T, x, z - 1), z - (z - 1));
            if (z + 1 < T->dimy) d = mk(x - x, (int)H(T, x, z + 1) - y, (z + 1) - z); /* sic: compares with dim.y, grid.h:167 */
            v3 nn = normalize(add(add(add(cross(u, l), cross(u, r)), cross(d, l)), cross(d, r)));
            out[k++] = (float)x; out[k++] = (float)y; out[k++] = (float)z;
            out[k++] = nn.x; out[k++] = nn.y; out[k++] = nn.z;
        }
}

/* genIndices, grid.h:118-136.  Returns the count; out may be NULL. */
long so_terrain_indices(const so_terrain* T, unsigned* out) {
    long k = 0;
    int dx = T->dimx, dz = T->dimz;
    for (int z = 0, j = dz - 1; z < dz && j >= 0; z++, j--)
        for (int x = 0, i = dx - 1; x < dx && i >= 0; x++, i--) {
            if (x + 1 < dx && z + 1 < dz) {
                if (out) { out[k] = z * dx + x; out[k + 1] = z * dx + (x + 1); out[k + 2] = (z + 1) * dx + x; }
                k += 3;
            }
            if (j - 1 >= 0 && i - 1 >= 0) {
                if (out) { out[k] = j * dx + i; out[k + 1] = j * dx + (i - 1); out[k + 2] = (j - 1) * dx + i; }
                k += 3;
            }
        }
    return k;
}

/* ======================= terrain stage of the step: contact response + erosion =======================
 * Specification (this project's):
 *   terrain coordinates t = (p - origin) / scale  (uniform scale: angles and normals are preserved)
 *   heights are fixed point: hfx = height * 4096 (exact for the reference's 8-bit heights)
 *   per particle, after integration and BEFORE the box collision (the order of the commented call,
 *   fluid_system.h:335-347):
 *     hit = Grid::collision(t(posCurr), t(posNext), velNext / scale)
 *     if hit and dt != 0:   d = |posNext - cp|, v' = v - (1 + cR*d/(dt*|v|)) * dot(v, n) * n,  posNext = cp
 *       (world units; fluid_system.h:337-339)
 *       erosion request at vertex (round(cp.x), round(cp.z)) with the PRE-response velocity v:
 *         vt  = | v - dot(v, n) n |                tangential speed (world units / s)
 *         cap = Kc * vt                            carrying capacity (height units)
 *         s   = sediment carried (fixed point)
 *         s > cap : deposit  q = rint((s/4096 - cap) * Kd * 4096)   (q <= s)
 *         s < cap : pick-up request q = rint((cap - s/4096) * Ke * 4096)
 *   then per vertex: grants are scaled so a vertex never drops below hmin:
 *         avail = max(hfx - hmin_fx, 0);  want = sum of requests
 *         grant = want <= avail ? q : floor(q * avail / want)            (64-bit integer)
 *   finally hfx += deposits - grants.  All sums are integer: results do not depend on the order. */
static inline int rint_fx(float v) { return (int)lrintf(v * 4096.0f); }

void so_terrain_stage(so_terrain* T, const so_erosion* E, int n, const float* pos_curr, float* pos_next, float* vel_next,
                      int* sediment, float dt, float cR, int* hit_out) {
    size_t cells = (size_t)T->rows * T->cols;
    long long* want = (long long*)calloc(cells, sizeof(long long));
    long long* delta = (long long*)calloc(cells, sizeof(long long));
    int* req_cell = (int*)malloc((size_t)n * sizeof(int));
    int* req_amt = (int*)malloc((size_t)n * sizeof(int));
    int* dep_cell = (int*)malloc((size_t)n * sizeof(int));
    int* dep_amt = (int*)malloc((size_t)n * sizeof(int));
    float inv = 1.0f / E->scale;
    /* per-particle part (independent, OpenMP for the timed CPU baseline); the per-vertex sums follow serially */
#pragma omp parallel for schedule(dynamic, 1024)
    for (int i = 0; i < n; i++) {
        req_cell[i] = -1; req_amt[i] = 0; dep_cell[i] = -1; dep_amt[i] = 0;
        if (hit_out) hit_out[i] = 0;
        float pc[3], pn[3], vn[3], cp[3] = { 0, 0, 0 }, nn[3] = { 0, 0, 0 };
        for (int a = 0; a < 3; a++) {
            pc[a] = (pos_curr[3 * i + a] - E->origin[a]) * inv;
            pn[a] = (pos_next[3 * i + a] - E->origin[a]) * inv;
            vn[a] = vel_next[3 * i + a] * inv;
        }
        if (!so_terrain_collision(T, pc, pn, vn, cp, nn) || dt == 0) continue;
        if (hit_out) hit_out[i] = 1;
        v3 v = mk(vel_next[3 * i], vel_next[3 * i + 1], vel_next[3 * i + 2]);
        v3 N = mk(nn[0], nn[1], nn[2]);
        v3 cw = mk(cp[0] * E->scale + E->origin[0], cp[1] * E->scale + E->origin[1], cp[2] * E->scale + E->origin[2]);
        v3 pw = mk(pos_next[3 * i], pos_next[3 * i + 1], pos_next[3 * i + 2]);
        float d = length(sub(pw, cw));
        float vn_ = dot(v, N);
        float k = (float)(1 + cR * (d / (dt * length(v))));
        v3 vt = sub(v, scl(vn_, N));
        float vtl = length(vt);
        v3 v2 = sub(v, scl(k * vn_, N));
        vel_next[3 * i] = v2.x; vel_next[3 * i + 1] = v2.y; vel_next[3 * i + 2] = v2.z;
        pos_next[3 * i] = cw.x; pos_next[3 * i + 1] = cw.y; pos_next[3 * i + 2] = cw.z;
        if (!E->enabled) continue;
        int vx = (int)floorf(cp[0] + 0.5f), vz = (int)floorf(cp[2] + 0.5f);
        if (vx < 0) vx = 0; if (vx >= T->rows) vx = T->rows - 1;
        if (vz < 0) vz = 0; if (vz >= T->cols) vz = T->cols - 1;
        int c = vx * T->cols + vz;
        float cap = E->Kc * vtl;
        float s = (float)sediment[i] * (1.0f / 4096.0f);
        if (s > cap) {
            int q = rint_fx((s - cap) * E->Kd);
            if (q > sediment[i]) q = sediment[i];
            if (q > 0) { sediment[i] -= q; dep_cell[i] = c; dep_amt[i] = q; }
        } else if (s < cap) {
            int q = rint_fx((cap - s) * E->Ke);
            if (q > E->max_pickup_fx) q = E->max_pickup_fx;
            if (q > 0) { req_cell[i] = c; req_amt[i] = q; }
        }
    }
    for (int i = 0; i < n; i++) {
        if (dep_cell[i] >= 0) delta[dep_cell[i]] += dep_amt[i];
        if (req_cell[i] >= 0) want[req_cell[i]] += req_amt[i];
    }
    for (int i = 0; i < n; i++) {
        int c = req_cell[i];
        if (c < 0) continue;
        long long avail = (long long)T->hfx[c] - E->hmin_fx;
        if (avail < 0) avail = 0;
        long long g = want[c] <= avail ? req_amt[i] : ((long long)req_amt[i] * avail) / want[c];
        sediment[i] += (int)g;
        delta[c] -= g;
    }
    for (size_t c = 0; c < cells; c++)
        if (delta[c]) { T->hfx[c] += (int)delta[c]; T->h[c] = (float)T->hfx[c] * (1.0f / 4096.0f); }
    free(want); free(delta); free(req_cell); free(req_amt); free(dep_cell); free(dep_amt);
}
