/* TEST INFRASTRUCTURE ONLY -- see sph_oracle.h for scope and parity status.
 *
 * Arithmetic notes (what makes this bit-identical to the compiled reference):
 *  - build: gcc -O2 -ffp-contract=off, x86-64 baseline (no FMA), FLT_EVAL_METHOD 0
 *  - glm 0.9.9.7 evaluation order: dot(a,b) = (ax*bx + ay*by) + az*bz
 *    (vendor/glm/glm/detail/func_geometric.inl:47-55), length = sqrt(dot) (:8-14),
 *    normalize = v * (1/sqrt(dot(v,v))) (:82-90, func_exponential.inl:135-139)
 *  - PI is the reference's 3.141592f (Erosion/sphere.h:8), not pi
 *  - powf(x, 3.0f) / powf(h, 9.0f) / powf(h, 6.0f) are libm calls exactly as in the reference;
 *    powf(x, 2.0f) is written as powf too so the compiler applies the same x*x strength
 *    reduction it applied to the reference build.
 */
#include "sph_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define REF_PI 3.141592f /* Erosion/sphere.h:8 */

typedef struct { float x, y, z; } v3;

static inline v3 ld3(const float* a, int i) { v3 r = { a[3 * i], a[3 * i + 1], a[3 * i + 2] }; return r; }
static inline void st3(float* a, int i, v3 v) { a[3 * i] = v.x; a[3 * i + 1] = v.y; a[3 * i + 2] = v.z; }
static inline v3 mk3(float x, float y, float z) { v3 r = { x, y, z }; return r; }
static inline v3 add3(v3 a, v3 b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline v3 sub3(v3 a, v3 b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline v3 mul3s(v3 a, float s) { return mk3(a.x * s, a.y * s, a.z * s); }
static inline v3 smul3(float s, v3 a) { return mk3(s * a.x, s * a.y, s * a.z); }
static inline v3 div3s(v3 a, float s) { return mk3(a.x / s, a.y / s, a.z / s); }
static inline v3 neg3(v3 a) { return mk3(-a.x, -a.y, -a.z); }
static inline float dot3(v3 a, v3 b) { float tx = a.x * b.x, ty = a.y * b.y, tz = a.z * b.z; return tx + ty + tz; }
static inline float len3(v3 a) { return sqrtf(dot3(a, a)); }
static inline v3 normalize3(v3 a) { float inv = 1.0f / sqrtf(dot3(a, a)); return mul3s(a, inv); }

void so_default_params(so_params* p) {
    /* Erosion/fluid_system.h:460-478 */
    p->g[0] = 0.0f; p->g[1] = -9.82f; p->g[2] = 0.0f;
    p->dt = 0.0f; p->p0 = 998.29f; p->mass = 0.02f; p->visc = 3.5f;
    p->surf_tens = 0.0728f; p->k = 3.0f; p->h = 0.0457f; p->len = 0.2f;
}

int so_lattice(int n, const float origin[3], float* pos) {
    /* Erosion/fluid_system.h:80-95 (Initialize) and :234-249 (AddParticles): the loop bound is
     * the double cbrt(n), so non-cube counts round up per axis. */
    int c = 0;
    for (int i = 0; i < cbrt(n); i++)
        for (int j = 0; j < cbrt(n); j++)
            for (int k = 0; k < cbrt(n); k++) {
                float x = -0.2 + i * 0.025;
                float y = -0.05 + j * 0.025;
                float z = -0.15 + k * 0.025;
                if (pos) { pos[3 * c] = x + origin[0]; pos[3 * c + 1] = y + origin[1]; pos[3 * c + 2] = z + origin[2]; }
                c++;
            }
    return c;
}

/* ---- smoothing kernels, Erosion/fluid_system.h:410-453 ---- */
static inline float kernDefault(float h, v3 r) {
    float len = len3(r);
    if (len > h) return 0.0f;
    return (float)(315.0f / (64.0f * REF_PI * powf(h, 9.0f))) * powf((h * h - len * len), 3.0f);
}
static inline v3 gradDefault(float h, v3 r) {
    float len = len3(r);
    return neg3(mul3s(mul3s(r, (float)(945.0f / (32.0f * REF_PI * powf(h, 9.0f)))), powf((h * h - len * len), 2.0f)));
}
static inline float laplDefault(float h, v3 r) {
    float len = len3(r);
    return -(945.0f / (32.0f * REF_PI * powf(h, 9.0f))) * (h * h - len * len) * (3 * h * h - 7 * len * len);
}
static inline v3 gradPressure(float h, v3 r) {
    float dist = len3(r);
    if (dist < 10e-5) {
        return neg3(mul3s(mul3s(normalize3(mk3(1.0f, 1.0f, 1.0f)), (float)(45.f / (REF_PI * powf(h, 6.0f)))), powf(h - dist, 2.0f)));
    } else {
        return neg3(mul3s(mul3s(normalize3(r), (float)(45.f / (REF_PI * powf(h, 6.0f)))), powf(h - dist, 2.0f)));
    }
}
static inline float laplVisc(float h, v3 r) {
    float len = len3(r);
    return (float)(45.0f / (REF_PI * powf(h, 6.0f))) * (h - len);
}

/* ---- per-pair bodies shared by the all-pairs and the grid step ---- */
/* pass 1 body, fluid_system.h:118-120 */
static inline void p1_pair(const so_params* P, const float* pos, int i, int j, float* density) {
    v3 d = sub3(ld3(pos, i), ld3(pos, j));
    float distance = len3(d);
    if (distance <= P->h) *density += P->mass * kernDefault(P->h, d);
}
/* pass 2 body, fluid_system.h:139-148 */
static inline void p2_pair(const so_params* P, const so_state* S, int i, int j, v3* fPress, v3* fVisc, v3* n, int* nid) {
    v3 d = sub3(ld3(S->pos, i), ld3(S->pos, j));
    float distance = len3(d);
    if (distance <= P->h && i != j) {
        *nid = j;
        float ci = S->pressure[i] / (S->density[i] * S->density[i]);
        float cj = S->pressure[j] / (S->density[j] * S->density[j]);
        *fPress = add3(*fPress, smul3((ci + cj) * P->mass, gradPressure(P->h, d)));
        v3 dv = sub3(ld3(S->vel, j), ld3(S->vel, i));
        *fVisc = add3(*fVisc, mul3s(mul3s(dv, (P->mass / S->density[j])), laplVisc(P->h, d)));
        *n = add3(*n, smul3((P->mass / S->density[j]), gradDefault(P->h, d)));
    }
}
/* pass 3 body, fluid_system.h:169-171 */
static inline void p3_pair(const so_params* P, const so_state* S, int i, int j, float* cfl) {
    v3 d = sub3(ld3(S->pos, i), ld3(S->pos, j));
    float distance = len3(d);
    if (distance <= P->h) *cfl += (P->mass / S->density[j]) * laplDefault(P->h, d);
}

/* box collision, fluid_system.h:355-407 (abs = float overload) */
static int collisionS(float len, v3 pos, v3* contactP, v3* normal) {
    if (fabsf(pos.x) < len && fabsf(pos.y) < len && fabsf(pos.z) < len) return 0;
    float x = pos.x, y = pos.y, z = pos.z;
    char max = 'x';
    float maxP = fabsf(x);
    if (maxP < fabsf(y)) { max = 'y'; maxP = fabsf(y); }
    if (maxP < fabsf(z)) { max = 'z'; maxP = fabsf(z); }
    *contactP = pos;
    switch (max) {
    case 'x':
        if (x < -len) { contactP->x = -len; *normal = mk3(1, 0, 0); }
        else { contactP->x = len; *normal = mk3(-1, 0, 0); }
        break;
    case 'y':
        if (y < -len) { contactP->y = -len; *normal = mk3(0, 1, 0); }
        else { contactP->y = len; *normal = mk3(0, -1, 0); }
        break;
    case 'z':
        if (z < -len) { contactP->z = -len; *normal = mk3(0, 0, 1); }
        else { contactP->z = len; *normal = mk3(0, 0, -1); }
        break;
    }
    return 1;
}

/* advance(), fluid_system.h:306-353, split in two so that the terrain stage (the commented call at
 * :335-340) can run between the integration and the box collision. */
static void integrate_one(const so_params* P, so_state* S, int i, v3* posNext, v3* velNext) {
    v3 fInternal = add3(ld3(S->fpress, i), ld3(S->fvisc, i));
    v3 fExternal = add3(ld3(S->fgrav, i), ld3(S->fsurf, i));
    v3 F = add3(fInternal, fExternal);
    v3 acc = div3s(F, S->density[i]);
    st3(S->acc, i, acc);
    float deltaT = P->dt;
    *velNext = add3(ld3(S->vel, i), mul3s(acc, deltaT));
    *posNext = add3(ld3(S->pos, i), mul3s(*velNext, deltaT));
}
static void box_one(const so_params* P, v3* posNext, v3* velNext) {
    float deltaT = P->dt;
    v3 contactP = mk3(0, 0, 0), norm = mk3(0, 0, 0);
    if (collisionS(P->len, *posNext, &contactP, &norm) && deltaT != 0) {
        float d = len3(sub3(*posNext, contactP));
        *velNext = sub3(*velNext, mul3s(mul3s(norm, (float)(1 + 0.5f * d / (deltaT * len3(*velNext)))), dot3(*velNext, norm)));
        *posNext = contactP;
    }
}
static void advance_one(const so_params* P, so_state* S, int i) {
    v3 posNext, velNext;
    integrate_one(P, S, i, &posNext, &velNext);
    box_one(P, &posNext, &velNext);
    st3(S->vel, i, velNext);
    st3(S->pos, i, posNext);
}

/* the two halves for callers that insert the terrain stage (tests of sphe_step with a terrain) */
void so_integrate(const so_params* P, so_state* S, float* pos_next, float* vel_next) {
    for (int i = 0; i < S->n; i++) { v3 p, v; integrate_one(P, S, i, &p, &v); st3(pos_next, i, p); st3(vel_next, i, v); }
}
void so_box(const so_params* P, int n, float* pos_next, float* vel_next) {
    for (int i = 0; i < n; i++) { v3 p = ld3(pos_next, i), v = ld3(vel_next, i); box_one(P, &p, &v); st3(pos_next, i, p); st3(vel_next, i, v); }
}

void so_step_allpairs(const so_params* P, so_state* S) {
    int n = S->n;
    /* pass 1, fluid_system.h:108-124 */
    for (int i = 0; i < n; i++) {
        float density = 0;
        for (int j = 0; j < n; j++) p1_pair(P, S->pos, i, j, &density);
        S->density[i] = density;
        S->pressure[i] = P->k * (S->density[i] - P->p0);
    }
    /* pass 2, fluid_system.h:128-155 */
    for (int i = 0; i < n; i++) {
        v3 fPress = mk3(0, 0, 0), fVisc = mk3(0, 0, 0), nn = mk3(0, 0, 0);
        int nid = S->neighb[i];
        for (int j = 0; j < n; j++) p2_pair(P, S, i, j, &fPress, &fVisc, &nn, &nid);
        S->neighb[i] = nid;
        st3(S->fpress, i, neg3(mul3s(fPress, S->density[i])));
        st3(S->fvisc, i, mul3s(fVisc, P->visc));
        st3(S->normal, i, nn);
    }
    /* pass 3, fluid_system.h:159-178 */
    for (int i = 0; i < n; i++) {
        st3(S->fgrav, i, smul3(S->density[i], mk3(P->g[0], P->g[1], P->g[2])));
        float cfl = 0.0;
        for (int j = 0; j < n; j++) p3_pair(P, S, i, j, &cfl);
        st3(S->fsurf, i, smul3(-P->surf_tens * cfl, ld3(S->normal, i)));
    }
    for (int i = 0; i < n; i++) advance_one(P, S, i);
}

/* ================= neighbour grid: this project's definition ================= */

void so_grid_for_box(const so_params* P, float lo[3], float hi[3], so_grid* G) {
    /* cell edge = h * (1 + 2^-10): strictly larger than h so that two particles within h can
     * never be more than one cell apart after the float rounding of (p - gmin) / cell. */
    G->cell = P->h * 1.0009765625f;
    for (int a = 0; a < 3; a++) {
        G->gmin[a] = lo[a];
        float ext = hi[a] - lo[a];
        int d = (int)ceilf(ext / G->cell);
        if (d < 1) d = 1;
        G->dim[a] = d;
    }
}

static inline int cell_axis(const so_grid* G, float p, int a) {
    float v = floorf((p - G->gmin[a]) / G->cell);
    /* clamp (also maps NaN to 0); clamping is monotone so neighbours stay within +-1 cell */
    if (!(v >= 0.0f)) return 0;
    if (v >= (float)G->dim[a]) return G->dim[a] - 1;
    return (int)v;
}

int so_cell_of(const so_grid* G, const float p[3]) {
    int cx = cell_axis(G, p[0], 0), cy = cell_axis(G, p[1], 1), cz = cell_axis(G, p[2], 2);
    return (cx * G->dim[1] + cy) * G->dim[2] + cz; /* x most significant: slabs are contiguous */
}

void so_bin(const so_grid* G, int n, const float* pos, int* cell_of, int* order, int* cell_start) {
    long ncells = (long)G->dim[0] * G->dim[1] * G->dim[2];
    memset(cell_start, 0, (size_t)(ncells + 1) * sizeof(int));
    for (int i = 0; i < n; i++) { cell_of[i] = so_cell_of(G, pos + 3 * i); cell_start[cell_of[i] + 1]++; }
    for (long c = 0; c < ncells; c++) cell_start[c + 1] += cell_start[c];
    int* cur = (int*)malloc((size_t)ncells * sizeof(int));
    memcpy(cur, cell_start, (size_t)ncells * sizeof(int));
    for (int i = 0; i < n; i++) order[cur[cell_of[i]]++] = i; /* stable: (cell, id) order */
    free(cur);
}

/* exact reference predicate: glm::length(xi - xj) <= h (fluid_system.h:118-119) */
static inline int is_neighbour(const so_params* P, const float* pos, int i, int j) {
    return len3(sub3(ld3(pos, i), ld3(pos, j))) <= P->h;
}

/* visit candidates of particle i in grid-walk order: dx, dy in -1..1 (outer to inner), then the
 * contiguous sorted range covering cells cz-1..cz+1 of column (cx+dx, cy+dy). */
#define WALK_BEGIN(G, cell_start, ci)                                                       \
    {                                                                                       \
        int _cz = (ci) % (G)->dim[2], _cy = ((ci) / (G)->dim[2]) % (G)->dim[1],             \
            _cx = (ci) / ((G)->dim[2] * (G)->dim[1]);                                       \
        int _z0 = _cz > 0 ? _cz - 1 : 0, _z1 = _cz < (G)->dim[2] - 1 ? _cz + 1 : _cz;       \
        for (int _dx = -1; _dx <= 1; _dx++) {                                               \
            int _x = _cx + _dx; if (_x < 0 || _x >= (G)->dim[0]) continue;                  \
            for (int _dy = -1; _dy <= 1; _dy++) {                                           \
                int _y = _cy + _dy; if (_y < 0 || _y >= (G)->dim[1]) continue;              \
                int _base = (_x * (G)->dim[1] + _y) * (G)->dim[2];                          \
                int _s = (cell_start)[_base + _z0], _e = (cell_start)[_base + _z1 + 1];     \
                for (int _k = _s; _k < _e; _k++) {
#define WALK_END }}}}

long so_neighbours(const so_params* P, const so_grid* G, int n, const float* pos,
                   const int* order, const int* cell_start, long* nbr_start, int* nbr) {
    long total = 0;
    for (int s = 0; s < n; s++) {
        int i = order[s];
        int ci = so_cell_of(G, pos + 3 * i);
        nbr_start[s] = total;
        WALK_BEGIN(G, cell_start, ci)
            int j = order[_k];
            if (is_neighbour(P, pos, i, j)) { if (nbr) nbr[total] = j; total++; }
        WALK_END
    }
    nbr_start[n] = total;
    return total;
}

static void isort(int* a, int m) {
    for (int i = 1; i < m; i++) { int v = a[i], k = i - 1; while (k >= 0 && a[k] > v) { a[k + 1] = a[k]; k--; } a[k + 1] = v; }
}

void so_forces_grid(const so_params* P, const so_grid* G, so_state* S) {
    int n = S->n;
    long ncells = (long)G->dim[0] * G->dim[1] * G->dim[2];
    int* cell_of = (int*)malloc((size_t)n * sizeof(int));
    int* order = (int*)malloc((size_t)n * sizeof(int));
    int* cell_start = (int*)malloc((size_t)(ncells + 1) * sizeof(int));
    so_bin(G, n, S->pos, cell_of, order, cell_start);

    /* neighbour lists by id, each sorted ascending so sums run in the reference's j order */
    long* ns = (long*)malloc((size_t)(n + 1) * sizeof(long));
    int* cnt = (int*)malloc((size_t)n * sizeof(int));
#pragma omp parallel for schedule(dynamic, 256)
    for (int i = 0; i < n; i++) {
        int c = 0;
        WALK_BEGIN(G, cell_start, cell_of[i])
            if (is_neighbour(P, S->pos, i, order[_k])) c++;
        WALK_END
        cnt[i] = c;
    }
    ns[0] = 0;
    for (int i = 0; i < n; i++) ns[i + 1] = ns[i] + cnt[i];
    int* nb = (int*)malloc((size_t)(ns[n] > 0 ? ns[n] : 1) * sizeof(int));
#pragma omp parallel for schedule(dynamic, 256)
    for (int i = 0; i < n; i++) {
        int* l = nb + ns[i]; int c = 0;
        WALK_BEGIN(G, cell_start, cell_of[i])
            int j = order[_k];
            if (is_neighbour(P, S->pos, i, j)) l[c++] = j;
        WALK_END
        isort(l, c);
    }

#pragma omp parallel for schedule(dynamic, 256)
    for (int i = 0; i < n; i++) {
        float density = 0;
        for (long q = ns[i]; q < ns[i + 1]; q++) p1_pair(P, S->pos, i, nb[q], &density);
        S->density[i] = density;
        S->pressure[i] = P->k * (S->density[i] - P->p0);
    }
#pragma omp parallel for schedule(dynamic, 256)
    for (int i = 0; i < n; i++) {
        v3 fPress = mk3(0, 0, 0), fVisc = mk3(0, 0, 0), nn = mk3(0, 0, 0);
        int nid = S->neighb[i];
        for (long q = ns[i]; q < ns[i + 1]; q++) p2_pair(P, S, i, nb[q], &fPress, &fVisc, &nn, &nid);
        S->neighb[i] = nid;
        st3(S->fpress, i, neg3(mul3s(fPress, S->density[i])));
        st3(S->fvisc, i, mul3s(fVisc, P->visc));
        st3(S->normal, i, nn);
    }
#pragma omp parallel for schedule(dynamic, 256)
    for (int i = 0; i < n; i++) {
        st3(S->fgrav, i, smul3(S->density[i], mk3(P->g[0], P->g[1], P->g[2])));
        float cfl = 0.0;
        for (long q = ns[i]; q < ns[i + 1]; q++) p3_pair(P, S, i, nb[q], &cfl);
        st3(S->fsurf, i, smul3(-P->surf_tens * cfl, ld3(S->normal, i)));
    }
    free(nb); free(cnt); free(ns); free(cell_start); free(order); free(cell_of);
}

void so_step_grid(const so_params* P, const so_grid* G, so_state* S) {
    int n = S->n;
    so_forces_grid(P, G, S);
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n; i++) advance_one(P, S, i);
}

int so_omp_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
