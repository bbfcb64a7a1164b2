// C ABI (include/sphe.h) over the CUDA hot path.  Host-side state of one FluidSystemSPH replacement.
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <new>
#include <vector>

#include "../../include/sphe.h"
#include "common.cuh"
#include "sim.h"

using namespace sphe;

static thread_local char g_err[512] = "";

static int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    return code;
}

#define CU(x)                                                                                   \
    do {                                                                                        \
        cudaError_t e_ = (x);                                                                   \
        if (e_ != cudaSuccess)                                                                  \
            return fail(SPHE_ERR_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #x, cudaGetErrorString(e_)); \
    } while (0)
#define TRY(x)                 \
    do {                       \
        int r_ = (x);          \
        if (r_ != SPHE_OK) return r_; \
    } while (0)

struct KTimer {
    int kind;
    cudaEvent_t a, b;
};

struct sphe_terrain {
    int dimx, dimy, dimz;          // Grid dims (grid.h:29-32)
    int rows = 512, cols = 512;    // heightfield storage; the reference hard-codes 512 x 512 (grid.h:78-81)
    int device = -1;
    bool ready = false;
    int *hfx = nullptr, *want = nullptr, *delta = nullptr, *hmax = nullptr, *lmax = nullptr;
    size_t cells_cap = 0;
    float origin[3] = {0.f, 0.f, 0.f};
    float scale = 1.0f;
    sphe_erosion E;
    float* d_surface = nullptr;
    unsigned* d_indices = nullptr;
    long long surface_floats = 0, index_count = 0;
    long long* d_sum = nullptr;   // [0] scratch sum, [1] cumulative contact count, [2] contacts outside the maintained rows
    int win0 = 0, win1 = 0;       // rows kept current on this rank (sphe_terrain_set_window); win1 <= win0: all rows
    sphe_terrain() { E.enabled = 0; E.Kc = 0.05f; E.Ke = 0.3f; E.Kd = 0.3f; E.hmin = 0.0f; E.max_pickup = 0.25f; }
};

struct sphe_sim {
    sphe_params P;
    float origin[3] = {0, 0, 0};
    int num = 10, init_num = 0, next_label = 0;  // fluid_system.h:473-474,479
    std::vector<int> labels;                     // reference `Id` labels when they differ from the index
    bool labels_identity = true;

    int device = -1;
    bool ready = false;
    cudaStream_t st = nullptr;
    bool own_stream = false, user_stream = false;
    // end-to-end path (sphe_step_host): a second stream carries the velocity upload and the density download while
    // the main stream computes
    cudaStream_t st_io = nullptr;
    cudaEvent_t ev_vel = nullptr, ev_density = nullptr, ev_io_done = nullptr;
    struct HostIO { bool wait_vel = false; float* density_out = nullptr; } io;

    int n = 0, cap = 0;
    float4 *posA = nullptr, *posB = nullptr, *posC = nullptr, *velA = nullptr, *velB = nullptr;
    int *idsA = nullptr, *idsB = nullptr;
    float *sedA = nullptr, *sedB = nullptr, *rho = nullptr;
    uint32_t *cell = nullptr, *cell_sorted = nullptr;
    uint2* tmp = nullptr;
    float* stage = nullptr;  // 10 * cap floats: id-order staging for uploads/downloads
    int* slot_of_id = nullptr;
    int *req_vertex = nullptr, *req_amount = nullptr;  // terrain stage: pending pick-up requests per survivor
    int *surv = nullptr, *surv_count = nullptr;        // terrain stage: survivors of the contact cull
    bool slot_valid = false;
    unsigned* nmask = nullptr;     // variant 20 (default): neighbour bit masks, [warp][row][lane] (stage.cu)
    size_t nmask_words = 0;
    bool masks_valid = false;      // the masks belong to the last step (sphe_debug_pair_lists)
    bool lists_valid = false;      // the pair index lists belong to the last step
    int* nlist = nullptr;   // variant 3: [nlist_cap][pairs_pad] neighbour indices
    int2* ncount = nullptr;
    size_t nlist_pairs = 0;
    int nlist_capacity = 128;      // rows allocated per pair list in HBM: doubled (up to 512) when pairs need more
    int nlist_smem = 64;           // entries of a list staged in shared memory by the density pass: 64 / 128 / 256
    int nlist_alloc_cap = 0;
    int overflow_cooldown = 0, smem_cooldown = 0;
    bool nlist_auto = true;        // choose the shared-memory entries from the spill statistics
    int* d_overflow = nullptr;     // pairs whose list overflowed in the current step
    int* h_overflow = nullptr;     // pinned: [0] the count of a recent step (copied without a sync), [1] pairs of that step

    long long ncells = 0, ncells_cap = 0;
    int *count = nullptr, *cell_start = nullptr, *cursor = nullptr, *tile_sum = nullptr;
    // single-launch scan (decoupled look-back): status words of the tiles, the tile ticket, and the host's running counts
    unsigned long long* scan_state = nullptr;
    unsigned* scan_ticket = nullptr;
    unsigned scan_ticket_base = 0, scan_epoch = 0;
    bool scan_onepass = true;
    GridP G{};
    bool grid_user = false;
    float glo[3], ghi[3];
    bool box_user = false;   // sphe_set_box: per-axis half-extents (multi-GPU channel scenes)
    float box[3] = {0, 0, 0};
    bool slab_on = false;    // multi-GPU x-slab mode (slab.cu)
    SlabP slab{};
    int n_owned = 0;
    int* slab_counters = nullptr;  // device int[8]
    int* slab_host = nullptr;      // pinned ring of counter snapshots, SLAB_RING x 8 ints
    float4* transit[2] = {nullptr, nullptr};   // records heading further left / right than the neighbour (slab.cu k_slab_forward)
    int* transit_n = nullptr;      // device int[4]: in transit left, right, -, forwarded so far
    int slab_cap_sent = 0;
    bool slab_unpacked = false;    // unpack already ran since the last pack (a repeat must clear its counters)
    int* d_n = nullptr;            // device word: exact EXTENT of the storage arrays after the last unpack (dead entries included)
    // In-place exchange (slab.cu k_slab_classify): between an unpack and the next step the storage arrays hold dead
    // entries and the appended records: `extent` (host upper bound) / d_n (exact) entries.  The step's binning compacts
    // them; from then on the LIVE count is on the device at cell_start[ncells] (live_dev) and s->n bounds it.
    bool extent_pending = false;
    int extent = 0, extent_base = 0;
    bool live_dev = false;
    // asynchronous unpack: results of the last SLAB_RING steps, read back without stalling the stream
    static const int SLAB_RING = 8;
    cudaEvent_t slab_ev[SLAB_RING] = {};
    long long slab_seq = 0;        // tickets issued
    long long slab_done = 0;       // tickets whose result has been folded into n
    int slab_add[SLAB_RING] = {};  // upper bound of the records appended by each ticket
    int slab_max[SLAB_RING][2] = {};
    int slab_pending = 0;          // > 0: s->n is an upper bound, kernels read the exact count from d_n
    // peer-memory exchange (NVLink): this slab's mailbox and the two neighbours' mailboxes
    char* mbox = nullptr;          // cudaMalloc'd: 4 record buffers [side][parity] + 4 flags
    size_t mbox_buf = 0;           // bytes of one record buffer
    int mbox_cap = 0;              // payload records per buffer
    char* peer_mbox[2] = {nullptr, nullptr};   // left / right neighbour's mailbox, mapped into this process
    bool peer_ipc[2] = {false, false};         // opened with cudaIpcOpenMemHandle (close on destroy)
    int peer_seq = 0;              // exchanges sent so far; sequence number of the current one
    // terrain zone sums over the same mailbox: [side][which = want / delta][parity] buffers of zone_ints ints + 8 flags
    int zone_ints = 0;
    size_t zone_base = 0;          // byte offset of the zone region in the mailbox
    int zone_seq[2] = {0, 0};      // zone sums done so far, per accumulator
    int* zone_done = nullptr;      // device word: blocks of the running k_zone_sum that have stored
    long long peer_timeout = 4000000000LL;     // clock cycles a consumer waits for its flags (~2 s)
    float grid_h = -1.f, grid_len = -1.f;

    bool diag = false;
    int diag_cap = 0;
    DiagOut D{};

    bool binned = false;  // debug hooks valid
    StepC lastC{};
    // 6 / 3 = pair index lists, density pass software-prefetched (default); 20 = TMA-staged candidates in shared memory +
    // bit-mask lists (stage.cu; measured slower); 0 = thread per particle (readable baseline)
    int variant_density = 6, variant_force = 3;

    bool timing = false;
    void* flush_buf = nullptr;
    size_t flush_bytes = 0;
    std::vector<KTimer> timers;
    int launches = 0;

    sphe_sim() {
        P.mass = 0.02f; P.visc = 3.5f; P.surf_tens = 0.0728f; P.p0 = 998.29f;
        P.g[0] = 0.0f; P.g[1] = -9.82f; P.g[2] = 0.0f;
        P.dt = 0.0f; P.k = 3.0f; P.h = 0.0457f; P.len = 0.2f; P.cR = 0.5f;
    }
};

// ------------------------------------------------------------------ device plumbing
static int ensure_device(sphe_sim* s) {
    if (s->ready) {
        CU(cudaSetDevice(s->device));
        return SPHE_OK;
    }
    int cnt = 0;
    cudaError_t e = cudaGetDeviceCount(&cnt);
    if (e != cudaSuccess || cnt == 0)
        return fail(SPHE_ERR_CUDA, "no CUDA device (%s); this library has no CPU fallback", cudaGetErrorString(e));
    if (s->device < 0) CU(cudaGetDevice(&s->device));
    CU(cudaSetDevice(s->device));
    if (!s->st && !s->user_stream) {
        CU(cudaStreamCreateWithFlags(&s->st, cudaStreamNonBlocking));
        s->own_stream = true;
    }
    s->ready = true;
    return SPHE_OK;
}

template <class T>
static int grow(T** p, size_t old_count, size_t new_count, cudaStream_t st, bool keep) {
    T* q = nullptr;
    cudaError_t e = cudaMalloc(&q, new_count * sizeof(T));
    if (e != cudaSuccess) return fail(SPHE_ERR_NOMEM, "cudaMalloc(%zu) failed: %s", new_count * sizeof(T), cudaGetErrorString(e));
    if (keep && *p && old_count) CU(cudaMemcpyAsync(q, *p, old_count * sizeof(T), cudaMemcpyDeviceToDevice, st));
    if (*p) { CU(cudaStreamSynchronize(st)); CU(cudaFree(*p)); }
    *p = q;
    return SPHE_OK;
}

static int reserve(sphe_sim* s, int need) {
    if (need <= s->cap) return SPHE_OK;
    size_t nc = std::max<size_t>((size_t)need, (size_t)s->cap * 3 / 2);
    nc = (nc + 255) & ~(size_t)255;
    size_t live = (size_t)s->n;
    TRY(grow(&s->posA, live, nc, s->st, true));
    TRY(grow(&s->velA, live, nc, s->st, true));
    TRY(grow(&s->idsA, live, nc, s->st, true));
    TRY(grow(&s->sedA, live, nc, s->st, true));
    TRY(grow(&s->posB, 0, nc, s->st, false));
    TRY(grow(&s->posC, 0, 2 * nc, s->st, false));  // 2x: doubles as the interleaved record array of variant 4
    TRY(grow(&s->velB, 0, nc, s->st, false));
    TRY(grow(&s->idsB, 0, nc, s->st, false));
    TRY(grow(&s->sedB, 0, nc, s->st, false));
    TRY(grow(&s->rho, live, nc, s->st, true));
    TRY(grow(&s->cell, 0, nc, s->st, false));
    TRY(grow(&s->cell_sorted, 0, nc, s->st, false));
    TRY(grow(&s->tmp, 0, nc, s->st, false));
    TRY(grow(&s->stage, 0, nc * 10, s->st, false));
    TRY(grow(&s->slot_of_id, 0, nc, s->st, false));
    // terrain stage: survivors per contact-path class (SPHE_SURV_CLASSES lists of nc entries), requests by padded survivor index
    TRY(grow(&s->req_vertex, 0, nc + 32 * SPHE_SURV_CLASSES, s->st, false));
    TRY(grow(&s->req_amount, 0, nc + 32 * SPHE_SURV_CLASSES, s->st, false));
    TRY(grow(&s->surv, 0, nc * SPHE_SURV_CLASSES, s->st, false));
    if (!s->surv_count) CU(cudaMalloc(&s->surv_count, 4 * sizeof(int)));
    s->cap = (int)nc;
    s->binned = false;
    s->slot_valid = false;
    return SPHE_OK;
}

static int reserve_diag(sphe_sim* s) {
    if (!s->diag || s->diag_cap >= s->cap) return SPHE_OK;
    size_t nc = (size_t)s->cap, old = (size_t)std::min(s->diag_cap, s->n);
    TRY(grow(&s->D.acc, old, nc, s->st, true));
    TRY(grow(&s->D.fpress, old, nc, s->st, true));
    TRY(grow(&s->D.fvisc, old, nc, s->st, true));
    TRY(grow(&s->D.fgrav, old, nc, s->st, true));
    TRY(grow(&s->D.fsurf, old, nc, s->st, true));
    TRY(grow(&s->D.normal, old, nc, s->st, true));
    TRY(grow(&s->D.neighb, old, nc, s->st, true));
    if (s->diag_cap == 0) {
        // the reference leaves these fields uninitialised until the first Run; we define them as 0
        CU(cudaMemsetAsync(s->D.acc, 0, nc * sizeof(float4), s->st));
        CU(cudaMemsetAsync(s->D.fpress, 0, nc * sizeof(float4), s->st));
        CU(cudaMemsetAsync(s->D.fvisc, 0, nc * sizeof(float4), s->st));
        CU(cudaMemsetAsync(s->D.fgrav, 0, nc * sizeof(float4), s->st));
        CU(cudaMemsetAsync(s->D.fsurf, 0, nc * sizeof(float4), s->st));
        CU(cudaMemsetAsync(s->D.normal, 0, nc * sizeof(float4), s->st));
        CU(cudaMemsetAsync(s->D.neighb, 0, nc * sizeof(int), s->st));
    }
    s->diag_cap = (int)nc;
    return SPHE_OK;
}

// ------------------------------------------------------------------ per-step constants and grid
static float sqrt_threshold(float h) {
    // largest float T with sqrtf(T) <= h  => (sqrtf(d2) <= h) == (d2 <= T) for every float d2
    float T = h * h;
    while (sqrtf(T) > h) T = nextafterf(T, 0.0f);
    while (sqrtf(nextafterf(T, INFINITY)) <= h) T = nextafterf(T, INFINITY);
    return T;
}

static StepC make_consts(const sphe_params& P) {
    const float PI_REF = 3.141592f;  // Erosion/sphere.h:8
    StepC C;
    float h = P.h;
    C.h = h; C.hh = h * h; C.T = sqrt_threshold(h);
    C.mass = P.mass; C.k = P.k; C.p0 = P.p0; C.visc = P.visc; C.surf = P.surf_tens;
    C.gx = P.g[0]; C.gy = P.g[1]; C.gz = P.g[2];
    C.dt = P.dt; C.len = P.len; C.cR = P.cR;
    C.lenx = C.leny = C.lenz = P.len; C.cube = 1; C.box = 1;
    C.t_lmax = nullptr; C.t_surv = nullptr; C.t_count = nullptr;
    float c315 = (float)(315.0f / (64.0f * PI_REF * powf(h, 9.0f)));  // fluid_system.h:415
    C.densK = P.mass * c315;
    C.c45 = (float)(45.f / (PI_REF * powf(h, 6.0f)));                 // :442, :452
    C.c945 = (float)(945.0f / (32.0f * PI_REF * powf(h, 9.0f)));      // :421, :427
    C.hh3 = 3 * h * h;
    return C;
}

static int setup_grid(sphe_sim* s) {
    const sphe_params& P = s->P;
    if (!(P.h > 0.0f) || !isfinite(P.h)) return fail(SPHE_ERR_ARG, "smoothing radius h must be positive");
    bool same = (s->grid_h == P.h) && (s->grid_user || s->box_user || s->grid_len == P.len) && s->ncells > 0;
    if (same) return SPHE_OK;
    // slab mode keeps the live particle count of the compacted arrays at cell_start[ncells]: move it to d_n before the
    // table is rebuilt (re-windowing after a re-cut, new h) and treat the arrays as a pending extent
    if (s->slab_on && s->live_dev && !s->extent_pending && s->cell_start && s->d_n) {
        CU(cudaMemcpyAsync(s->d_n, s->cell_start + s->ncells, sizeof(int), cudaMemcpyDeviceToDevice, s->st));
        s->extent = s->n; s->extent_pending = true;
    }
    s->live_dev = false;
    float cell = P.h * 1.0009765625f;  // h * (1 + 2^-10), see oracle so_grid_for_box
    float lo[3], hi[3];
    for (int a = 0; a < 3; a++) {
        float half = s->box_user ? s->box[a] : P.len;
        if (s->grid_user) { lo[a] = s->glo[a]; hi[a] = s->ghi[a]; }
        else { lo[a] = -half - 2.0f * cell; hi[a] = half + 2.0f * cell; }
    }
    int dim[3];
    for (int a = 0; a < 3; a++) {
        float ext = hi[a] - lo[a];
        int d = (int)ceilf(ext / cell);
        if (d < 1) d = 1;
        dim[a] = d;
    }
    // a slab bins into the window [x0 - halo, x1 + halo) of the global grid
    int gnx = dim[0], xoff = 0;
    if (s->slab_on) {
        int w0 = std::max(s->slab.x0 - s->slab.halo, 0), w1 = std::min(s->slab.x1 + s->slab.halo, gnx);
        if (!s->slab.has_left) w0 = 0;
        if (!s->slab.has_right) w1 = gnx;
        if (w1 <= w0) return fail(SPHE_ERR_ARG, "slab [%d,%d) lies outside the global grid (%d columns)", s->slab.x0, s->slab.x1, gnx);
        xoff = w0; dim[0] = w1 - w0;
    }
    long long nc = (long long)dim[0] * dim[1] * dim[2];
    if (nc >= (1LL << 31) - 8) return fail(SPHE_ERR_ARG, "neighbour grid too large: %d x %d x %d cells", dim[0], dim[1], dim[2]);
    s->G.gx = lo[0]; s->G.gy = lo[1]; s->G.gz = lo[2]; s->G.cell = cell;
    s->G.nx = dim[0]; s->G.ny = dim[1]; s->G.nz = dim[2];
    s->G.gnx = gnx; s->G.xoff = xoff;
    if (nc > s->ncells_cap) {
        size_t padded = ((size_t)nc + 1 + 63) & ~(size_t)63;
        TRY(grow(&s->count, 0, padded, s->st, false));
        TRY(grow(&s->cell_start, 0, padded, s->st, false));
        TRY(grow(&s->cursor, 0, padded, s->st, false));
        TRY(grow(&s->tile_sum, 0, (size_t)scan_tiles_for(nc) + 1, s->st, false));
        TRY(grow(&s->scan_state, 0, (size_t)scan_tiles_for(nc) + 1, s->st, false));
        CU(cudaMemsetAsync(s->scan_state, 0, ((size_t)scan_tiles_for(nc) + 1) * sizeof(unsigned long long), s->st));   // epoch 0 = never valid
        if (!s->scan_ticket) { CU(cudaMalloc(&s->scan_ticket, sizeof(unsigned))); CU(cudaMemsetAsync(s->scan_ticket, 0, sizeof(unsigned), s->st)); s->scan_ticket_base = 0; }
        s->ncells_cap = nc;
    }
    CU(cudaMemsetAsync(s->count, 0, ((size_t)nc + 1) * sizeof(int), s->st));
    s->ncells = nc;
    s->grid_h = P.h; s->grid_len = P.len;
    s->binned = false;
    return SPHE_OK;
}

// ------------------------------------------------------------------ timing helpers
struct Scope {
    sphe_sim* s; int kind; KTimer t{}; bool on;
    Scope(sphe_sim* s_, int kind_, int nlaunch = 1) : s(s_), kind(kind_), on(s_->timing) {
        s->launches += nlaunch;
        if (on) { t.kind = kind; cudaEventCreate(&t.a); cudaEventCreate(&t.b); cudaEventRecord(t.a, s->st); }
    }
    ~Scope() { if (on) { cudaEventRecord(t.b, s->st); s->timers.push_back(t); } }
};

// ------------------------------------------------------------------ the step
static int terrain_ready(sphe_terrain* t);
extern "C" { static int slab_settle(sphe_sim* s); }
// Extent of the storage arrays in slab mode: host upper bound and the device word holding the exact value (NULL: the
// bound is exact).  After a step the arrays are compact: s->n bounds the live count the scan left at cell_start[ncells].
static int slab_extent_bound(const sphe_sim* s) { return s->extent_pending ? s->extent : s->n; }
static const int* slab_extent_dev(const sphe_sim* s) {
    return s->extent_pending ? s->d_n : (s->live_dev ? s->cell_start + s->ncells : nullptr);
}
static TerrainDev terrain_view(const sphe_terrain* t);

static int step_device(sphe_sim* s, sphe_terrain* t, int terrain_phases = 7) {
    TRY(ensure_device(s));
    if (t) TRY(terrain_ready(t));
    if (s->n == 0) return SPHE_OK;
    TRY(setup_grid(s));
    TRY(reserve_diag(s));
    StepC C = make_consts(s->P);
    if (s->box_user) {
        C.lenx = s->box[0]; C.leny = s->box[1]; C.lenz = s->box[2];
        C.cube = (C.lenx == C.leny && C.leny == C.lenz) ? 1 : 0;
    }
    if (s->slab_on && s->diag) return fail(SPHE_ERR_STATE, "per-particle diagnostics are indexed by local id and are not available in slab mode");
    C.t_lmax = nullptr;
    if (t) {
        // exact contact cull in the force epilogue; survivors skip the box there and get it after the contact search
        C.t_lmax = t->lmax; C.t_surv = s->surv; C.t_count = s->surv_count; C.t_cap = s->cap;
        C.t_rows = t->rows; C.t_cols = t->cols; C.t_dimx = t->dimx; C.t_dimz = t->dimz;
        C.t_ox = t->origin[0]; C.t_oy = t->origin[1]; C.t_oz = t->origin[2]; C.t_inv = 1.0f / t->scale;
        CU(cudaMemsetAsync(s->surv_count, 0, 4 * sizeof(int), s->st));
    }
    s->lastC = C;
    int n = s->n;
    // Slab mode: the storage arrays may hold dead entries + appended records (extent_pending): hash and scatter walk that
    // extent (exact count on the device), every later kernel works on the live count the scan leaves at cell_start[ncells].
    const int n_in = (s->slab_on && s->extent_pending) ? s->extent : n;
    const int* nd_in = !s->slab_on ? nullptr : (s->extent_pending ? s->d_n : (s->live_dev ? s->cell_start + s->ncells : nullptr));
    const int* nd = s->slab_on ? s->cell_start + s->ncells : nullptr;
    { Scope k(s, SPHE_K_HASH); launch_hash(s->st, n_in, nd_in, s->posA, s->slab_on ? s->idsA : nullptr, s->G, s->cell, s->count); }
    if (s->scan_onepass) {
        Scope k(s, SPHE_K_SCAN, 1);
        const int ntiles = scan_tiles_for(s->ncells);
        // epochs run 1 .. 2^30 - 1 and then start over after clearing the status words (one memset every ~10^9 steps)
        if (++s->scan_epoch >= (1u << 30)) {
            s->scan_epoch = 1;
            CU(cudaMemsetAsync(s->scan_state, 0, ((size_t)scan_tiles_for(s->ncells_cap) + 1) * sizeof(unsigned long long), s->st));
        }
        launch_scan_onepass(s->st, s->ncells, s->count, s->cell_start, s->cursor, s->scan_state, s->scan_ticket, s->scan_ticket_base, s->scan_epoch);
        s->scan_ticket_base += (unsigned)ntiles;    // wraps together with the device word
    } else {
        Scope k(s, SPHE_K_SCAN, 2); launch_scan(s->st, s->ncells, s->count, s->tile_sum, s->cell_start, s->cursor);
    }
    { Scope k(s, SPHE_K_SCATTER); launch_scatter(s->st, n_in, nd_in, s->cell, s->idsA, s->cursor, s->tmp, s->cell_sorted); }
    // sphe_step_host: the velocities are still arriving on the io stream.  Nothing before the force pass reads them, so the
    // reorder only records the permutation and they are gathered after the density pass (below), which hides their upload.
    int* const late_vel = s->io.wait_vel ? s->slot_of_id : nullptr;
    { Scope k(s, SPHE_K_REORDER);
      launch_rank_reorder(s->st, n, nd, s->tmp, s->cell_sorted, s->cell_start, s->posA, s->velA, s->sedA, s->posB, s->velB, s->sedB,
                          s->idsB, late_vel); }
    const int vd = s->variant_density, vf = s->variant_force;
    if (vf == 20 && vd != 20) return fail(SPHE_ERR_ARG, "force variant 20 reads the masks of density variant 20");
    s->masks_valid = false; s->lists_valid = false;
    if (vd == 20) {
        size_t need = stage_mask_words(s->cap);
        if (need > s->nmask_words) { TRY(grow(&s->nmask, 0, need, s->st, false)); s->nmask_words = need; }
    }
    if ((vd >= 3 && vd != 20) || (vf >= 3 && vf != 20)) {
        // the plain pair-list format is shared by 3 (plain), 6 (software-prefetched), 7/9 (quad density) and 31-33 (unroll A/B):
        // any of those may be combined; the record (4) and sub-list (52/54/58) formats need the same variant in both passes
        auto plain = [](int v) { return v == 3 || v == 6 || v == 7 || v == 9 || v == 10 || v == 11 || (v >= 31 && v <= 33); };
        if (vd != vf && !(plain(vd) && plain(vf)) && !(vf < 3 && plain(vd)) && !(vd == 20 && vf < 3))
            return fail(SPHE_ERR_ARG, "neighbour-list variants with different list formats cannot be combined");
        // capacity for every list variant: S sub-lists of slist_entries(S) entries per pair, S <= 8 -> <= 96 ints per pair
        if (!s->d_overflow) {
            CU(cudaMalloc(&s->d_overflow, 8 * sizeof(int))); CU(cudaMemsetAsync(s->d_overflow, 0, 8 * sizeof(int), s->st));
            CU(cudaMallocHost(&s->h_overflow, 8 * sizeof(int))); memset(s->h_overflow, 0, 8 * sizeof(int));
        }
        // List sizing from the counters of a recent step (copied back without a sync, so a step or two old):
        // h_overflow = {pairs beyond the allocated rows, pairs that spilled out of shared memory, pairs that would
        // spill at half the shared-memory capacity, pairs of that step}.
        //  * rows (HBM, per pair): a pair beyond them falls back to the direct walk in the force pass, ~10x the cost
        //    of a list walk and it stalls its whole warp -> double the rows as soon as 0.1 % of the pairs need it.
        //  * shared-memory entries: 64 keeps the density pass at 6 CTAs/SM; spilling re-walks the saturated runs, so
        //    when MOST pairs spill (dense scenes, 60-120 neighbours) the 128 / 256-entry instantiations win.
        const int dv = s->variant_density;
        const bool sized = dv == 3 || dv == 6 || dv == 10 || dv == 11 || (dv >= 31 && dv <= 33);
        // The host may be many steps ahead of the GPU, so the counts can be old: every block of counts carries the
        // sizing it was produced with ([3] staged entries, [4] rows) and is ignored unless that is still the current one.
        const long long pairs = std::max((s->n + 1) / 2, 1);
        const bool rows_current = s->h_overflow[4] == s->nlist_capacity;
        const bool smem_current = s->h_overflow[3] == s->nlist_smem;
        if (sized && rows_current && s->nlist_capacity < 512 && s->h_overflow[0] * 1000LL > pairs) s->nlist_capacity *= 2;
        // staged entries: 64 (32-bit entries) -> 128 (16-bit entries, same shared memory, ~10 % more instructions) once 10 %
        // of the pairs spill (a spill stalls its warp) -> 256 (32-bit, 66 KB per 64-thread CTA) once 60 % spill even then
        if (s->nlist_auto && (dv == 3 || dv == 6) && smem_current) {
            const long long spilled = s->h_overflow[1], half = s->h_overflow[2];
            const int m = s->nlist_smem;
            if (m == 64 && spilled * 10 > pairs) s->nlist_smem = 128;
            else if (m == 128 && spilled * 10 > pairs * 6) s->nlist_smem = 256;
            else if (m == 128 && half * 20 < pairs) s->nlist_smem = 64;
            else if (m == 256 && half * 10 < pairs * 3) s->nlist_smem = 128;
        }
        size_t pp = (size_t)nlist_pairs_pad(s->cap) + 128;
        const int entries = std::max(std::max(96, s->nlist_capacity), s->nlist_smem);
        if (pp > s->nlist_pairs || entries > s->nlist_alloc_cap) {
            pp = std::max(pp, s->nlist_pairs);
            TRY(grow(&s->nlist, 0, pp * (size_t)entries, s->st, false));
            TRY(grow(&s->ncount, 0, pp * 8, s->st, false));
            s->nlist_pairs = pp; s->nlist_alloc_cap = entries;
        }
    }
    if (vd == 20) {
        Scope k(s, SPHE_K_DENSITY);
        launch_density_stage(s->st, n, nd, s->posB, s->posC, s->velB, s->cell_sorted, s->cell_start, s->G, C, s->rho, s->nmask);
        s->masks_valid = true;
    } else {
      Scope k(s, SPHE_K_DENSITY);
      launch_density(s->st, s->variant_density, n, nd, s->posB, s->posC, s->velB, s->cell_sorted, s->cell_start, s->G, C, s->rho,
                     s->nlist, s->ncount, s->nlist_capacity, s->d_overflow, s->nlist_smem);
      { auto plain = [](int v) { return v == 3 || v == 4 || v == 6 || v == 7 || v == 9 || v == 10 || v == 11 || (v >= 31 && v <= 33); };
        s->lists_valid = plain(vd) && vf >= 3; }
      if (s->d_overflow) {
          CU(cudaMemcpyAsync(s->h_overflow, s->d_overflow, 5 * sizeof(int), cudaMemcpyDeviceToHost, s->st));
          CU(cudaMemsetAsync(s->d_overflow, 0, 3 * sizeof(int), s->st));
      } }
    if (late_vel) {
        CU(cudaStreamWaitEvent(s->st, s->ev_vel, 0));
        Scope k(s, SPHE_K_REORDER);
        launch_gather_vel(s->st, n, late_vel, s->velA, s->velB);
    }
    if (s->io.density_out) {
        // the densities leave for the host while the force pass runs (idsB = this step's sorted ids)
        CU(cudaEventRecord(s->ev_density, s->st));
        CU(cudaStreamWaitEvent(s->st_io, s->ev_density, 0));
        float* drho = s->stage + 6 * (size_t)s->cap;
        launch_unsort_f1(s->st_io, n, s->rho, s->idsB, drho);
        CU(cudaMemcpyAsync(s->io.density_out, drho, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost, s->st_io));
    }
    if (vf == 20) {
        Scope k(s, SPHE_K_FORCE);
        launch_force_stage(s->st, n, nd, s->posC, s->velB, s->rho, s->idsB, s->cell_sorted, s->cell_start, s->G, C, s->nmask, s->posA, s->velA,
                           s->diag ? &s->D : nullptr);
    } else {
      Scope k(s, SPHE_K_FORCE);
      launch_force(s->st, vf, n, nd, s->posC, s->velB, s->rho, s->idsB, s->cell_sorted, s->cell_start, s->G, C,
                   s->posA, s->velA, s->diag ? &s->D : nullptr, s->nlist, s->ncount); }
    if (t) {
        TerrainDev T = terrain_view(t);
        Scope k(s, SPHE_K_TERRAIN, terrain_stage_launches(C, T));
        launch_terrain_stage(s->st, s->surv, s->surv_count, s->cap, s->posB, s->posA, s->velA, (int*)s->sedB, C, T, 1, s->req_vertex, s->req_amount, nullptr,
                             terrain_phases);
    }
    std::swap(s->idsA, s->idsB);
    std::swap(s->sedA, s->sedB);
    s->binned = true;
    s->slot_valid = false;
    if (s->slab_on) { s->extent_pending = false; s->live_dev = true; }
    CU(cudaGetLastError());
    return SPHE_OK;
}

// append m lattice particles (fluid_system.h:80-95 / :234-249) to the storage arrays
static int append_lattice(sphe_sim* s, int count_arg) {
    std::vector<float> pos;
    for (int i = 0; i < cbrt(count_arg); i++)
        for (int j = 0; j < cbrt(count_arg); j++)
            for (int k = 0; k < cbrt(count_arg); k++) {
                float x = -0.2 + i * 0.025;
                float y = -0.05 + j * 0.025;
                float z = -0.15 + k * 0.025;
                pos.push_back(x + s->origin[0]);
                pos.push_back(y + s->origin[1]);
                pos.push_back(z + s->origin[2]);
            }
    int m = (int)(pos.size() / 3);
    if (m == 0) return SPHE_OK;
    TRY(ensure_device(s));
    TRY(reserve(s, s->n + m));
    std::vector<float4> p4(m), v4(m, make_float4(0, 0, 0, 0));
    std::vector<int> ids(m);
    for (int i = 0; i < m; i++) {
        p4[i] = make_float4(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2], 0.f);
        ids[i] = s->n + i;  // internal id == index in the reference's vector
    }
    CU(cudaMemcpyAsync(s->posA + s->n, p4.data(), m * sizeof(float4), cudaMemcpyHostToDevice, s->st));
    CU(cudaMemcpyAsync(s->velA + s->n, v4.data(), m * sizeof(float4), cudaMemcpyHostToDevice, s->st));
    CU(cudaMemcpyAsync(s->idsA + s->n, ids.data(), m * sizeof(int), cudaMemcpyHostToDevice, s->st));
    CU(cudaMemsetAsync(s->sedA + s->n, 0, m * sizeof(float), s->st));
    CU(cudaMemsetAsync(s->rho + s->n, 0, m * sizeof(float), s->st));
    CU(cudaStreamSynchronize(s->st));
    // reference labels: Id = id++ (may restart at 0 if Initialize is called on a non-empty system)
    if (s->next_label != s->n) s->labels_identity = false;
    if (!s->labels_identity || !s->labels.empty()) {
        if ((int)s->labels.size() < s->n) { int o = (int)s->labels.size(); s->labels.resize(s->n); for (int i = o; i < s->n; i++) s->labels[i] = i; }
        for (int i = 0; i < m; i++) s->labels.push_back(s->next_label + i);
    }
    s->next_label += m;
    s->n += m;
    s->binned = false;
    s->slot_valid = false;
    if (s->diag) TRY(reserve_diag(s));
    return SPHE_OK;
}

static int ensure_slots(sphe_sim* s) {
    if (s->slot_valid) return SPHE_OK;
    launch_slot_of_id(s->st, s->n, s->idsA, s->slot_of_id);
    s->slot_valid = true;
    return SPHE_OK;
}

// ================================================================== C ABI
extern "C" {

const char* sphe_last_error(void) { return g_err; }
int sphe_abi_version(void) { return 1; }

int sphe_create(sphe_sim** out) {
    if (!out) return fail(SPHE_ERR_ARG, "out is NULL");
    *out = new sphe_sim();
    return SPHE_OK;
}

void sphe_destroy(sphe_sim* s) {
    if (!s) return;
    if (s->ready) {
        cudaSetDevice(s->device);
        cudaStreamSynchronize(s->st);
        void* ptrs[] = {s->posA, s->posB, s->posC, s->velA, s->velB, s->idsA, s->idsB, s->sedA, s->sedB, s->rho, s->cell,
                        s->cell_sorted, s->tmp, s->stage, s->slot_of_id, s->req_vertex, s->req_amount, s->surv, s->surv_count, s->nlist, s->ncount, s->count, s->cell_start, s->cursor, s->tile_sum,
                        s->flush_buf, s->slab_counters, s->scan_state, s->scan_ticket, s->nmask, s->zone_done, s->D.acc, s->D.fpress, s->D.fvisc, s->D.fgrav, s->D.fsurf, s->D.normal, s->D.neighb};
        for (void* p : ptrs) if (p) cudaFree(p);
        for (int k = 0; k < 2; k++) if (s->peer_mbox[k] && s->peer_ipc[k]) cudaIpcCloseMemHandle(s->peer_mbox[k]);
        if (s->mbox) cudaFree(s->mbox);
        if (s->slab_host) cudaFreeHost(s->slab_host);
        cudaFree(s->transit[0]); cudaFree(s->transit[1]); cudaFree(s->transit_n);
        cudaFree(s->d_overflow); if (s->h_overflow) cudaFreeHost(s->h_overflow);
        if (s->st_io) { cudaStreamDestroy(s->st_io); cudaEventDestroy(s->ev_vel); cudaEventDestroy(s->ev_density); cudaEventDestroy(s->ev_io_done); }
        if (s->d_n) cudaFree(s->d_n);
        for (auto& e : s->slab_ev) if (e) cudaEventDestroy(e);
        if (s->own_stream && s->st) cudaStreamDestroy(s->st);
    }
    delete s;
}

int sphe_set_device(sphe_sim* s, int device) {
    if (!s) return fail(SPHE_ERR_ARG, "NULL handle");
    if (s->ready) return fail(SPHE_ERR_STATE, "device already initialised");
    s->device = device;
    return SPHE_OK;
}

int sphe_set_stream(sphe_sim* s, void* stream) {
    if (!s) return fail(SPHE_ERR_ARG, "NULL handle");
    if (s->ready && s->own_stream) { cudaStreamSynchronize(s->st); cudaStreamDestroy(s->st); }
    s->st = (cudaStream_t)stream;
    s->own_stream = false;
    s->user_stream = true;
    return SPHE_OK;
}

int sphe_initialize(sphe_sim* s, int n_parts) {
    if (!s) return fail(SPHE_ERR_ARG, "NULL handle");
    // fluid_system.h:76-78
    s->num = n_parts; s->init_num = n_parts; s->next_label = 0;
    return append_lattice(s, n_parts);
}

int sphe_add_particles(sphe_sim* s, int n) {
    if (!s) return fail(SPHE_ERR_ARG, "NULL handle");
    TRY(append_lattice(s, n));
    s->num += n;  // fluid_system.h:250
    return SPHE_OK;
}

int sphe_reset(sphe_sim* s) {
    if (!s) return fail(SPHE_ERR_ARG, "NULL handle");
    // fluid_system.h:255-258
    s->n = 0; s->num = 0; s->next_label = 0;
    s->labels.clear(); s->labels_identity = true;
    s->binned = false; s->slot_valid = false;
    TRY(append_lattice(s, s->init_num));
    s->num += s->init_num;
    return SPHE_OK;
}

int sphe_set_origin(sphe_sim* s, const float o[3]) {
    if (!s || !o) return fail(SPHE_ERR_ARG, "NULL argument");
    memcpy(s->origin, o, sizeof s->origin);
    return SPHE_OK;
}
int sphe_get_origin(sphe_sim* s, float o[3]) {
    if (!s || !o) return fail(SPHE_ERR_ARG, "NULL argument");
    memcpy(o, s->origin, sizeof s->origin);
    return SPHE_OK;
}
int sphe_set_dt(sphe_sim* s, float dt) { if (!s) return fail(SPHE_ERR_ARG, "NULL handle"); s->P.dt = dt; return SPHE_OK; }
float sphe_get_dt(sphe_sim* s) { return s ? s->P.dt : 0.0f; }
sphe_params* sphe_params_ptr(sphe_sim* s) { return s ? &s->P : nullptr; }
int sphe_count(sphe_sim* s) { return s ? s->n : 0; }
int sphe_num(sphe_sim* s) { return s ? s->num : 0; }

int sphe_set_grid_bounds(sphe_sim* s, const float lo[3], const float hi[3]) {
    if (!s) return fail(SPHE_ERR_ARG, "NULL handle");
    if (!lo || !hi) { s->grid_user = false; s->grid_h = -1.f; return SPHE_OK; }
    for (int a = 0; a < 3; a++) {
        if (!(hi[a] > lo[a])) return fail(SPHE_ERR_ARG, "grid bounds: hi must exceed lo on axis %d", a);
        s->glo[a] = lo[a]; s->ghi[a] = hi[a];
    }
    s->grid_user = true;
    s->grid_h = -1.f;  // force re-setup
    return SPHE_OK;
}

int sphe_grid_info_get(sphe_sim* s, sphe_grid_info* out) {
    if (!s || !out) return fail(SPHE_ERR_ARG, "NULL argument");
    TRY(ensure_device(s));
    TRY(setup_grid(s));
    out->gmin[0] = s->G.gx; out->gmin[1] = s->G.gy; out->gmin[2] = s->G.gz;
    out->cell = s->G.cell;
    out->dim[0] = s->G.nx; out->dim[1] = s->G.ny; out->dim[2] = s->G.nz;
    return SPHE_OK;
}

static int upload_from_host(sphe_sim* s, int n, const float* pos, const float* vel) {
    TRY(ensure_device(s));
    s->n = 0;
    TRY(reserve(s, n));
    float* dpos = s->stage;
    float* dvel = s->stage + 3 * (size_t)s->cap;
    CU(cudaMemcpyAsync(dpos, pos, 3 * (size_t)n * sizeof(float), cudaMemcpyHostToDevice, s->st));
    CU(cudaMemcpyAsync(dvel, vel, 3 * (size_t)n * sizeof(float), cudaMemcpyHostToDevice, s->st));
    launch_pack_state(s->st, n, dpos, dvel, s->posA, s->velA, s->idsA, s->sedA);
    s->n = n;
    s->binned = false; s->slot_valid = false;
    return SPHE_OK;
}

int sphe_upload_state(sphe_sim* s, int n, const float* pos, const float* vel) {
    if (!s || n < 0 || (n > 0 && (!pos || !vel))) return fail(SPHE_ERR_ARG, "bad arguments");
    TRY(upload_from_host(s, n, pos, vel));
    CU(cudaMemsetAsync(s->rho, 0, (size_t)std::max(n, 1) * sizeof(float), s->st));
    s->num = n; s->init_num = n; s->next_label = n;
    s->labels.clear(); s->labels_identity = true;
    if (s->diag) { s->diag_cap = 0; TRY(reserve_diag(s)); }
    CU(cudaStreamSynchronize(s->st));
    return SPHE_OK;
}

int sphe_step(sphe_sim* s, sphe_terrain* t) {
    if (!s) return fail(SPHE_ERR_ARG, "NULL handle");
    return step_device(s, t);
}

int sphe_sync(sphe_sim* s) {
    if (!s) return fail(SPHE_ERR_ARG, "NULL handle");
    if (!s->ready) return SPHE_OK;
    CU(cudaSetDevice(s->device));
    CU(cudaStreamSynchronize(s->st));
    return SPHE_OK;
}

int sphe_step_host(sphe_sim* s, sphe_terrain* t, int n, const float* pos_in, const float* vel_in, float* pos_out,
                   float* vel_out, float* density_out) {
    if (!s || n <= 0 || !pos_in || !vel_in || !pos_out || !vel_out) return fail(SPHE_ERR_ARG, "bad arguments");
    if (s->slab_on) return fail(SPHE_ERR_STATE, "not available in slab mode (sphe_slab_upload / sphe_slab_download)");
    TRY(ensure_device(s));
    if (!s->st_io) {
        CU(cudaStreamCreateWithFlags(&s->st_io, cudaStreamNonBlocking));
        CU(cudaEventCreateWithFlags(&s->ev_vel, cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&s->ev_density, cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&s->ev_io_done, cudaEventDisableTiming));
    }
    // PCIe is the bottleneck of this path (52 B/particle/step against a 0.37 us/particle... step): keep both DMA
    // directions busy while the kernels run.  Positions first (the binning needs only them); the velocities
    // follow on the io stream and are awaited right before the reorder gathers them; the densities go back while
    // the force pass runs.
    // The signature carries no sediment: what the particles picked up in earlier calls stays on the device, indexed by id
    // (the terrain keeps its eroded heights, so dropping the load would break  sum(heights) + sum(carried) = const).
    // A call with a different particle count is a new system and starts with empty loads.
    const bool keep_sed = t && t->E.enabled && s->n == n && s->cap >= n;
    s->n = 0;
    TRY(reserve(s, n));
    float* dpos = s->stage;
    float* dvel = s->stage + 3 * (size_t)s->cap;
    float* dsed = s->stage + 7 * (size_t)s->cap;
    CU(cudaStreamSynchronize(s->st));          // the staging area may still be read by an earlier download
    if (keep_sed) launch_unsort_f1(s->st, n, s->sedA, s->idsA, dsed);   // sorted slots -> id order, before the ids are reset
    CU(cudaMemcpyAsync(dpos, pos_in, 3 * (size_t)n * sizeof(float), cudaMemcpyHostToDevice, s->st));
    CU(cudaEventRecord(s->ev_io_done, s->st));
    launch_pack_state(s->st, n, dpos, nullptr, s->posA, s->velA, s->idsA, keep_sed ? nullptr : s->sedA);
    if (keep_sed) CU(cudaMemcpyAsync(s->sedA, dsed, (size_t)n * sizeof(float), cudaMemcpyDeviceToDevice, s->st));
    // the velocity copy starts when the position copy has finished (two concurrent copies would only share the
    // link and delay the positions the binning is waiting for)
    CU(cudaStreamWaitEvent(s->st_io, s->ev_io_done, 0));
    CU(cudaMemcpyAsync(dvel, vel_in, 3 * (size_t)n * sizeof(float), cudaMemcpyHostToDevice, s->st_io));
    launch_pack_state(s->st_io, n, nullptr, dvel, s->posA, s->velA, s->idsA, nullptr);
    CU(cudaEventRecord(s->ev_vel, s->st_io));
    s->n = n;
    s->binned = false; s->slot_valid = false;
    s->io.wait_vel = true; s->io.density_out = density_out;
    int rc = step_device(s, t);
    s->io.wait_vel = false; s->io.density_out = nullptr;
    if (rc != SPHE_OK) { cudaStreamSynchronize(s->st_io); return rc; }
    launch_unsort_f4(s->st, n, s->posA, s->idsA, dpos);
    CU(cudaMemcpyAsync(pos_out, dpos, 3 * (size_t)n * sizeof(float), cudaMemcpyDeviceToHost, s->st));
    launch_unsort_f4(s->st, n, s->velA, s->idsA, dvel);
    CU(cudaMemcpyAsync(vel_out, dvel, 3 * (size_t)n * sizeof(float), cudaMemcpyDeviceToHost, s->st));
    CU(cudaStreamSynchronize(s->st));
    CU(cudaStreamSynchronize(s->st_io));
    return SPHE_OK;
}

int sphe_set_l2_flush(sphe_sim* s, long long bytes) {
    if (!s || bytes < 0) return fail(SPHE_ERR_ARG, "bad arguments");
    TRY(ensure_device(s));
    if (s->flush_buf) { CU(cudaStreamSynchronize(s->st)); CU(cudaFree(s->flush_buf)); s->flush_buf = nullptr; }
    s->flush_bytes = (size_t)bytes;
    if (bytes > 0) {
        cudaError_t e = cudaMalloc(&s->flush_buf, (size_t)bytes);
        if (e != cudaSuccess) { s->flush_bytes = 0; return fail(SPHE_ERR_NOMEM, "flush buffer: %s", cudaGetErrorString(e)); }
    }
    return SPHE_OK;
}

int sphe_timed_steps(sphe_sim* s, sphe_terrain* t, int steps, float* ms_total, float* ms_kernels, int* launches) {
    if (!s || steps < 0) return fail(SPHE_ERR_ARG, "bad arguments");
    TRY(ensure_device(s));
    std::vector<cudaEvent_t> ev((size_t)steps * 2);
    for (auto& e : ev) CU(cudaEventCreate(&e));
    s->timing = (ms_kernels != nullptr);
    s->timers.clear();
    s->launches = 0;
    CU(cudaStreamSynchronize(s->st));
    int rc = SPHE_OK;
    for (int i = 0; i < steps && rc == SPHE_OK; i++) {
        // optional L2 flush between timed steps, outside the per-step event bracket
        if (s->flush_buf) CU(cudaMemsetAsync(s->flush_buf, i & 0xff, s->flush_bytes, s->st));
        CU(cudaEventRecord(ev[2 * i], s->st));
        rc = step_device(s, t);
        CU(cudaEventRecord(ev[2 * i + 1], s->st));
    }
    CU(cudaStreamSynchronize(s->st));
    s->timing = false;
    float ms = 0.f;
    if (rc == SPHE_OK)
        for (int i = 0; i < steps; i++) { float m = 0.f; CU(cudaEventElapsedTime(&m, ev[2 * i], ev[2 * i + 1])); ms += m; }
    if (ms_total) *ms_total = ms;
    if (ms_kernels) {
        for (int k = 0; k < SPHE_K_COUNT; k++) ms_kernels[k] = 0.f;
        for (auto& kt : s->timers) {
            float m = 0.f;
            cudaEventElapsedTime(&m, kt.a, kt.b);
            ms_kernels[kt.kind] += m;
            cudaEventDestroy(kt.a); cudaEventDestroy(kt.b);
        }
        s->timers.clear();
    }
    if (launches) *launches = s->launches;
    for (auto& e : ev) cudaEventDestroy(e);
    return rc;
}

int sphe_kernel_timing(sphe_sim* s, int on) {
    if (!s) return fail(SPHE_ERR_ARG, "NULL handle");
    TRY(ensure_device(s));
    for (auto& kt : s->timers) { cudaEventDestroy(kt.a); cudaEventDestroy(kt.b); }
    s->timers.clear();
    s->launches = 0;
    s->timing = on != 0;
    return SPHE_OK;
}

int sphe_kernel_times(sphe_sim* s, float* ms_kernels, int* launches) {
    if (!s) return fail(SPHE_ERR_ARG, "NULL handle");
    TRY(ensure_device(s));
    CU(cudaStreamSynchronize(s->st));
    if (ms_kernels) for (int k = 0; k < SPHE_K_COUNT; k++) ms_kernels[k] = 0.f;
    for (auto& kt : s->timers) {
        float m = 0.f;
        cudaEventElapsedTime(&m, kt.a, kt.b);
        if (ms_kernels) ms_kernels[kt.kind] += m;
        cudaEventDestroy(kt.a); cudaEventDestroy(kt.b);
    }
    s->timers.clear();
    if (launches) *launches = s->launches;
    s->launches = 0;
    return SPHE_OK;
}

int sphe_set_diagnostics(sphe_sim* s, int on) {
    if (!s) return fail(SPHE_ERR_ARG, "NULL handle");
    s->diag = on != 0;
    if (s->diag && s->ready) TRY(reserve_diag(s));
    return SPHE_OK;
}

int sphe_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int sphe_nlist_capacity(sphe_sim* s) { return s ? s->nlist_capacity : 0; }
int sphe_nlist_overflowed(sphe_sim* s) { return (s && s->h_overflow) ? s->h_overflow[0] : 0; }

int sphe_nlist_smem_entries(sphe_sim* s) { return s ? s->nlist_smem : 0; }

int sphe_set_nlist_capacity(sphe_sim* s, int entries) {
    if (!s) return fail(SPHE_ERR_ARG, "NULL handle");
    if (entries == 0) { s->nlist_auto = true; return SPHE_OK; }
    if (entries != 64 && entries != 128 && entries != 256) return fail(SPHE_ERR_ARG, "shared-memory list entries must be 64, 128, 256 or 0 (chosen from the spill statistics)");
    s->nlist_auto = false; s->nlist_smem = entries;
    s->nlist_capacity = std::max(s->nlist_capacity, entries);
    return SPHE_OK;
}

int sphe_set_variant(sphe_sim* s, int density_variant, int force_variant) {
    if (!s) return fail(SPHE_ERR_ARG, "NULL handle");
    if (!variant_supported(density_variant, force_variant))
        return fail(SPHE_ERR_ARG, "kernel variant (%d, %d) is not in this build (0 tpp, 3 / 6 / 10 index lists, 20 staged; the round-1 experiments need "
                                  "SPHE_WITH_EXPERIMENTS=1 python sph-erosion_b200/build.py)", density_variant, force_variant);
    s->variant_density = density_variant; s->variant_force = force_variant;
    return SPHE_OK;
}

static int copy_f4_by_id(sphe_sim* s, const float4* by_id, float* host_xyz) {
    // diagnostics are already indexed by id: strip the 4th lane on the host
    std::vector<float4> tmp((size_t)s->n);
    CU(cudaMemcpyAsync(tmp.data(), by_id, (size_t)s->n * sizeof(float4), cudaMemcpyDeviceToHost, s->st));
    CU(cudaStreamSynchronize(s->st));
    for (int i = 0; i < s->n; i++) { host_xyz[3 * i] = tmp[i].x; host_xyz[3 * i + 1] = tmp[i].y; host_xyz[3 * i + 2] = tmp[i].z; }
    return SPHE_OK;
}

// Slab handles keep GLOBAL ids (ghost copies carry SPHE_GHOST_BIT) and, while exchange results are in flight, only an upper
// bound of the particle count: the id-indexed accessors below would write outside their staging area.
static int not_in_slab_mode(sphe_sim* s, const char* what) {
    if (s->slab_on) return fail(SPHE_ERR_STATE, "%s is indexed by local particle id and is not available in slab mode (use sphe_slab_download)", what);
    return SPHE_OK;
}

int sphe_download(sphe_sim* s, int field, void* out) {
    if (!s || !out) return fail(SPHE_ERR_ARG, "NULL argument");
    TRY(not_in_slab_mode(s, "sphe_download"));
    if (s->n == 0) return SPHE_OK;
    TRY(ensure_device(s));
    int n = s->n;
    float* stage = s->stage;
    switch (field) {
    case SPHE_F_POS:
    case SPHE_F_VEL:
        launch_unsort_f4(s->st, n, field == SPHE_F_POS ? s->posA : s->velA, s->idsA, stage);
        CU(cudaMemcpyAsync(out, stage, 3 * (size_t)n * sizeof(float), cudaMemcpyDeviceToHost, s->st));
        break;
    case SPHE_F_DENSITY:
        launch_unsort_f1(s->st, n, s->rho, s->idsA, stage);
        CU(cudaMemcpyAsync(out, stage, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost, s->st));
        break;
    case SPHE_F_PRESSURE: {
        launch_unsort_f1(s->st, n, s->rho, s->idsA, stage);
        CU(cudaMemcpyAsync(out, stage, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost, s->st));
        CU(cudaStreamSynchronize(s->st));
        float* f = (float*)out;
        // Pressure = k * (Density - p0) (fluid_system.h:123) with the parameters of the last step
        for (int i = 0; i < n; i++) f[i] = s->lastC.k * (f[i] - s->lastC.p0);
        break;
    }
    case SPHE_F_SEDIMENT: {
        // carried sediment is fixed point (1/4096 height units) on the device
        launch_unsort_f1(s->st, n, s->sedA, s->idsA, stage);
        CU(cudaMemcpyAsync(out, stage, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost, s->st));
        CU(cudaStreamSynchronize(s->st));
        float* f = (float*)out;
        const int* q = (const int*)out;
        for (int i = 0; i < n; i++) f[i] = (float)q[i] * (1.0f / 4096.0f);
        break;
    }
    case SPHE_F_ID: {
        int* o = (int*)out;
        if (s->labels.empty()) for (int i = 0; i < n; i++) o[i] = i;
        else memcpy(o, s->labels.data(), (size_t)n * sizeof(int));
        return SPHE_OK;
    }
    case SPHE_F_ACC: case SPHE_F_FPRESS: case SPHE_F_FVISC: case SPHE_F_FGRAV: case SPHE_F_FSURF: case SPHE_F_NORMAL: {
        if (!s->diag || s->diag_cap < n) return fail(SPHE_ERR_STATE, "diagnostics are off: call sphe_set_diagnostics(s, 1) before stepping");
        const float4* src = field == SPHE_F_ACC ? s->D.acc : field == SPHE_F_FPRESS ? s->D.fpress : field == SPHE_F_FVISC ? s->D.fvisc
                          : field == SPHE_F_FGRAV ? s->D.fgrav : field == SPHE_F_FSURF ? s->D.fsurf : s->D.normal;
        return copy_f4_by_id(s, src, (float*)out);
    }
    case SPHE_F_NEIGHB:
        if (!s->diag || s->diag_cap < n) return fail(SPHE_ERR_STATE, "diagnostics are off");
        CU(cudaMemcpyAsync(out, s->D.neighb, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, s->st));
        break;
    default:
        return fail(SPHE_ERR_ARG, "unknown field %d", field);
    }
    CU(cudaStreamSynchronize(s->st));
    return SPHE_OK;
}

int sphe_download_positions(sphe_sim* s, float* host_xyz) { return sphe_download(s, SPHE_F_POS, host_xyz); }

// Renderer hand-off without the PCIe round trip: packed xyz positions in id order written straight into a DEVICE buffer
// the caller owns -- typically an OpenGL vertex buffer mapped with cudaGraphicsGLRegisterBuffer /
// cudaGraphicsResourceGetMappedPointer, drawn as one instanced call instead of the
// reference's one glDrawElements per particle (fluid_system.h:185-204).  Returns after the copy kernel has finished.
int sphe_write_positions_device(sphe_sim* s, void* device_xyz, long long capacity_floats) {
    if (!s || !device_xyz) return fail(SPHE_ERR_ARG, "NULL argument");
    TRY(not_in_slab_mode(s, "sphe_write_positions_device"));
    if (capacity_floats < 3LL * s->n) return fail(SPHE_ERR_ARG, "device buffer of %lld floats < 3 x %d particles", capacity_floats, s->n);
    if (s->n == 0) return SPHE_OK;
    TRY(ensure_device(s));
    cudaPointerAttributes a{};
    if (cudaPointerGetAttributes(&a, device_xyz) != cudaSuccess || (a.type != cudaMemoryTypeDevice && a.type != cudaMemoryTypeManaged)) {
        cudaGetLastError();
        return fail(SPHE_ERR_ARG, "device_xyz is not a device pointer");
    }
    launch_unsort_f4(s->st, s->n, s->posA, s->idsA, (float*)device_xyz);
    CU(cudaStreamSynchronize(s->st));
    return SPHE_OK;
}

int sphe_get_particle(sphe_sim* s, int id, sphe_particle* out) {
    if (!s || !out) return fail(SPHE_ERR_ARG, "NULL argument");
    TRY(not_in_slab_mode(s, "sphe_get_particle"));
    // the reference indexes unchecked (fluid_system.h:286-289); we report instead of reading out of bounds
    if (id < 0 || id >= s->n) return fail(SPHE_ERR_ARG, "particle id %d out of range [0,%d)", id, s->n);
    TRY(ensure_device(s));
    if (!s->diag) { s->diag = true; TRY(reserve_diag(s)); }
    TRY(ensure_slots(s));
    int slot = 0;
    CU(cudaMemcpyAsync(&slot, s->slot_of_id + id, sizeof(int), cudaMemcpyDeviceToHost, s->st));
    CU(cudaStreamSynchronize(s->st));
    float4 p, v, a, fp, fv, fg, fs, nn;
    float rho;
    int nb;
    CU(cudaMemcpyAsync(&p, s->posA + slot, sizeof p, cudaMemcpyDeviceToHost, s->st));
    CU(cudaMemcpyAsync(&v, s->velA + slot, sizeof v, cudaMemcpyDeviceToHost, s->st));
    CU(cudaMemcpyAsync(&rho, s->rho + slot, sizeof rho, cudaMemcpyDeviceToHost, s->st));
    CU(cudaMemcpyAsync(&a, s->D.acc + id, sizeof a, cudaMemcpyDeviceToHost, s->st));
    CU(cudaMemcpyAsync(&fp, s->D.fpress + id, sizeof fp, cudaMemcpyDeviceToHost, s->st));
    CU(cudaMemcpyAsync(&fv, s->D.fvisc + id, sizeof fv, cudaMemcpyDeviceToHost, s->st));
    CU(cudaMemcpyAsync(&fg, s->D.fgrav + id, sizeof fg, cudaMemcpyDeviceToHost, s->st));
    CU(cudaMemcpyAsync(&fs, s->D.fsurf + id, sizeof fs, cudaMemcpyDeviceToHost, s->st));
    CU(cudaMemcpyAsync(&nn, s->D.normal + id, sizeof nn, cudaMemcpyDeviceToHost, s->st));
    CU(cudaMemcpyAsync(&nb, s->D.neighb + id, sizeof nb, cudaMemcpyDeviceToHost, s->st));
    CU(cudaStreamSynchronize(s->st));
    out->id = s->labels.empty() ? id : s->labels[id];
    out->position[0] = p.x; out->position[1] = p.y; out->position[2] = p.z;
    out->velocity[0] = v.x; out->velocity[1] = v.y; out->velocity[2] = v.z;
    out->acceleration[0] = a.x; out->acceleration[1] = a.y; out->acceleration[2] = a.z;
    out->density = rho;
    out->pressure = s->binned ? s->lastC.k * (rho - s->lastC.p0) : 0.0f;
    out->pressure_force[0] = fp.x; out->pressure_force[1] = fp.y; out->pressure_force[2] = fp.z;
    out->viscosity_force[0] = fv.x; out->viscosity_force[1] = fv.y; out->viscosity_force[2] = fv.z;
    out->gravity_force[0] = fg.x; out->gravity_force[1] = fg.y; out->gravity_force[2] = fg.z;
    out->surface_force[0] = fs.x; out->surface_force[1] = fs.y; out->surface_force[2] = fs.z;
    out->surface_normal[0] = nn.x; out->surface_normal[1] = nn.y; out->surface_normal[2] = nn.z;
    out->neighb_id = nb;
    return SPHE_OK;
}

// ---- neighbour-grid test hooks
static int need_binned(sphe_sim* s) {
    if (!s) return fail(SPHE_ERR_ARG, "NULL handle");
    if (!s->binned) return fail(SPHE_ERR_STATE, "no binning available: call sphe_step first");
    return ensure_device(s);
}

int sphe_debug_cells(sphe_sim* s, int* cell_of_id) {
    TRY(need_binned(s));
    TRY(not_in_slab_mode(s, "sphe_debug_cells"));
    int* d = (int*)s->stage;
    launch_unsort_u32(s->st, s->n, s->cell_sorted, s->idsA, d);
    CU(cudaMemcpyAsync(cell_of_id, d, (size_t)s->n * sizeof(int), cudaMemcpyDeviceToHost, s->st));
    CU(cudaStreamSynchronize(s->st));
    return SPHE_OK;
}

int sphe_debug_sorted_order(sphe_sim* s, int* ids_sorted) {
    TRY(need_binned(s));
    CU(cudaMemcpyAsync(ids_sorted, s->idsA, (size_t)s->n * sizeof(int), cudaMemcpyDeviceToHost, s->st));
    CU(cudaStreamSynchronize(s->st));
    return SPHE_OK;
}

int sphe_debug_cell_start(sphe_sim* s, int* cell_start) {
    TRY(need_binned(s));
    CU(cudaMemcpyAsync(cell_start, s->cell_start, ((size_t)s->ncells + 1) * sizeof(int), cudaMemcpyDeviceToHost, s->st));
    CU(cudaStreamSynchronize(s->st));
    return SPHE_OK;
}

int sphe_debug_neighbours(sphe_sim* s, long long* nbr_start, int* nbr, long long cap, long long* total) {
    TRY(need_binned(s));
    TRY(not_in_slab_mode(s, "sphe_debug_neighbours"));
    int n = s->n;
    int* dcount = nullptr;
    CU(cudaMalloc(&dcount, (size_t)n * sizeof(int)));
    // posB still holds the sorted pre-integration positions of the last step
    launch_neighbour_count(s->st, n, s->posB, s->cell_sorted, s->cell_start, s->G, s->lastC, dcount);
    std::vector<int> counts((size_t)n);
    CU(cudaMemcpyAsync(counts.data(), dcount, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, s->st));
    CU(cudaStreamSynchronize(s->st));
    CU(cudaFree(dcount));
    std::vector<long long> starts((size_t)n + 1);
    starts[0] = 0;
    for (int i = 0; i < n; i++) starts[i + 1] = starts[i] + counts[i];
    if (total) *total = starts[n];
    if (nbr_start) memcpy(nbr_start, starts.data(), ((size_t)n + 1) * sizeof(long long));
    if (!nbr) return SPHE_OK;
    if (cap < starts[n]) return fail(SPHE_ERR_ARG, "neighbour buffer too small: need %lld", starts[n]);
    long long* dstart = nullptr;
    int* dn = nullptr;
    CU(cudaMalloc(&dstart, ((size_t)n + 1) * sizeof(long long)));
    CU(cudaMalloc(&dn, (size_t)std::max<long long>(starts[n], 1) * sizeof(int)));
    CU(cudaMemcpyAsync(dstart, starts.data(), ((size_t)n + 1) * sizeof(long long), cudaMemcpyHostToDevice, s->st));
    launch_neighbour_fill(s->st, n, s->posB, s->idsA, s->cell_sorted, s->cell_start, s->G, s->lastC, dstart, dn);
    CU(cudaMemcpyAsync(nbr, dn, (size_t)starts[n] * sizeof(int), cudaMemcpyDeviceToHost, s->st));
    CU(cudaStreamSynchronize(s->st));
    CU(cudaFree(dstart)); CU(cudaFree(dn));
    return SPHE_OK;
}

int sphe_debug_pair_lists(sphe_sim* s, int cap, int* counts, int* entries) {
    TRY(need_binned(s));
    TRY(not_in_slab_mode(s, "sphe_debug_pair_lists"));
    if (!counts || !entries || cap <= 0) return fail(SPHE_ERR_ARG, "bad arguments");
    if (!s->masks_valid && !s->lists_valid) return fail(SPHE_ERR_STATE, "the last step left no pair lists or masks (thread-per-particle kernels?)");
    int n = s->n;
    int *dc = nullptr, *de = nullptr;
    CU(cudaMalloc(&dc, (size_t)n * sizeof(int)));
    CU(cudaMalloc(&de, (size_t)n * cap * sizeof(int)));
    // posB / cell_sorted / cell_start still hold the sorted pre-integration state the masks were recorded against
    if (s->masks_valid) launch_stage_decode(s->st, n, s->posB, s->cell_sorted, s->cell_start, s->G, s->nmask, cap, dc, de);
    else launch_list_decode(s->st, n, s->nlist, s->ncount, cap, dc, de);
    CU(cudaMemcpyAsync(counts, dc, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, s->st));
    CU(cudaMemcpyAsync(entries, de, (size_t)n * cap * sizeof(int), cudaMemcpyDeviceToHost, s->st));
    CU(cudaStreamSynchronize(s->st));
    CU(cudaFree(dc)); CU(cudaFree(de));
    CU(cudaGetLastError());
    return SPHE_OK;
}

int sphe_set_box(sphe_sim* s, const float half[3]) {
    if (!s) return fail(SPHE_ERR_ARG, "NULL handle");
    if (!half) { s->box_user = false; s->grid_h = -1.f; return SPHE_OK; }
    for (int a = 0; a < 3; a++) {
        if (!(half[a] > 0.0f)) return fail(SPHE_ERR_ARG, "box half-extent %d must be positive", a);
        s->box[a] = half[a];
    }
    s->box_user = true;
    s->grid_h = -1.f;  // force re-setup
    return SPHE_OK;
}

// ---- multi-GPU x-slabs
int sphe_slab_configure(sphe_sim* s, int x0, int x1, int has_left, int has_right) {
    if (!s) return fail(SPHE_ERR_ARG, "NULL handle");
    if (x1 <= x0) return fail(SPHE_ERR_ARG, "slab needs x1 > x0");
    if (s->diag) return fail(SPHE_ERR_STATE, "switch diagnostics off before slab mode");
    TRY(ensure_device(s));
    s->slab.x0 = x0; s->slab.x1 = x1; s->slab.halo = 2;
    s->slab.has_left = has_left != 0; s->slab.has_right = has_right != 0;
    s->slab.wrap_left = s->slab.wrap_right = 0; s->slab.far_x0 = 0x7fffffff;
    s->slab_on = true;
    s->grid_h = -1.f;  // re-window the grid
    if (!s->slab_counters) { CU(cudaMalloc(&s->slab_counters, 16 * sizeof(int))); CU(cudaMemset(s->slab_counters, 0, 16 * sizeof(int))); }
    if (!s->d_n) CU(cudaMalloc(&s->d_n, sizeof(int)));
    if (!s->slab_host) CU(cudaMallocHost(&s->slab_host, sphe_sim::SLAB_RING * 8 * sizeof(int)));
    for (int k = 0; k < 2; k++) if (!s->transit[k]) CU(cudaMalloc(&s->transit[k], 2 * SPHE_TRANSIT_CAP * sizeof(float4)));
    // (the handle's stream is non-blocking: a memset on the legacy stream is NOT ordered before kernels on it -- synchronise)
    if (!s->transit_n) { CU(cudaMalloc(&s->transit_n, 4 * sizeof(int))); CU(cudaMemset(s->transit_n, 0, 4 * sizeof(int))); CU(cudaDeviceSynchronize()); }
    for (auto& e : s->slab_ev) if (!e) CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    return SPHE_OK;
}

int sphe_slab_ring(sphe_sim* s, int wrap_left, int wrap_right, int far_x0) {
    if (!s) return fail(SPHE_ERR_ARG, "NULL handle");
    if (!s->slab_on) return fail(SPHE_ERR_STATE, "call sphe_slab_configure first");
    if (wrap_left && wrap_right) return fail(SPHE_ERR_ARG, "a ring needs at least 3 slabs: only one link of a slab can wrap");
    if ((wrap_left && !s->slab.has_left) || (wrap_right && !s->slab.has_right))
        return fail(SPHE_ERR_ARG, "a wrap link is a link: configure the slab with a neighbour on that side");
    if (wrap_left && far_x0 < s->slab.x1 + s->slab.halo) return fail(SPHE_ERR_ARG, "far_x0 must lie beyond this slab's halo zone");
    s->slab.wrap_left = wrap_left != 0; s->slab.wrap_right = wrap_right != 0;
    s->slab.far_x0 = wrap_left ? far_x0 : 0x7fffffff;
    return SPHE_OK;
}

int sphe_slab_info(sphe_sim* s, int* gnx, int* xoff, int* n_total, int* n_owned) {
    if (!s) return fail(SPHE_ERR_ARG, "NULL handle");
    TRY(ensure_device(s));
    TRY(setup_grid(s));
    TRY(slab_settle(s));
    if (gnx) *gnx = s->G.gnx;
    if (xoff) *xoff = s->G.xoff;
    if (n_total) *n_total = s->n;
    if (n_owned) *n_owned = s->slab_on ? s->n_owned : s->n;
    return SPHE_OK;
}

int sphe_slab_transit(sphe_sim* s, int out[3]) {
    if (!s || !out) return fail(SPHE_ERR_ARG, "bad arguments");
    if (!s->slab_on) return fail(SPHE_ERR_STATE, "call sphe_slab_configure first");
    TRY(ensure_device(s));
    int h[4];
    CU(cudaMemcpyAsync(h, s->transit_n, sizeof h, cudaMemcpyDeviceToHost, s->st));
    CU(cudaStreamSynchronize(s->st));
    out[0] = h[0]; out[1] = h[1]; out[2] = h[3];
    return SPHE_OK;
}

int sphe_slab_upload(sphe_sim* s, int n, const float* pos, const float* vel, const int* ids) {
    if (!s || n < 0 || (n > 0 && (!pos || !vel || !ids))) return fail(SPHE_ERR_ARG, "bad arguments");
    if (!s->slab_on) return fail(SPHE_ERR_STATE, "call sphe_slab_configure first");
    TRY(ensure_device(s));
    s->n = 0;
    TRY(reserve(s, std::max(n, 1)));
    float* dpos = s->stage;
    float* dvel = s->stage + 3 * (size_t)s->cap;
    int* dids = (int*)(s->stage + 6 * (size_t)s->cap);
    CU(cudaMemcpyAsync(dpos, pos, 3 * (size_t)n * sizeof(float), cudaMemcpyHostToDevice, s->st));
    CU(cudaMemcpyAsync(dvel, vel, 3 * (size_t)n * sizeof(float), cudaMemcpyHostToDevice, s->st));
    CU(cudaMemcpyAsync(dids, ids, (size_t)n * sizeof(int), cudaMemcpyHostToDevice, s->st));
    launch_pack_state_ids(s->st, n, dpos, dvel, dids, s->posA, s->velA, s->idsA, s->sedA);
    CU(cudaStreamSynchronize(s->st));
    CU(cudaMemsetAsync(s->transit_n, 0, 4 * sizeof(int), s->st));
    CU(cudaStreamSynchronize(s->st));
    s->n = n; s->n_owned = n;
    s->extent_pending = false; s->live_dev = false;
    s->slab_done = s->slab_seq; s->slab_pending = 0;
    s->num = n; s->init_num = n; s->next_label = n;
    s->labels.clear(); s->labels_identity = true;
    s->binned = false; s->slot_valid = false;
    return SPHE_OK;
}

int sphe_slab_pack(sphe_sim* s, void* dev_send_left, void* dev_send_right, int cap_records, int reserve_incoming) {
    if (!s || !dev_send_left || !dev_send_right || cap_records < 0 || reserve_incoming < 0) return fail(SPHE_ERR_ARG, "bad arguments");
    if (!s->slab_on) return fail(SPHE_ERR_STATE, "call sphe_slab_configure first");
    TRY(ensure_device(s));
    TRY(setup_grid(s));
    // room for everything that can arrive, so nothing has to grow between pack and unpack
    TRY(reserve(s, std::max(slab_extent_bound(s) + reserve_incoming, 1)));
    CU(cudaMemsetAsync(s->slab_counters, 0, 15 * sizeof(int), s->st));   // [15] = sticky zone-sum error, folded into [7] by k_slab_headers
    launch_slab_classify(s->st, slab_extent_bound(s), slab_extent_dev(s), s->posA, s->velA, s->idsA, s->sedA, s->G, s->slab,
                         (float4*)dev_send_left, (float4*)dev_send_right, cap_records, s->slab_counters);
    launch_slab_forward(s->st, s->transit[0], s->transit[1], s->transit_n, s->slab.has_left ? (float4*)dev_send_left : nullptr,
                        s->slab.has_right ? (float4*)dev_send_right : nullptr, cap_records, s->slab_counters, false);
    launch_slab_headers(s->st, s->slab_counters, (float4*)dev_send_left, (float4*)dev_send_right);
    s->launches += 3;
    s->extent_base = slab_extent_bound(s);   // the appended records go behind it (sphe_slab_unpack*)
    s->binned = false; s->slot_valid = false;
    s->slab_cap_sent = cap_records;
    s->slab_unpacked = false;
    CU(cudaGetLastError());
    return SPHE_OK;
}

// Folds the result of ticket `t` (which must have completed) into the host-side bookkeeping.
static int slab_fold(sphe_sim* s, long long t, int out[6]) {
    const int slot = (int)(t % sphe_sim::SLAB_RING);
    const int* c = s->slab_host + 8 * slot;  // kept, to_left, to_right, owned(kept), owned(appended), from_left, from_right, error
    if (c[7] == 2) return fail(SPHE_ERR_NOMEM, "more than %d records in transit to a slab beyond the neighbour", SPHE_TRANSIT_CAP);
    if (c[7]) return fail(SPHE_ERR_STATE, "peer exchange: a neighbour's records did not arrive within %lld clock cycles", s->peer_timeout);
    if (c[1] > s->slab_cap_sent || c[2] > s->slab_cap_sent)
        return fail(SPHE_ERR_NOMEM, "slab send overflow: %d left / %d right records > buffer capacity %d", c[1], c[2], s->slab_cap_sent);
    int got_l = std::min(c[5], s->slab_max[slot][0]), got_r = std::min(c[6], s->slab_max[slot][1]);
    long long n = (long long)c[0] + got_l + got_r;
    if (n > s->cap) return fail(SPHE_ERR_NOMEM, "slab particle capacity %d exceeded (%lld)", s->cap, n);
    if (t + 1 > s->slab_done) {
        // exact count at ticket t + what later tickets may have added since
        long long hi = n;
        for (long long u = t + 1; u < s->slab_seq; u++) hi += s->slab_add[u % sphe_sim::SLAB_RING];
        s->n = (int)std::min<long long>(hi, s->cap);
        s->n_owned = c[3] + c[4];
        s->slab_done = t + 1;
        s->slab_pending = (int)(s->slab_seq - s->slab_done);
    }
    if (out) { out[0] = (int)n; out[1] = c[3] + c[4]; out[2] = c[1]; out[3] = c[2]; out[4] = c[5]; out[5] = c[6]; }
    return SPHE_OK;
}

// Blocks until every issued ticket has completed and s->n is exact again.
static int slab_settle(sphe_sim* s) {
    if (!s->slab_on || s->slab_seq == s->slab_done) return SPHE_OK;
    long long t = s->slab_seq - 1;
    CU(cudaEventSynchronize(s->slab_ev[t % sphe_sim::SLAB_RING]));
    return slab_fold(s, t, nullptr);
}

int sphe_slab_unpack_async(sphe_sim* s, const void* dev_recv_left, int max_left, const void* dev_recv_right, int max_right,
                           long long* ticket) {
    if (!s || max_left < 0 || max_right < 0) return fail(SPHE_ERR_ARG, "bad arguments");
    if (!s->slab_on) return fail(SPHE_ERR_STATE, "call sphe_slab_configure first");
    TRY(ensure_device(s));
    if (!dev_recv_left) max_left = 0;
    if (!dev_recv_right) max_right = 0;
    // never let the ring wrap over a result that has not been folded yet
    if (s->slab_seq - s->slab_done >= sphe_sim::SLAB_RING - 1) {
        long long t = s->slab_seq - 2;
        CU(cudaEventSynchronize(s->slab_ev[t % sphe_sim::SLAB_RING]));
        TRY(slab_fold(s, t, nullptr));
    }
    const long long t = s->slab_seq;
    const int slot = (int)(t % sphe_sim::SLAB_RING);
    if (s->slab_unpacked) {  // repeated after a re-send
        CU(cudaMemsetAsync(s->slab_counters + 4, 0, 3 * sizeof(int), s->st));
        CU(cudaMemsetAsync(s->transit_n, 0, 2 * sizeof(int), s->st));
    }
    s->slab_unpacked = true;
    launch_slab_append(s->st, max_left, max_right, (const float4*)dev_recv_left, (const float4*)dev_recv_right, s->G, s->slab,
                       s->cap, s->posA, s->velA, s->idsA, s->sedA, s->slab_counters, s->d_n, s->transit[0], s->transit[1], s->transit_n);
    s->launches += 1;
    CU(cudaMemcpyAsync(s->slab_host + 8 * slot, s->slab_counters, 8 * sizeof(int), cudaMemcpyDeviceToHost, s->st));
    CU(cudaEventRecord(s->slab_ev[slot], s->st));
    s->slab_add[slot] = max_left + max_right;
    s->slab_max[slot][0] = max_left; s->slab_max[slot][1] = max_right;
    s->slab_seq = t + 1;
    s->slab_pending = (int)(s->slab_seq - s->slab_done);
    // until the result is folded, n is an upper bound (pack made room for it) and kernels read the device counts
    s->n = (int)std::min<long long>((long long)s->n + max_left + max_right, s->cap);
    s->extent = (int)std::min<long long>((long long)s->extent_base + max_left + max_right, s->cap);
    s->extent_pending = true;
    s->binned = false; s->slot_valid = false;
    if (ticket) *ticket = t;
    CU(cudaGetLastError());
    return SPHE_OK;
}

int sphe_slab_result(sphe_sim* s, long long ticket, int wait, int out[6]) {
    if (!s || ticket < 0 || ticket >= s->slab_seq) return fail(SPHE_ERR_ARG, "unknown ticket");
    if (s->slab_seq - ticket > sphe_sim::SLAB_RING) return fail(SPHE_ERR_STATE, "ticket %lld is too old", ticket);
    TRY(ensure_device(s));
    cudaEvent_t ev = s->slab_ev[ticket % sphe_sim::SLAB_RING];
    if (wait) CU(cudaEventSynchronize(ev));
    else {
        cudaError_t e = cudaEventQuery(ev);
        if (e == cudaErrorNotReady) return 1;  // not an error: result not available yet
        CU(e);
    }
    return slab_fold(s, ticket, out);
}

int sphe_slab_unpack(sphe_sim* s, const void* dev_recv_left, int max_left, const void* dev_recv_right, int max_right, int out[6]) {
    long long t = 0;
    TRY(sphe_slab_unpack_async(s, dev_recv_left, max_left, dev_recv_right, max_right, &t));
    return sphe_slab_result(s, t, 1, out);   // the synchronous form: one host sync per step
}

int sphe_slab_column_histogram(sphe_sim* s, int gnx, int* hist) {
    if (!s || !hist || gnx <= 0) return fail(SPHE_ERR_ARG, "bad arguments");
    if (!s->slab_on) return fail(SPHE_ERR_STATE, "call sphe_slab_configure first");
    TRY(ensure_device(s));
    TRY(setup_grid(s));
    if (gnx != s->G.gnx) return fail(SPHE_ERR_ARG, "the global grid has %d columns, not %d", s->G.gnx, gnx);
    TRY(slab_settle(s));
    memset(hist, 0, (size_t)gnx * sizeof(int));
    const int n = slab_extent_bound(s);
    if (n == 0) return SPHE_OK;
    int* d = nullptr;
    CU(cudaMalloc(&d, (size_t)gnx * sizeof(int)));
    CU(cudaMemsetAsync(d, 0, (size_t)gnx * sizeof(int), s->st));
    if (launch_slab_column_hist(s->st, n, slab_extent_dev(s), s->posA, s->idsA, s->G, d) != 0) { cudaFree(d); return fail(SPHE_ERR_ARG, "too many cell columns (%d) for the histogram kernel", gnx); }
    CU(cudaMemcpyAsync(hist, d, (size_t)gnx * sizeof(int), cudaMemcpyDeviceToHost, s->st));
    CU(cudaStreamSynchronize(s->st));
    CU(cudaFree(d));
    CU(cudaGetLastError());
    return SPHE_OK;
}

int sphe_slab_download(sphe_sim* s, int cap, int* ids, float* pos, float* vel, float* rho, float* sed, int* n_out) {
    if (!s || !ids || !pos || !vel || !n_out) return fail(SPHE_ERR_ARG, "bad arguments");
    TRY(ensure_device(s));
    TRY(slab_settle(s));
    int n = slab_extent_bound(s);
    *n_out = 0;
    if (n == 0) return SPHE_OK;
    if (!s->slab_counters) CU(cudaMalloc(&s->slab_counters, 16 * sizeof(int)));
    size_t cp = (size_t)s->cap;
    int* d_cnt = s->slab_counters + 7;
    int* d_ids = (int*)s->stage;
    float *d_pos = s->stage + cp, *d_vel = s->stage + 4 * cp, *d_rho = s->stage + 7 * cp, *d_sed = s->stage + 8 * cp;
    CU(cudaMemsetAsync(d_cnt, 0, sizeof(int), s->st));
    launch_slab_gather_owned(s->st, n, slab_extent_dev(s), s->posA, s->velA, s->rho, s->sedA, s->idsA, d_cnt, d_ids, d_pos, d_vel, d_rho, d_sed);
    int m = 0;
    CU(cudaMemcpyAsync(&m, d_cnt, sizeof(int), cudaMemcpyDeviceToHost, s->st));
    CU(cudaStreamSynchronize(s->st));
    int rc = SPHE_OK;
    if (m > cap) rc = fail(SPHE_ERR_ARG, "output capacity %d < %d owned particles", cap, m);
    else {
        cudaMemcpyAsync(ids, d_ids, (size_t)m * sizeof(int), cudaMemcpyDeviceToHost, s->st);
        cudaMemcpyAsync(pos, d_pos, 3 * (size_t)m * sizeof(float), cudaMemcpyDeviceToHost, s->st);
        cudaMemcpyAsync(vel, d_vel, 3 * (size_t)m * sizeof(float), cudaMemcpyDeviceToHost, s->st);
        if (rho) cudaMemcpyAsync(rho, d_rho, (size_t)m * sizeof(float), cudaMemcpyDeviceToHost, s->st);
        if (sed) cudaMemcpyAsync(sed, d_sed, (size_t)m * sizeof(float), cudaMemcpyDeviceToHost, s->st);
        CU(cudaStreamSynchronize(s->st));
        *n_out = m;
    }
    return rc;
}


// ---- peer-memory exchange: records are stored straight into the neighbour's mailbox over NVLink
// Mailbox layout (one cudaMalloc, exportable as ONE cudaIpcMemHandle):
//   buffer(side, parity) at (2*side + parity) * mbox_buf, side 0 = records coming from the LEFT neighbour,
//   side 1 = from the RIGHT; each buffer = header record + mbox_cap payload records of 32 bytes;
//   flag(side, parity)   at 4 * mbox_buf + (2*side + parity) * 128: sequence number of the exchange whose
//   payload is complete in that buffer.
// Exchange q uses parity q & 1.  A producer may only overwrite buffer parity p at exchange q + 2 after it has
// itself consumed exchange q + 1 of the same neighbour, which that neighbour published after finishing its
// own append of exchange q (stream order) -- so two buffers per direction are enough and nobody ever waits
// for a consumer.
static inline float4* mbox_buffer(char* base, size_t buf, int side, int parity) { return (float4*)(base + (size_t)(2 * side + parity) * buf); }
static inline int* mbox_flag(char* base, size_t buf, int side, int parity) { return (int*)(base + 4 * buf + (size_t)(2 * side + parity) * 128); }

static inline int* zone_buffer(char* base, size_t zone_base, int zone_ints, int side, int which, int parity) {
    return (int*)(base + zone_base + (size_t)((side * 2 + which) * 2 + parity) * (((size_t)zone_ints * 4 + 255) & ~(size_t)255));
}
static inline int* zone_flag(char* base, size_t zone_base, int zone_ints, int side, int which, int parity) {
    return (int*)(base + zone_base + 8 * (((size_t)zone_ints * 4 + 255) & ~(size_t)255) + (size_t)((side * 2 + which) * 2 + parity) * 128);
}

int sphe_slab_peer_setup(sphe_sim* s, int cap_records, int reserve_particles) { return sphe_slab_peer_setup_zones(s, cap_records, reserve_particles, 0); }

int sphe_slab_peer_setup_zones(sphe_sim* s, int cap_records, int reserve_particles, int zone_ints) {
    if (!s || cap_records < 1 || zone_ints < 0) return fail(SPHE_ERR_ARG, "bad arguments");
    if (!s->slab_on) return fail(SPHE_ERR_STATE, "call sphe_slab_configure first");
    TRY(ensure_device(s));
    CU(cudaStreamSynchronize(s->st));
    for (int k = 0; k < 2; k++) {
        if (s->peer_mbox[k] && s->peer_ipc[k]) cudaIpcCloseMemHandle(s->peer_mbox[k]);
        s->peer_mbox[k] = nullptr; s->peer_ipc[k] = false;
    }
    if (s->mbox) { CU(cudaFree(s->mbox)); s->mbox = nullptr; }
    s->mbox_cap = cap_records;
    s->mbox_buf = (((size_t)cap_records + 1) * 32 + 255) & ~(size_t)255;
    size_t bytes = 4 * s->mbox_buf + 4 * 128;
    s->zone_ints = zone_ints;
    s->zone_base = (bytes + 255) & ~(size_t)255;
    s->zone_seq[0] = s->zone_seq[1] = 0;
    if (zone_ints > 0) bytes = s->zone_base + 8 * ((((size_t)zone_ints * 4 + 255) & ~(size_t)255) + 128);
    if (!s->zone_done) { CU(cudaMalloc(&s->zone_done, sizeof(int))); CU(cudaMemset(s->zone_done, 0, sizeof(int))); }
    cudaError_t e = cudaMalloc(&s->mbox, bytes);
    if (e != cudaSuccess) return fail(SPHE_ERR_NOMEM, "mailbox of %zu bytes: %s", bytes, cudaGetErrorString(e));
    CU(cudaMemset(s->mbox, 0, bytes));
    CU(cudaDeviceSynchronize());   // flags and headers must be zero before ANY stream (ours or a neighbour's) touches the mailbox
    s->peer_seq = 0;
    if (reserve_particles > 0) TRY(reserve(s, reserve_particles));
    return SPHE_OK;
}

int sphe_slab_peer_handle(sphe_sim* s, void* handle64) {
    if (!s || !handle64) return fail(SPHE_ERR_ARG, "bad arguments");
    if (!s->mbox) return fail(SPHE_ERR_STATE, "call sphe_slab_peer_setup first");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    TRY(ensure_device(s));
    cudaIpcMemHandle_t h;
    CU(cudaIpcGetMemHandle(&h, s->mbox));
    memcpy(handle64, &h, sizeof h);
    return SPHE_OK;
}

// left / right: the 64-byte handles of the neighbours' mailboxes (another process, any GPU of the node), NULL
// where there is no neighbour.  Their mailboxes must have been set up with the same cap_records.
int sphe_slab_peer_connect(sphe_sim* s, const void* left_handle64, const void* right_handle64) {
    if (!s) return fail(SPHE_ERR_ARG, "NULL handle");
    if (!s->mbox) return fail(SPHE_ERR_STATE, "call sphe_slab_peer_setup first");
    TRY(ensure_device(s));
    const void* hs[2] = {left_handle64, right_handle64};
    for (int k = 0; k < 2; k++) {
        if (s->peer_mbox[k] && s->peer_ipc[k]) cudaIpcCloseMemHandle(s->peer_mbox[k]);
        s->peer_mbox[k] = nullptr; s->peer_ipc[k] = false;
        if (!hs[k]) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, hs[k], sizeof h);
        void* p = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            cudaGetLastError();
            return fail(SPHE_ERR_CUDA, "cudaIpcOpenMemHandle(%s neighbour): %s", k ? "right" : "left", cudaGetErrorString(e));
        }
        s->peer_mbox[k] = (char*)p; s->peer_ipc[k] = true;
    }
    return SPHE_OK;
}

// Same-process form (several slabs driven by one process on one GPU: tests, single-GPU domain splitting).
int sphe_slab_peer_connect_local(sphe_sim* s, sphe_sim* left, sphe_sim* right) {
    if (!s) return fail(SPHE_ERR_ARG, "NULL handle");
    if (!s->mbox) return fail(SPHE_ERR_STATE, "call sphe_slab_peer_setup first");
    sphe_sim* ns[2] = {left, right};
    for (int k = 0; k < 2; k++) {
        if (s->peer_mbox[k] && s->peer_ipc[k]) cudaIpcCloseMemHandle(s->peer_mbox[k]);
        s->peer_mbox[k] = nullptr; s->peer_ipc[k] = false;
        if (!ns[k]) continue;
        if (!ns[k]->mbox || ns[k]->mbox_cap != s->mbox_cap) return fail(SPHE_ERR_STATE, "neighbour mailbox missing or of a different capacity");
        if (ns[k]->device != s->device) {
            // another GPU of the same process: direct peer access over NVLink instead of an IPC mapping
            TRY(ensure_device(s));
            int can = 0;
            CU(cudaDeviceCanAccessPeer(&can, s->device, ns[k]->device));
            if (!can) return fail(SPHE_ERR_CUDA, "device %d cannot access device %d as a peer", s->device, ns[k]->device);
            cudaError_t e = cudaDeviceEnablePeerAccess(ns[k]->device, 0);
            if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
            else if (e != cudaSuccess) return fail(SPHE_ERR_CUDA, "cudaDeviceEnablePeerAccess(%d): %s", ns[k]->device, cudaGetErrorString(e));
        }
        s->peer_mbox[k] = ns[k]->mbox;
    }
    return SPHE_OK;
}

int sphe_slab_peer_timeout(sphe_sim* s, long long clock_cycles) {
    if (!s || clock_cycles < 1) return fail(SPHE_ERR_ARG, "bad arguments");
    s->peer_timeout = clock_cycles;
    return SPHE_OK;
}

// folds every finished ticket without blocking, so the launch bound s->n stays close to the exact count
static int slab_fold_ready(sphe_sim* s) {
    while (s->slab_done < s->slab_seq) {
        long long t = s->slab_done;
        cudaError_t e = cudaEventQuery(s->slab_ev[t % sphe_sim::SLAB_RING]);
        if (e == cudaErrorNotReady) break;
        CU(e);
        TRY(slab_fold(s, t, nullptr));
    }
    return SPHE_OK;
}

// Exchange, producer half: drop ghosts, compact, and store migrants + halo straight into the neighbours'
// mailboxes (k_slab_classify<true>), then publish counts and flags.  Never waits for anybody.
int sphe_slab_send(sphe_sim* s) {
    if (!s) return fail(SPHE_ERR_ARG, "NULL handle");
    if (!s->slab_on || !s->mbox) return fail(SPHE_ERR_STATE, "call sphe_slab_configure and sphe_slab_peer_setup first");
    if ((s->slab.has_left && !s->peer_mbox[0]) || (s->slab.has_right && !s->peer_mbox[1]))
        return fail(SPHE_ERR_STATE, "neighbour mailbox not connected");
    TRY(ensure_device(s));
    TRY(setup_grid(s));            // the classification needs the global grid (a host that never asked for slab_info has none yet)
    TRY(slab_fold_ready(s));
    const int incoming = 2 * s->mbox_cap;
    if ((long long)slab_extent_bound(s) + incoming > s->cap) TRY(slab_settle(s));   // exact count before deciding to grow
    TRY(reserve(s, std::max(slab_extent_bound(s) + incoming, 1)));
    const int q = ++s->peer_seq, par = q & 1;
    // what goes LEFT lands in the left neighbour's "from the right" buffer, and vice versa
    float4* dl = s->slab.has_left ? mbox_buffer(s->peer_mbox[0], s->mbox_buf, 1, par) : nullptr;
    float4* dr = s->slab.has_right ? mbox_buffer(s->peer_mbox[1], s->mbox_buf, 0, par) : nullptr;
    int* fl = s->slab.has_left ? mbox_flag(s->peer_mbox[0], s->mbox_buf, 1, par) : nullptr;
    int* fr = s->slab.has_right ? mbox_flag(s->peer_mbox[1], s->mbox_buf, 0, par) : nullptr;
    CU(cudaMemsetAsync(s->slab_counters, 0, 15 * sizeof(int), s->st));   // [15] = sticky zone-sum error, folded into [7] by k_slab_headers
    launch_slab_classify(s->st, slab_extent_bound(s), slab_extent_dev(s), s->posA, s->velA, s->idsA, s->sedA, s->G, s->slab,
                         dl, dr, s->mbox_cap, s->slab_counters, true);
    launch_slab_forward(s->st, s->transit[0], s->transit[1], s->transit_n, dl, dr, s->mbox_cap, s->slab_counters, true);
    launch_slab_headers(s->st, s->slab_counters, dl, dr, fl, fr, q);
    s->launches += 3;
    s->extent_base = slab_extent_bound(s);
    s->binned = false; s->slot_valid = false;
    s->slab_cap_sent = s->mbox_cap;
    s->slab_unpacked = false;
    CU(cudaGetLastError());
    return SPHE_OK;
}

// Exchange, consumer half: k_slab_append<true> waits on the device for this exchange's flags and appends the
// mailbox payload.  No host sync; the counts come back later through the ticket (sphe_slab_result).
int sphe_slab_recv(sphe_sim* s, long long* ticket) {
    if (!s) return fail(SPHE_ERR_ARG, "NULL handle");
    if (!s->slab_on || !s->mbox || s->peer_seq < 1) return fail(SPHE_ERR_STATE, "sphe_slab_send has not run");
    if (s->slab_unpacked) return fail(SPHE_ERR_STATE, "sphe_slab_recv already ran for this exchange");
    TRY(ensure_device(s));
    if (s->slab_seq - s->slab_done >= sphe_sim::SLAB_RING - 1) {
        long long t = s->slab_seq - 2;
        CU(cudaEventSynchronize(s->slab_ev[t % sphe_sim::SLAB_RING]));
        TRY(slab_fold(s, t, nullptr));
    }
    const int q = s->peer_seq, par = q & 1;
    const int ml = s->slab.has_left ? s->mbox_cap : 0, mr = s->slab.has_right ? s->mbox_cap : 0;
    const long long t = s->slab_seq;
    const int slot = (int)(t % sphe_sim::SLAB_RING);
    s->slab_unpacked = true;
    launch_slab_append(s->st, ml, mr, s->slab.has_left ? mbox_buffer(s->mbox, s->mbox_buf, 0, par) : nullptr,
                       s->slab.has_right ? mbox_buffer(s->mbox, s->mbox_buf, 1, par) : nullptr, s->G, s->slab, s->cap, s->posA, s->velA,
                       s->idsA, s->sedA, s->slab_counters, s->d_n, s->transit[0], s->transit[1], s->transit_n,
                       s->slab.has_left ? mbox_flag(s->mbox, s->mbox_buf, 0, par) : nullptr,
                       s->slab.has_right ? mbox_flag(s->mbox, s->mbox_buf, 1, par) : nullptr, q, s->peer_timeout);
    s->launches += 1;
    CU(cudaMemcpyAsync(s->slab_host + 8 * slot, s->slab_counters, 8 * sizeof(int), cudaMemcpyDeviceToHost, s->st));
    CU(cudaEventRecord(s->slab_ev[slot], s->st));
    s->slab_add[slot] = ml + mr;
    s->slab_max[slot][0] = ml; s->slab_max[slot][1] = mr;
    s->slab_seq = t + 1;
    s->slab_pending = (int)(s->slab_seq - s->slab_done);
    s->n = (int)std::min<long long>((long long)s->n + ml + mr, s->cap);
    s->extent = (int)std::min<long long>((long long)s->extent_base + ml + mr, s->cap);
    s->extent_pending = true;
    s->binned = false; s->slot_valid = false;
    if (ticket) *ticket = t;
    CU(cudaGetLastError());
    return SPHE_OK;
}

// Sum of an erosion accumulator over the boundary zones with the two x-neighbours, through the peer mailboxes
// (k_zone_sum): which = 0 `want` (between phase 0 and 1), 1 `delta` (between phase 1 and 2); off_left / off_right = first
// element of the zone shared with that neighbour in the accumulator array, -1 = none; count = ints per zone.
int sphe_slab_zone_sum(sphe_sim* s, sphe_terrain* t, int which, long long off_left, long long off_right, int count) {
    if (!s || !t || which < 0 || which > 1 || count < 0) return fail(SPHE_ERR_ARG, "bad arguments");
    if (!s->slab_on || !s->mbox || s->zone_ints <= 0) return fail(SPHE_ERR_STATE, "call sphe_slab_peer_setup_zones first");
    if (count > s->zone_ints) return fail(SPHE_ERR_ARG, "zone of %d ints > %d reserved in the mailbox", count, s->zone_ints);
    if ((off_left >= 0 && !s->peer_mbox[0]) || (off_right >= 0 && !s->peer_mbox[1])) return fail(SPHE_ERR_STATE, "neighbour mailbox not connected");
    const long long cells = (long long)t->rows * t->cols;
    if ((off_left >= 0 && off_left + count > cells) || (off_right >= 0 && off_right + count > cells)) return fail(SPHE_ERR_ARG, "zone outside the terrain");
    TRY(ensure_device(s));
    TRY(terrain_ready(t));
    const int q = ++s->zone_seq[which], par = q & 1;
    int* arr = which ? t->delta : t->want;
    // what goes LEFT lands in the left neighbour's "from the right" buffer, and vice versa
    int* out_l = off_left >= 0 ? zone_buffer(s->peer_mbox[0], s->zone_base, s->zone_ints, 1, which, par) : nullptr;
    int* out_r = off_right >= 0 ? zone_buffer(s->peer_mbox[1], s->zone_base, s->zone_ints, 0, which, par) : nullptr;
    int* fo_l = off_left >= 0 ? zone_flag(s->peer_mbox[0], s->zone_base, s->zone_ints, 1, which, par) : nullptr;
    int* fo_r = off_right >= 0 ? zone_flag(s->peer_mbox[1], s->zone_base, s->zone_ints, 0, which, par) : nullptr;
    launch_zone_sum(s->st, arr, (int)off_left, (int)off_right, count, out_l, out_r, fo_l, fo_r,
                    zone_buffer(s->mbox, s->zone_base, s->zone_ints, 0, which, par), zone_buffer(s->mbox, s->zone_base, s->zone_ints, 1, which, par),
                    zone_flag(s->mbox, s->zone_base, s->zone_ints, 0, which, par), zone_flag(s->mbox, s->zone_base, s->zone_ints, 1, which, par),
                    q, s->peer_timeout, s->zone_done, s->slab_counters + 15);
    s->launches += 1;
    CU(cudaGetLastError());
    return SPHE_OK;
}

// One step of the terrain-coupled simulation in phases, for runs where several slabs share the erosion of
// one terrain: phase 0 = binning .. forces .. contact response + erosion requests; phase 1 = grants;
// phase 2 = apply + cull map.  Between 0 and 1 the per-vertex `want` array must hold the sum over all slabs,
// between 1 and 2 the `delta` array (sphe_terrain_accumulators; integer sums, any order).
int sphe_step_phase(sphe_sim* s, sphe_terrain* t, int phase) {
    if (!s || phase < 0 || phase > 2) return fail(SPHE_ERR_ARG, "bad arguments");
    if (phase == 0) return step_device(s, t, TERRAIN_CONTACT);
    if (!t || s->n == 0) return SPHE_OK;
    TRY(ensure_device(s));
    TRY(terrain_ready(t));
    TerrainDev T = terrain_view(t);
    Scope k(s, SPHE_K_TERRAIN, phase == 1 ? 1 : 2);
    // the step's sediment array is sedA after phase 0 swapped the buffers
    launch_terrain_stage(s->st, s->surv, s->surv_count, s->cap, nullptr, nullptr, nullptr, (int*)s->sedA, s->lastC, T, 1, s->req_vertex,
                         s->req_amount, nullptr, phase == 1 ? TERRAIN_GRANT : TERRAIN_APPLY);
    CU(cudaGetLastError());
    return SPHE_OK;
}

void* sphe_device_ptr(sphe_sim* s, int which) {
    if (!s) return nullptr;
    switch (which) {
    case SPHE_D_POSQ: return s->posA;
    case SPHE_D_VELV: return s->velA;
    case SPHE_D_IDS: return s->idsA;
    case SPHE_D_RHO: return s->rho;
    }
    return nullptr;
}

}  // extern "C"

// ================================================================== terrain (replacement of Grid)
static int terrain_alloc(sphe_terrain* t, int rows, int cols) {
    size_t cells = (size_t)rows * cols;
    if (cells > t->cells_cap) {
        cudaFree(t->hfx); cudaFree(t->want); cudaFree(t->delta); cudaFree(t->lmax);
        t->hfx = t->want = t->delta = t->lmax = nullptr;
        cudaError_t e = cudaMalloc(&t->hfx, cells * sizeof(int));
        if (e == cudaSuccess) e = cudaMalloc(&t->want, cells * sizeof(int));
        if (e == cudaSuccess) e = cudaMalloc(&t->delta, cells * sizeof(int));
        if (e == cudaSuccess) e = cudaMalloc(&t->lmax, cells * sizeof(int));
        if (e != cudaSuccess) return fail(SPHE_ERR_NOMEM, "terrain %d x %d: %s", rows, cols, cudaGetErrorString(e));
        t->cells_cap = cells;
    }
    t->rows = rows; t->cols = cols;
    CU(cudaMemset(t->hfx, 0, cells * sizeof(int)));
    CU(cudaMemset(t->want, 0, cells * sizeof(int)));
    CU(cudaMemset(t->delta, 0, cells * sizeof(int)));
    CU(cudaMemset(t->lmax, 0, cells * sizeof(int)));
    CU(cudaMemset(t->hmax, 0, sizeof(int)));
    return SPHE_OK;
}

static int terrain_ready(sphe_terrain* t) {
    if (!t) return fail(SPHE_ERR_ARG, "NULL terrain");
    if (t->ready) { CU(cudaSetDevice(t->device)); return SPHE_OK; }
    int cnt = 0;
    cudaError_t e = cudaGetDeviceCount(&cnt);
    if (e != cudaSuccess || cnt == 0)
        return fail(SPHE_ERR_CUDA, "no CUDA device (%s); this library has no CPU fallback", cudaGetErrorString(e));
    if (t->device < 0) CU(cudaGetDevice(&t->device));
    CU(cudaSetDevice(t->device));
    CU(cudaMalloc(&t->hmax, sizeof(int)));
    CU(cudaMalloc(&t->d_sum, 3 * sizeof(long long)));
    CU(cudaMemset(t->d_sum, 0, 3 * sizeof(long long)));
    TRY(terrain_alloc(t, t->rows, t->cols));  // Grid(): 512 x 512 heightfield, zeroed (grid.h:78-81)
    t->ready = true;
    return SPHE_OK;
}

static TerrainDev terrain_view(const sphe_terrain* t) {
    TerrainDev T;
    T.rows = t->rows; T.cols = t->cols; T.dimx = t->dimx; T.dimy = t->dimy; T.dimz = t->dimz;
    T.hfx = t->hfx; T.hfx_rw = t->hfx; T.want = t->want; T.delta = t->delta; T.hmax_fx = t->hmax; T.hmax_rw = t->hmax; T.lmax = t->lmax; T.lmax_rw = t->lmax;
    T.ox = t->origin[0]; T.oy = t->origin[1]; T.oz = t->origin[2]; T.scale = t->scale; T.inv_scale = 1.0f / t->scale;
    T.Kc = t->E.Kc; T.Ke = t->E.Ke; T.Kd = t->E.Kd;
    T.hmin_fx = (int)lrint((double)t->E.hmin * 4096.0); T.max_pickup_fx = (int)lrint((double)t->E.max_pickup * 4096.0);
    T.erosion = t->E.enabled ? 1 : 0;
    T.contacts = (unsigned long long*)(t->d_sum + 1);
    T.violations = (unsigned long long*)(t->d_sum + 2);
    const bool windowed = t->win1 > t->win0;
    T.win0 = windowed ? std::max(t->win0, 0) : 0;
    T.win1 = windowed ? std::min(t->win1, t->rows) : t->rows;
    return T;
}

extern "C" {

int sphe_terrain_create(sphe_terrain** out, int dimx, int dimy, int dimz) {
    if (!out || dimx < 1 || dimy < 1 || dimz < 1) return fail(SPHE_ERR_ARG, "bad terrain dimensions");
    sphe_terrain* t = new sphe_terrain();
    t->dimx = dimx; t->dimy = dimy; t->dimz = dimz;
    *out = t;
    return SPHE_OK;
}

void sphe_terrain_destroy(sphe_terrain* t) {
    if (!t) return;
    if (t->ready) {
        cudaSetDevice(t->device);
        cudaDeviceSynchronize();
        void* ptrs[] = {t->hfx, t->want, t->delta, t->hmax, t->lmax, t->d_surface, t->d_indices, t->d_sum};
        for (void* p : ptrs) if (p) cudaFree(p);
    }
    delete t;
}

int sphe_terrain_load_heightfield_ex(sphe_terrain* t, const unsigned char* img, int rows, int cols) {
    if (!t || !img || rows < 2 || cols < 2) return fail(SPHE_ERR_ARG, "bad arguments");
    TRY(terrain_ready(t));
    CU(cudaDeviceSynchronize());
    TRY(terrain_alloc(t, rows, cols));
    size_t cells = (size_t)rows * cols;
    unsigned char* d = nullptr;
    CU(cudaMalloc(&d, cells));
    CU(cudaMemcpy(d, img, cells, cudaMemcpyHostToDevice));  // the caller keeps ownership of img (main.cpp:102-103)
    launch_heights_from_u8(0, (int)cells, d, t->hfx, t->hmax);
    launch_terrain_lmax(0, terrain_view(t));
    CU(cudaDeviceSynchronize());
    CU(cudaFree(d));
    return SPHE_OK;
}

int sphe_terrain_load_heightfield(sphe_terrain* t, const unsigned char* img) {
    return sphe_terrain_load_heightfield_ex(t, img, 512, 512);  // LoadHeightfield, grid.h:98-102
}

int sphe_terrain_set_heights(sphe_terrain* t, const float* h, int rows, int cols) {
    if (!t || !h || rows < 2 || cols < 2) return fail(SPHE_ERR_ARG, "bad arguments");
    TRY(terrain_ready(t));
    CU(cudaDeviceSynchronize());
    TRY(terrain_alloc(t, rows, cols));
    size_t cells = (size_t)rows * cols;
    float* d = nullptr;
    CU(cudaMalloc(&d, cells * sizeof(float)));
    CU(cudaMemcpy(d, h, cells * sizeof(float), cudaMemcpyHostToDevice));
    launch_heights_from_f32(0, (int)cells, d, t->hfx, t->hmax);
    launch_terrain_lmax(0, terrain_view(t));
    CU(cudaDeviceSynchronize());
    CU(cudaFree(d));
    return SPHE_OK;
}

int sphe_terrain_get_heights(sphe_terrain* t, float* h) {
    if (!t || !h) return fail(SPHE_ERR_ARG, "bad arguments");
    TRY(terrain_ready(t));
    CU(cudaDeviceSynchronize());
    size_t cells = (size_t)t->rows * t->cols;
    float* d = nullptr;
    CU(cudaMalloc(&d, cells * sizeof(float)));
    launch_heights_to_f32(0, (int)cells, t->hfx, d);
    CU(cudaMemcpy(h, d, cells * sizeof(float), cudaMemcpyDeviceToHost));
    CU(cudaFree(d));
    return SPHE_OK;
}

int sphe_terrain_get_heights_fx(sphe_terrain* t, int* hfx) {
    if (!t || !hfx) return fail(SPHE_ERR_ARG, "bad arguments");
    TRY(terrain_ready(t));
    CU(cudaDeviceSynchronize());
    CU(cudaMemcpy(hfx, t->hfx, (size_t)t->rows * t->cols * sizeof(int), cudaMemcpyDeviceToHost));
    return SPHE_OK;
}

int sphe_terrain_size(sphe_terrain* t, int* rows, int* cols, int dims[3]) {
    if (!t) return fail(SPHE_ERR_ARG, "NULL terrain");
    if (rows) *rows = t->rows;
    if (cols) *cols = t->cols;
    if (dims) { dims[0] = t->dimx; dims[1] = t->dimy; dims[2] = t->dimz; }
    return SPHE_OK;
}

int sphe_terrain_height_at(sphe_terrain* t, int x, int y) {
    if (!t || x < 0 || y < 0 || x >= t->rows || y >= t->cols) { fail(SPHE_ERR_ARG, "height index out of range"); return -1; }
    if (terrain_ready(t) != SPHE_OK) return -1;
    int v = 0;
    if (cudaDeviceSynchronize() != cudaSuccess || cudaMemcpy(&v, t->hfx + (size_t)t->cols * x + y, sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) {
        fail(SPHE_ERR_CUDA, "height read failed");
        return -1;
    }
    v >>= 12;  // GetHeightfieldAt returns unsigned char (grid.h:104-107): integer part, saturated
    return v < 0 ? 0 : (v > 255 ? 255 : v);
}

int sphe_terrain_set_transform(sphe_terrain* t, const float origin[3], float scale) {
    if (!t || !origin || !(scale > 0.0f)) return fail(SPHE_ERR_ARG, "bad arguments");
    memcpy(t->origin, origin, sizeof t->origin);
    t->scale = scale;
    return SPHE_OK;
}

sphe_erosion* sphe_terrain_erosion_ptr(sphe_terrain* t) { return t ? &t->E : nullptr; }

int sphe_terrain_update_grid(sphe_terrain* t, int dimx, int dimy, int dimz) {
    if (!t || dimx < 1 || dimy < 1 || dimz < 1) return fail(SPHE_ERR_ARG, "bad terrain dimensions");
    TRY(terrain_ready(t));
    CU(cudaDeviceSynchronize());
    t->dimx = dimx; t->dimy = dimy; t->dimz = dimz;
    long long nv = (long long)dimx * dimz, nq = (long long)(dimx - 1) * (dimz - 1);
    cudaFree(t->d_surface); cudaFree(t->d_indices);
    t->d_surface = nullptr; t->d_indices = nullptr;
    CU(cudaMalloc(&t->d_surface, (size_t)std::max<long long>(nv * 6, 1) * sizeof(float)));
    CU(cudaMalloc(&t->d_indices, (size_t)std::max<long long>(nq * 6, 1) * sizeof(unsigned)));
    TerrainDev T = terrain_view(t);
    launch_terrain_surface(0, T, t->d_surface);
    launch_terrain_indices(0, dimx, dimz, t->d_indices);
    CU(cudaDeviceSynchronize());
    t->surface_floats = nv * 6; t->index_count = nq * 6;
    return SPHE_OK;
}

long long sphe_terrain_surface_size(sphe_terrain* t) { return t ? t->surface_floats : 0; }
long long sphe_terrain_indices_size(sphe_terrain* t) { return t ? t->index_count : 0; }

int sphe_terrain_get_surface(sphe_terrain* t, float* out) {
    if (!t || !out) return fail(SPHE_ERR_ARG, "bad arguments");
    if (!t->d_surface) return fail(SPHE_ERR_STATE, "call sphe_terrain_update_grid first");
    CU(cudaSetDevice(t->device));
    CU(cudaMemcpy(out, t->d_surface, (size_t)t->surface_floats * sizeof(float), cudaMemcpyDeviceToHost));
    return SPHE_OK;
}

int sphe_terrain_get_indices(sphe_terrain* t, unsigned* out) {
    if (!t || !out) return fail(SPHE_ERR_ARG, "bad arguments");
    if (!t->d_indices) return fail(SPHE_ERR_STATE, "call sphe_terrain_update_grid first");
    CU(cudaSetDevice(t->device));
    CU(cudaMemcpy(out, t->d_indices, (size_t)t->index_count * sizeof(unsigned), cudaMemcpyDeviceToHost));
    return SPHE_OK;
}

int sphe_terrain_collision(sphe_terrain* t, int n, const float* pos_curr, const float* pos_next, const float* vel_next,
                           int* hit, float* contact, float* normal) {
    if (!t || n < 0 || (n > 0 && (!pos_curr || !pos_next || !vel_next || !hit || !contact || !normal)))
        return fail(SPHE_ERR_ARG, "bad arguments");
    if (n == 0) return SPHE_OK;
    TRY(terrain_ready(t));
    CU(cudaDeviceSynchronize());
    float* d = nullptr;
    int* dh = nullptr;
    size_t v = 3 * (size_t)n;
    CU(cudaMalloc(&d, 5 * v * sizeof(float)));
    CU(cudaMalloc(&dh, (size_t)n * sizeof(int)));
    CU(cudaMemcpy(d, pos_curr, v * sizeof(float), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(d + v, pos_next, v * sizeof(float), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(d + 2 * v, vel_next, v * sizeof(float), cudaMemcpyHostToDevice));
    launch_terrain_collide(0, n, d, d + v, d + 2 * v, terrain_view(t), dh, d + 3 * v, d + 4 * v);
    CU(cudaMemcpy(hit, dh, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(contact, d + 3 * v, v * sizeof(float), cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(normal, d + 4 * v, v * sizeof(float), cudaMemcpyDeviceToHost));
    CU(cudaFree(d)); CU(cudaFree(dh));
    return SPHE_OK;
}

int sphe_terrain_stage_host(sphe_terrain* t, int n, const float* pos_curr, float* pos_next, float* vel_next, int* sediment,
                            float dt, float cR, int* hit) {
    if (!t || n <= 0 || !pos_curr || !pos_next || !vel_next || !sediment) return fail(SPHE_ERR_ARG, "bad arguments");
    TRY(terrain_ready(t));
    CU(cudaDeviceSynchronize());
    float4 *po = nullptr, *pn = nullptr, *vn = nullptr;
    int *sd = nullptr, *rq = nullptr, *dh = nullptr;
    CU(cudaMalloc(&po, (size_t)n * sizeof(float4))); CU(cudaMalloc(&pn, (size_t)n * sizeof(float4)));
    CU(cudaMalloc(&vn, (size_t)n * sizeof(float4))); CU(cudaMalloc(&sd, (size_t)n * sizeof(int)));
    CU(cudaMalloc(&rq, (3 * (size_t)n + 4 + 64 * SPHE_SURV_CLASSES) * sizeof(int))); CU(cudaMalloc(&dh, (size_t)n * sizeof(int)));
    CU(cudaMemset(dh, 0, (size_t)n * sizeof(int)));
    std::vector<float4> a((size_t)n), b((size_t)n), c((size_t)n);
    for (int i = 0; i < n; i++) {
        a[i] = make_float4(pos_curr[3 * i], pos_curr[3 * i + 1], pos_curr[3 * i + 2], 0.f);
        b[i] = make_float4(pos_next[3 * i], pos_next[3 * i + 1], pos_next[3 * i + 2], 0.f);
        c[i] = make_float4(vel_next[3 * i], vel_next[3 * i + 1], vel_next[3 * i + 2], 0.f);
    }
    CU(cudaMemcpy(po, a.data(), (size_t)n * sizeof(float4), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(pn, b.data(), (size_t)n * sizeof(float4), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(vn, c.data(), (size_t)n * sizeof(float4), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(sd, sediment, (size_t)n * sizeof(int), cudaMemcpyHostToDevice));
    StepC C{};
    C.dt = dt; C.cR = cR; C.box = 0; C.cube = 1; C.lenx = C.leny = C.lenz = C.len = 3.0e38f;
    // layout of rq: request vertices [n + 32 K], request amounts [n + 32 K], survivor list [n] (class 0 only), counts [4]
    const size_t rn = (size_t)n + 32 * SPHE_SURV_CLASSES;
    launch_iota(0, n, rq + 2 * rn, rq + 2 * rn + n);   // no cull in front of the hook: everyone is a survivor
    launch_terrain_stage(0, rq + 2 * rn, rq + 2 * rn + n, n, po, pn, vn, sd, C, terrain_view(t), 0, rq, rq + rn, dh);
    CU(cudaDeviceSynchronize());
    CU(cudaMemcpy(b.data(), pn, (size_t)n * sizeof(float4), cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(c.data(), vn, (size_t)n * sizeof(float4), cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(sediment, sd, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost));
    if (hit) CU(cudaMemcpy(hit, dh, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost));
    for (int i = 0; i < n; i++) {
        pos_next[3 * i] = b[i].x; pos_next[3 * i + 1] = b[i].y; pos_next[3 * i + 2] = b[i].z;
        vel_next[3 * i] = c[i].x; vel_next[3 * i + 1] = c[i].y; vel_next[3 * i + 2] = c[i].z;
    }
    cudaFree(po); cudaFree(pn); cudaFree(vn); cudaFree(sd); cudaFree(rq); cudaFree(dh);
    return SPHE_OK;
}

int sphe_terrain_accumulators(sphe_terrain* t, void** want, void** delta, long long* cells) {
    if (!t) return fail(SPHE_ERR_ARG, "NULL terrain");
    TRY(terrain_ready(t));
    if (want) *want = t->want;
    if (delta) *delta = t->delta;
    if (cells) *cells = (long long)t->rows * t->cols;
    return SPHE_OK;
}

int sphe_terrain_heights_device(sphe_terrain* t, void** hfx, long long* cells) {
    if (!t || !hfx) return fail(SPHE_ERR_ARG, "bad arguments");
    TRY(terrain_ready(t));
    *hfx = t->hfx;
    if (cells) *cells = (long long)t->rows * t->cols;
    return SPHE_OK;
}

int sphe_terrain_refresh(sphe_terrain* t) {
    if (!t) return fail(SPHE_ERR_ARG, "NULL terrain");
    TRY(terrain_ready(t));
    CU(cudaDeviceSynchronize());
    launch_terrain_lmax(0, terrain_view(t));
    CU(cudaDeviceSynchronize());
    return SPHE_OK;
}

int sphe_terrain_total_fx_rows(sphe_terrain* t, int row0, int row1, long long* sum) {
    if (!t || !sum) return fail(SPHE_ERR_ARG, "bad arguments");
    TRY(terrain_ready(t));
    row0 = std::max(row0, 0); row1 = std::min(row1, t->rows);
    *sum = 0;
    if (row1 <= row0) return SPHE_OK;
    CU(cudaDeviceSynchronize());
    CU(cudaMemset(t->d_sum, 0, sizeof(long long)));
    launch_sum_i32(0, (row1 - row0) * t->cols, t->hfx + (size_t)row0 * t->cols, nullptr, t->d_sum);
    CU(cudaMemcpy(sum, t->d_sum, sizeof(long long), cudaMemcpyDeviceToHost));
    return SPHE_OK;
}

int sphe_terrain_total_fx(sphe_terrain* t, long long* sum) {
    if (!t) return fail(SPHE_ERR_ARG, "bad arguments");
    return sphe_terrain_total_fx_rows(t, 0, t->rows, sum);
}

int sphe_terrain_set_window(sphe_terrain* t, int row0, int row1) {
    if (!t) return fail(SPHE_ERR_ARG, "NULL terrain");
    if (row1 > row0 && row1 - row0 < 8) return fail(SPHE_ERR_ARG, "terrain window [%d,%d) is narrower than 8 rows", row0, row1);
    t->win0 = row0; t->win1 = row1;
    return SPHE_OK;
}

int sphe_terrain_window_violations(sphe_terrain* t, long long* count) {
    if (!t || !count) return fail(SPHE_ERR_ARG, "bad arguments");
    TRY(terrain_ready(t));
    CU(cudaDeviceSynchronize());
    CU(cudaMemcpy(count, t->d_sum + 2, sizeof(long long), cudaMemcpyDeviceToHost));
    return SPHE_OK;
}

int sphe_terrain_contacts(sphe_terrain* t, long long* total, int reset) {
    if (!t || !total) return fail(SPHE_ERR_ARG, "bad arguments");
    TRY(terrain_ready(t));
    CU(cudaDeviceSynchronize());
    CU(cudaMemcpy(total, t->d_sum + 1, sizeof(long long), cudaMemcpyDeviceToHost));
    if (reset) CU(cudaMemset(t->d_sum + 1, 0, sizeof(long long)));
    return SPHE_OK;
}

// Survivors of the contact cull in the last step, per Grid::collision path class (same cell / one axis / both axes).
int sphe_terrain_survivors(sphe_sim* s, int out[3]) {
    if (!s || !out) return fail(SPHE_ERR_ARG, "bad arguments");
    out[0] = out[1] = out[2] = 0;
    if (!s->surv_count) return SPHE_OK;
    TRY(ensure_device(s));
    CU(cudaMemcpyAsync(out, s->surv_count, 3 * sizeof(int), cudaMemcpyDeviceToHost, s->st));
    CU(cudaStreamSynchronize(s->st));
    return SPHE_OK;
}

int sphe_sediment_total_fx(sphe_sim* s, long long* sum) {
    if (!s || !sum) return fail(SPHE_ERR_ARG, "bad arguments");
    TRY(ensure_device(s));
    TRY(slab_settle(s));
    *sum = 0;
    if (s->n == 0) return SPHE_OK;
    long long* d = nullptr;
    CU(cudaMalloc(&d, sizeof(long long)));
    CU(cudaMemsetAsync(d, 0, sizeof(long long), s->st));
    launch_sum_i32(s->st, s->slab_on ? slab_extent_bound(s) : s->n, (const int*)s->sedA, s->slab_on ? s->idsA : nullptr, d,
                   s->slab_on ? slab_extent_dev(s) : nullptr);
    CU(cudaMemcpyAsync(sum, d, sizeof(long long), cudaMemcpyDeviceToHost, s->st));
    CU(cudaStreamSynchronize(s->st));
    CU(cudaFree(d));
    return SPHE_OK;
}

int sphe_set_sediment_fx(sphe_sim* s, const int* sediment_by_id) {
    if (!s || !sediment_by_id) return fail(SPHE_ERR_ARG, "bad arguments");
    if (s->slab_on) return fail(SPHE_ERR_STATE, "not available in slab mode");
    TRY(ensure_device(s));
    if (s->n == 0) return SPHE_OK;
    std::vector<int> ids((size_t)s->n), out((size_t)s->n);
    CU(cudaMemcpyAsync(ids.data(), s->idsA, (size_t)s->n * sizeof(int), cudaMemcpyDeviceToHost, s->st));
    CU(cudaStreamSynchronize(s->st));
    for (int i = 0; i < s->n; i++) out[i] = sediment_by_id[ids[i]];
    CU(cudaMemcpyAsync(s->sedA, out.data(), (size_t)s->n * sizeof(int), cudaMemcpyHostToDevice, s->st));
    CU(cudaStreamSynchronize(s->st));
    return SPHE_OK;
}


// ---- on-disk state (the reference has no format).  Little-endian, everything a bit-exact resume
// needs: parameters, bookkeeping, positions/velocities (id order), carried sediment (fixed point), and -- when a
// terrain is given -- its fixed-point heights, transform and erosion parameters.  Results do not depend on the
// storage order of the particles (the binning makes the order canonical), so a resumed run continues bit for bit.
namespace {
struct StateHeader {
    char magic[8];          // "SPHESTA1"
    int n, has_terrain, n_labels, box_user;
    sphe_params P;
    float origin[3], box[3];
    int num, init_num, next_label, reserved;
};
struct TerrainHeader {
    int rows, cols, dims[3];
    float origin[3], scale;
    sphe_erosion E;
};
struct File {
    FILE* f = nullptr;
    ~File() { if (f) fclose(f); }
};
}  // namespace

static int save_state_to(sphe_sim* s, sphe_terrain* t, const char* path) {
    const int n = s->n;
    StateHeader H{};
    memcpy(H.magic, "SPHESTA1", 8);
    H.n = n; H.has_terrain = t ? 1 : 0; H.n_labels = (int)s->labels.size(); H.box_user = s->box_user ? 1 : 0;
    H.P = s->P;
    for (int a = 0; a < 3; a++) { H.origin[a] = s->origin[a]; H.box[a] = s->box[a]; }
    H.num = s->num; H.init_num = s->init_num; H.next_label = s->next_label;
    std::vector<float> pos(3 * (size_t)n), vel(3 * (size_t)n);
    std::vector<int> sed((size_t)n);
    if (n > 0) {
        TRY(sphe_download(s, SPHE_F_POS, pos.data()));
        TRY(sphe_download(s, SPHE_F_VEL, vel.data()));
        launch_unsort_f1(s->st, n, s->sedA, s->idsA, s->stage);    // raw fixed-point words, id order
        CU(cudaMemcpyAsync(sed.data(), s->stage, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, s->st));
        CU(cudaStreamSynchronize(s->st));
    }
    // everything that can fail on the device happens BEFORE the file is opened
    TerrainHeader T{};
    std::vector<int> hfx;
    if (t) {
        TRY(terrain_ready(t));
        T.rows = t->rows; T.cols = t->cols; T.dims[0] = t->dimx; T.dims[1] = t->dimy; T.dims[2] = t->dimz;
        for (int a = 0; a < 3; a++) T.origin[a] = t->origin[a];
        T.scale = t->scale; T.E = t->E;
        hfx.resize((size_t)t->rows * t->cols);
        TRY(sphe_terrain_get_heights_fx(t, hfx.data()));
    }
    File F; F.f = fopen(path, "wb");
    if (!F.f) return fail(SPHE_ERR_ARG, "cannot write %s", path);
    bool ok = fwrite(&H, sizeof H, 1, F.f) == 1;
    ok = ok && (s->labels.empty() || fwrite(s->labels.data(), sizeof(int), s->labels.size(), F.f) == s->labels.size());
    ok = ok && (n == 0 || (fwrite(pos.data(), sizeof(float), pos.size(), F.f) == pos.size() &&
                           fwrite(vel.data(), sizeof(float), vel.size(), F.f) == vel.size() &&
                           fwrite(sed.data(), sizeof(int), sed.size(), F.f) == sed.size()));
    if (ok && t) ok = fwrite(&T, sizeof T, 1, F.f) == 1 && fwrite(hfx.data(), sizeof(int), hfx.size(), F.f) == hfx.size();
    ok = ok && fflush(F.f) == 0;
    if (!ok) return fail(SPHE_ERR_ARG, "short write to %s", path);
    return SPHE_OK;
}

// Writes to "<path>.tmp" and renames: a failure never leaves a partial file under `path`.
int sphe_save_state(sphe_sim* s, sphe_terrain* t, const char* path) {
    if (!s || !path) return fail(SPHE_ERR_ARG, "bad arguments");
    if (s->slab_on) return fail(SPHE_ERR_STATE, "not available in slab mode (use sphe_slab_download per rank)");
    TRY(ensure_device(s));
    try {
        const std::string tmp = std::string(path) + ".tmp";
        int rc = save_state_to(s, t, tmp.c_str());
        if (rc != SPHE_OK) { remove(tmp.c_str()); return rc; }
        if (rename(tmp.c_str(), path) != 0) { remove(tmp.c_str()); return fail(SPHE_ERR_ARG, "cannot rename %s to %s", tmp.c_str(), path); }
    } catch (const std::bad_alloc&) {
        return fail(SPHE_ERR_NOMEM, "out of host memory while saving %s", path);
    }
    return SPHE_OK;
}

// The whole file is read and checked against its own length BEFORE the simulation or the terrain is touched: a corrupt or
// truncated file leaves both exactly as they were.
static int load_state_from(sphe_sim* s, sphe_terrain* t, const char* path) {
    File F; F.f = fopen(path, "rb");
    if (!F.f) return fail(SPHE_ERR_ARG, "cannot read %s", path);
    if (fseek(F.f, 0, SEEK_END) != 0) return fail(SPHE_ERR_ARG, "cannot seek in %s", path);
    const long long file_bytes = ftell(F.f);
    rewind(F.f);
    StateHeader H;
    if (file_bytes < (long long)sizeof H || fread(&H, sizeof H, 1, F.f) != 1 || memcmp(H.magic, "SPHESTA1", 8) != 0)
        return fail(SPHE_ERR_ARG, "%s is not a sphe state file", path);
    if (H.n < 0 || H.n_labels < 0 || (H.n_labels != 0 && H.n_labels != H.n)) return fail(SPHE_ERR_ARG, "%s: corrupt header", path);
    if (H.has_terrain && !t) return fail(SPHE_ERR_ARG, "%s holds a terrain: pass a terrain handle to receive it", path);
    const size_t n = (size_t)H.n;
    // sizes from the file are only trusted as far as the file is long
    long long need = (long long)sizeof H + (long long)H.n_labels * 4 + (long long)n * 28;
    if (need > file_bytes) return fail(SPHE_ERR_ARG, "%s: truncated particle data (%lld bytes for %d particles, file has %lld)", path, need, H.n, file_bytes);
    std::vector<int> labels((size_t)H.n_labels), sed(n);
    std::vector<float> pos(3 * n), vel(3 * n);
    bool ok = labels.empty() || fread(labels.data(), sizeof(int), labels.size(), F.f) == labels.size();
    ok = ok && (n == 0 || (fread(pos.data(), sizeof(float), pos.size(), F.f) == pos.size() &&
                           fread(vel.data(), sizeof(float), vel.size(), F.f) == vel.size() &&
                           fread(sed.data(), sizeof(int), sed.size(), F.f) == sed.size()));
    if (!ok) return fail(SPHE_ERR_ARG, "%s: truncated particle data", path);
    TerrainHeader T{};
    std::vector<float> h;
    if (H.has_terrain) {
        if (fread(&T, sizeof T, 1, F.f) != 1) return fail(SPHE_ERR_ARG, "%s: truncated terrain header", path);
        if (T.rows < 2 || T.cols < 2 || T.rows > 65536 || T.cols > 65536 || (long long)T.rows * T.cols > (1LL << 30))
            return fail(SPHE_ERR_ARG, "%s: terrain of %d x %d vertices", path, T.rows, T.cols);
        need += (long long)sizeof T + (long long)T.rows * T.cols * 4;
        if (need > file_bytes) return fail(SPHE_ERR_ARG, "%s: truncated terrain heights", path);
        std::vector<int> hfx((size_t)T.rows * T.cols);
        if (fread(hfx.data(), sizeof(int), hfx.size(), F.f) != hfx.size()) return fail(SPHE_ERR_ARG, "%s: truncated terrain heights", path);
        h.resize(hfx.size());
        for (size_t i = 0; i < hfx.size(); i++) {
            if (hfx[i] > (1 << 24) || hfx[i] < -(1 << 24)) return fail(SPHE_ERR_ARG, "%s: height out of the exactly representable range", path);
            h[i] = (float)hfx[i] * (1.0f / 4096.0f);     // exact: |hfx| < 2^24, and set_heights rounds h * 4096 back to hfx
        }
    }
    // ---- nothing above touched s or t
    s->P = H.P;
    for (int a = 0; a < 3; a++) { s->origin[a] = H.origin[a]; s->box[a] = H.box[a]; }
    s->box_user = H.box_user != 0;
    s->grid_h = -1.f;   // re-derive the neighbour grid from the loaded parameters
    TRY(sphe_upload_state(s, H.n, pos.data(), vel.data()));
    s->num = H.num; s->init_num = H.init_num; s->next_label = H.next_label;
    s->labels = labels; s->labels_identity = labels.empty();
    if (n > 0) TRY(sphe_set_sediment_fx(s, sed.data()));
    if (H.has_terrain) {
        t->dimx = T.dims[0]; t->dimy = T.dims[1]; t->dimz = T.dims[2];
        TRY(sphe_terrain_set_heights(t, h.data(), T.rows, T.cols));
        TRY(sphe_terrain_set_transform(t, T.origin, T.scale));
        t->E = T.E;
    }
    return SPHE_OK;
}

int sphe_load_state(sphe_sim* s, sphe_terrain* t, const char* path) {
    if (!s || !path) return fail(SPHE_ERR_ARG, "bad arguments");
    if (s->slab_on) return fail(SPHE_ERR_STATE, "not available in slab mode");
    try {
        return load_state_from(s, t, path);
    } catch (const std::bad_alloc&) {
        return fail(SPHE_ERR_NOMEM, "out of host memory while loading %s", path);
    }
}

}  // extern "C"
