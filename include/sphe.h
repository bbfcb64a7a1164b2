/* sphe.h -- C ABI of the B200-native SPH-Erosion hot path (libsphe_b200.so).
 *
 * The reference has no FFI layer: its boundary is the public C++ surface of `FluidSystemSPH`
 * (Erosion/fluid_system.h:66-289) and `Grid` (Erosion/grid.h:26-841) as used by Erosion/main.cpp.
 * Every entry point below names the reference member it replaces.  The header-only shim classes in
 * sph-erosion_b200/host/{fluid_system.h,grid.h} re-create that C++ surface on top of this ABI, so
 * main.cpp (or a headless driver) compiles against them unchanged -- see INTEGRATION.md.
 *
 * Conventions: plain pointers and sizes only; every call returns SPHE_OK (0) or a negative error
 * code and records a message retrievable with sphe_last_error(); all host arrays are in PARTICLE-ID
 * order (the order of the reference's std::vector<FluidParticle>); vec3 arrays are packed xyz float32.
 * Threading: like the reference (single render thread, main.cpp:197-357) a handle is not re-entrant.
 * There is no CPU fallback: without a CUDA device every compute call fails with SPHE_ERR_CUDA.
 */
#ifndef SPHE_H
#define SPHE_H

#ifdef __cplusplus
extern "C" {
#endif

#define SPHE_OK 0
#define SPHE_ERR_CUDA (-1)     /* CUDA runtime error / no device */
#define SPHE_ERR_ARG (-2)      /* bad argument */
#define SPHE_ERR_STATE (-3)    /* call out of order (e.g. debug hook before any step) */
#define SPHE_ERR_NOMEM (-4)

typedef struct sphe_sim sphe_sim;         /* replaces a FluidSystemSPH instance */
typedef struct sphe_terrain sphe_terrain; /* replaces a Grid instance */

/* Host-resident simulation parameters.  The reference hands ImGui raw pointers into the object
 * (GetMass/GetVisc/GetSurfTen/Getp0/GetGrav, fluid_system.h:261-284, used at main.cpp:278-290), so
 * these live in host memory owned by the handle, are stable for its lifetime and are re-read by
 * every sphe_step().  Defaults: fluid_system.h:460-478. */
typedef struct sphe_params {
    float mass;      /* MASS      0.02    */
    float visc;      /* visc      3.5     */
    float surf_tens; /* surf_tens 0.0728  */
    float p0;        /* p0        998.29  */
    float g[3];      /* g         (0,-9.82,0) */
    float dt;        /* deltaT    0 (paused; main.cpp:476-485 toggles 0.01) */
    float k;         /* k         3.0  (private, no setter in the reference) */
    float h;         /* h = smoothRadius 0.0457 (private) */
    float len;       /* box half extent 0.2 (private) */
    float cR;        /* terrain restitution 0.5 (fluid_system.h:470, used only by the commented call :335-340) */
} sphe_params;

/* Mirror of `struct FluidParticle` (fluid_system.h:49-64), 112 bytes, what GetParticle returns. */
typedef struct sphe_particle {
    int id;
    float position[3], velocity[3], acceleration[3];
    float density, pressure;
    float pressure_force[3], viscosity_force[3], gravity_force[3], surface_force[3], surface_normal[3];
    int neighb_id;
} sphe_particle;

/* Field selectors for sphe_download (element type / width in brackets). */
enum {
    SPHE_F_POS = 0,      /* float[3] */
    SPHE_F_VEL = 1,      /* float[3] */
    SPHE_F_ACC = 2,      /* float[3]  diagnostics */
    SPHE_F_DENSITY = 3,  /* float     */
    SPHE_F_PRESSURE = 4, /* float     */
    SPHE_F_FPRESS = 5,   /* float[3]  diagnostics */
    SPHE_F_FVISC = 6,    /* float[3]  diagnostics */
    SPHE_F_FGRAV = 7,    /* float[3]  diagnostics */
    SPHE_F_FSURF = 8,    /* float[3]  diagnostics */
    SPHE_F_NORMAL = 9,   /* float[3]  diagnostics */
    SPHE_F_ID = 10,      /* int       */
    SPHE_F_NEIGHB = 11,  /* int       diagnostics */
    SPHE_F_SEDIMENT = 12 /* float     carried sediment (erosion model, not in the reference) */
};

/* Neighbour-grid description (this project's definition; the reference is all-pairs). */
typedef struct sphe_grid_info {
    float gmin[3];
    float cell;   /* edge = h * (1 + 2^-10) */
    int dim[3];   /* cell id = (cx*dim[1] + cy)*dim[2] + cz, coordinates clamped into the grid
                   (slab mode: the window's dims, cx relative to sphe_slab_info's xoff) */
} sphe_grid_info;

/* Per-kernel device times of the timed steps (CUDA events on the launch stream). */
enum { SPHE_K_HASH = 0, SPHE_K_SCAN, SPHE_K_SCATTER, SPHE_K_REORDER, SPHE_K_DENSITY, SPHE_K_FORCE,
       SPHE_K_TERRAIN, SPHE_K_COUNT };

const char* sphe_last_error(void);
int sphe_abi_version(void);

/* ---- lifetime ---- */
/* FluidSystemSPH() (fluid_system.h:69-72).  Does NO CUDA work: the reference object is a global
 * constructed before main() (main.cpp:50).  Device state is created lazily by the first call that
 * needs it. */
int sphe_create(sphe_sim** out);
void sphe_destroy(sphe_sim* s);
int sphe_set_device(sphe_sim* s, int device); /* before first use; default = current device */
int sphe_device_count(void);                  /* visible CUDA devices (0 when there is none) */

/* ---- scene ---- */
int sphe_initialize(sphe_sim* s, int n_parts);   /* Initialize(int)   fluid_system.h:74-102  */
int sphe_add_particles(sphe_sim* s, int n);      /* AddParticles(int) fluid_system.h:232-251 */
int sphe_reset(sphe_sim* s);                     /* Reset()           fluid_system.h:253-259 */
int sphe_set_origin(sphe_sim* s, const float o[3]); /* SetOrigin      fluid_system.h:206-209 */
int sphe_get_origin(sphe_sim* s, float o[3]);       /* GetOrigin      fluid_system.h:211-214 */
int sphe_set_dt(sphe_sim* s, float dt);             /* SetDeltaTime   fluid_system.h:216-219 */
float sphe_get_dt(sphe_sim* s);                     /* GetDeltaTime   fluid_system.h:221-224 */
sphe_params* sphe_params_ptr(sphe_sim* s);          /* GetMass..GetGrav fluid_system.h:261-284 */
int sphe_count(sphe_sim* s);                        /* m_Particles.size() */
int sphe_num(sphe_sim* s);                          /* `num` (what PrintCoords iterates, :226-230) */

/* Neighbour-grid bounds.  Default: the box [-len,len]^3 padded by 2 cells.  Particles outside are
 * clamped into the border cells (correctness is unaffected, only balance). */
int sphe_set_grid_bounds(sphe_sim* s, const float lo[3], const float hi[3]);
int sphe_grid_info_get(sphe_sim* s, sphe_grid_info* out);

/* Per-axis box half-extents (the reference box is the cube `len`, fluid_system.h:478; this is the
 * generalisation the multi-GPU channel scenes need).  NULL restores the cube.  With unequal extents
 * collisionS picks the axis of largest overshoot |coord| - half instead of largest |coord|. */
int sphe_set_box(sphe_sim* s, const float half[3]);

/* Replace the whole state (parity tests, checkpoints, the e2e host-buffer path); ids become 0..n-1. */
int sphe_upload_state(sphe_sim* s, int n, const float* pos, const float* vel);

/* ---- the hot path ---- */
/* Run(Grid&) (fluid_system.h:104-183 + advance :306-353 + collisionS :355-407).  `t` may be NULL
 * (the reference ignores its Grid argument: the call into Grid::collision is commented out,
 * fluid_system.h:335-340).  With a terrain attached the particle-terrain contact of Grid::collision
 * (grid.h:462-805) and the erosion model run after the box collision. */
int sphe_step(sphe_sim* s, sphe_terrain* t);
int sphe_sync(sphe_sim* s);

/* One step with HOST buffers: upload pos/vel (id order), step, download pos/vel (+density if non-NULL).
 * This is the end-to-end call bench.py times ("e2e"). */
int sphe_step_host(sphe_sim* s, sphe_terrain* t, int n, const float* pos_in, const float* vel_in,
                   float* pos_out, float* vel_out, float* density_out);

/* Run `steps` steps, timing every step (summed into ms_total) and every kernel with CUDA events on the launch stream.
 * ms_kernels[SPHE_K_COUNT] receives the summed device time per kernel class; launches receives the
 * number of kernel launches issued. */
/* Timing hygiene for sphe_timed_steps: write `bytes` (> L2 size) of scratch between timed steps, outside
 * the timed brackets.  0 disables. */
int sphe_set_l2_flush(sphe_sim* s, long long bytes);
int sphe_timed_steps(sphe_sim* s, sphe_terrain* t, int steps, float* ms_total, float* ms_kernels, int* launches);

/* Free-running variant for callers that own the step loop (the multi-GPU driver): switch per-kernel
 * event timing on, step, then collect the summed times / launch count since the last collect. */
int sphe_kernel_timing(sphe_sim* s, int on);
int sphe_kernel_times(sphe_sim* s, float* ms_kernels, int* launches);

/* ---- accessors ---- */
/* Store the per-particle debug fields the reference keeps in FluidParticle (forces, normal,
 * acceleration, NeighbId).  Off by default; sphe_get_particle switches it on. */
int sphe_set_diagnostics(sphe_sim* s, int on);
int sphe_get_particle(sphe_sim* s, int id, sphe_particle* out); /* GetParticle fluid_system.h:286-289 */
int sphe_download(sphe_sim* s, int field, void* host_out);      /* whole field, id order */
/* Packed positions for Draw() (fluid_system.h:185-204): xyz float32, id order. */
int sphe_download_positions(sphe_sim* s, float* host_xyz);
/* The same into a DEVICE buffer of the caller (e.g. a mapped OpenGL vertex buffer: CUDA-GL interop, no PCIe round trip);
 * capacity_floats >= 3 * sphe_count(s).  One instanced draw call then replaces the reference's draw call per particle. */
int sphe_write_positions_device(sphe_sim* s, void* device_xyz, long long capacity_floats);

/* ---- neighbour-grid test hooks (state of the LAST step's binning) ---- */
int sphe_debug_cells(sphe_sim* s, int* cell_of_id);          /* [n]  cell id per particle id       */
int sphe_debug_sorted_order(sphe_sim* s, int* ids_sorted);   /* [n]  ids in (cell,id) order        */
int sphe_debug_cell_start(sphe_sim* s, int* cell_start);     /* [ncells+1]                         */
/* CSR neighbour lists by sorted slot, ids in grid-walk order, self included.  Call with nbr=NULL to
 * get the total in *total. */
int sphe_debug_neighbours(sphe_sim* s, long long* nbr_start, int* nbr, long long cap, long long* total);
/* The PRODUCTION neighbour lists of the last step (the pair index lists, or the bit masks of the staged variant, that
 * the force pass walked), decoded: for the particle in sorted slot i, entries[i*cap .. i*cap+counts[i]) are the sorted
 * slots of the candidates recorded for it (a pair of targets shares one list, so this is a superset of its neighbours;
 * extra entries have zero weight); counts[i] = -1: the force pass walked the cells directly for this particle (list
 * beyond its rows / cell neighbourhood larger than the stage). */
int sphe_debug_pair_lists(sphe_sim* s, int cap, int* counts, int* entries);

/* ---- multi-GPU x-slabs (no reference counterpart) ----
 * One handle per GPU owns the global cell columns [x0, x1) of the neighbour grid and bins into the
 * window [x0-2, x1+2).  Per step the driver (sph-erosion_b200/slabs.py, NCCL P2P) does
 *     pack -> exchange records with the two x-neighbours -> unpack -> sphe_step.
 * Records are 32 bytes: (x, y, z, sediment) (vx, vy, vz, id bits); both migrants and the 2-layer
 * halo travel in the same buffer, the receiver classifies each record by its own cell column.
 * Particle ids are global, < 2^30 (bit 30 marks ghost copies). */
int sphe_slab_configure(sphe_sim* s, int x0, int x1, int has_left, int has_right);
/* Owned particles per global cell column, hist[gnx] (load balancing: re-cut the slabs by particle-count quantiles).
 * Computed on the device; only the gnx counts cross PCIe. */
int sphe_slab_column_histogram(sphe_sim* s, int gnx, int* hist);
/* Ring closure for 3 or more slabs.  The reference's box response sends a particle that sits EXACTLY on the
 * -x wall to the +x wall (collisionS, fluid_system.h:375-382: x == -len takes the `else` branch), i.e. from
 * the first slab to the last one in a single step.  wrap_left (first slab; configure it with has_left = 1):
 * the left link goes to the last slab and carries those particles (cell column >= far_x0 = the last slab's
 * x0) and nothing else.  wrap_right (last slab; has_right = 1): the right link goes to the first slab and
 * carries no payload; it keeps the flag protocol of the mailboxes symmetric.  Wrap links carry no halo. */
int sphe_slab_ring(sphe_sim* s, int wrap_left, int wrap_right, int far_x0);
int sphe_slab_info(sphe_sim* s, int* gnx, int* xoff, int* n_total, int* n_owned);
int sphe_slab_upload(sphe_sim* s, int n, const float* pos, const float* vel, const int* ids);
/* A particle that crosses MORE than one slab in a step (the reference's contact response can eject particles at
 * hundreds of box units per second) is received by a slab that is not its owner: that slab hands the record on
 * in the same direction at the next exchange, so the particle reaches its owner one step per extra slab later
 * instead of being lost.  out[3] = {records waiting to go left, waiting to go right, forwarded so far};
 * owned + in-transit particles over all slabs is conserved. */
int sphe_slab_transit(sphe_sim* s, int out[3]);
/* Exchange buffers hold cap_records + 1 records; record 0 is a header whose first int is the payload
 * count, so a buffer can be sent with a size both sides agree on beforehand and no count has to reach
 * the host before the transfer is posted.
 * pack:   drops ghosts, compacts what stays, fills the two DEVICE send buffers.  Asynchronous.
 *         reserve_incoming = upper bound of the records that can arrive (array growth happens here).
 * unpack: appends the payload of the two received DEVICE buffers (NULL = no neighbour on that side;
 *         at most max_left / max_right records were transferred), then performs the step's only host
 *         sync and returns out[6] = {n_total, n_owned, sent_left, sent_right, got_left, got_right}. */
int sphe_slab_pack(sphe_sim* s, void* dev_send_left, void* dev_send_right, int cap_records, int reserve_incoming);
int sphe_slab_unpack(sphe_sim* s, const void* dev_recv_left, int max_left, const void* dev_recv_right, int max_right, int out[6]);
/* Asynchronous form: no host sync at all.  The append kernel leaves the exact particle count in device
 * memory, the following sphe_step launches with an upper bound and its kernels read the exact count
 * there; the counters travel to pinned host memory behind an event.  sphe_slab_result(ticket) returns the
 * same out[6] later (wait = 0: returns 1 if the step has not finished yet).  At most 6 tickets may be
 * outstanding; every accessor that needs the exact count settles them first. */
int sphe_slab_unpack_async(sphe_sim* s, const void* dev_recv_left, int max_left, const void* dev_recv_right, int max_right,
                           long long* ticket);
int sphe_slab_result(sphe_sim* s, long long ticket, int wait, int out[6]);
/* Peer-memory exchange (the multi-GPU default): no transport library on the data path.  Every slab owns a
 * MAILBOX in its GPU's memory (4 record buffers -- from the left / from the right neighbour, double buffered
 * by exchange parity -- and 4 flags).  sphe_slab_send runs the pack kernel with the NEIGHBOURS' mailboxes as
 * its output (stores over NVLink, peer memory mapped with cudaIpcOpenMemHandle) and then publishes the
 * record counts and a sequence flag (release, system scope); sphe_slab_recv launches the append kernel, which
 * waits ON THE DEVICE for its own mailbox flags (acquire) and appends the payload.  Neither call syncs the
 * host; sphe_slab_result(ticket) returns the counts later.  A flag that does not arrive within the timeout
 * (default 4e9 clock cycles) is reported as an error by sphe_slab_result instead of hanging the GPU.
 *   setup:   allocates the mailbox for cap_records payload records per buffer (same on every slab) and
 *            reserves particle storage (so the arrays never grow mid-run);
 *   handle:  64-byte cudaIpcMemHandle of the mailbox, to be handed to the two neighbour processes;
 *   connect: maps the neighbours' mailboxes (NULL = no neighbour on that side);
 *   connect_local: the same for slabs that live in ONE process -- on one GPU (tests) or on several GPUs of the
 *            node (direct peer access is enabled; host/headless_slabs.cpp drives N GPUs from one C++ thread,
 *            every call only enqueues). */
int sphe_slab_peer_setup(sphe_sim* s, int cap_records, int reserve_particles);
int sphe_slab_peer_handle(sphe_sim* s, void* handle64);
int sphe_slab_peer_connect(sphe_sim* s, const void* left_handle64, const void* right_handle64);
int sphe_slab_peer_connect_local(sphe_sim* s, sphe_sim* left, sphe_sim* right);
/* The same mailbox with room for terrain zone sums (slab-local terrain): zone_ints = ints of one boundary zone of the erosion
 * accumulators (2W rows x cols).  sphe_slab_zone_sum adds the x-neighbours' copies of the zones into `want` (which = 0,
 * between sphe_step_phase 0 and 1) or `delta` (which = 1, between 1 and 2) in ONE launch: remote stores + flags, no host sync.
 * off_left / off_right: first element of the zone shared with that neighbour, -1 = none. */
int sphe_slab_peer_setup_zones(sphe_sim* s, int cap_records, int reserve_particles, int zone_ints);
int sphe_slab_zone_sum(sphe_sim* s, sphe_terrain* t, int which, long long off_left, long long off_right, int count);
int sphe_slab_peer_timeout(sphe_sim* s, long long clock_cycles);
int sphe_slab_send(sphe_sim* s);
int sphe_slab_recv(sphe_sim* s, long long* ticket);
/* Owned particles only, storage order; rho/sed may be NULL. */
int sphe_slab_download(sphe_sim* s, int cap, int* ids, float* pos, float* vel, float* rho, float* sed, int* n_out);

/* ---- terrain: replacement of the reference's Grid (Erosion/grid.h) ----
 * Heightfield storage is rows x cols with H(x, z) = map[cols * x + z] (GetHeightfieldAt, grid.h:104-107;
 * the reference hard-codes 512 x 512, grid.h:78-81); dimx/dimy/dimz are the Grid dimensions used for the
 * range checks of collision() and for the render mesh (main.cpp:100-105 uses 50, 255, 50).
 * Heights live on the device in fixed point (1/4096) so that erosion is exactly conservative.
 * World <-> terrain coordinates: world = origin + scale * terrain (uniform; default identity, like the
 * reference where 1 cell = 1 world unit). */
typedef struct sphe_erosion {   /* this project's erosion model (no reference code exists) */
    int enabled;        /* 0 (default): contact response only */
    float Kc;           /* carrying capacity per unit tangential speed   [height units / (world unit / s)] */
    float Ke, Kd;       /* pick-up and deposit rates per step, 0..1 */
    float hmin;         /* bedrock height: nothing is picked up below it */
    float max_pickup;   /* per particle per step, height units */
} sphe_erosion;

int sphe_terrain_create(sphe_terrain** out, int dimx, int dimy, int dimz);              /* Grid(int,int,int) grid.h:72-82 */
void sphe_terrain_destroy(sphe_terrain* t);
int sphe_terrain_load_heightfield(sphe_terrain* t, const unsigned char* img_512x512);   /* LoadHeightfield grid.h:98-102 */
int sphe_terrain_load_heightfield_ex(sphe_terrain* t, const unsigned char* img, int rows, int cols);
int sphe_terrain_set_heights(sphe_terrain* t, const float* h, int rows, int cols);
int sphe_terrain_get_heights(sphe_terrain* t, float* h);                                /* rows*cols floats */
int sphe_terrain_get_heights_fx(sphe_terrain* t, int* hfx);                             /* exact fixed-point heights */
int sphe_terrain_size(sphe_terrain* t, int* rows, int* cols, int dims[3]);
int sphe_terrain_height_at(sphe_terrain* t, int x, int y);                              /* GetHeightfieldAt grid.h:104-107; -1 on error */
int sphe_terrain_update_grid(sphe_terrain* t, int dimx, int dimy, int dimz);            /* UpdateGrid grid.h:138-176 */
long long sphe_terrain_surface_size(sphe_terrain* t);                                   /* GetSurfacePartsSize grid.h:827 */
long long sphe_terrain_indices_size(sphe_terrain* t);                                   /* GetIndicesSize grid.h:812 */
int sphe_terrain_get_surface(sphe_terrain* t, float* out);                              /* GetSurfaceParts grid.h:822: x y z nx ny nz per vertex */
int sphe_terrain_get_indices(sphe_terrain* t, unsigned* out);                           /* GetIndices grid.h:807 */
/* Grid::collision (grid.h:462-805), batched: arrays of n packed xyz in TERRAIN coordinates. */
int sphe_terrain_collision(sphe_terrain* t, int n, const float* pos_curr, const float* pos_next, const float* vel_next,
                           int* hit, float* contact, float* normal);
int sphe_terrain_set_transform(sphe_terrain* t, const float origin[3], float scale);
sphe_erosion* sphe_terrain_erosion_ptr(sphe_terrain* t);   /* host-resident, re-read by every step (like sphe_params_ptr) */
/* The terrain stage of sphe_step on caller-provided arrays (world coordinates; sediment fixed point):
 * contact response (the call commented out at fluid_system.h:335-340) + erosion, no box collision. */
int sphe_terrain_stage_host(sphe_terrain* t, int n, const float* pos_curr, float* pos_next, float* vel_next, int* sediment,
                            float dt, float cR, int* hit);
/* Slabs eroding ONE terrain (every rank holds a replica): run the step in phases and sum the integer
 * accumulators over the ranks in between -- phase 0 (binning .. forces, contact response, erosion requests of
 * OWNED particles; ghost copies never enter the terrain stage), sum `want`, phase 1 (grants), sum `delta`,
 * phase 2 (apply + cull map).  All sums are integers: the replicas stay bit-identical and equal to a
 * single-GPU run.  sphe_terrain_accumulators returns the two DEVICE arrays (rows*cols int32 each). */
int sphe_step_phase(sphe_sim* s, sphe_terrain* t, int phase);
int sphe_terrain_accumulators(sphe_terrain* t, void** want, void** delta, long long* cells);
int sphe_terrain_total_fx(sphe_terrain* t, long long* sum);     /* sum of all heights, fixed point */
int sphe_terrain_total_fx_rows(sphe_terrain* t, int row0, int row1, long long* sum);   /* ... of the rows [row0, row1) */
/* Slab-local terrain.  Rows map to x, so a slab's owned particles only ever touch the rows under the slab plus
 * a margin.  set_window: this replica keeps only the rows [row0, row1) current -- apply and the cull map run over
 * them alone, so the terrain cost of a rank does not grow with the number of slabs -- and the caller sums the
 * accumulators with its x-neighbours over the rows both can touch instead of all-reducing whole arrays
 * (sph-erosion_b200/slabs.py TerrainWindowShare).  row1 <= row0 restores the whole terrain.  A contact that
 * reads or writes within 2 rows of an interior window edge is counted: window_violations must stay 0. */
int sphe_terrain_set_window(sphe_terrain* t, int row0, int row1);
int sphe_terrain_window_violations(sphe_terrain* t, long long* count);
/* Moving a window (the slabs were re-cut): a replica is current only inside its window, so the caller first brings every
 * row up to date from its owner -- heights_device is the DEVICE array of fixed-point heights (rows*cols int32) the ranks
 * sum their owned rows into (slabs.py TerrainWindowShare.recut) -- then sets the new window and calls refresh, which
 * rebuilds the derived cull map over it.  Between two steps only. */
int sphe_terrain_heights_device(sphe_terrain* t, void** hfx, long long* cells);
int sphe_terrain_refresh(sphe_terrain* t);
/* survivors of the exact contact cull in the last step, per Grid::collision path class (same cell / one axis / both axes) */
int sphe_terrain_survivors(sphe_sim* s, int out[3]);
int sphe_terrain_contacts(sphe_terrain* t, long long* total, int reset); /* particle-terrain contacts since the last reset */
int sphe_sediment_total_fx(sphe_sim* s, long long* sum);        /* sum of carried sediment (owned particles), fixed point */
int sphe_set_sediment_fx(sphe_sim* s, const int* sediment_by_id);

/* ---- on-disk state (the reference has no format).  One little-endian file with everything a bit-exact resume
 * needs: parameters and bookkeeping, positions / velocities in id order, carried sediment (fixed point) and, when
 * `t` is given, the terrain's fixed-point heights, transform and erosion parameters.  A run resumed from a file
 * continues bit for bit (results do not depend on the storage order of the particles). */
int sphe_save_state(sphe_sim* s, sphe_terrain* t /* or NULL */, const char* path);
int sphe_load_state(sphe_sim* s, sphe_terrain* t /* required when the file holds a terrain */, const char* path);

/* ---- raw device access for multi-GPU plumbing (halo exchange lives above this ABI) ---- */
enum { SPHE_D_POSQ = 0, SPHE_D_VELV = 1, SPHE_D_IDS = 2, SPHE_D_RHO = 3 };
void* sphe_device_ptr(sphe_sim* s, int which);
int sphe_set_stream(sphe_sim* s, void* cuda_stream); /* run on a caller-provided cudaStream_t */
/* Kernel-variant selector for the two neighbour passes (tuning / ncu A-B runs).  3 = neighbour lists
 * (default, must be set for both passes), 1 = packed pair, 0 = thread per particle. */
int sphe_set_variant(sphe_sim* s, int density_variant, int force_variant);
/* Sizing of the neighbour lists of the default (list) kernels; all of it is automatic, these are for tests/tuning.
 * A list has `sphe_nlist_capacity` rows per particle pair in HBM (128, doubled up to 512 as soon as 0.1 % of the
 * pairs need more: a pair beyond its rows is handled by a direct walk in the force pass -- still exact, but ~10x
 * slower and it stalls its warp).  The density pass stages the first `sphe_nlist_smem_entries` entries in shared
 * memory (64: 6 CTAs/SM) and writes longer lists straight to their rows (a saturated run is walked twice);
 * when most pairs spill (60-120 neighbours) it switches to 128 / 256 staged entries, and back.
 * sphe_set_nlist_capacity(64|128|256) pins the staged entries, 0 returns to the automatic choice. */
int sphe_nlist_capacity(sphe_sim* s);
int sphe_nlist_smem_entries(sphe_sim* s);
int sphe_set_nlist_capacity(sphe_sim* s, int entries);
int sphe_nlist_overflowed(sphe_sim* s);   /* particle pairs whose list overflowed in a recent step (read back without a sync) */

#ifdef __cplusplus
}
#endif
#endif /* SPHE_H */
