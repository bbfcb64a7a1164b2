/* TEST INFRASTRUCTURE ONLY -- CPU oracle for the SPH-Erosion hot path.
 *
 * Plain-C restatement of the reference algorithm (Erosion/fluid_system.h,
 * Erosion/grid.h).  Each function cites the reference lines it follows.
 * PARITY STATUS
 *   - SPH step (density/pressure, forces, integration, box collision): PINNED.
 *     Bit-identical to the compiled reference (oracle/_ref) -- tests/test_oracle_vs_ref.py,
 *     and to the committed golden vectors in tests/golden/.
 *   - Neighbour grid (cell index, sorted order, cell-start, neighbour lists): the reference
 *     has none; this file DEFINES them.  Pinned indirectly: the neighbour sets
 *     equal the reference predicate's and the grid step is bit-identical to the all-pairs step.
 *   - Grid::collision: PINNED against oracle/_ref on lena_gray + synthetic fields.
 *   - Erosion / sediment transport: PARITY UNPINNED (no reference code exists);
 *     the model is specified in DESIGN.md and self-tested by invariants.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may use this.
 */
#ifndef SPH_ORACLE_H
#define SPH_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
    float mass, visc, surf_tens, p0, k, h, len, dt;
    float g[3];
} so_params;

/* uniform neighbour grid definition (this project's own spec) */
typedef struct {
    float gmin[3];
    float cell;      /* cell edge, >= h*(1+2^-10) */
    int dim[3];
} so_grid;

/* per-particle state + diagnostics, SoA, all in particle-id order */
typedef struct {
    int n;
    float *pos, *vel, *acc;            /* 3n */
    float *density, *pressure;         /* n  */
    float *fpress, *fvisc, *fgrav, *fsurf, *normal; /* 3n */
    int *neighb;                       /* n, last neighbour id (fluid_system.h:144) */
} so_state;

void so_default_params(so_params* p);

/* Initialize/AddParticles lattice (fluid_system.h:74-102, 232-251). Returns count written
 * (ceil(cbrt(n))^3); pos may be NULL to query the count. */
int so_lattice(int n, const float origin[3], float* pos);

/* Reference step, all pairs, j = 0..n-1 order (fluid_system.h:104-183, 306-407). */
void so_step_allpairs(const so_params* P, so_state* S);

/* grid helpers */
void so_grid_for_box(const so_params* P, float lo[3], float hi[3], so_grid* G);
int so_cell_of(const so_grid* G, const float p[3]);                 /* linear cell id */
/* cell_of[n] (by id), order[n] (ids sorted by (cell,id)), cell_start[ncells+1] */
void so_bin(const so_grid* G, int n, const float* pos, int* cell_of, int* order, int* cell_start);
/* neighbour lists in grid-walk order, CSR: nbr_start[n+1] (by sorted slot), nbr[] holds ids.
 * Returns total count; pass nbr=NULL to only count. Includes self. */
long so_neighbours(const so_params* P, const so_grid* G, int n, const float* pos,
                   const int* order, const int* cell_start, long* nbr_start, int* nbr);

/* Same step as so_step_allpairs but using the grid; neighbours are accumulated in ascending id
 * order so every sum is bit-identical to the all-pairs loops.  OpenMP over particles. */
void so_step_grid(const so_params* P, const so_grid* G, so_state* S);

/* so_step_grid in pieces: passes 1-3 only; integration without the box; the box collision alone.
 * so_step_grid == so_forces_grid + so_integrate + so_box + store. */
void so_forces_grid(const so_params* P, const so_grid* G, so_state* S);
void so_integrate(const so_params* P, so_state* S, float* pos_next, float* vel_next);
void so_box(const so_params* P, int n, float* pos_next, float* vel_next);

int so_omp_threads(void);

#ifdef __cplusplus
}
#endif
#endif
