// Shared device/host definitions for the sm_100a SPH-Erosion hot path.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define SPHE_FULL 0xffffffffu

// Uniform neighbour grid (this project's definition; the reference is all-pairs).
// cell id = (cx*ny + cy)*nz + cz -- x most significant so an x-slab is one contiguous range of the
// sorted arrays (multi-GPU slabs), z least significant so the three z-neighbours of a cell are one
// contiguous run of the sorted arrays (the 27-cell walk is 9 runs).
// Multi-GPU x-slabs bin into a WINDOW of the global grid: gnx is the global number of cell columns
// along x, xoff the first global column of the window and nx the window width (single GPU: xoff = 0,
// nx = gnx).  Window cell id = ((cx_global - xoff)*ny + cy)*nz + cz.
struct GridP {
    float gx, gy, gz, cell;
    int nx, ny, nz;
    int gnx, xoff;
};

// Ghost copies of a neighbour slab's particles carry this bit in their id; the canonical in-cell
// order ignores it, so a slab sums its neighbours in the same order as a single-GPU run.
#define SPHE_GHOST_BIT 0x40000000
#define SPHE_ID_MASK 0x3fffffff

// Per-step constants, computed on the HOST with the same libm calls as the reference
// (powf(h,9), powf(h,6): Erosion/fluid_system.h:415-452) so they are bit-identical.
struct StepC {
    float h, hh, T;           // T = largest float with sqrtf(T) <= h  (exact neighbour predicate)
    float mass, k, p0, visc, surf;
    float gx, gy, gz;
    float dt, len, cR;
    float lenx, leny, lenz;   // per-axis box half-extents (all = len unless sphe_set_box was called)
    int box;                  // 0: the box collision is applied by the terrain stage instead (terrain.cu)
    int cube;                 // 1: the reference's cubic box, compare |coord| directly (fluid_system.h:362-371)
    float densK;              // mass * 315/(64 PI h^9)
    float c45;                // 45/(PI h^6)
    float c945;               // 945/(32 PI h^9)
    float hh3;                // 3*h*h
    // Terrain attached: the force kernels' epilogue runs the exact contact cull (terrain.cu) on the state
    // it has in registers and appends the survivors to t_surv; only those go through the contact search.
    const int* t_lmax;        // NULL: no terrain
    int* t_surv;              // survivor slots: SPHE_SURV_CLASSES lists of t_cap entries each, one per contact-path class
    int* t_count;             // survivors per class
    int t_cap;
    int t_rows, t_cols, t_dimx, t_dimz;
    float t_ox, t_oy, t_oz, t_inv;
};

// Exact contact culling (see k_terrain_contact): nothing above the local maxima of the cells of posCurr and
// posNext can touch the heightfield.  Same arithmetic in sph.cu (FMA contraction allowed) and terrain.cu
// (not allowed): (p - o) * inv has nothing to contract.
// Returns -1 (cannot touch) or the CLASS of the path Grid::collision will take (grid.h:476, :623, :640): 0 = posCurr and
// posNext in the same terrain cell, 1 = cells differ along one axis, 2 = along both (straight to the corner fans).
// Survivors are listed per class so a warp of the contact kernel runs ONE of the three branches instead of all of them.
#define SPHE_SURV_CLASSES 3
__device__ __forceinline__ int terrain_may_touch(const StepC& C, float ox_, float oy_, float oz_, float px, float py, float pz) {
    float cx = floorf((ox_ - C.t_ox) * C.t_inv), cz = floorf((oz_ - C.t_oz) * C.t_inv);
    float nx = floorf((px - C.t_ox) * C.t_inv), nz = floorf((pz - C.t_oz) * C.t_inv);
    float mx = (float)(C.t_dimx - 1), mz = (float)(C.t_dimz - 1);
    if (cx < 0.0f || cx >= mx || cz < 0.0f || cz >= mz || nx < 0.0f || nx >= mx || nz < 0.0f || nz >= mz) return -1;
    int ia = min((int)cx, C.t_rows - 1) * C.t_cols + min((int)cz, C.t_cols - 1);
    int ib = min((int)nx, C.t_rows - 1) * C.t_cols + min((int)nz, C.t_cols - 1);
    int top = max(__ldg(&C.t_lmax[ia]), __ldg(&C.t_lmax[ib]));
    if (!((py - C.t_oy) * C.t_inv <= (float)top * (1.0f / 4096.0f) + 0.01f)) return -1;
    return (cx != nx ? 1 : 0) + (cz != nz ? 1 : 0);
}

__device__ __forceinline__ int cell_axis(float p, float gmin, float cell, int dim) {
    // bit-identical to oracle/sph_oracle.c cell_axis(): IEEE sub, div, floor; clamp; NaN -> 0
    float v = floorf(__fdiv_rn(__fsub_rn(p, gmin), cell));
    if (!(v >= 0.0f)) return 0;
    if (v >= (float)dim) return dim - 1;
    return (int)v;
}

__device__ __forceinline__ void cell_coords(const GridP& G, float x, float y, float z, int& cx, int& cy, int& cz) {
    cx = cell_axis(x, G.gx, G.cell, G.gnx) - G.xoff;
    cx = cx < 0 ? 0 : (cx >= G.nx ? G.nx - 1 : cx);
    cy = cell_axis(y, G.gy, G.cell, G.ny);
    cz = cell_axis(z, G.gz, G.cell, G.nz);
}

// Exact restatement of glm::length(xi - xj)^2 before the sqrt: ((dx*dx + dy*dy) + dz*dz) with
// every operation rounded separately (no FMA contraction), vendor/glm func_geometric.inl:47-55.
__device__ __forceinline__ float dist2_exact(float dx, float dy, float dz) {
    return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

// Box collision + response, fluid_system.h:342-347 + collisionS :355-407.  Shared by the force kernels
// (sph.cu) and the terrain stage (terrain.cu, built without FMA contraction): written with explicitly
// rounded operations so both translation units produce the same bits.
// The reference box is a cube of half-extent len: the violated axis is the one with the largest
// |coordinate| (ties x -> y -> z by strict <, :362-371).  sphe_set_box generalises it to per-axis
// half-extents for the multi-GPU channel scenes; then the axis with the largest overshoot wins.
__device__ __forceinline__ bool box_collide(const StepC& C, float& px, float& py, float& pz, float& vx, float& vy, float& vz) {
    float ax = fabsf(px), ay = fabsf(py), az = fabsf(pz);
    if (ax < C.lenx && ay < C.leny && az < C.lenz) return false;
    if (C.dt == 0.0f) return false;
    if (!C.cube) { ax = __fsub_rn(ax, C.lenx); ay = __fsub_rn(ay, C.leny); az = __fsub_rn(az, C.lenz); }
    int axis = 0;
    float m = ax;
    if (m < ay) { axis = 1; m = ay; }
    if (m < az) { axis = 2; m = az; }
    float cx = px, cy = py, cz = pz, nx = 0.0f, ny = 0.0f, nz = 0.0f;
    if (axis == 0) { if (px < -C.lenx) { cx = -C.lenx; nx = 1.0f; } else { cx = C.lenx; nx = -1.0f; } }
    else if (axis == 1) { if (py < -C.leny) { cy = -C.leny; ny = 1.0f; } else { cy = C.leny; ny = -1.0f; } }
    else { if (pz < -C.lenz) { cz = -C.lenz; nz = 1.0f; } else { cz = C.lenz; nz = -1.0f; } }
    float ex = __fsub_rn(px, cx), ey = __fsub_rn(py, cy), ez = __fsub_rn(pz, cz);
    float d = __fsqrt_rn(dist2_exact(ex, ey, ez));
    float vlen = __fsqrt_rn(dist2_exact(vx, vy, vz));
    float sc = __fadd_rn(1.0f, __fdiv_rn(__fmul_rn(0.5f, d), __fmul_rn(C.dt, vlen)));
    float vn = __fadd_rn(__fadd_rn(__fmul_rn(vx, nx), __fmul_rn(vy, ny)), __fmul_rn(vz, nz));
    vx = __fsub_rn(vx, __fmul_rn(__fmul_rn(nx, sc), vn));
    vy = __fsub_rn(vy, __fmul_rn(__fmul_rn(ny, sc), vn));
    vz = __fsub_rn(vz, __fmul_rn(__fmul_rn(nz, sc), vn));
    px = cx; py = cy; pz = cz;
    return true;
}
