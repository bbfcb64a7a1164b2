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
    int cube;                 // 1: the reference's cubic box, compare |coord| directly (fluid_system.h:362-371)
    float densK;              // mass * 315/(64 PI h^9)
    float c45;                // 45/(PI h^6)
    float c945;               // 945/(32 PI h^9)
    float hh3;                // 3*h*h
};

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
