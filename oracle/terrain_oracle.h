/* TEST INFRASTRUCTURE ONLY -- see terrain_oracle.c for scope and parity status. */
#ifndef TERRAIN_ORACLE_H
#define TERRAIN_ORACLE_H
#ifdef __cplusplus
extern "C" {
#endif

/* Grid (Erosion/grid.h:26-51): heightfield storage rows x cols (the reference hard-codes 512 x 512,
 * grid.h:78-81), H(x, z) = h[cols * x + z]; dimx/dimy/dimz are the Grid dimensions used for range
 * checks and the render mesh (main.cpp:100-105 uses 50, 255, 50). */
typedef struct {
    int rows, cols;
    int dimx, dimy, dimz;
    float* h;   /* heights, float (the reference stores unsigned char) */
    int* hfx;   /* heights in fixed point, h * 4096 -- authoritative for the erosion model */
} so_terrain;

/* this project's erosion model parameters (DESIGN.md "erosion model") */
typedef struct {
    int enabled;
    float origin[3];
    float scale;          /* world units per terrain cell */
    float Kc, Ke, Kd;     /* capacity per unit tangential speed, pick-up rate, deposit rate */
    int hmin_fx;          /* bedrock level, fixed point */
    int max_pickup_fx;    /* per particle per step, fixed point */
} so_erosion;

int so_terrain_collision(const so_terrain* T, const float posCurr[3], const float posNext[3], const float velNext[3],
                         float contact[3], float normal[3]);
void so_terrain_surface(const so_terrain* T, float* out6);
long so_terrain_indices(const so_terrain* T, unsigned* out);
void so_terrain_stage(so_terrain* T, const so_erosion* E, int n, const float* pos_curr, float* pos_next, float* vel_next,
                      int* sediment, float dt, float cR, int* hit_out);

#ifdef __cplusplus
}
#endif
#endif
