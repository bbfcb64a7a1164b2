// Internal declarations shared by the .cu translation units (not part of the C ABI).
// Launchers of the per-step kernels take the particle count twice: `n` sizes the grid (an upper bound is
// enough) and `n_dev`, when not NULL, is a device word holding the exact count -- in slab mode the host
// launches a step before it knows how many halo records arrived.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "common.cuh"

namespace sphe {

// ---- binning.cu
// ids != NULL (slab mode): entries with id == -1 are dead (dropped by the exchange): no cell, not counted, not scattered
void launch_hash(cudaStream_t st, int n, const int* n_dev, const float4* posq, const int* ids, const GridP& G, uint32_t* cell, int* count);
int scan_tiles_for(long long ncells);
void launch_scan_onepass(cudaStream_t st, long long ncells, int* count, int* cell_start, int* cursor, unsigned long long* state,
                         unsigned* ticket, unsigned ticket_base, unsigned epoch);
// cell_start[ncells] = number of LIVE particles (the sum of the counts): the count every later kernel of the step works on
void launch_scan(cudaStream_t st, long long ncells, int* count, int* tile_sum, int* cell_start, int* cursor);
// also writes the FINAL sorted cell ids (every slot of a cell's range holds that cell's id)
void launch_scatter(cudaStream_t st, int n, const int* n_dev, const uint32_t* cell, const int* ids, int* cursor, uint2* tmp, uint32_t* cell_sorted);
void launch_rank_reorder(cudaStream_t st, int n, const int* n_dev, const uint2* tmp, const uint32_t* cell_sorted, const int* cell_start,
                         const float4* posq_in, const float4* velv_in, const float* sed_in,
                         float4* posq_out, float4* velv_out, float* sed_out, int* ids_out, int* src_of_slot = nullptr);
void launch_gather_vel(cudaStream_t st, int n, const int* src_of_slot, const float4* velv_in, float4* velv_out);

// ---- terrain.cu
// Device view of a terrain (replacement of the reference's Grid, Erosion/grid.h:26-51).
struct TerrainDev {
    int rows, cols;          // heightfield storage, H(x, z) = hfx[cols * x + z] / 4096
    int dimx, dimy, dimz;    // Grid dimensions (range checks, render mesh)
    const int* hfx;          // heights, fixed point 1/4096
    int* hfx_rw;
    int* want;               // per-vertex sum of pick-up requests of the current step
    int* delta;              // per-vertex deposits - grants of the current step
    const int* hmax_fx;      // upper bound of all heights (contact culling)
    int* hmax_rw;
    const int* lmax;         // per cell: max height of the 4x4 vertex neighbourhood (exact contact culling)
    int* lmax_rw;
    unsigned long long* contacts;  // cumulative number of particle-terrain contacts
    // slab-local maintenance (sphe_terrain_set_window): only the rows [win0, win1) are kept current on this rank
    // (apply + cull map run over them only); a contact that reaches outside them is counted in *violations
    int win0, win1;
    unsigned long long* violations;
    float ox, oy, oz, scale, inv_scale;   // world = origin + scale * terrain coordinates
    float Kc, Ke, Kd;
    int hmin_fx, max_pickup_fx;
    int erosion;
};
void launch_terrain_stage(cudaStream_t st, const int* surv, const int* surv_count, int surv_cap, const float4* pos_old, float4* posq, float4* velv,
                          int* sediment, const StepC& C, const TerrainDev& T, int apply_box, int* req_vertex, int* req_amount,
                          int* hit_out, int phases = 7);
enum { TERRAIN_CONTACT = 1, TERRAIN_GRANT = 2, TERRAIN_APPLY = 4 };  // phases of the terrain stage (multi-GPU: reduce between them)
void launch_iota(cudaStream_t st, int n, int* a, int* count);
int terrain_stage_launches(const StepC& C, const TerrainDev& T);
void launch_terrain_lmax(cudaStream_t st, const TerrainDev& T);
void launch_terrain_surface(cudaStream_t st, const TerrainDev& T, float* out);
void launch_terrain_indices(cudaStream_t st, int dimx, int dimz, unsigned* out);
void launch_terrain_collide(cudaStream_t st, int n, const float* pc, const float* pn, const float* vn, const TerrainDev& T,
                            int* hit, float* cp, float* nrm);
void launch_heights_from_u8(cudaStream_t st, int cells, const unsigned char* img, int* hfx, int* hmax);
void launch_heights_from_f32(cudaStream_t st, int cells, const float* src, int* hfx, int* hmax);
void launch_heights_to_f32(cudaStream_t st, int cells, const int* hfx, float* dst);
void launch_sum_i32(cudaStream_t st, int n, const int* a, const int* ghost_ids, long long* out, const int* n_dev = nullptr);

// ---- slab.cu (multi-GPU x-slabs)
struct SlabP {
    int x0, x1;      // owned global cell columns [x0, x1)
    int halo;        // cell layers mirrored from each neighbour
    int has_left, has_right;
    // ring closure (sphe_slab_ring): the left link of the FIRST slab / the right link of the LAST slab goes to
    // the slab at the other end of the channel.  A wrap link carries no halo, only the particles the reference's
    // box quirk moved from the -x wall to the +x wall (cell column >= far_x0, the last slab's x0).
    int wrap_left = 0, wrap_right = 0, far_x0 = 0x7fffffff;
};
void launch_slab_classify(cudaStream_t st, int n, const int* n_dev, const float4* posq, const float4* velv, int* ids, const float* sed,
                          const GridP& G, const SlabP& S, float4* send_left, float4* send_right, int cap_records, int* counters, bool remote = false);
// flag_* != NULL: peer-memory exchange, the buffers and flags live in the neighbour GPUs' mailboxes
void launch_slab_headers(cudaStream_t st, int* counters, float4* send_left, float4* send_right, int* flag_left = nullptr,
                         int* flag_right = nullptr, int seq = 0);
// records in transit: received from one side, owned further along (k_slab_append -> k_slab_forward at the next pack)
#define SPHE_TRANSIT_CAP 8192
void launch_zone_sum(cudaStream_t st, int* arr, int off_l, int off_r, int n, int* out_l, int* out_r, int* flag_out_l, int* flag_out_r,
                     const int* in_l, const int* in_r, const int* flag_in_l, const int* flag_in_r, int seq, long long timeout, int* done, int* err);
void launch_slab_forward(cudaStream_t st, const float4* transit_l, const float4* transit_r, int* transit_n, float4* send_left,
                         float4* send_right, int cap_records, int* counters, bool remote);
void launch_slab_append(cudaStream_t st, int max_l, int max_r, const float4* rec_l, const float4* rec_r, const GridP& G,
                        const SlabP& S, int cap_particles, float4* posq, float4* velv, int* ids, float* sed, int* counters, int* n_out,
                        float4* transit_l, float4* transit_r, int* transit_n,
                        const int* flag_l = nullptr, const int* flag_r = nullptr, int seq = 0, long long timeout_cycles = 0);
void launch_slab_gather_owned(cudaStream_t st, int n, const int* n_dev, const float4* posq, const float4* velv, const float* rho, const float* sed,
                              const int* ids, int* counter, int* out_ids, float* out_pos, float* out_vel, float* out_rho,
                              float* out_sed);
int launch_slab_column_hist(cudaStream_t st, int n, const int* n_dev, const float4* posq, const int* ids, const GridP& G, int* hist);
void launch_pack_state_ids(cudaStream_t st, int n, const float* pos, const float* vel, const int* ids_in, float4* posq, float4* velv,
                           int* ids, float* sed);

// ---- sph.cu
// Optional per-particle debug fields of the reference's FluidParticle (fluid_system.h:49-64),
// indexed by PARTICLE ID (not by sorted slot).
struct DiagOut {
    float4* acc;
    float4* fpress;
    float4* fvisc;
    float4* fgrav;
    float4* fsurf;
    float4* normal;
    int* neighb;
};


void launch_density(cudaStream_t st, int variant, int n, const int* n_dev, const float4* posq, float4* posq_q, float4* velv,
                    const uint32_t* cell_sorted, const int* cell_start, const GridP& G, const StepC& C, float* rho,
                    int* nlist, int2* ncount, int cap = 64, int* overflow = nullptr, int smem_cap = 64);
int nlist_cap();
bool variant_supported(int density, int force);
int nlist_pairs_pad(int n);
void launch_force(cudaStream_t st, int variant, int n, const int* n_dev, const float4* posq, const float4* velv, const float* rho,
                  const int* ids, const uint32_t* cell_sorted, const int* cell_start, const GridP& G, const StepC& C,
                  float4* posq_out, float4* velv_out, const DiagOut* diag, const int* nlist, const int2* ncount);
// ---- stage.cu: cell-cooperative passes with shared-memory-staged candidates and bit-mask neighbour lists (variant 20)
int stage_pairs_pad(int n);
size_t stage_mask_words(int n);
void launch_density_stage(cudaStream_t st, int n, const int* n_dev, const float4* posq, float4* posq_q, float4* velv,
                          const uint32_t* cell_sorted, const int* cell_start, const GridP& G, const StepC& C, float* rho, unsigned* nmask);
void launch_force_stage(cudaStream_t st, int n, const int* n_dev, const float4* posq_q, const float4* velv, const float* rho,
                        const int* ids, const uint32_t* cell_sorted, const int* cell_start, const GridP& G, const StepC& C,
                        const unsigned* nmask, float4* posq_out, float4* velv_out, const DiagOut* diag);
void launch_list_decode(cudaStream_t st, int n, const int* nlist, const int2* ncount, int cap, int* counts, int* out);
void launch_neighb_id(cudaStream_t st, int n, const int* n_dev, const float4* posq, const int* ids, const uint32_t* cell_sorted,
                      const int* cell_start, const GridP& G, const StepC& C, int* neighb_by_id);
void launch_stage_decode(cudaStream_t st, int n, const float4* posq, const uint32_t* cell_sorted, const int* cell_start,
                         const GridP& G, const unsigned* nmask, int cap, int* counts, int* out);
void launch_neighbour_count(cudaStream_t st, int n, const float4* posq, const uint32_t* cell_sorted, const int* cell_start,
                            const GridP& G, const StepC& C, int* counts);
void launch_neighbour_fill(cudaStream_t st, int n, const float4* posq, const int* ids, const uint32_t* cell_sorted,
                           const int* cell_start, const GridP& G, const StepC& C, const long long* nbr_start, int* nbr);
void launch_unsort_f4(cudaStream_t st, int n, const float4* src, const int* ids, float* dst_xyz);
void launch_unsort_f1(cudaStream_t st, int n, const float* src, const int* ids, float* dst);
void launch_unsort_u32(cudaStream_t st, int n, const uint32_t* src, const int* ids, int* dst);
void launch_pack_state(cudaStream_t st, int n, const float* pos_xyz, const float* vel_xyz, float4* posq, float4* velv,
                       int* ids, float* sed);
void launch_slot_of_id(cudaStream_t st, int n, const int* ids, int* slot_of_id);

}  // namespace sphe
