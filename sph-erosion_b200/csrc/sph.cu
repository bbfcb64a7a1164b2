// SPH neighbour passes over the sorted SoA arrays.
//
//   k_density : pass 1 of FluidSystemSPH::Run (Erosion/fluid_system.h:108-124, kernDefault :410-416)
//   k_force   : passes 2+3 fused (:128-178; gradPressure :430-447, laplVisc :449-453,
//               gradDefault :418-422, laplDefault :424-428) + advance() (:306-353) + collisionS (:355-407)
//
// Pass 3 only needs particle i's own normal from pass 2, so 2 and 3 share one neighbour walk
// (SURVEY.md section 3.1).  advance() overwrites what neighbours still read, so the new state goes to
// a second buffer.
//
// Arithmetic: the neighbour PREDICATE is bit-exact (dist2_exact + threshold T, see common.cuh).  The
// sums use FMA and factor the kernel constants out of the loop; they match the reference within the
// fp32 tolerance stated in tests/test_gpu_parity.py (every SPH kernel vanishes at r = h, so the sums
// are continuous in the predicate).
#include "common.cuh"
#include "sim.h"

namespace sphe {

// Visit the candidates of a particle in cell c in canonical grid-walk order: dx = -1..1, dy = -1..1,
// then the contiguous sorted range covering cells cz-1..cz+1 of column (cx+dx, cy+dy).
// Identical to WALK_BEGIN/WALK_END in oracle/sph_oracle.c.
template <class F>
__device__ __forceinline__ void walk27(const GridP& G, const int* __restrict__ cell_start, uint32_t c, F&& f) {
    int cz = (int)(c % (uint32_t)G.nz);
    uint32_t t = c / (uint32_t)G.nz;
    int cy = (int)(t % (uint32_t)G.ny);
    int cx = (int)(t / (uint32_t)G.ny);
    int z0 = cz > 0 ? cz - 1 : 0;
    int z1 = cz < G.nz - 1 ? cz + 1 : cz;
#pragma unroll 1
    for (int dx = -1; dx <= 1; dx++) {
        int x = cx + dx;
        if (x < 0 || x >= G.nx) continue;
#pragma unroll 1
        for (int dy = -1; dy <= 1; dy++) {
            int y = cy + dy;
            if (y < 0 || y >= G.ny) continue;
            int base = (x * G.ny + y) * G.nz;
            int s = __ldg(&cell_start[base + z0]);
            int e = __ldg(&cell_start[base + z1 + 1]);
            for (int k = s; k < e; k++) f(k);
        }
    }
}

// ------------------------------------------------------------------ pass 1: density + pressure
// Writes rho, and packs what pass 2 needs per neighbour into the arrays it will stream anyway:
//   posq_q[i] = (x, y, z, P_i / rho_i^2)      velv[i].w = mass / rho_i
__global__ void __launch_bounds__(128) k_density_tpp(int n, const float4* __restrict__ posq, float4* __restrict__ posq_q,
                                                     float4* __restrict__ velv, const uint32_t* __restrict__ cell_sorted,
                                                     const int* __restrict__ cell_start, GridP G, StepC C,
                                                     float* __restrict__ rho) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 pi = posq[i];
    float sum = 0.0f;
    walk27(G, cell_start, cell_sorted[i], [&](int k) {
        float4 pj = __ldg(&posq[k]);
        float d2 = dist2_exact(pi.x - pj.x, pi.y - pj.y, pi.z - pj.z);
        if (d2 <= C.T) {
            float w = C.hh - d2;
            sum = fmaf(w * w, w, sum);
        }
    });
    float r = C.densK * sum;
    float P = C.k * (r - C.p0);
    rho[i] = r;
    posq_q[i] = make_float4(pi.x, pi.y, pi.z, P / (r * r));
    velv[i].w = C.mass / r;
}

// ------------------------------------------------------------------ box collision, fluid_system.h:355-407
__device__ __forceinline__ bool collision_box(float len, float x, float y, float z, float& cx, float& cy, float& cz,
                                              float& nx, float& ny, float& nz) {
    float ax = fabsf(x), ay = fabsf(y), az = fabsf(z);
    if (ax < len && ay < len && az < len) return false;
    int axis = 0;
    float m = ax;
    if (m < ay) { axis = 1; m = ay; }
    if (m < az) { axis = 2; m = az; }
    cx = x; cy = y; cz = z;
    nx = ny = nz = 0.0f;
    if (axis == 0) { if (x < -len) { cx = -len; nx = 1.0f; } else { cx = len; nx = -1.0f; } }
    else if (axis == 1) { if (y < -len) { cy = -len; ny = 1.0f; } else { cy = len; ny = -1.0f; } }
    else { if (z < -len) { cz = -len; nz = 1.0f; } else { cz = len; nz = -1.0f; } }
    return true;
}

// ------------------------------------------------------------------ passes 2+3 + integrate + collide
template <bool DIAG>
__global__ void __launch_bounds__(128) k_force_tpp(int n, const float4* __restrict__ posq_q, const float4* __restrict__ velv,
                                                   const float* __restrict__ rho, const int* __restrict__ ids,
                                                   const uint32_t* __restrict__ cell_sorted,
                                                   const int* __restrict__ cell_start, GridP G, StepC C,
                                                   float4* __restrict__ posq_out, float4* __restrict__ velv_out, DiagOut D) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 pi = posq_q[i];
    const float4 vi = velv[i];
    const float rho_i = rho[i];
    const float inv_sqrt3 = 0.57735026f;  // glm::normalize(vec3(1)) = 1 * (1/sqrt(3)), fluid_system.h:439

    float ax = 0.f, ay = 0.f, az = 0.f;     // sum (q_i+q_j) (h-r)^2 dir
    float fx = 0.f, fy = 0.f, fz = 0.f;     // sum (v_j-v_i) vol_j (h-r)
    float nx = 0.f, ny = 0.f, nz = 0.f;     // sum vol_j (h^2-r^2)^2 d
    float cf = 0.f;                         // sum vol_j (h^2-r^2)(3h^2-7r^2)   (self included)
    int maxid = -1;

    walk27(G, cell_start, cell_sorted[i], [&](int k) {
        float4 pj = __ldg(&posq_q[k]);
        float dx = pi.x - pj.x, dy = pi.y - pj.y, dz = pi.z - pj.z;
        float d2 = dist2_exact(dx, dy, dz);
        if (d2 <= C.T) {
            float4 vj = __ldg(&velv[k]);
            float w = C.hh - d2;
            float vw = vj.w * w;
            cf = fmaf(vw, fmaf(-7.0f, d2, C.hh3), cf);
            float vww = vw * w;
            nx = fmaf(vww, dx, nx); ny = fmaf(vww, dy, ny); nz = fmaf(vww, dz, nz);
            float rinv = rsqrtf(fmaxf(d2, 1e-30f));
            float r = d2 * rinv;
            float hm = C.h - r;
            float tv = vj.w * hm;
            fx = fmaf(tv, vj.x - vi.x, fx); fy = fmaf(tv, vj.y - vi.y, fy); fz = fmaf(tv, vj.z - vi.z, fz);
            // pressure: j != i only (fluid_system.h:142); coincident pairs use direction (1,1,1)/sqrt(3)
            float s = (k != i) ? (pi.w + pj.w) * hm * hm : 0.0f;
            bool tiny = r <= 1e-4f;  // == (double)r < 10e-5, fluid_system.h:438
            float ux = tiny ? inv_sqrt3 : dx * rinv;
            float uy = tiny ? inv_sqrt3 : dy * rinv;
            float uz = tiny ? inv_sqrt3 : dz * rinv;
            ax = fmaf(s, ux, ax); ay = fmaf(s, uy, ay); az = fmaf(s, uz, az);
            if (DIAG) { if (k != i) maxid = max(maxid, __ldg(&ids[k])); }
        }
    });

    // PressureForce = -(fPress * rho_i), fPress = -mass*c45 * A        (fluid_system.h:145,151)
    float kp = rho_i * C.mass * C.c45;
    float Fpx = kp * ax, Fpy = kp * ay, Fpz = kp * az;
    // ViscosityForce = fVisc * visc, fVisc = c45 * F                    (:146,153)
    float kv = C.visc * C.c45;
    float Fvx = kv * fx, Fvy = kv * fy, Fvz = kv * fz;
    // SurfaceNormal = -c945 * N                                         (:147,154)
    float Nx = -C.c945 * nx, Ny = -C.c945 * ny, Nz = -C.c945 * nz;
    // colorFieldLapl = -c945 * cf ; SurfaceForce = -surf_tens * cfl * n (:171,177)
    float cfl = -C.c945 * cf;
    float ks = -C.surf * cfl;
    float Fsx = ks * Nx, Fsy = ks * Ny, Fsz = ks * Nz;
    // GravityForce = rho_i * g                                          (:163)
    float Fgx = rho_i * C.gx, Fgy = rho_i * C.gy, Fgz = rho_i * C.gz;

    // advance(), fluid_system.h:318-350
    float Fx = (Fpx + Fvx) + (Fgx + Fsx), Fy = (Fpy + Fvy) + (Fgy + Fsy), Fz = (Fpz + Fvz) + (Fgz + Fsz);
    float acx = Fx / rho_i, acy = Fy / rho_i, acz = Fz / rho_i;
    float dt = C.dt;
    float vx = fmaf(acx, dt, vi.x), vy = fmaf(acy, dt, vi.y), vz = fmaf(acz, dt, vi.z);
    float px = fmaf(vx, dt, pi.x), py = fmaf(vy, dt, pi.y), pz = fmaf(vz, dt, pi.z);

    float cx, cy, cz, bx, by, bz;
    if (collision_box(C.len, px, py, pz, cx, cy, cz, bx, by, bz) && dt != 0.0f) {
        float ex = px - cx, ey = py - cy, ez = pz - cz;
        float d = sqrtf(ex * ex + ey * ey + ez * ez);
        float vlen = sqrtf(vx * vx + vy * vy + vz * vz);
        float sc = 1.0f + 0.5f * d / (dt * vlen);
        float vn = vx * bx + vy * by + vz * bz;
        vx -= bx * sc * vn; vy -= by * sc * vn; vz -= bz * sc * vn;
        px = cx; py = cy; pz = cz;
    }
    posq_out[i] = make_float4(px, py, pz, 0.0f);
    velv_out[i] = make_float4(vx, vy, vz, 0.0f);

    if (DIAG) {
        int id = ids[i];
        D.acc[id] = make_float4(acx, acy, acz, 0.f);
        D.fpress[id] = make_float4(Fpx, Fpy, Fpz, 0.f);
        D.fvisc[id] = make_float4(Fvx, Fvy, Fvz, 0.f);
        D.fgrav[id] = make_float4(Fgx, Fgy, Fgz, 0.f);
        D.fsurf[id] = make_float4(Fsx, Fsy, Fsz, 0.f);
        D.normal[id] = make_float4(Nx, Ny, Nz, 0.f);
        if (maxid >= 0) D.neighb[id] = maxid;  // last neighbour in ascending-id order (:144)
    }
}

// ------------------------------------------------------------------ neighbour-list test hooks
__global__ void __launch_bounds__(128) k_nbr_count(int n, const float4* __restrict__ posq, const uint32_t* __restrict__ cell_sorted,
                                                   const int* __restrict__ cell_start, GridP G, StepC C, int* __restrict__ counts) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 pi = posq[i];
    int c = 0;
    walk27(G, cell_start, cell_sorted[i], [&](int k) {
        float4 pj = __ldg(&posq[k]);
        if (dist2_exact(pi.x - pj.x, pi.y - pj.y, pi.z - pj.z) <= C.T) c++;
    });
    counts[i] = c;
}

__global__ void __launch_bounds__(128) k_nbr_fill(int n, const float4* __restrict__ posq, const int* __restrict__ ids,
                                                  const uint32_t* __restrict__ cell_sorted, const int* __restrict__ cell_start,
                                                  GridP G, StepC C, const long long* __restrict__ nbr_start, int* __restrict__ nbr) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 pi = posq[i];
    long long o = nbr_start[i];
    walk27(G, cell_start, cell_sorted[i], [&](int k) {
        float4 pj = __ldg(&posq[k]);
        if (dist2_exact(pi.x - pj.x, pi.y - pj.y, pi.z - pj.z) <= C.T) nbr[o++] = ids[k];
    });
}

// ------------------------------------------------------------------ id-order gathers / packing
__global__ void k_unsort_f4(int n, const float4* __restrict__ src, const int* __restrict__ ids, float* __restrict__ dst) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 v = src[i];
    size_t o = 3 * (size_t)ids[i];
    dst[o] = v.x; dst[o + 1] = v.y; dst[o + 2] = v.z;
}
__global__ void k_unsort_f1(int n, const float* __restrict__ src, const int* __restrict__ ids, float* __restrict__ dst) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[ids[i]] = src[i];
}
__global__ void k_unsort_u32(int n, const uint32_t* __restrict__ src, const int* __restrict__ ids, int* __restrict__ dst) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[ids[i]] = (int)src[i];
}
__global__ void k_pack_state(int n, const float* __restrict__ pos, const float* __restrict__ vel, float4* __restrict__ posq,
                             float4* __restrict__ velv, int* __restrict__ ids, float* __restrict__ sed) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    posq[i] = make_float4(pos[3 * (size_t)i], pos[3 * (size_t)i + 1], pos[3 * (size_t)i + 2], 0.f);
    velv[i] = make_float4(vel[3 * (size_t)i], vel[3 * (size_t)i + 1], vel[3 * (size_t)i + 2], 0.f);
    ids[i] = i;
    if (sed) sed[i] = 0.f;
}
__global__ void k_slot_of_id(int n, const int* __restrict__ ids, int* __restrict__ slot) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) slot[ids[i]] = i;
}

// ------------------------------------------------------------------ launch wrappers
static inline int nblk(int n, int b) { return (n + b - 1) / b; }

void launch_density(cudaStream_t st, int variant, int n, const float4* posq, float4* posq_q, float4* velv,
                    const uint32_t* cell_sorted, const int* cell_start, const GridP& G, const StepC& C, float* rho) {
    if (n <= 0) return;
    (void)variant;
    k_density_tpp<<<nblk(n, 128), 128, 0, st>>>(n, posq, posq_q, velv, cell_sorted, cell_start, G, C, rho);
}

void launch_force(cudaStream_t st, int variant, int n, const float4* posq_q, const float4* velv, const float* rho,
                  const int* ids, const uint32_t* cell_sorted, const int* cell_start, const GridP& G, const StepC& C,
                  float4* posq_out, float4* velv_out, const DiagOut* diag) {
    if (n <= 0) return;
    (void)variant;
    if (diag) k_force_tpp<true><<<nblk(n, 128), 128, 0, st>>>(n, posq_q, velv, rho, ids, cell_sorted, cell_start, G, C, posq_out, velv_out, *diag);
    else k_force_tpp<false><<<nblk(n, 128), 128, 0, st>>>(n, posq_q, velv, rho, ids, cell_sorted, cell_start, G, C, posq_out, velv_out, DiagOut{});
}

void launch_neighbour_count(cudaStream_t st, int n, const float4* posq, const uint32_t* cell_sorted, const int* cell_start,
                            const GridP& G, const StepC& C, int* counts) {
    if (n > 0) k_nbr_count<<<nblk(n, 128), 128, 0, st>>>(n, posq, cell_sorted, cell_start, G, C, counts);
}
void launch_neighbour_fill(cudaStream_t st, int n, const float4* posq, const int* ids, const uint32_t* cell_sorted,
                           const int* cell_start, const GridP& G, const StepC& C, const long long* nbr_start, int* nbr) {
    if (n > 0) k_nbr_fill<<<nblk(n, 128), 128, 0, st>>>(n, posq, ids, cell_sorted, cell_start, G, C, nbr_start, nbr);
}
void launch_unsort_f4(cudaStream_t st, int n, const float4* src, const int* ids, float* dst_xyz) {
    if (n > 0) k_unsort_f4<<<nblk(n, 256), 256, 0, st>>>(n, src, ids, dst_xyz);
}
void launch_unsort_f1(cudaStream_t st, int n, const float* src, const int* ids, float* dst) {
    if (n > 0) k_unsort_f1<<<nblk(n, 256), 256, 0, st>>>(n, src, ids, dst);
}
void launch_unsort_u32(cudaStream_t st, int n, const uint32_t* src, const int* ids, int* dst) {
    if (n > 0) k_unsort_u32<<<nblk(n, 256), 256, 0, st>>>(n, src, ids, dst);
}
void launch_pack_state(cudaStream_t st, int n, const float* pos_xyz, const float* vel_xyz, float4* posq, float4* velv,
                       int* ids, float* sed) {
    if (n > 0) k_pack_state<<<nblk(n, 256), 256, 0, st>>>(n, pos_xyz, vel_xyz, posq, velv, ids, sed);
}
void launch_slot_of_id(cudaStream_t st, int n, const int* ids, int* slot_of_id) {
    if (n > 0) k_slot_of_id<<<nblk(n, 256), 256, 0, st>>>(n, ids, slot_of_id);
}

}  // namespace sphe
