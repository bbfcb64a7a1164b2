// Device code shared by the neighbour-pass translation units (sph.cu, stage.cu): the per-particle epilogue of the force
// pass (forces -> advance() -> collisionS, fluid_system.h:145-177, :306-353) and small helpers.
#pragma once
#include "common.cuh"
#include "sim.h"

namespace sphe {

// ------------------------------------------------------------------ force -> integrate -> collide (per particle)
// PressureForce = -(fPress*rho_i), fPress = -mass*c45*A (fluid_system.h:145,151); ViscosityForce = c45*visc*F
// (:146,153); SurfaceNormal = -c945*N (:147,154); colorFieldLapl = -c945*cf, SurfaceForce = -surf_tens*cfl*n
// (:171,177); GravityForce = rho_i*g (:163); then advance() (:318-350) with collisionS (:342-347).
template <bool DIAG>
__device__ __forceinline__ void force_epilogue(int i, float4 pi, float4 vi, float rho_i, float ax, float ay, float az,
                                               float fx, float fy, float fz, float nx, float ny, float nz, float cf, int maxid,
                                               const StepC& C, const int* __restrict__ ids, float4* __restrict__ posq_out,
                                               float4* __restrict__ velv_out, const DiagOut& D) {
    float kp = rho_i * C.mass * C.c45;
    float Fpx = kp * ax, Fpy = kp * ay, Fpz = kp * az;
    float kv = C.visc * C.c45;
    float Fvx = kv * fx, Fvy = kv * fy, Fvz = kv * fz;
    float Nx = -C.c945 * nx, Ny = -C.c945 * ny, Nz = -C.c945 * nz;
    float cfl = -C.c945 * cf;
    float ks = -C.surf * cfl;
    float Fsx = ks * Nx, Fsy = ks * Ny, Fsz = ks * Nz;
    float Fgx = rho_i * C.gx, Fgy = rho_i * C.gy, Fgz = rho_i * C.gz;
    float Fx = (Fpx + Fvx) + (Fgx + Fsx), Fy = (Fpy + Fvy) + (Fgy + Fsy), Fz = (Fpz + Fvz) + (Fgz + Fsz);
    float acx = Fx / rho_i, acy = Fy / rho_i, acz = Fz / rho_i;
    float dt = C.dt;
    float vx = fmaf(acx, dt, vi.x), vy = fmaf(acy, dt, vi.y), vz = fmaf(acz, dt, vi.z);
    float px = fmaf(vx, dt, pi.x), py = fmaf(vy, dt, pi.y), pz = fmaf(vz, dt, pi.z);
    // With a terrain, particles that may touch it keep their un-boxed state: the contact search
    // (k_terrain_contact) runs on them first and applies the box afterwards (fluid_system.h:335-347).
    bool surv = false;
    if (C.t_lmax) {
        // ghost copies (slab mode) never enter the terrain stage: their owner rank resolves the contact and
        // files the erosion request, the copy is dropped at the next exchange
        const int cls = (dt != 0.0f && !(__ldg(&ids[i]) & SPHE_GHOST_BIT)) ? terrain_may_touch(C, pi.x, pi.y, pi.z, px, py, pz) : -1;
        surv = cls >= 0;
        // one atomic per class per warp: lanes of the same class take consecutive slots of that class's list
        unsigned act = __activemask();
        unsigned peers = __match_any_sync(act, cls);
        if (surv) {
            int lane = threadIdx.x & 31, leader = __ffs(peers) - 1, base = 0;
            if (lane == leader) base = atomicAdd(C.t_count + cls, __popc(peers));
            base = __shfl_sync(peers, base, leader);
            C.t_surv[(size_t)cls * C.t_cap + base + __popc(peers & ((1u << lane) - 1u))] = i;
        }
    }
    if (C.box && !surv) box_collide(C, px, py, pz, vx, vy, vz);
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

__device__ __forceinline__ float rsqrt_ftz(float x) {
    float y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

}  // namespace sphe
