// TEST INFRASTRUCTURE ONLY -- never part of the product path.
//
// Thin extern "C" harness around the UNMODIFIED reference headers, compiled
// in place from /root/reference (see oracle/Makefile; nothing is copied into
// this repo).  It exists to (1) validate oracle/sph_oracle.c bit-for-bit and
// (2) serve as the "reference" CPU baseline in bench.py.
//
// The reference keeps its state private (Erosion/fluid_system.h:456-484), so
// state injection uses the `#define private public` trick recommended in
// standard/glm headers are included first so
// only the reference's own classes are affected.
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include <cfloat>
#include <vector>
#include <list>
#include <memory>
#include <string>
#include <fstream>
#include <sstream>
#include <iostream>
#include <iomanip>
#include <omp.h>
#include <glm/glm.hpp>
#include <glm/gtc/matrix_transform.hpp>
#include <glm/gtc/type_ptr.hpp>
#include <GL/glew.h>

#define private public
#include "fluid_system.h"
#undef private

// ~Sphere (Erosion/sphere.h:87-95) references two GLEW entry points; the
// Sphere is never constructed headless, so null pointers suffice.
extern "C" {
PFNGLDELETEVERTEXARRAYSPROC __glewDeleteVertexArrays = nullptr;
PFNGLDELETEBUFFERSPROC __glewDeleteBuffers = nullptr;
}

struct RefSim {
    FluidSystemSPH sim;
    Grid grid;
    RefSim() : grid(1, 1, 1) {}
};

extern "C" {

int ref_sizeof_particle() { return (int)sizeof(FluidParticle); }

void* ref_create() { return new RefSim(); }
void ref_destroy(void* h) { delete (RefSim*)h; }

void ref_initialize(void* h, int n) { ((RefSim*)h)->sim.Initialize(n); }
void ref_add_particles(void* h, int n) { ((RefSim*)h)->sim.AddParticles(n); }
void ref_reset(void* h) { ((RefSim*)h)->sim.Reset(); }
void ref_set_origin(void* h, float x, float y, float z) { ((RefSim*)h)->sim.SetOrigin(glm::vec3(x, y, z)); }
void ref_set_dt(void* h, float dt) { ((RefSim*)h)->sim.SetDeltaTime(dt); }
float ref_get_dt(void* h) { return ((RefSim*)h)->sim.GetDeltaTime(); }
int ref_count(void* h) { return (int)((RefSim*)h)->sim.m_Particles.size(); }

// private-member injection (no setter exists in the reference)
void ref_set_len(void* h, float len) { ((RefSim*)h)->sim.len = len; }
void ref_set_h(void* h, float hh) { ((RefSim*)h)->sim.h = hh; ((RefSim*)h)->sim.smoothRadius = hh; }
void ref_set_k(void* h, float k) { ((RefSim*)h)->sim.k = k; }
void ref_set_params(void* h, float mass, float visc, float surf, float p0, const float* g) {
    FluidSystemSPH& s = ((RefSim*)h)->sim;
    *s.GetMass() = mass; *s.GetVisc() = visc; *s.GetSurfTen() = surf; *s.Getp0() = p0;
    *s.GetGrav() = glm::vec3(g[0], g[1], g[2]);
}
void ref_get_params(void* h, float* out /*mass visc surf p0 gx gy gz k h len*/) {
    FluidSystemSPH& s = ((RefSim*)h)->sim;
    out[0] = s.MASS; out[1] = s.visc; out[2] = s.surf_tens; out[3] = s.p0;
    out[4] = s.g.x; out[5] = s.g.y; out[6] = s.g.z; out[7] = s.k; out[8] = s.h; out[9] = s.len;
}

void ref_set_state(void* h, int n, const float* pos, const float* vel) {
    FluidSystemSPH& s = ((RefSim*)h)->sim;
    s.m_Particles.clear();
    s.m_Particles.resize(n);
    for (int i = 0; i < n; i++) {
        FluidParticle p;
        memset(&p, 0, sizeof p);
        p.Id = i;
        p.Position = glm::vec3(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]);
        p.Velocity = glm::vec3(vel[3 * i], vel[3 * i + 1], vel[3 * i + 2]);
        s.m_Particles[i] = p;
    }
    s.num = n; s.init_num = n; s.id = n;
    s.smoothRadius = s.h;
}

void ref_run(void* h, int steps) {
    RefSim* r = (RefSim*)h;
    for (int i = 0; i < steps; i++) r->sim.Run(r->grid);
}

// field ids: 0 pos 1 vel 2 acc 3 density 4 pressure 5 fpress 6 fvisc 7 fgrav
// 8 fsurf 9 normal (float outputs); 10 id 11 neighb id (int outputs)
void ref_get_field(void* h, int field, void* out) {
    FluidSystemSPH& s = ((RefSim*)h)->sim;
    int n = (int)s.m_Particles.size();
    float* f = (float*)out; int* ii = (int*)out;
    for (int i = 0; i < n; i++) {
        const FluidParticle& p = s.m_Particles[i];
        const glm::vec3* v = nullptr;
        switch (field) {
            case 0: v = &p.Position; break;
            case 1: v = &p.Velocity; break;
            case 2: v = &p.Acceleration; break;
            case 3: f[i] = p.Density; break;
            case 4: f[i] = p.Pressure; break;
            case 5: v = &p.PressureForce; break;
            case 6: v = &p.ViscosityForce; break;
            case 7: v = &p.GravityForce; break;
            case 8: v = &p.SurfaceForce; break;
            case 9: v = &p.SurfaceNormal; break;
            case 10: ii[i] = p.Id; break;
            case 11: ii[i] = p.NeighbId; break;
        }
        if (v) { f[3 * i] = v->x; f[3 * i + 1] = v->y; f[3 * i + 2] = v->z; }
    }
}

// ---- Grid (Erosion/grid.h) ----
void* ref_grid_create(int dx, int dy, int dz) { return new Grid(dx, dy, dz); }
void ref_grid_destroy(void* g) { delete (Grid*)g; }
void ref_grid_load_heightfield(void* g, const unsigned char* img) { ((Grid*)g)->LoadHeightfield((unsigned char*)img); }
int ref_grid_height_at(void* g, int x, int y) { return ((Grid*)g)->GetHeightfieldAt(x, y); }
void ref_grid_update(void* g, int dx, int dy, int dz) { ((Grid*)g)->UpdateGrid(dx, dy, dz); }
long ref_grid_surface_size(void* g) { return (long)((Grid*)g)->GetSurfacePartsSize(); }
long ref_grid_indices_size(void* g) { return (long)((Grid*)g)->GetIndicesSize(); }
void ref_grid_get_surface(void* g, float* out) { auto v = ((Grid*)g)->GetSurfaceParts(); memcpy(out, v.data(), v.size() * sizeof(float)); }
void ref_grid_get_indices(void* g, unsigned* out) { auto v = ((Grid*)g)->GetIndices(); memcpy(out, v.data(), v.size() * sizeof(unsigned)); }
int ref_grid_voxel_type(void* g, int x, int y, int z) { return (int)((Grid*)g)->GetVoxel(x, y, z).type; }

// batched Grid::collision (Erosion/grid.h:462-805)
void ref_grid_collision(void* g, int n, const float* pc, const float* pn, const float* vn,
                        int* hit, float* cp, float* nrm) {
    Grid* G = (Grid*)g;
    for (int i = 0; i < n; i++) {
        glm::vec3 c(0.0f), m(0.0f);
        bool r = G->collision(glm::vec3(pc[3 * i], pc[3 * i + 1], pc[3 * i + 2]),
                              glm::vec3(pn[3 * i], pn[3 * i + 1], pn[3 * i + 2]),
                              glm::vec3(vn[3 * i], vn[3 * i + 1], vn[3 * i + 2]), c, m);
        hit[i] = r ? 1 : 0;
        cp[3 * i] = c.x; cp[3 * i + 1] = c.y; cp[3 * i + 2] = c.z;
        nrm[3 * i] = m.x; nrm[3 * i + 1] = m.y; nrm[3 * i + 2] = m.z;
    }
}

int ref_omp_max_threads() {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

}  // extern "C"
