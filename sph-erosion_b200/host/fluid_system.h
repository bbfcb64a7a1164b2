// Drop-in for the reference's Erosion/fluid_system.h: struct FluidParticle and class FluidSystemSPH
// with the public surface main.cpp uses (fluid_system.h:49-64, :66-289), implemented over the C ABI of
// libsphe_b200.so (include/sphe.h).  Particle state lives on the GPU; parameters live in host memory
// owned by the handle, so the raw pointers ImGui writes through (main.cpp:278-290) keep working and
// are re-read by every Run().  The reference reports no errors on this path; failures of the C ABI
// are printed to stderr and otherwise ignored to stay drop-in.
#pragma once
#include <cstdio>
#include <iostream>
#include <vector>

#include "grid.h"
#include "sphe.h"
#include "sphe_glm_compat.h"

struct FluidParticle {   // fluid_system.h:49-64 (112 bytes)
    int Id;
    glm::vec3 Position;
    glm::vec3 Velocity;
    glm::vec3 Acceleration;
    float Density;
    float Pressure;
    glm::vec3 PressureForce;
    glm::vec3 ViscosityForce;
    glm::vec3 GravityForce;
    glm::vec3 SurfaceForce;
    glm::vec3 SurfaceNormal;
    int NeighbId;
};

class Shader;   // only Draw() needs it (rendering)

class FluidSystemSPH {
public:
    // fluid_system.h:69-72.  The reference object is a global constructed before main() (main.cpp:50),
    // before any CUDA/GL initialisation: sphe_create does no CUDA work.
    FluidSystemSPH() { check(sphe_create(&m_S), "FluidSystemSPH"); }
    ~FluidSystemSPH() { sphe_destroy(m_S); }
    FluidSystemSPH(const FluidSystemSPH&) = delete;
    FluidSystemSPH& operator=(const FluidSystemSPH&) = delete;

    void Initialize(int nParts) { check(sphe_initialize(m_S, nParts), "Initialize"); }   // :74-102
    void Run(Grid& grid) { check(sphe_step(m_S, m_UseTerrain ? grid.handle() : nullptr), "Run"); }   // :104-183
    void SetOrigin(const glm::vec3& new_origin) { float o[3] = {new_origin.x, new_origin.y, new_origin.z}; sphe_set_origin(m_S, o); }   // :206-209
    glm::vec3 GetOrigin() const { float o[3] = {0, 0, 0}; sphe_get_origin(m_S, o); return glm::vec3(o[0], o[1], o[2]); }             // :211-214
    void SetDeltaTime(float dt) { sphe_set_dt(m_S, dt); }      // :216-219
    float GetDeltaTime() const { return sphe_get_dt(m_S); }    // :221-224

    void PrintCoords() const {   // :226-230
        int n = sphe_count(m_S), num = sphe_num(m_S);
        std::vector<float> p(3 * (size_t)(n > 0 ? n : 1));
        if (n > 0) check(sphe_download_positions(m_S, p.data()), "PrintCoords");
        for (int i = 0; i < num && i < n; i++)
            std::cout << "[" << i << "] " << p[3 * i] << " " << p[3 * i + 1] << " " << p[3 * i + 2] << std::endl;
    }

    void AddParticles(int n) { check(sphe_add_particles(m_S, n), "AddParticles"); }   // :232-251
    void Reset() { check(sphe_reset(m_S), "Reset"); }                                // :253-259

    // :261-284 -- pointers into host storage owned by the handle, stable for its lifetime
    float* GetMass() { return &sphe_params_ptr(m_S)->mass; }
    float* GetVisc() { return &sphe_params_ptr(m_S)->visc; }
    float* GetSurfTen() { return &sphe_params_ptr(m_S)->surf_tens; }
    float* Getp0() { return &sphe_params_ptr(m_S)->p0; }
    glm::vec3* GetGrav() {
        static_assert(sizeof(glm::vec3) == 3 * sizeof(float), "glm::vec3 must be 3 packed floats");
        return reinterpret_cast<glm::vec3*>(sphe_params_ptr(m_S)->g);
    }

    FluidParticle GetParticle(int id) {   // :286-289 (the reference indexes unchecked; out of range returns zeros here)
        sphe_particle q{};
        FluidParticle p{};
        if (sphe_get_particle(m_S, id, &q) != SPHE_OK) return p;
        p.Id = q.id;
        p.Position = v(q.position); p.Velocity = v(q.velocity); p.Acceleration = v(q.acceleration);
        p.Density = q.density; p.Pressure = q.pressure;
        p.PressureForce = v(q.pressure_force); p.ViscosityForce = v(q.viscosity_force); p.GravityForce = v(q.gravity_force);
        p.SurfaceForce = v(q.surface_force); p.SurfaceNormal = v(q.surface_normal);
        p.NeighbId = q.neighb_id;
        return p;
    }

    // :185-204.  Rendering is outside the hot path: with SPHE_WITH_GL (and the reference's shader.h /
    // sphere.h on the include path) this reproduces the reference's one-sphere-per-particle loop from a
    // single packed position download; without it Draw is a no-op and Positions() is the hand-off.
#ifdef SPHE_WITH_GL
    void Draw(const Shader& shader, int selected_part);
#else
    void Draw(const Shader&, int) {}
#endif
    // Renderer hand-off without the PCIe round trip: packed xyz positions (id order) into a DEVICE buffer, e.g. an OpenGL
    // vertex buffer mapped through CUDA-GL interop (DrawInstanced below); capacity in floats.
    void PositionsToDevice(void* device_xyz, long long capacity_floats) {
        check(sphe_write_positions_device(m_S, device_xyz, capacity_floats), "PositionsToDevice");
    }
#ifdef SPHE_WITH_GL_INTEROP
    // One instanced draw instead of the reference's glDrawElements per particle (fluid_system.h:190-203).  The instance
    // buffer is registered with CUDA once (cudaGraphicsGLRegisterBuffer) and filled on the device every frame; the vertex
    // shader adds the per-instance offset (attribute `instance_attr`, divisor 1) to the unit sphere scaled by 0.01.
    // Needs <cuda_gl_interop.h> and GLEW on the include path (not in this image: compiled by nobody here).
    void DrawInstanced(GLuint sphere_vao, GLsizei sphere_index_count, GLuint instance_attr = 3);
#endif
    const std::vector<float>& Positions() {
        int n = sphe_count(m_S);
        m_Pos.resize(3 * (size_t)n);
        if (n > 0) check(sphe_download_positions(m_S, m_Pos.data()), "Positions");
        return m_Pos;
    }

    // ---- beyond the reference
    // The reference's call into Grid::collision is commented out (fluid_system.h:335-340), so by default
    // Run(grid) ignores the grid exactly like the reference.  UseTerrain(true) makes it live: terrain
    // contact + erosion run between the integration and the box collision.
    void UseTerrain(bool on) { m_UseTerrain = on; }
    sphe_sim* handle() const { return m_S; }
    int Count() const { return sphe_count(m_S); }
    // Checkpoint / resume (the reference has no on-disk format): particles + parameters, and the terrain with its
    // eroded heights when `grid` is given.  A resumed run continues bit for bit.
    bool Save(const char* path, Grid* grid = nullptr) { int rc = sphe_save_state(m_S, grid ? grid->handle() : nullptr, path); check(rc, "Save"); return rc == SPHE_OK; }
    bool Load(const char* path, Grid* grid = nullptr) { int rc = sphe_load_state(m_S, grid ? grid->handle() : nullptr, path); check(rc, "Load"); return rc == SPHE_OK; }

private:
    static glm::vec3 v(const float* a) { return glm::vec3(a[0], a[1], a[2]); }
    static void check(int rc, const char* what) {
        if (rc != SPHE_OK) std::fprintf(stderr, "sphe: %s failed (%d): %s\n", what, rc, sphe_last_error());
    }
    sphe_sim* m_S = nullptr;
    bool m_UseTerrain = false;
    std::vector<float> m_Pos;
};

#ifdef SPHE_WITH_GL_INTEROP
#include <cuda_gl_interop.h>
inline void FluidSystemSPH::DrawInstanced(GLuint sphere_vao, GLsizei sphere_index_count, GLuint instance_attr) {
    static GLuint vbo = 0;
    static cudaGraphicsResource* res = nullptr;
    static size_t cap_floats = 0;
    const size_t need = 3 * (size_t)sphe_count(m_S);
    if (need == 0) return;
    if (need > cap_floats) {       // (re)create and register the instance buffer when the particle count grows
        if (res) { cudaGraphicsUnregisterResource(res); res = nullptr; }
        if (!vbo) glGenBuffers(1, &vbo);
        glBindBuffer(GL_ARRAY_BUFFER, vbo);
        glBufferData(GL_ARRAY_BUFFER, need * sizeof(float), nullptr, GL_DYNAMIC_DRAW);
        cudaGraphicsGLRegisterBuffer(&res, vbo, cudaGraphicsRegisterFlagsWriteDiscard);
        cap_floats = need;
        glBindVertexArray(sphere_vao);
        glEnableVertexAttribArray(instance_attr);
        glVertexAttribPointer(instance_attr, 3, GL_FLOAT, GL_FALSE, 3 * sizeof(float), (void*)0);
        glVertexAttribDivisor(instance_attr, 1);
    }
    void* dev = nullptr; size_t bytes = 0;
    cudaGraphicsMapResources(1, &res, 0);
    cudaGraphicsResourceGetMappedPointer(&dev, &bytes, res);
    PositionsToDevice(dev, (long long)(bytes / sizeof(float)));
    cudaGraphicsUnmapResources(1, &res, 0);
    glBindVertexArray(sphere_vao);
    glDrawElementsInstanced(GL_TRIANGLES, sphere_index_count, GL_UNSIGNED_INT, nullptr, (GLsizei)(need / 3));
}
#endif

#ifdef SPHE_WITH_GL
#include "shader.h"
#include "sphere.h"
#include <glm/gtc/matrix_transform.hpp>
#include <memory>
inline void FluidSystemSPH::Draw(const Shader& shader, int selected_part) {
    static std::unique_ptr<Sphere> sphere;
    if (!sphere) sphere = std::make_unique<Sphere>(10, 10, 1, glm::vec3(0.0, 0.0, 0.0));
    const std::vector<float>& p = Positions();
    for (size_t i = 0; i < p.size() / 3; i++) {
        glm::mat4 model = glm::scale(glm::translate(glm::mat4(1.0), glm::vec3(p[3 * i], p[3 * i + 1], p[3 * i + 2])), glm::vec3(0.01f));
        shader.setMat4("model", model);
        shader.setVec3("myColor", (int)i == selected_part ? glm::vec3(1.0, 1.0, 0.0) : glm::vec3(0.0, 0.0, 1.0));
        sphere->Draw();
    }
}
#endif
