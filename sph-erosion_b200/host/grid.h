// Drop-in for the reference's Erosion/grid.h: class Grid with the public surface main.cpp uses
// (grid.h:72-841), implemented over the C ABI of libsphe_b200.so (include/sphe.h, "terrain").
// The heightfield, the render mesh and the particle-terrain contact search live on the GPU.
//
// Differences a maintainer should know (all documented in INTEGRATION.md):
//  * no dense Voxel array is allocated (the reference allocates dimX*dimY*dimZ voxels of 32 bytes and
//    only ever writes the surface voxel of each column, grid.h:89-96,150): GetVoxel derives the same
//    answer from the heightfield;
//  * GetIndices returns one generation of indices (the reference never clears `indices`, so every
//    UpdateGrid call appends another copy, grid.h:53-58 vs :118-136);
//  * heights can change while the simulation runs (erosion); GetHeightfieldAt / GetSurfaceParts read the
//    current device state.
#pragma once
#include <cstddef>
#include <cstdio>
#include <vector>

#include "sphe.h"
#include "voxel.h"

class Grid {
public:
    Grid(int dimX = 512, int dimY = 512, int dimZ = 512) : m_Dim((float)dimX, (float)dimY, (float)dimZ) {
        check(sphe_terrain_create(&m_T, dimX, dimY, dimZ), "Grid");   // no CUDA work here
    }
    ~Grid() { sphe_terrain_destroy(m_T); }
    Grid(const Grid&) = delete;
    Grid& operator=(const Grid&) = delete;

    // grid.h:84-96
    Voxel GetVoxel(int x, int y, int z) const {
        Voxel v;
        int h = sphe_terrain_height_at(m_T, x, z);
        int top = h >= (int)m_Dim.y ? (int)m_Dim.y - 1 : h;
        if (h >= 0 && y == top) { v.type = VoxelType::VOXEL_MAT; v.position = glm::vec3((float)x, (float)y, (float)z); }
        return v;
    }

    void LoadHeightfield(unsigned char* img) { check(sphe_terrain_load_heightfield(m_T, img), "LoadHeightfield"); }  // grid.h:98-102
    unsigned char GetHeightfieldAt(int x, int y) { int v = sphe_terrain_height_at(m_T, x, y); return (unsigned char)(v < 0 ? 0 : v); }  // :104-107

    void UpdateGrid(int dimx, int dimy, int dimz) {   // grid.h:138-176
        check(sphe_terrain_update_grid(m_T, dimx, dimy, dimz), "UpdateGrid");
        m_Dim = glm::vec3((float)dimx, (float)dimy, (float)dimz);
    }

    // grid.h:462-805; coordinates in terrain units (1 cell = 1 unit), like the reference
    bool collision(const glm::vec3& posCurr, const glm::vec3& posNext, const glm::vec3& velNext, glm::vec3& contactP, glm::vec3& norm) {
        float pc[3] = {posCurr.x, posCurr.y, posCurr.z}, pn[3] = {posNext.x, posNext.y, posNext.z}, vn[3] = {velNext.x, velNext.y, velNext.z};
        float cp[3], nn[3];
        int hit = 0;
        check(sphe_terrain_collision(m_T, 1, pc, pn, vn, &hit, cp, nn), "collision");
        if (hit) { contactP = glm::vec3(cp[0], cp[1], cp[2]); norm = glm::vec3(nn[0], nn[1], nn[2]); }
        return hit != 0;
    }

    std::vector<unsigned int> GetIndices() {          // grid.h:807-810
        std::vector<unsigned int> v((size_t)sphe_terrain_indices_size(m_T));
        if (!v.empty()) check(sphe_terrain_get_indices(m_T, v.data()), "GetIndices");
        return v;
    }
    size_t GetIndicesSize() const { return (size_t)sphe_terrain_indices_size(m_T); }
    glm::vec3 GetDim() const { return m_Dim; }
    std::vector<float> GetSurfaceParts() {            // grid.h:822-825: x y z nx ny nz per vertex, z-major
        std::vector<float> v((size_t)sphe_terrain_surface_size(m_T));
        if (!v.empty()) check(sphe_terrain_get_surface(m_T, v.data()), "GetSurfaceParts");
        return v;
    }
    size_t GetSurfacePartsSize() const { return (size_t)sphe_terrain_surface_size(m_T); }
    std::vector<float> GetFluidParts() { return std::vector<float>(); }   // never filled by the reference either (grid.h:40,832)
    size_t GetFluidPartsSize() const { return 0; }

    // ---- beyond the reference: what the erosion scenes need
    sphe_terrain* handle() const { return m_T; }
    sphe_erosion* Erosion() { return sphe_terrain_erosion_ptr(m_T); }
    void SetTransform(const glm::vec3& origin, float scale) { float o[3] = {origin.x, origin.y, origin.z}; check(sphe_terrain_set_transform(m_T, o, scale), "SetTransform"); }

private:
    static void check(int rc, const char* what) {
        if (rc != SPHE_OK) std::fprintf(stderr, "sphe: %s failed (%d): %s\n", what, rc, sphe_last_error());
    }
    sphe_terrain* m_T = nullptr;
    glm::vec3 m_Dim;
};
