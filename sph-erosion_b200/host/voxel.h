// Drop-in for the reference's Erosion/voxel.h: same enum and struct (voxel.h:1-27).
#pragma once
#include "sphe_glm_compat.h"

enum class VoxelType { VOXEL_AIR = 0, VOXEL_WAT, VOXEL_MAT };

struct Voxel {
    float density;
    glm::vec3 position;
    glm::vec3 velocity;
    VoxelType type;
    Voxel() : density(0), position(0.0f), velocity(0.0f), type(VoxelType::VOXEL_AIR) {}
    Voxel(VoxelType t) : density(0), position(0.0f), velocity(0.0f), type(t) {}
};
