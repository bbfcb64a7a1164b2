// Headless driver: the call pattern of the reference's Erosion/main.cpp without the window --
// global FluidSystemSPH (main.cpp:50), Grid(50,255,50) + LoadHeightfield + UpdateGrid (:100-110),
// SetOrigin + Initialize(1000) (:170,183), then per "frame" Run(grid) (:313), with the F key
// (dt 0 <-> 0.01, :476-485), mouse buttons (AddParticles(125) / Reset, :499-515), the ImGui parameter
// pointers (:278-290) and the particle inspector (:292-305) driven from the command line.
//
//   headless [--steps N] [--particles N] [--terrain raw512.bin] [--erosion] [--add-at STEP] [--dump file]
//            [--save file] [--load file]      checkpoint after the run / resume from a checkpoint instead of Initialize
//
// --dump writes the final state as raw little-endian float32: count (int32), then pos[3n], vel[3n], density[n].
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "fluid_system.h"
#include "grid.h"

FluidSystemSPH fluidsph;   // constructed before main(), like main.cpp:50

int main(int argc, char** argv) {
    int steps = 100, particles = 1000, add_at = -1;
    bool erosion = false;
    std::string terrain_file, dump_file, save_file, load_file;
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        auto next = [&](const char* what) -> const char* { if (i + 1 >= argc) { std::fprintf(stderr, "missing value for %s\n", what); std::exit(2); } return argv[++i]; };
        if (a == "--steps") steps = std::atoi(next("--steps"));
        else if (a == "--particles") particles = std::atoi(next("--particles"));
        else if (a == "--terrain") terrain_file = next("--terrain");
        else if (a == "--erosion") erosion = true;
        else if (a == "--add-at") add_at = std::atoi(next("--add-at"));
        else if (a == "--dump") dump_file = next("--dump");
        else if (a == "--save") save_file = next("--save");
        else if (a == "--load") load_file = next("--load");
        else { std::fprintf(stderr, "unknown argument %s\n", a.c_str()); return 2; }
    }

    float dimensions[3] = {50, 255, 50};
    Grid grid((int)dimensions[0], (int)dimensions[1], (int)dimensions[2]);
    if (!terrain_file.empty()) {
        std::vector<unsigned char> img(512 * 512);
        FILE* f = std::fopen(terrain_file.c_str(), "rb");
        if (!f || std::fread(img.data(), 1, img.size(), f) != img.size()) { std::fprintf(stderr, "cannot read %s (512x512 raw 8-bit)\n", terrain_file.c_str()); return 1; }
        std::fclose(f);
        grid.LoadHeightfield(img.data());
        grid.UpdateGrid((int)dimensions[0], (int)dimensions[1], (int)dimensions[2]);
        std::printf("terrain: %zu surface floats, %zu indices, H(5,7) = %d\n", grid.GetSurfaceParts().size(), grid.GetIndices().size(), (int)grid.GetHeightfieldAt(5, 7));
        if (erosion) {
            // drop the default scene onto the top-left corner of the terrain: 1 cell = 0.0125 world units
            grid.SetTransform(glm::vec3(-0.25f, -0.2f - 0.0125f * 150.0f, -0.25f), 0.0125f);
            grid.Erosion()->enabled = 1;
            fluidsph.UseTerrain(true);
        }
    }

    fluidsph.SetOrigin(glm::vec3(0.0f));
    if (!load_file.empty()) {
        if (!fluidsph.Load(load_file.c_str(), terrain_file.empty() ? nullptr : &grid)) return 1;
    } else {
        fluidsph.Initialize(particles);
        if (sphe_count(fluidsph.handle()) == 0) { std::fprintf(stderr, "no particles: %s\n", sphe_last_error()); return 1; }
        fluidsph.Run(grid);                       // paused frame: deltaT = 0 recomputes forces, moves nothing
        if (fluidsph.GetDeltaTime() == 0) fluidsph.SetDeltaTime(0.01f);   // the F key
        *fluidsph.GetVisc() = 3.5f;               // ImGui-style write through the parameter pointers
    }
    for (int s = 0; s < steps; s++) {
        if (s == add_at) fluidsph.AddParticles(125);
        fluidsph.Run(grid);
    }
    FluidParticle p = fluidsph.GetParticle(0);
    std::printf("particles %d  particle 0: pos %.9g %.9g %.9g  density %.9g  pressure %.9g\n", fluidsph.Count(),
                p.Position.x, p.Position.y, p.Position.z, p.Density, p.Pressure);
    if (!save_file.empty() && !fluidsph.Save(save_file.c_str(), terrain_file.empty() ? nullptr : &grid)) return 1;
    if (!dump_file.empty()) {
        int n = fluidsph.Count();
        std::vector<float> pos(3 * (size_t)n), vel(3 * (size_t)n), rho((size_t)n);
        sphe_download(fluidsph.handle(), SPHE_F_POS, pos.data());
        sphe_download(fluidsph.handle(), SPHE_F_VEL, vel.data());
        sphe_download(fluidsph.handle(), SPHE_F_DENSITY, rho.data());
        FILE* f = std::fopen(dump_file.c_str(), "wb");
        if (!f) { std::fprintf(stderr, "cannot write %s\n", dump_file.c_str()); return 1; }
        std::fwrite(&n, sizeof n, 1, f);
        std::fwrite(pos.data(), sizeof(float), pos.size(), f);
        std::fwrite(vel.data(), sizeof(float), vel.size(), f);
        std::fwrite(rho.data(), sizeof(float), rho.size(), f);
        std::fclose(f);
    }
    return 0;
}
