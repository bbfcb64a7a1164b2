// Multi-GPU headless driver in C++ over the C ABI (include/sphe.h "multi-GPU x-slabs"): ONE process, ONE host thread,
// K x-slabs round-robin over the visible GPUs.  Every call of the step only enqueues work (the slab exchange has no
// host sync: the pack kernel stores into the neighbours' mailboxes over NVLink, the append kernel waits on device
// flags), so a single thread keeps all GPUs busy; no MPI, no NCCL, no Python.
//
//   headless_slabs [--slabs K] [--axis N] [--steps S] [--check]
//
// Scene: the weak-scaling dam break of bench.py (slabs.channel_block, layout "contiguous"): one block of K*N x N x N
// particles at spacing 0.025 at the left end of a channel of half-extents (K*L, L, L), L = 0.02*N.
// --check also runs the whole scene on ONE handle and requires the slabs to reproduce it BIT FOR BIT.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "sphe.h"

#define OK(call) do { int rc_ = (call); if (rc_ != SPHE_OK) { std::fprintf(stderr, "%s failed (%d): %s\n", #call, rc_, sphe_last_error()); std::exit(1); } } while (0)

static const float SPACING = 0.025f, H = 0.0457f;

int main(int argc, char** argv) {
    int K = 2, n_axis = 40, steps = 50, devices = 0;
    bool check = false, debug = false;
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        auto next = [&]() { if (i + 1 >= argc) { std::fprintf(stderr, "missing value for %s\n", a.c_str()); std::exit(2); } return std::atoi(argv[++i]); };
        if (a == "--slabs") K = next();
        else if (a == "--axis") n_axis = next();
        else if (a == "--steps") steps = next();
        else if (a == "--devices") devices = next();
        else if (a == "--check") check = true;
        else if (a == "--debug") debug = true;
        else { std::fprintf(stderr, "unknown argument %s\n", a.c_str()); return 2; }
    }
    if (K < 1 || n_axis < 8) { std::fprintf(stderr, "need --slabs >= 1 and --axis >= 8\n"); return 2; }
    if (devices <= 0) devices = std::max(sphe_device_count(), 1);
    const float L = 0.02f * n_axis;
    const float box[3] = {L * K, L, L};
    const long long per = (long long)n_axis * n_axis * n_axis, n_total = per * K;

    auto make = [&](int device) {
        sphe_sim* s = nullptr;
        OK(sphe_create(&s));
        OK(sphe_set_device(s, device));
        sphe_params* P = sphe_params_ptr(s);
        P->len = L; P->dt = 0.01f; P->g[1] = -9.82f * 10.0f / n_axis;      // bench.scene_gravity
        OK(sphe_set_box(s, box));
        return s;
    };
    auto block = [&](int r, std::vector<float>& pos, std::vector<int>& ids) {   // slab r's share of the lattice
        pos.resize(3 * (size_t)per); ids.resize((size_t)per);
        size_t k = 0;
        for (int i = 0; i < n_axis; i++) for (int j = 0; j < n_axis; j++) for (int l = 0; l < n_axis; l++, k++) {
            pos[3 * k] = (float)(-(double)box[0] + (double)(r * n_axis + i) * 0.025);
            pos[3 * k + 1] = (float)(-(double)L / 4 + j * 0.025);
            pos[3 * k + 2] = (float)(-0.75 * (double)L + l * 0.025);
            ids[k] = (int)(r * per + (long long)k);
        }
    };

    // global grid (any handle knows it once the box is set): column cuts at the lattice mid-planes between the blocks
    std::vector<sphe_sim*> sims(K);
    for (int r = 0; r < K; r++) sims[r] = make(r % devices);
    int gnx = 0;
    OK(sphe_slab_configure(sims[0], 0, 8, 0, 0));            // provisional, to read the global grid
    OK(sphe_slab_info(sims[0], &gnx, nullptr, nullptr, nullptr));
    sphe_grid_info gi;
    OK(sphe_grid_info_get(sims[0], &gi));
    std::vector<int> edge(K + 1, 0);
    edge[K] = gnx;
    for (int r = 1; r < K; r++) edge[r] = (int)std::floor(((float)(-(double)box[0] + (r * n_axis - 0.5) * 0.025) - gi.gmin[0]) / gi.cell);
    const int layer = (int)(n_axis * n_axis * (H * 1.001f / SPACING + 1));
    const int cap = std::max(4 * layer, 1 << 14);
    std::vector<float> pos, vel;
    std::vector<int> ids;
    for (int r = 0; r < K; r++) {
        const bool ring = K >= 3;
        OK(sphe_slab_configure(sims[r], edge[r], edge[r + 1], ring || r > 0, ring || r < K - 1));
        if (ring && (r == 0 || r == K - 1)) OK(sphe_slab_ring(sims[r], r == 0, r == K - 1, edge[K - 1]));
        block(r, pos, ids);
        vel.assign(pos.size(), 0.0f);
        OK(sphe_slab_upload(sims[r], (int)per, pos.data(), vel.data(), ids.data()));
        OK(sphe_slab_peer_setup(sims[r], cap, (int)(per * 13 / 10) + 6 * cap));
    }
    for (int r = 0; r < K; r++) {
        sphe_sim* l = K >= 3 ? sims[(r + K - 1) % K] : (r > 0 ? sims[r - 1] : nullptr);
        sphe_sim* rt = K >= 3 ? sims[(r + 1) % K] : (r < K - 1 ? sims[r + 1] : nullptr);
        OK(sphe_slab_peer_connect_local(sims[r], l, rt));
    }

    std::vector<long long> ticket(K, -1);
    auto step_all = [&]() {
        if (K > 1) {
            for (int r = 0; r < K; r++) OK(sphe_slab_send(sims[r]));             // all producers first: nobody waits for a launch
            for (int r = 0; r < K; r++) OK(sphe_slab_recv(sims[r], &ticket[r])); // that has not been enqueued yet
        }
        for (int r = 0; r < K; r++) OK(sphe_step(sims[r], nullptr));
    };
    if (debug && K > 1) {   // one exchange, then what every slab kept / sent / received / has in transit
        for (int r = 0; r < K; r++) OK(sphe_slab_send(sims[r]));
        for (int r = 0; r < K; r++) OK(sphe_slab_recv(sims[r], &ticket[r]));
        for (int r = 0; r < K; r++) {
            int out[6] = {0, 0, 0, 0, 0, 0}, tr[3] = {0, 0, 0};
            int rc = sphe_slab_result(sims[r], ticket[r], 1, out);
            sphe_slab_transit(sims[r], tr);
            std::printf("slab %d [%d,%d): rc %d (%s) n_total %d owned %d to_left %d to_right %d from_left %d from_right %d | transit %d %d forwarded %d\n",
                        r, edge[r], edge[r + 1], rc, rc ? sphe_last_error() : "ok", out[0], out[1], out[2], out[3], out[4], out[5], tr[0], tr[1], tr[2]);
        }
        for (int r = 0; r < K; r++) OK(sphe_step(sims[r], nullptr));
    }
    for (int w = 0; w < 5; w++) step_all();
    for (int r = 0; r < K; r++) OK(sphe_sync(sims[r]));
    auto t0 = std::chrono::steady_clock::now();
    for (int s = 0; s < steps; s++) step_all();
    for (int r = 0; r < K; r++) OK(sphe_sync(sims[r]));
    const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    std::printf("%d slabs on %d device(s), %lld particles: %.4f ms/step, %.3e particle-updates/s (host wall clock around %d steps)\n",
                K, std::min(devices, K), n_total, 1e3 * sec / steps, (double)n_total * steps / sec, steps);

    // census: every particle owned exactly once (+ records in transit)
    std::vector<float> all_pos(3 * (size_t)n_total, 0.0f);
    std::vector<int> seen((size_t)n_total, 0);
    long long owned = 0, transit = 0;
    for (int r = 0; r < K; r++) {
        int out[6];
        if (K > 1) OK(sphe_slab_result(sims[r], ticket[r], 1, out));
        int m = 0, tr[3] = {0, 0, 0};
        std::vector<int> oi((size_t)per * 2 + cap);
        std::vector<float> op(3 * oi.size()), ov(3 * oi.size());
        OK(sphe_slab_download(sims[r], (int)oi.size(), oi.data(), op.data(), ov.data(), nullptr, nullptr, &m));
        OK(sphe_slab_transit(sims[r], tr));
        transit += tr[0] + tr[1];
        for (int k = 0; k < m; k++) { seen[oi[k]]++; std::memcpy(&all_pos[3 * (size_t)oi[k]], &op[3 * (size_t)k], 3 * sizeof(float)); }
        owned += m;
    }
    long long bad = 0;
    for (long long i = 0; i < n_total; i++) bad += seen[i] > 1;
    std::printf("census: %lld owned + %lld in transit of %lld, %lld owned twice\n", owned, transit, n_total, bad);
    bool ok = bad == 0 && owned + transit == n_total;

    if (check) {
        sphe_sim* one = make(0);
        std::vector<float> p1(3 * (size_t)n_total), v1(3 * (size_t)n_total, 0.0f);
        for (int r = 0; r < K; r++) { block(r, pos, ids); std::memcpy(&p1[3 * (size_t)(r * per)], pos.data(), pos.size() * sizeof(float)); }
        OK(sphe_upload_state(one, (int)n_total, p1.data(), v1.data()));
        const int total_steps = steps + 5 + ((debug && K > 1) ? 1 : 0);
        for (int s = 0; s < total_steps; s++) OK(sphe_step(one, nullptr));
        OK(sphe_download(one, SPHE_F_POS, p1.data()));
        const bool same = transit == 0 && std::memcmp(p1.data(), all_pos.data(), p1.size() * sizeof(float)) == 0;
        std::printf("positions after %d steps bit-equal to the single-handle run: %s\n", total_steps, same ? "yes" : "NO");
        ok = ok && same;
        sphe_destroy(one);
    }
    for (auto s : sims) sphe_destroy(s);
    std::printf("HEADLESS_SLABS %s\n", ok ? "OK" : "FAILED");
    return ok ? 0 : 1;
}
