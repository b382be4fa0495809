// ORACLE — TEST INFRASTRUCTURE ONLY (see rtw_oracle.hpp).
// Command-line driver: renders the seeded scenes::simple restatement on the CPU and writes the PPM
// exactly like bin/src/main.rs:89-104 (P3, rows reversed so the top row comes first).
//   oracle_cli [--width W] [--height H] [--spp S] [--depth D] [--seed N] [--tmin X] [--threads T]
//              [--faithful-bvh] [--rng w64|w32] [--grid N] [--ground 0|1|2] [--out file.ppm]
#include "rtw_oracle.hpp"

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <string>

using namespace orc;

int main(int argc, char** argv) {
    uint32_t W = 400, H = 225, spp = 10, depth = 50;
    uint64_t seed = 20261018;
    int grid = 11, ground = 0;
    Options opt;
    std::string out;
    for (int i = 1; i < argc; ++i) {
        std::string a = argv[i];
        auto next = [&]() { return std::string(i + 1 < argc ? argv[++i] : "0"); };
        if (a == "--width") W = std::stoul(next());
        else if (a == "--height") H = std::stoul(next());
        else if (a == "--spp") spp = std::stoul(next());
        else if (a == "--depth") depth = std::stoul(next());
        else if (a == "--seed") seed = std::stoull(next());
        else if (a == "--tmin") opt.tmin = std::stod(next());
        else if (a == "--threads") opt.threads = std::stoi(next());
        else if (a == "--faithful-bvh") opt.faithful_bvh = true;
        else if (a == "--rng") opt.rng_mode = next() == "w32" ? W32 : W64;
        else if (a == "--grid") grid = std::stoi(next());
        else if (a == "--ground") ground = std::stoi(next());
        else if (a == "--out") out = next();
        else { std::fprintf(stderr, "unknown argument %s\n", a.c_str()); return 2; }
    }
    opt.seed = seed;
    SceneDesc d = scene_simple(seed, grid, 0.8, 0.95, ground);
    auto sc = scene_from_arrays(d.sphere_mat.size(), d.spheres.data(), d.sphere_mat.data(), d.materials.size(), d.materials.data(),
                                d.plane_mat.size(), d.planes.data(), d.plane_mat.data(), d.lights.size() / 4, d.lights.data());
    CameraBuilder cb = d.cam;
    cb.vfov = 40.;                                   // main.rs:73
    cb.aspect_ratio = (double)W / (double)H;
    cb.image_width = W; cb.image_height = H; cb.samples_per_pixel = spp; cb.max_depth = depth;
    Camera cam = camera_build(cb);
    std::vector<double> img((size_t)W * H * 3);
    Counters c; bool pan = false;
    auto t0 = std::chrono::steady_clock::now();
    render(*sc, cam, opt, img.data(), &c, &pan);
    double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    int nt = opt.threads > 0 ? opt.threads : (int)std::thread::hardware_concurrency();
    std::printf("{\"impl\": \"oracle (C++ restatement of the reference, not the reference binary)\", \"width\": %u, \"height\": %u, "
                "\"spp\": %u, \"depth\": %u, \"threads\": %d, \"faithful_bvh\": %s, \"seconds\": %.4f, \"mpaths_per_s\": %.4f, "
                "\"mrays_per_s\": %.4f, \"rays_per_path\": %.4f, \"bvh_nodes\": %zu, \"bvh_leaves\": %zu, \"bvh_depth\": %zu, "
                "\"box_tests_per_ray\": %.2f, \"box_builds_per_ray\": %.2f, \"sphere_tests_per_ray\": %.2f, \"panicked\": %s}\n",
                W, H, spp, depth, nt, opt.faithful_bvh ? "true" : "false", sec, c.paths / sec * 1e-6, c.rays / sec * 1e-6,
                (double)c.rays / c.paths, sc->world->node_count(), sc->world->leaf_count(), sc->world->depth(),
                (double)c.box_tests / c.rays, (double)c.box_builds / c.rays, (double)c.sphere_tests / c.rays, pan ? "true" : "false");
    if (!out.empty()) {
        FILE* f = std::fopen(out.c_str(), "w");
        if (!f) return 1;
        std::fprintf(f, "P3\n%u %u\n255\n", W, H);
        for (uint32_t j = H; j-- > 0;)
            for (uint32_t i = 0; i < W; ++i) {
                const double* p = &img[3 * ((size_t)j * W + i)];
                std::fprintf(f, "%u %u %u\n", quantise(p[0], spp), quantise(p[1], spp), quantise(p[2], spp));
            }
        std::fclose(f);
    }
    return 0;
}
