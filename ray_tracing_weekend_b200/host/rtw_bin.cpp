// rtw_bin.cpp — the reference's `bin` (bin/src/main.rs:54-105) with the CUDA backend behind
// Camera::render:   rtw_bin <simple|simple-light|cornell-box|debug|simple-transform|checkered-spheres> [--backend cuda] [--width W --height H --spp S --depth D]
//                           [--seed N] [--precision f32|f64] [--tmin X] [--out image.ppm]
// Config.toml parsing is replaced by flags (defaults = the reference's Config.toml:7-11).
// Writes ASCII P3 with rows reversed exactly like main.rs:89-104.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <string>

#include "rtw_host.hpp"

using namespace rtw_host;

int main(int argc, char** argv) {
    std::string scene = "simple", backend = "cuda", out = "image.ppm";
    uint32_t width = 400, height = 400, spp = 1000, depth = 50;
    RenderOptions opt;
    for (int i = 1; i < argc; ++i) {
        std::string a = argv[i];
        auto next = [&]() { return std::string(i + 1 < argc ? argv[++i] : ""); };
        if (a == "--backend") backend = next();
        else if (a == "--width") width = std::stoul(next());
        else if (a == "--height") height = std::stoul(next());
        else if (a == "--spp") spp = std::stoul(next());
        else if (a == "--depth") depth = std::stoul(next());
        else if (a == "--seed") opt.seed = std::stoull(next());
        else if (a == "--tmin") opt.tmin = std::stod(next());
        else if (a == "--precision") opt.precision = next() == "f64" ? Precision::F64 : Precision::F32;
        else if (a == "--out") out = next();
        else if (a[0] != '-') scene = a;
        else { std::fprintf(stderr, "unknown argument %s\n", a.c_str()); return 2; }
    }
    if (backend != "cuda") { std::fprintf(stderr, "this binary only carries the CUDA backend (--backend cuda); the CPU renderer is the reference's own\n"); return 2; }
    try {
        scenes::Output simple_sc;
        scenes::GeneralOutput general_sc;
        bool general = true;
        CameraBuilder cb;
        if (scene == "simple") { simple_sc = scenes::simple(opt.seed); cb = simple_sc.cam; general = false; }
        else if (scene == "simple-light" || scene == "simple_light") general_sc = scenes::simple_light(opt.seed);
        else if (scene == "cornell-box" || scene == "cornell_box") general_sc = scenes::cornell_box();
        else if (scene == "debug") general_sc = scenes::debugging_scene(opt.seed);
        else if (scene == "simple-transform" || scene == "simple_transform") general_sc = scenes::simple_transform(opt.seed);
        else if (scene == "checkered-spheres" || scene == "checkered_spheres") general_sc = scenes::checkered_spheres();
        else if (scene == "plane") general_sc = scenes::plane();
        else {
            // perlin-spheres pairs Lambertian spheres in plain view with an EMPTY lights list: the reference panics on the first light
            // sample (hittable_list.rs:414-419)
            std::fprintf(stderr, "scene '%s' is not provided (simple, simple-light, cornell-box, debug, simple-transform, checkered-spheres, plane)\n", scene.c_str());
            return 2;
        }
        if (general) cb = general_sc.cam;
        // main.rs:72-79
        Camera cam = cb.with_vfov(40.).with_aspect_ratio((double)width / (double)height).with_max_depth(depth)
                         .with_image_width(width).with_image_height(height).with_samples_per_pixel((uint16_t)spp).build();
        rtw_stats st{};
        auto t0 = std::chrono::steady_clock::now();
        auto img = general ? cam.render(general_sc.world_ref(), general_sc.lights_ref(), opt, &st)
                           : cam.render(simple_sc.world, simple_sc.lights, opt, &st);
        double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        std::fprintf(stderr, "rendered %ux%u spp %u in %.3f s (kernel %.3f ms): %.1f Mpaths/s, %.1f Mrays/s\n", width, height, spp, sec,
                     st.kernel_ms, st.paths / st.kernel_ms * 1e-3, st.rays / st.kernel_ms * 1e-3);
        FILE* f = std::fopen(out.c_str(), "w");
        if (!f) { std::perror("fopen"); return 1; }
        std::fprintf(f, "P3\n%u %u\n255\n", width, height);
        for (size_t j = img.size(); j-- > 0;)
            for (const SampledColour& c : img[j]) std::fprintf(f, "%s\n", c.to_string().c_str());
        std::fclose(f);
    } catch (const std::exception& e) {
        std::fprintf(stderr, "error: %s\n", e.what());
        return 1;
    }
    return 0;
}
