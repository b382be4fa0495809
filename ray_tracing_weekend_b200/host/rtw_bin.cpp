// rtw_bin.cpp — the reference's `bin` (bin/src/main.rs:54-105) with the CUDA backend behind
// Camera::render:   rtw_bin <simple|simple-light|cornell-box|debug|simple-transform|checkered-spheres> [--backend cuda] [--width W --height H --spp S --depth D]
//                           [--seed N] [--precision f32|f64] [--tmin X] [--out image.ppm]
//                           [--gpus N [--collective auto|peer|nccl]] [--mode wavefront|megakernel]
//                   [--config Config.toml]  [--format p3|p6|png]  [--passes K [--checkpoint FILE] [--resume]]
// Image parameters come from flags (defaults = the reference's Config.toml:7-11) or, with --config, from the [image] table of a
// Config.toml like the reference's (bin/src/config.rs).  The default output is ASCII P3 with rows reversed exactly like
// main.rs:89-104; --format adds binary P6 and PNG.  --passes renders progressively into fixed-point accumulators, writing a
// checkpoint after every pass; --resume continues an interrupted render from it — the image is the one-shot image bit for bit.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <string>

#include "rtw_host.hpp"

using namespace rtw_host;

int main(int argc, char** argv) {
    std::string scene = "simple", backend = "cuda", out = "image.ppm", config, format = "p3", checkpoint;
    uint32_t width = 400, height = 400, spp = 1000, depth = 50, passes = 0, stop_after = 0;
    bool resume = false;
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
        else if (a == "--gpus") opt.n_gpus = std::stoi(next());                    // Camera::render on N GPUs (rtw_render_multi)
        else if (a == "--collective") { std::string c = next(); opt.collective = c == "peer" ? RTW_COLLECTIVE_PEER : c == "nccl" ? RTW_COLLECTIVE_NCCL : RTW_COLLECTIVE_AUTO; }
        else if (a == "--mode") { std::string m = next(); opt.mode = m == "megakernel" ? RTW_MEGAKERNEL : RTW_WAVEFRONT; }
        else if (a == "--out") out = next();
        else if (a == "--config") config = next();
        else if (a == "--format") format = next();
        else if (a == "--passes") passes = std::stoul(next());
        else if (a == "--checkpoint") checkpoint = next();
        else if (a == "--resume") resume = true;
        else if (a == "--stop-after") stop_after = std::stoul(next());      // stop after this many passes (the checkpoint stays)
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
        double aspect = (double)width / (double)height;
        if (!config.empty()) {              // main.rs:57-65
            Image im = read_config(config);
            width = im.image_width; height = im.image_height; spp = im.samples_per_pixel; depth = im.max_depth; aspect = im.aspect_ratio;
        }
        // main.rs:72-79
        Camera cam = cb.with_vfov(40.).with_aspect_ratio(aspect).with_max_depth(depth)
                         .with_image_width(width).with_image_height(height).with_samples_per_pixel((uint16_t)spp).build();
        rtw_stats st{};
        auto t0 = std::chrono::steady_clock::now();
        World wref = general ? general_sc.world_ref() : World(simple_sc.world);
        World lref = general ? general_sc.lights_ref() : World(simple_sc.lights);
        auto img = passes ? cam.render_progressive(wref, lref, opt, passes, checkpoint, resume, &st, stop_after) : cam.render(wref, lref, opt, &st);
        double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        std::fprintf(stderr, "rendered %ux%u spp %u on %d GPU(s) in %.3f s (kernel %.3f ms): %.1f Mpaths/s, %.1f Mrays/s\n", width, height, spp,
                     opt.n_gpus, sec, st.kernel_ms, st.paths / st.kernel_ms * 1e-3, st.rays / st.kernel_ms * 1e-3);
        if (format == "p6") write_p6(out, img);
        else if (format == "png") write_png(out, img);
        else write_p3(out, img);
    } catch (const std::exception& e) {
        std::fprintf(stderr, "error: %s\n", e.what());
        return 1;
    }
    return 0;
}
