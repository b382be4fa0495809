// CPU-side check of the C++ host mirror's image writers and Config.toml reader (no GPU involved).
// usage: writers_and_config <out_dir> <Config.toml>
#include <cstdio>
#include "../../ray_tracing_weekend_b200/host/rtw_host.hpp"
using namespace rtw_host;
int main(int argc, char** argv) {
    if (argc < 3) return 2;
    std::string dir = argv[1];
    const uint32_t w = 5, h = 3;
    std::vector<std::vector<SampledColour>> img(h, std::vector<SampledColour>(w));
    for (uint32_t j = 0; j < h; ++j)          // row j = 0 is the BOTTOM row, like Camera::render's output
        for (uint32_t i = 0; i < w; ++i) { img[j][i].rgb[0] = (uint8_t)(10 * i); img[j][i].rgb[1] = (uint8_t)(100 + j); img[j][i].rgb[2] = (uint8_t)(i * j + 7); }
    write_p3(dir + "/a.ppm", img); write_p6(dir + "/b.ppm", img); write_png(dir + "/c.png", img);
    try {
        Image im = read_config(argv[2]);
        std::printf("%.17g %u %u %u %u\n", im.aspect_ratio, im.image_width, im.image_height, (unsigned)im.samples_per_pixel, (unsigned)im.max_depth);
    } catch (const std::exception& e) { std::printf("error: %s\n", e.what()); }
    return 0;
}
