// CPU-side check of the C++ host mirror's scene generators (no GPU involved): writes the plain-data description of one scene —
// exactly what it would hand to rtw_scene_create_general — to a file, so that the test can compare it byte for byte with the
// Python mirror's (scenes.py) description of the same scene.
// usage: dump_scenes <scene> <out_file>
#include <cstdio>
#include <cstring>
#include "../../ray_tracing_weekend_b200/host/rtw_host.hpp"
using namespace rtw_host;
template <class V> static void put(FILE* f, const V& v) {
    uint64_t n = v.size();
    std::fwrite(&n, sizeof(n), 1, f);
    if (n) std::fwrite(v.data(), sizeof(v[0]), n, f);
}
int main(int argc, char** argv) {
    if (argc < 3) return 2;
    const std::string name = argv[1];
    const uint64_t seed = 20261018;
    scenes::GeneralOutput g;
    if (name == "simple_light") g = scenes::simple_light(seed);
    else if (name == "cornell_box") g = scenes::cornell_box();
    else if (name == "debugging_scene") g = scenes::debugging_scene(seed);
    else if (name == "simple_transform") g = scenes::simple_transform(seed);
    else if (name == "checkered_spheres") g = scenes::checkered_spheres();
    else if (name == "plane") g = scenes::plane();
    else return 3;
    SceneDescription d(g.world_ref(), g.lights_ref());
    FILE* f = std::fopen(argv[2], "wb");
    if (!f) return 4;
    put(f, d.spheres); put(f, d.planes); put(f, d.quads); put(f, d.cuboids); put(f, d.transforms); put(f, d.materials);
    put(f, d.textures); put(f, d.perlins); put(f, d.world); put(f, d.lights);
    const uint32_t flags[2] = {d.pod.world_is_bvh, d.pod.lights_is_bvh};
    std::fwrite(flags, sizeof(flags), 1, f);
    rtw_camera_builder cb = g.cam.pod();
    std::fwrite(&cb, sizeof(cb), 1, f);
    std::fclose(f);
    return 0;
}
