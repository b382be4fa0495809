// rtw_host_capi.cpp — exports the C++ host mirror's scene generator (rtw_host::scenes::simple) as
// plain arrays so the Python harness (tests, bench) can build the same scene the C++ CLI renders.
#include "../../include/rtw_host.h"
#include "rtw_host.hpp"

struct rtwh_scene_desc { rtw_host::scenes::Output out; };

extern "C" {

rtwh_scene_desc* rtwh_scene_simple(uint64_t seed, int32_t n, double p_lambertian, double p_metal, int32_t ground) {
    return new rtwh_scene_desc{rtw_host::scenes::simple(seed, n, p_lambertian, p_metal, ground)};
}
void rtwh_scene_desc_destroy(rtwh_scene_desc* d) { delete d; }
void rtwh_scene_desc_counts(const rtwh_scene_desc* d, uint64_t out[3]) {
    out[0] = d->out.world.list().spheres().size(); out[1] = d->out.world.list().planes().size(); out[2] = d->out.lights.spheres().size();
}
void rtwh_scene_desc_copy(const rtwh_scene_desc* d, rtw_sphere* spheres, rtw_material* sphere_materials, rtw_plane* planes,
                          rtw_material* plane_materials, rtw_sphere* lights, rtw_camera_builder* camera) {
    size_t k = 0;
    for (const auto& s : d->out.world.list().spheres()) { spheres[k] = {s.center.x, s.center.y, s.center.z, s.radius}; sphere_materials[k] = s.mat->pod; ++k; }
    k = 0;
    for (const auto& p : d->out.world.list().planes()) { planes[k] = {p.point.x, p.point.y, p.point.z, p.normal.x, p.normal.y, p.normal.z}; plane_materials[k] = p.mat->pod; ++k; }
    k = 0;
    for (const auto& s : d->out.lights.spheres()) lights[k++] = {s.center.x, s.center.y, s.center.z, s.radius};
    if (camera) *camera = d->out.cam.pod();
}

}  // extern "C"
