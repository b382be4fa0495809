/* rtw_host.h — C view of the host-side scene generator that stands in for the reference's `scenes`
 * crate (scenes/src/lib.rs) while no Rust toolchain is available.  Host only, no GPU needed.
 * In the Rust workspace this is NOT bound: `scenes::simple` stays Rust and feeds rtw.h directly. */
#ifndef RTW_HOST_H
#define RTW_HOST_H
#include "rtw.h"
#ifdef __cplusplus
extern "C" {
#endif

typedef struct rtwh_scene_desc rtwh_scene_desc;

/* scenes::simple (scenes/src/lib.rs:155-233), seeded.  Reference parameters: n = 11,
 * p_lambertian = 0.8, p_metal = 0.95, ground = 0 (one-sided Plane); ground 1 = book-1 ground sphere,
 * ground 2 = none. */
RTW_API rtwh_scene_desc* rtwh_scene_simple(uint64_t seed, int32_t n, double p_lambertian, double p_metal, int32_t ground);
RTW_API void rtwh_scene_desc_destroy(rtwh_scene_desc* d);
/* counts: spheres, planes, lights */
RTW_API void rtwh_scene_desc_counts(const rtwh_scene_desc* d, uint64_t out[3]);
/* one material per primitive (sphere_material[i] = i, plane materials follow the spheres') */
RTW_API void rtwh_scene_desc_copy(const rtwh_scene_desc* d, rtw_sphere* spheres, rtw_material* sphere_materials,
                                  rtw_plane* planes, rtw_material* plane_materials, rtw_sphere* lights,
                                  rtw_camera_builder* camera);
#ifdef __cplusplus
}
#endif
#endif
