// kernels_f32.cu — the fast (FP32) instantiation of every kernel.  Compiled with FMA contraction on.
#include <algorithm>
#include "rtw_launch.cuh"
namespace rtw {
RTW_DEFINE_LAUNCHERS(f32, float, false)
cudaError_t launch_render_pool_f32(RenderParams<float> P, PoolParams Q, bool count, int sm_count, cudaStream_t s, LaunchInfo* info) {
    return count ? launch_render_pool_impl<true>(P, Q, sm_count, s, info) : launch_render_pool_impl<false>(P, Q, sm_count, s, info);
}
// chunk = G pixel slots x spp paths; aim for >= 512 paths per queue transaction
uint32_t pool_pixels_per_chunk(uint32_t spp) {
    if (spp == 0) return 1;
    uint32_t g = (512 + spp - 1) / spp;
    return g < 1 ? 1 : (g > 256 ? 256 : g);
}
}
