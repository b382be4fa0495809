// rtw_launch.hpp — host-callable launchers exported by kernels_f32.cu (fast) and kernels_f64.cu (exact).
#pragma once
#include "rtw_kernels.cuh"

namespace rtw {

struct LaunchInfo { int grid = 0, block = 0; size_t smem = 0; int blocks_per_sm = 0; };

#define RTW_DECLARE_LAUNCHERS(SUFFIX, T)                                                                         \
    cudaError_t launch_render_##SUFFIX(RenderParams<T> P, bool count, int sm_count, cudaStream_t s, LaunchInfo*); \
    cudaError_t launch_trace_##SUFFIX(const BatchParams<T>& P, cudaStream_t s);                                   \
    cudaError_t launch_scatter_##SUFFIX(const BatchParams<T>& P, cudaStream_t s);                                 \
    cudaError_t launch_shade_##SUFFIX(const ShadeParams<T>& P, cudaStream_t s);                                   \
    cudaError_t launch_get_rays_##SUFFIX(const BatchParams<T>& P, double* o, double* d, cudaStream_t s);          \
    cudaError_t launch_path_radiance_##SUFFIX(const BatchParams<T>& P, cudaStream_t s);                           \
    cudaError_t launch_render_general_##SUFFIX(RenderParams<T, SceneViewG<T>> P, bool count, int sm_count, cudaStream_t s, LaunchInfo*); \
    cudaError_t launch_trace_general_##SUFFIX(const BatchParams<T, SceneViewG<T>>& P, cudaStream_t s);            \
    cudaError_t launch_scatter_general_##SUFFIX(const BatchParams<T, SceneViewG<T>>& P, cudaStream_t s);          \
    cudaError_t launch_path_radiance_general_##SUFFIX(const BatchParams<T, SceneViewG<T>>& P, cudaStream_t s);    \
    cudaError_t launch_untile_##SUFFIX(const T* tiles, uint32_t width, uint32_t height, uint32_t world,           \
                                       uint32_t tiles_per_rank, uint32_t spp, double* rgb_sum, uint8_t* rgb8, cudaStream_t s);

RTW_DECLARE_LAUNCHERS(f32, float)
cudaError_t launch_render_pool_f32(RenderParams<float> P, PoolParams Q, bool count, int sm_count, cudaStream_t s, LaunchInfo* info);
uint32_t pool_pixels_per_chunk(uint32_t spp);
float pool_sample_cap(uint32_t spp_total);
// scratch: primary_candidates_scratch_bytes(width, height) bytes of device memory (the per-block lists of the first level)
size_t primary_candidates_scratch_bytes(uint32_t width, uint32_t height);
cudaError_t launch_primary_candidates_f32(const SceneView<float>& scene, const CameraT<float>& cam, void* scratch, uint4* cand, cudaStream_t s);
// order: chunk_order_words(n_chunks, cap) words, cap = n_chunks + chunk_order_extra(split_chunks) queue positions (layout: order_words() in
// rtw_kernels.cuh); own_rank / own_world: this GPU's share of the chunks (0 / 1: all of them)
size_t chunk_order_words(uint32_t n_chunks, uint32_t cap);
uint32_t chunk_order_extra(uint32_t split_chunks);
cudaError_t launch_chunk_order_f32(const uint4* cand, const SceneView<float>& scene, const CameraT<float>& cam, uint32_t rank, uint32_t world,
                                   uint32_t tiles_x, uint32_t tiles_total, uint32_t n_slots, uint32_t pixels_per_chunk, uint32_t n_chunks, uint32_t cap,
                                   uint32_t own_rank, uint32_t own_world, uint32_t* order, cudaStream_t s);
// after launch_chunk_order_f32 (always): write the queue — the costly chunks (the last split_chunks of them in pieces), the last tail_chunks
// background-only chunks, end marks; the rest of the background-only chunks is what launch_render_background_f32 renders — with the same
// RenderParams / PoolParams as the wavefront (order + cap + 2 n_chunks + 2: queue length)
cudaError_t launch_chunk_split_f32(uint32_t* order, uint32_t n_chunks, uint32_t cap, uint32_t tail_chunks, uint32_t split_chunks, cudaStream_t s);
// the calling thread's next wavefront launches run render_background_kernel on `stream`, forked from / joined to the render stream with the two
// events (all NULL: on the render stream itself, after the wavefront kernel)
void set_background_side_stream(cudaStream_t stream, cudaEvent_t fork, cudaEvent_t join);
cudaError_t launch_render_background_f32(const RenderParams<float>& P, const PoolParams& Q, const uint32_t* order, int sm_count, cudaStream_t s);
cudaError_t launch_render_pool_general_f32(RenderParams<float, SceneViewG<float>> P, PoolParams Q, bool count, int sm_count, cudaStream_t s, LaunchInfo* info);
cudaError_t launch_render_wavefront_f32(RenderParams<float> P, PoolParams Q, uint32_t bvh_depth, bool count, int sm_count, cudaStream_t s, LaunchInfo* info);
uint32_t wavefront_max_bvh_depth();
cudaError_t launch_render_wavefront_general_f32(RenderParams<float, SceneViewG<float>> P, PoolParams Q, uint32_t bvh_depth, bool count, int sm_count,
                                                cudaStream_t s, LaunchInfo* info);
cudaError_t launch_resolve_accum_f32(const unsigned long long* accum, const uint32_t* poison, uint32_t width, uint32_t height, uint32_t spp,
                                     double* rgb_sum, uint8_t* rgb8, cudaStream_t s);
cudaError_t launch_peer_reduce_resolve_f32(const PeerBlocks& B, uint32_t slot_begin, uint32_t slot_end, uint32_t width, uint32_t height,
                                           uint32_t spp, double* rgb_sum, uint8_t* rgb8, cudaStream_t s);
RTW_DECLARE_LAUNCHERS(f64, double)

}  // namespace rtw
