// rtw_launch.cuh — launcher bodies, instantiated once per arithmetic policy.
#pragma once
#include <cstdlib>
#include "rtw_launch.hpp"

namespace rtw {

constexpr int kRenderBlock = 256;
constexpr int kBatchBlock = 128;
constexpr size_t kSmemSceneBudget = 64 * 1024;   // per-CTA budget for staged scene data (fast path)

// decide which scene sections the fast kernels stage in shared memory; returns the dynamic smem size.
// A scene that fits as a whole is staged with its nodes at the padded stride (bank-conflict-free LDS.128, see kShNodeStridePadded)
// when that copy exists and still fits; RTW_SH_NODE_STRIDE=64 in the environment forces the unpadded copy (A/B measurements).
inline bool padded_nodes_allowed() {
    static const bool ok = [] { const char* e = std::getenv("RTW_SH_NODE_STRIDE"); return !(e && std::atoi(e) == 64); }();
    return ok;
}
template <class T, bool EXACT>
size_t plan_smem(RenderParams<T>& P, int block, bool* all_shared = nullptr, size_t scene_budget = kSmemSceneBudget) {
    if (P.stack_depth == 0 || P.stack_depth > (uint32_t)kStackDepth) P.stack_depth = kStackDepth;
    size_t smem = sizeof(int32_t) * P.stack_depth * block;
    P.smem_nodes = P.smem_spheres = P.smem_lights = 0;
    P.sh_node_stride = 64;
    if (all_shared) *all_shared = false;
    if (!EXACT) {
        size_t budget = scene_budget;
        size_t lights = (size_t)P.scene.n_lights * sizeof(Vec4T<T>);
        if (lights && lights <= budget && P.scene.n_light_nodes == 0) { P.smem_lights = (uint32_t)lights; budget -= lights; }
        size_t sph = (size_t)P.scene.n_spheres * sizeof(Vec4T<T>);
        size_t sph_total = 2 * sph + ((size_t)P.scene.n_spheres * 4 + 15) / 16 * 16;
        size_t nodes = (size_t)P.scene.n_nodes * sizeof(Node<T>);
        size_t nodes_padded = (size_t)P.scene.n_nodes * kShNodeStridePadded;
        const bool lights_ok = P.smem_lights || !P.scene.n_lights || P.scene.n_light_nodes > 0;
        if (nodes + sph_total <= budget) {
            P.smem_nodes = (uint32_t)nodes; P.smem_spheres = (uint32_t)sph;
            if (lights_ok && all_shared && P.scene.nodes_staged && nodes_padded + sph_total <= budget && padded_nodes_allowed()) {
                P.smem_nodes = (uint32_t)nodes_padded; P.sh_node_stride = kShNodeStridePadded;      // only the all-shared kernels know the stride
            }
        } else {
            // large scene: pin as many top levels (BFS prefix) as fit
            size_t top = std::min(nodes, budget) / sizeof(Node<T>) * sizeof(Node<T>);
            P.smem_nodes = (uint32_t)top;
        }
        if (all_shared)
            *all_shared = (P.smem_nodes == nodes || P.sh_node_stride != 64) && (P.smem_spheres || !P.scene.n_spheres) && lights_ok;
        smem += P.smem_nodes + (P.smem_spheres ? 2 * (size_t)P.smem_spheres + ((size_t)P.scene.n_spheres * 4 + 15) / 16 * 16 : 0) + P.smem_lights;
    }
    return smem;
}

template <class K>
cudaError_t persistent_grid(K kernel, int block, size_t smem, int sm_count, int* grid, LaunchInfo* info) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, block, smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    *grid = sm_count * per_sm;
    if (info) { info->grid = *grid; info->block = block; info->smem = smem; info->blocks_per_sm = per_sm; }
    return cudaSuccess;
}

template <class K, class P>
cudaError_t launch_persistent(K kernel, const P& params, int block, size_t smem, int sm_count, cudaStream_t s, LaunchInfo* info) {
    int grid = 0;
    cudaError_t e = persistent_grid(kernel, block, smem, sm_count, &grid, info);
    if (e != cudaSuccess) return e;
    kernel<<<grid, block, smem, s>>>(params);
    return cudaGetLastError();
}

template <class T, bool EXACT, bool COUNT>
cudaError_t launch_render_impl(RenderParams<T> P, int sm_count, cudaStream_t s, LaunchInfo* info) {
    bool sh = false;
    size_t smem = plan_smem<T, EXACT>(P, kRenderBlock, &sh);
    if constexpr (!EXACT) {
        if (sh) return launch_persistent(render_mega_kernel<T, EXACT, COUNT, kRenderBlock, true>, P, kRenderBlock, smem, sm_count, s, info);
    }
    return launch_persistent(render_mega_kernel<T, EXACT, COUNT, kRenderBlock, false>, P, kRenderBlock, smem, sm_count, s, info);
}

// general scenes: lane-per-pixel megakernel, scene in global memory (small tables: L1-resident)
template <class T, bool EXACT>
cudaError_t launch_render_general_t(RenderParams<T, SceneViewG<T>> P, bool count, int sm_count, cudaStream_t s, LaunchInfo* info) {
    if (P.stack_depth == 0 || P.stack_depth > (uint32_t)kStackDepth) P.stack_depth = kStackDepth;
    size_t smem = sizeof(int32_t) * P.stack_depth * kRenderBlock;
    if (count) return launch_persistent(render_mega_kernel<T, EXACT, true, kRenderBlock, false, SceneViewG<T>>, P, kRenderBlock, smem, sm_count, s, info);
    return launch_persistent(render_mega_kernel<T, EXACT, false, kRenderBlock, false, SceneViewG<T>>, P, kRenderBlock, smem, sm_count, s, info);
}
template <class T, bool EXACT> cudaError_t launch_trace_general_t(const BatchParams<T, SceneViewG<T>>& P, cudaStream_t s) {
    if (P.n == 0) return cudaSuccess;
    trace_batch_kernel<T, EXACT, kBatchBlock, SceneViewG<T>><<<(int)((P.n + kBatchBlock - 1) / kBatchBlock), kBatchBlock, 0, s>>>(P);
    return cudaGetLastError();
}
template <class T, bool EXACT> cudaError_t launch_scatter_general_t(const BatchParams<T, SceneViewG<T>>& P, cudaStream_t s) {
    if (P.n == 0) return cudaSuccess;
    scatter_batch_kernel<T, EXACT, kBatchBlock, SceneViewG<T>><<<(int)((P.n + kBatchBlock - 1) / kBatchBlock), kBatchBlock, 0, s>>>(P);
    return cudaGetLastError();
}
template <class T, bool EXACT> cudaError_t launch_path_radiance_general_t(const BatchParams<T, SceneViewG<T>>& P, cudaStream_t s) {
    if (P.n == 0) return cudaSuccess;
    path_radiance_kernel<T, EXACT, kBatchBlock, SceneViewG<T>><<<(int)((P.n + kBatchBlock - 1) / kBatchBlock), kBatchBlock, 0, s>>>(P);
    return cudaGetLastError();
}

template <class T, bool EXACT>
cudaError_t launch_render_t(RenderParams<T> P, bool count, int sm_count, cudaStream_t s, LaunchInfo* info) {
    return count ? launch_render_impl<T, EXACT, true>(P, sm_count, s, info) : launch_render_impl<T, EXACT, false>(P, sm_count, s, info);
}

inline cudaError_t launch_resolve_accum(const unsigned long long* accum, const uint32_t* poison, uint32_t width, uint32_t height, uint32_t spp,
                                        double* rgb_sum, uint8_t* rgb8, cudaStream_t s) {
    dim3 block(32, 8), grid((width + 31) / 32, (height + 7) / 8);
    resolve_accum_kernel<0><<<grid, block, 0, s>>>(accum, poison, width, height, (width + kTileW - 1) / kTileW, spp, rgb_sum, rgb8);
    return cudaGetLastError();
}
inline int batch_grid(size_t n) { return (int)((n + kBatchBlock - 1) / kBatchBlock); }

template <class T, bool EXACT> cudaError_t launch_trace_t(const BatchParams<T>& P, cudaStream_t s) {
    if (P.n == 0) return cudaSuccess;
    trace_batch_kernel<T, EXACT, kBatchBlock><<<batch_grid(P.n), kBatchBlock, 0, s>>>(P);
    return cudaGetLastError();
}
template <class T, bool EXACT> cudaError_t launch_scatter_t(const BatchParams<T>& P, cudaStream_t s) {
    if (P.n == 0) return cudaSuccess;
    scatter_batch_kernel<T, EXACT, kBatchBlock><<<batch_grid(P.n), kBatchBlock, 0, s>>>(P);
    return cudaGetLastError();
}
template <class T, bool EXACT> cudaError_t launch_shade_t(const ShadeParams<T>& P, cudaStream_t s) {
    if (P.n == 0) return cudaSuccess;
    shade_batch_kernel<T, EXACT, kBatchBlock><<<batch_grid(P.n), kBatchBlock, 0, s>>>(P);
    return cudaGetLastError();
}
template <class T, bool EXACT> cudaError_t launch_get_rays_t(const BatchParams<T>& P, double* o, double* d, cudaStream_t s) {
    if (P.n == 0) return cudaSuccess;
    get_rays_kernel<T, EXACT, kBatchBlock><<<batch_grid(P.n), kBatchBlock, 0, s>>>(P, o, d);
    return cudaGetLastError();
}
template <class T, bool EXACT> cudaError_t launch_path_radiance_t(const BatchParams<T>& P, cudaStream_t s) {
    if (P.n == 0) return cudaSuccess;
    path_radiance_kernel<T, EXACT, kBatchBlock><<<batch_grid(P.n), kBatchBlock, 0, s>>>(P);
    return cudaGetLastError();
}
template <class T>
cudaError_t launch_untile_t(const T* tiles, uint32_t width, uint32_t height, uint32_t world, uint32_t tiles_per_rank, uint32_t spp,
                            double* rgb_sum, uint8_t* rgb8, cudaStream_t s) {
    dim3 block(32, 8), grid((width + 31) / 32, (height + 7) / 8);
    uint32_t tiles_x = (width + kTileW - 1) / kTileW;
    untile_resolve_kernel<T><<<grid, block, 0, s>>>(tiles, width, height, world, tiles_per_rank, tiles_x, spp, rgb_sum, rgb8);
    return cudaGetLastError();
}

template <bool COUNT, bool SH>
cudaError_t launch_render_pool_sh(RenderParams<float> P, PoolParams Q, size_t smem, int sm_count, cudaStream_t s, LaunchInfo* info) {
    auto kernel = render_pool_kernel<COUNT, kRenderBlock, SH>;
    int grid = 0;
    cudaError_t e = persistent_grid(kernel, kRenderBlock, smem, sm_count, &grid, info);
    if (e != cudaSuccess) return e;
    kernel<<<grid, kRenderBlock, smem, s>>>(P, Q);
    return cudaGetLastError();
}

template <class PARAMS> inline cudaError_t pool_clear(const PARAMS& P, const PoolParams& Q, cudaStream_t s) {
    uint32_t n_slots = P.n_local_tiles * (kTileW * kTileH);
    cudaError_t e = cudaMemsetAsync(Q.accum, 0, (size_t)n_slots * 3 * sizeof(unsigned long long), s);
    if (e != cudaSuccess) return e;
    return cudaMemsetAsync(Q.poison, 0, (size_t)n_slots * sizeof(uint32_t), s);
}
template <class PARAMS> inline cudaError_t pool_finalize(const PARAMS& P, const PoolParams& Q, cudaStream_t s) {
    if (!P.tiles) return cudaSuccess;               // sample partition: the caller reduces and resolves the accumulators itself
    uint32_t n_slots = P.n_local_tiles * (kTileW * kTileH);
    pool_finalize_kernel<0><<<(n_slots + 255) / 256, 256, 0, s>>>(Q.accum, Q.poison, P.tiles, n_slots);
    return cudaGetLastError();
}

template <bool COUNT>
cudaError_t launch_render_pool_impl(RenderParams<float> P, PoolParams Q, int sm_count, cudaStream_t s, LaunchInfo* info) {
    bool sh = false;
    size_t smem = plan_smem<float, false>(P, kRenderBlock, &sh);
    cudaError_t e = pool_clear(P, Q, s);
    if (e != cudaSuccess) return e;
    e = sh ? launch_render_pool_sh<COUNT, true>(P, Q, smem, sm_count, s, info) : launch_render_pool_sh<COUNT, false>(P, Q, smem, sm_count, s, info);
    if (e != cudaSuccess) return e;
    return pool_finalize(P, Q, s);
}

// general scenes, FP32: the pooled path stream (fixed-point accumulation: the image does not depend on scheduling)
template <bool COUNT>
cudaError_t launch_render_pool_general_impl(RenderParams<float, SceneViewG<float>> P, PoolParams Q, int sm_count, cudaStream_t s, LaunchInfo* info) {
    if (P.stack_depth == 0 || P.stack_depth > (uint32_t)kStackDepth) P.stack_depth = kStackDepth;
    size_t smem = sizeof(int32_t) * P.stack_depth * kRenderBlock;
    cudaError_t e = pool_clear(P, Q, s);
    if (e != cudaSuccess) return e;
    auto kernel = render_pool_kernel<COUNT, kRenderBlock, false, SceneViewG<float>>;
    int grid = 0;
    e = persistent_grid(kernel, kRenderBlock, smem, sm_count, &grid, info);
    if (e != cudaSuccess) return e;
    kernel<<<grid, kRenderBlock, smem, s>>>(P, Q);
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    return pool_finalize(P, Q, s);
}

#define RTW_DEFINE_LAUNCHERS(SUFFIX, T, EXACT)                                                                                   \
    cudaError_t launch_render_##SUFFIX(RenderParams<T> P, bool count, int sm_count, cudaStream_t s, LaunchInfo* info) {           \
        return launch_render_t<T, EXACT>(P, count, sm_count, s, info);                                                            \
    }                                                                                                                            \
    cudaError_t launch_trace_##SUFFIX(const BatchParams<T>& P, cudaStream_t s) { return launch_trace_t<T, EXACT>(P, s); }         \
    cudaError_t launch_scatter_##SUFFIX(const BatchParams<T>& P, cudaStream_t s) { return launch_scatter_t<T, EXACT>(P, s); }     \
    cudaError_t launch_shade_##SUFFIX(const ShadeParams<T>& P, cudaStream_t s) { return launch_shade_t<T, EXACT>(P, s); }         \
    cudaError_t launch_get_rays_##SUFFIX(const BatchParams<T>& P, double* o, double* d, cudaStream_t s) {                         \
        return launch_get_rays_t<T, EXACT>(P, o, d, s);                                                                           \
    }                                                                                                                            \
    cudaError_t launch_path_radiance_##SUFFIX(const BatchParams<T>& P, cudaStream_t s) { return launch_path_radiance_t<T, EXACT>(P, s); } \
    cudaError_t launch_render_general_##SUFFIX(RenderParams<T, SceneViewG<T>> P, bool count, int sm_count, cudaStream_t s, LaunchInfo* info) { \
        return launch_render_general_t<T, EXACT>(P, count, sm_count, s, info);                                                    \
    }                                                                                                                            \
    cudaError_t launch_trace_general_##SUFFIX(const BatchParams<T, SceneViewG<T>>& P, cudaStream_t s) { return launch_trace_general_t<T, EXACT>(P, s); } \
    cudaError_t launch_scatter_general_##SUFFIX(const BatchParams<T, SceneViewG<T>>& P, cudaStream_t s) { return launch_scatter_general_t<T, EXACT>(P, s); } \
    cudaError_t launch_path_radiance_general_##SUFFIX(const BatchParams<T, SceneViewG<T>>& P, cudaStream_t s) { return launch_path_radiance_general_t<T, EXACT>(P, s); } \
    cudaError_t launch_untile_##SUFFIX(const T* tiles, uint32_t width, uint32_t height, uint32_t world, uint32_t tiles_per_rank,  \
                                       uint32_t spp, double* rgb_sum, uint8_t* rgb8, cudaStream_t s) {                            \
        return launch_untile_t<T>(tiles, width, height, world, tiles_per_rank, spp, rgb_sum, rgb8, s);                            \
    }

}  // namespace rtw
