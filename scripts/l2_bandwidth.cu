// L2 bandwidth micro-benchmark for the BVH-fetch roofline (MEASURED_PEAKS.json has no L2 figure).
//   stream : every thread reads consecutive float4s of an L2-resident buffer (coalesced upper bound)
//   gather : every thread reads one 64-byte node (4 x float4) at a pseudo-random index — the access pattern of
//            incoherent BVH traversal out of L2 (scene of BASELINE config C4: 23 MB of nodes + 32 MB of spheres)
// Build + run on the GPU box:  nvcc -O3 -gencode arch=compute_100a,code=sm_100a scripts/l2_bandwidth.cu -o /tmp/l2bw && /tmp/l2bw
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void stream_kernel(const float4* __restrict__ p, size_t n, int reps, float* sink) {
    float acc = 0.f;
    for (int r = 0; r < reps; ++r)
        for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
            float4 v = __ldcg(p + i);
            acc += v.x + v.y + v.z + v.w;
        }
    if (acc == 123.456f) *sink = acc;
}
__global__ void gather_kernel(const float4* __restrict__ p, uint32_t n_nodes, int per_thread, float* sink) {
    uint32_t s = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
    float acc = 0.f;
    for (int k = 0; k < per_thread; ++k) {
        s = s * 1664525u + 1013904223u;
        const float4* q = p + (size_t)((s >> 8) % n_nodes) * 4;
        float4 a = __ldcg(q), b = __ldcg(q + 1), c = __ldcg(q + 2), d = __ldcg(q + 3);
        acc += a.x + b.y + c.z + d.w;
        s += __float_as_uint(acc) & 1u;          // make the next address depend on the data (like a traversal)
    }
    if (acc == 123.456f) *sink = acc;
}

int main() {
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    float* sink; cudaMalloc(&sink, 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    printf("{\"device\": \"%s\", \"l2_bytes\": %d, \"results\": [", prop.name, prop.l2CacheSize);
    bool first = true;
    for (size_t mb : {16, 32, 64, 96, 512}) {
        size_t bytes = mb << 20, n = bytes / 16;
        float4* buf; cudaMalloc(&buf, bytes); cudaMemset(buf, 0, bytes);
        int grid = prop.multiProcessorCount * 8, reps = mb <= 96 ? 20 : 4;
        stream_kernel<<<grid, 512>>>(buf, n, 2, sink);                       // warm L2
        cudaEventRecord(e0); stream_kernel<<<grid, 512>>>(buf, n, reps, sink); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double stream_gbs = (double)bytes * reps / (ms * 1e-3) / 1e9;
        uint32_t n_nodes = (uint32_t)(bytes / 64); int per_thread = 256;
        gather_kernel<<<grid, 512>>>(buf, n_nodes, 32, sink);
        cudaEventRecord(e0); gather_kernel<<<grid, 512>>>(buf, n_nodes, per_thread, sink); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        double gather_gbs = (double)grid * 512 * per_thread * 64 / (ms * 1e-3) / 1e9;
        printf("%s{\"buffer_mb\": %zu, \"stream_read_gbs\": %.0f, \"gather_64B_dependent_gbs\": %.0f}", first ? "" : ", ", mb, stream_gbs, gather_gbs);
        first = false;
        cudaFree(buf);
    }
    printf("]}\n");
    return 0;
}
