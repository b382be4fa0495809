// bvh_device.hpp — device-side LBVH construction (bvh_device.cu), SURVEY §8 row f4.
#pragma once
#include "rtw_device.cuh"

namespace rtw {

struct DeviceBvhInfo { uint32_t inner_nodes = 0, leaves = 0, depth = 0, node_slots = 0; };

// d_spheres / d_mats: [n] double4 in INPUT order ((cx, cy, cz, r) / (albedo, param)); d_info: [n] prim_id << 2 | kind.
// lo / hi: bounds of the sphere centres.  Writes out->node_slots (<= n - 1) nodes of both precisions in breadth-first level order
// (root = 0; the output arrays must hold n - 1 nodes) and the Morton-sorted sphere / material / info arrays the leaf ranges index.
cudaError_t build_lbvh_device(const double* d_spheres, const double* d_mats, const uint32_t* d_info, size_t n, const double lo[3], const double hi[3],
                              int max_leaf, Node<double>* nodes64, Node<float>* nodes32, Vec4T<double>* s64, Vec4T<double>* m64,
                              Vec4T<float>* s32, Vec4T<float>* m32, uint32_t* info_sorted, DeviceBvhInfo* out, cudaStream_t st);

}  // namespace rtw
