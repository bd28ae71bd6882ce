// lbvh.cuh — SURVEY.md §8f rank 1: the device-walked tree built ON the device (LBVH: Morton order + Karras'
// hierarchy + bottom-up boxes), for producers that cannot afford the host SAH build (1 s for 1 M triangles) or whose
// geometry changes.  It replaces only the tree `k_trace_fast` walks INSIDE a mesh; the reference's decisions are
// untouched — every candidate is still admitted by the reference's triangle test and its `tri_gate` box, and the
// winner is still the minimum of (t, top-level rank, triangle DFS index) — so ids and t stay bit-identical
// (tests/test_gpu_parity.py::test_device_built_tree_gives_the_same_answers).  LBVH quality is below binned SAH
// (one triangle per leaf, no surface-area reasoning): it is an option (b200rt_scene_build_bvh_device), not the default.
//
// The pure index arithmetic is __host__ __device__ so that tests/test_abi.py can check it on the CPU.
#pragma once
#include <cstdint>

#include "../../include/b200rt.h"

#ifdef __CUDACC__
#define LBVH_HD __host__ __device__
#else
#define LBVH_HD
#endif

// 21 bits per axis, interleaved into a 63-bit code
LBVH_HD inline uint64_t lbvh_expand21(uint64_t v) {
  v &= 0x1FFFFFull;
  v = (v | (v << 32)) & 0x1F00000000FFFFull;
  v = (v | (v << 16)) & 0x1F0000FF0000FFull;
  v = (v | (v << 8)) & 0x100F00F00F00F00Full;
  v = (v | (v << 4)) & 0x10C30C30C30C30C3ull;
  v = (v | (v << 2)) & 0x1249249249249249ull;
  return v;
}
LBVH_HD inline uint64_t lbvh_morton63(float x, float y, float z) {  // x, y, z in [0, 1]
  const float s = 2097152.0f, top = 2097151.0f;
  const uint64_t xi = (uint64_t)(x * s < 0.f ? 0.f : (x * s > top ? top : x * s));
  const uint64_t yi = (uint64_t)(y * s < 0.f ? 0.f : (y * s > top ? top : y * s));
  const uint64_t zi = (uint64_t)(z * s < 0.f ? 0.f : (z * s > top ? top : z * s));
  return (lbvh_expand21(xi) << 2) | (lbvh_expand21(yi) << 1) | lbvh_expand21(zi);
}
LBVH_HD inline int lbvh_clz64(uint64_t v) {
#ifdef __CUDA_ARCH__
  return __clzll((long long)v);
#else
  return v ? __builtin_clzll(v) : 64;
#endif
}
// length of the common prefix of sorted keys i and j, -1 outside [0, n); equal codes are told apart by their
// positions (Karras 2012, section 4), so duplicate Morton codes still give a proper binary tree
LBVH_HD inline int lbvh_delta(const uint64_t* keys, int n, int i, int j) {
  if (j < 0 || j >= n) return -1;
  const uint64_t x = keys[i] ^ keys[j];
  if (x == 0) return 64 + lbvh_clz64((uint64_t)(uint32_t)(i ^ j)) - 32;
  return lbvh_clz64(x);
}
// Karras 2012: internal node i of n-1 covers the key range [first, last] and splits after `split`;
// its children are internal node `split` / leaf `split` and internal node / leaf `split + 1`.
LBVH_HD inline void lbvh_range(const uint64_t* keys, int n, int i, int* first, int* last, int* split) {
  const int d = (lbvh_delta(keys, n, i, i + 1) - lbvh_delta(keys, n, i, i - 1)) >= 0 ? 1 : -1;
  const int dmin = lbvh_delta(keys, n, i, i - d);
  int lmax = 2;
  while (lbvh_delta(keys, n, i, i + lmax * d) > dmin) lmax *= 2;
  int l = 0;
  for (int t = lmax / 2; t >= 1; t /= 2)
    if (lbvh_delta(keys, n, i, i + (l + t) * d) > dmin) l += t;
  const int j = i + l * d;
  const int dnode = lbvh_delta(keys, n, i, j);
  int s = 0;
  for (int t = (l + 1) / 2;; t = (t + 1) / 2) {
    if (lbvh_delta(keys, n, i, i + (s + t) * d) > dnode) s += t;
    if (t == 1) break;
  }
  const int g = i + s * d + (d < 0 ? -1 : 0);
  *first = i < j ? i : j;
  *last = i < j ? j : i;
  *split = g;
}

#ifdef __CUDACC__
// ---- kernels (one mesh at a time: triangles [tri_begin, tri_begin + n) of S.tris, nodes [node_base, node_base + n - 1)) ----
__global__ void k_lbvh_keys(const float4* tris, int tri_begin, int n, float3 lo, float3 inv_ext, uint64_t* keys,
                            uint32_t* idx) {
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
    const float4 a = tris[3 * (size_t)(tri_begin + k)], b = tris[3 * (size_t)(tri_begin + k) + 1],
                 c = tris[3 * (size_t)(tri_begin + k) + 2];
    // v0, a1 = v0 - v1, a2 = v0 - v2: centroid = v0 - (a1 + a2) / 3
    const float cx = a.x - (a.w + b.z) * (1.0f / 3.0f), cy = a.y - (b.x + b.w) * (1.0f / 3.0f),
                cz = a.z - (b.y + c.x) * (1.0f / 3.0f);
    keys[k] = lbvh_morton63((cx - lo.x) * inv_ext.x, (cy - lo.y) * inv_ext.y, (cz - lo.z) * inv_ext.z);
    idx[k] = (uint32_t)k;
  }
}
// leaves: gather the triangles into Morton order, their boxes; parent links are filled by k_lbvh_hierarchy
__global__ void k_lbvh_leaves(const float4* tris, int tri_begin, int n, const uint32_t* order, int leaf_base,
                              float4* fast_tris, int* fast_tri_ref, float* leaf_box) {
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
    const int t = tri_begin + (int)order[k];
    const float4 a = tris[3 * (size_t)t], b = tris[3 * (size_t)t + 1], c = tris[3 * (size_t)t + 2];
    fast_tris[3 * (size_t)(leaf_base + k)] = a;
    fast_tris[3 * (size_t)(leaf_base + k) + 1] = b;
    fast_tris[3 * (size_t)(leaf_base + k) + 2] = c;
    fast_tri_ref[leaf_base + k] = t;
    const float v0[3] = {a.x, a.y, a.z}, a1[3] = {a.w, b.x, b.y}, a2[3] = {b.z, b.w, c.x};
    for (int q = 0; q < 3; q++) {
      const float p1 = v0[q] - a1[q], p2 = v0[q] - a2[q];
      leaf_box[6 * (size_t)k + q] = fminf(v0[q], fminf(p1, p2));
      leaf_box[6 * (size_t)k + 3 + q] = fmaxf(v0[q], fmaxf(p1, p2));
    }
  }
}
__global__ void k_lbvh_hierarchy(const uint64_t* keys, int n, int node_base, int leaf_base, b200rt_bvh_node* nodes,
                                 int* parent /* 2n-1 entries: internal [0, n-1), leaves [n-1, 2n-1); (node << 1) | side */) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n - 1; i += gridDim.x * blockDim.x) {
    int first, last, split;
    lbvh_range(keys, n, i, &first, &last, &split);
    const bool left_leaf = split == first, right_leaf = split + 1 == last;
    b200rt_bvh_node& nd = nodes[node_base + i];
    nd.left = left_leaf ? B200RT_REF(B200RT_KIND_LEAF1, leaf_base + split) : B200RT_REF(B200RT_KIND_FNODE, node_base + split);
    nd.right = right_leaf ? B200RT_REF(B200RT_KIND_LEAF1, leaf_base + split + 1)
                          : B200RT_REF(B200RT_KIND_FNODE, node_base + split + 1);
    nd.pad[0] = 0; nd.pad[1] = 0;
    parent[left_leaf ? (n - 1 + split) : split] = (i << 1);
    parent[right_leaf ? (n - 1 + split + 1) : (split + 1)] = (i << 1) | 1;
    if (i == 0) parent[0] = -1;
  }
}
// bottom-up: the second thread to arrive at a node owns it (both children are complete then)
__global__ void k_lbvh_refit(int n, int node_base, const int* parent, const float* leaf_box, b200rt_bvh_node* nodes,
                             unsigned int* visits, float grow) {
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
    float box[6];
    for (int q = 0; q < 6; q++) box[q] = leaf_box[6 * (size_t)k + q];
    int code = parent[n - 1 + k];
    while (code >= 0) {
      const int p = code >> 1;
      volatile b200rt_bvh_node* nd = nodes + node_base + p;
      volatile float* mn = (code & 1) ? nd->rmin : nd->lmin;
      volatile float* mx = (code & 1) ? nd->rmax : nd->lmax;
      for (int q = 0; q < 3; q++) { mn[q] = box[q] - grow; mx[q] = box[3 + q] + grow; }   // child boxes are stored grown
      __threadfence();
      if (atomicAdd(&visits[p], 1u) == 0u) break;   // the sibling is not done yet: it will carry on from here
      __threadfence();
      for (int q = 0; q < 3; q++) {
        box[q] = fminf(nd->lmin[q], nd->rmin[q]) + grow;       // the node's own tight box
        box[3 + q] = fmaxf(nd->lmax[q], nd->rmax[q]) - grow;
      }
      code = parent[p];
    }
  }
}
#endif
