// TEST INFRASTRUCTURE ONLY — the default traversal's per-lane source (csrc/rt_wide.cuh: top-level gate mask, object
// selection, instance entry, the 8-wide node step with its quantised boxes, leaf admission through tri_gate) and the
// wide-tree builder (csrc/wide_bvh.cpp) compiled for the HOST, so that the whole closest-hit / any-hit answer of a ray
// can be compared with the oracle's on a box without a GPU.  What differs on the device is only the warp scheduling
// around these functions (k_trace_wide).  Nothing in the product links this file: tests/test_wide_bvh.py compiles it
// into build/libwide_host.so with g++ (-ffp-contract=off, the host twin of -fmad=false).
#include <cuda_runtime.h>   // float4, uint2, make_float4 ... (host-compatible header)

#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>
#undef __noinline__
#define __noinline__
template <class T> static inline T __ldg(const T* p) { return *p; }
template <class T> static inline T __ldcs(const T* p) { return *p; }
static inline float __int_as_float(int i) { float f; std::memcpy(&f, &i, 4); return f; }
static inline float __uint_as_float(unsigned i) { float f; std::memcpy(&f, &i, 4); return f; }
static inline unsigned __float_as_uint(float f) { unsigned i; std::memcpy(&i, &f, 4); return i; }
static inline int __float_as_int(float f) { int i; std::memcpy(&i, &f, 4); return i; }
static inline float __fmaf_rn(float a, float b, float c) { return std::fmaf(a, b, c); }
static inline int __clz(int x) { return x == 0 ? 32 : __builtin_clz((unsigned)x); }
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline int __ffs(int x) { return __builtin_ffs(x); }
static inline int __ffsll(long long x) { return __builtin_ffsll(x); }

#include "../ceng795_b200/csrc/rt_wide.cuh"
#include "../ceng795_b200/csrc/wide_bvh.cpp"

extern "C" {

// rays: 8 floats per ray (o.xyz, time, d.xyz, tlimit — tlimit only read when any != 0).
// ids_out[2k] = top-level rank (-1 miss), ids_out[2k+1] = triangle index (-1); t_out[k] = t (inf on miss).
// any != 0: ids_out[2k] = 1 when something lies in (0, tlimit), else 0.
// stats_out[0..5] = node visits, triangle tests, deepest stack, wide nodes, wide depth, top-level objects.
int widehost_trace(const B200SceneDesc* d, const float* rays, int64_t n, int any, int32_t* ids_out, float* t_out,
                   int64_t* stats_out) {
  if (!d || d->semantics == B200RT_SEM_HW1) return -1;
  std::vector<float> flat;
  if (!b200rt_wide::flatten_top(d, flat)) return -2;
  const int n_top = (int)(flat.size() / 8);
  if (n_top == 0 || n_top > RT_TOP_LINEAR_MAX) return -3;
  if (d->n_fast_tris == 0 && d->n_tris > 1) return -4;   // the description must carry the SAH tree (the builder's does)

  std::vector<b200rt_wide::Node> nodes;
  std::vector<int32_t> order;
  std::vector<int> wroot((size_t)d->n_meshes, -1);
  b200rt_wide::Stats ws = {0, 0, 0, 0};
  for (int m = 0; m < d->n_meshes; m++) {
    const b200rt_mesh& M = d->meshes[m];
    if (M.tri_count < 2 || M.fast_root == B200RT_REF_NONE) continue;
    float ext = 0.f;
    for (int a = 0; a < 3; a++) {
      ext = fmaxf(ext, M.root_box[3 + a] - M.root_box[a]);
      ext = fmaxf(ext, fmaxf(fabsf(M.root_box[a]), fabsf(M.root_box[3 + a])));
    }
    const float grow = ext * (3e-5f + 4.0f * d->isect_eps) + 1e-30f;
    wroot[m] = b200rt_wide::build(d->fast_nodes, M.fast_root, d->fast_tri_geom, grow, nodes, order, &ws);
  }
  std::vector<float4> wtris(order.size() * 3 + 3);
  std::vector<int> wref(order.size() + 1);
  for (size_t k = 0; k < order.size(); k++) {
    std::memcpy(&wtris[3 * k], d->fast_tri_geom + 12 * (size_t)order[k], 48);
    wref[k] = d->fast_tri_ref[order[k]];
  }

  DScene S;
  std::memset(&S, 0, sizeof(S));
  S.semantics = d->semantics; S.isect_eps = d->isect_eps; S.root = d->root;
  std::memcpy(S.root_box, d->root_box, sizeof(S.root_box));
  S.n_top = n_top;
  S.tris = reinterpret_cast<const float4*>(d->tri_geom);
  S.tri_gate = d->tri_gate;
  S.wide_nodes = reinterpret_cast<const float4*>(nodes.data());
  S.wide_tris = wtris.data();
  S.wide_tri_ref = wref.data();
  S.wide_root = wroot.data();
  S.meshes = d->meshes; S.inst = d->instances; S.spheres = d->spheres; S.triprims = d->triprims;
  const float4* top = reinterpret_cast<const float4*>(flat.data());
  const bool mode7 = d->semantics >= B200RT_SEM_HW4;

  TravCounters tc = {0, 0, 0, 0};
  int max_sp = 0;
  unsigned long long nodes_total = 0, tris_total = 0;
  for (int64_t k = 0; k < n; k++) {
    const float* r = rays + 8 * k;
    const v3 o = mk3(r[0], r[1], r[2]), dir = mk3(r[4], r[5], r[6]);
    HitRec h;
    tc.nodes = 0; tc.tris = 0;
    bool f;
    if (any) {
      f = mode7 ? wide_trace_one<true, 7>(S, top, o, dir, r[3], r[7], h, tc, &max_sp)
                : wide_trace_one<true, 2>(S, top, o, dir, r[3], r[7], h, tc, &max_sp);
      ids_out[2 * k] = f ? 1 : 0; ids_out[2 * k + 1] = -1; t_out[k] = 0.f;
    } else {
      f = mode7 ? wide_trace_one<false, 7>(S, top, o, dir, r[3], RT_INF, h, tc, &max_sp)
                : wide_trace_one<false, 2>(S, top, o, dir, r[3], RT_INF, h, tc, &max_sp);
      ids_out[2 * k] = f ? h.top_rank : -1; ids_out[2 * k + 1] = f ? h.tri : -1; t_out[k] = f ? h.t : RT_INF;
    }
    nodes_total += tc.nodes; tris_total += tc.tris;
  }
  if (stats_out) {
    stats_out[0] = (int64_t)nodes_total; stats_out[1] = (int64_t)tris_total; stats_out[2] = max_sp;
    stats_out[3] = (int64_t)nodes.size(); stats_out[4] = ws.depth; stats_out[5] = n_top;
  }
  return 0;
}

}
