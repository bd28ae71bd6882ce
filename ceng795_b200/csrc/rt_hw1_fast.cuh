// rt_hw1_fast.cuh — HW1 without the brute force, with HW1's order-dependent answer.
//
// HW1 has no BVH: Scene::trace_ray walks every triangle, then every sphere, and keeps a hit iff
// `t < best - 1e-6 && t > 0` (HW1/Scene.cpp:35-54).  That rule depends on the ORDER of primitives whose t lie within
// 1e-6 of each other, so "closest hit through a BVH" is not the same function.  It is, on a small set:
//
//   Sort the hits of a ray by t: t_1 <= t_2 <= ...  Let j be the first index with  t_j < fl(t_{j+1} - eps)  (a gap),
//   S = {1..j}.  (a) A member of S met while `best` is outside S (or inf) is always taken: t_s <= t_j < fl(t_k - eps)
//   for every k outside S (fl(. - eps) is monotone).  (b) Once `best` is in S no outsider is taken: t_k >= t_s >
//   fl(t_s - eps).  So `best` ends in S and the members of S are accepted or rejected exactly as if the loop ran
//   over S alone, in index order.
//
// The kernel walks a SAH tree over all triangles (built by b200rt_scene_create), keeps the hits within a window above
// the running minimum, adds the spheres, finds the gap and replays S in index order.  A ray whose window overflows,
// or whose chain of near-ties does not end inside the window, takes the brute-force loop (never seen on cfg1).
// The shadow predicate `any t in (0, limit)` (HW1/Scene.cpp:72-101) is order-free: any-hit through the same tree.
#pragma once
#include "rt_fast.cuh"

#define HW1_CAND 8

struct Hw1Cands {
  float t[HW1_CAND];
  int idx[HW1_CAND];
  int n;
  float tmin;
  bool overflow;
};
__device__ __forceinline__ float hw1_window(float tmin, float eps) { return tmin + 16.0f * fmaxf(eps, fabsf(tmin) * 2.4e-7f); }
__device__ __forceinline__ void hw1_add(Hw1Cands& c, float t, int idx, float eps) {
  if (c.n > 0 && t > hw1_window(c.tmin, eps)) return;
  if (t < c.tmin) {
    c.tmin = t;
    const float w = hw1_window(t, eps);
    int m = 0;
    for (int i = 0; i < c.n; i++)
      if (c.t[i] <= w) { c.t[m] = c.t[i]; c.idx[m] = c.idx[i]; m++; }
    c.n = m;
  }
  if (c.n == HW1_CAND) { c.overflow = true; return; }
  c.t[c.n] = t; c.idx[c.n] = idx; c.n++;
}

// closest hit with HW1's rule.  Returns false when the ray has to take the brute-force loop.
__device__ __forceinline__ bool hw1_resolve(Hw1Cands& c, float eps, float* t_out, int* idx_out) {
  if (c.overflow) return false;
  if (c.n == 0) { *t_out = RT_INF; *idx_out = -1; return true; }
  for (int i = 1; i < c.n; i++) {                     // sort by t
    const float t = c.t[i]; const int id = c.idx[i];
    int j = i;
    while (j > 0 && c.t[j - 1] > t) { c.t[j] = c.t[j - 1]; c.idx[j] = c.idx[j - 1]; j--; }
    c.t[j] = t; c.idx[j] = id;
  }
  int m = -1;
  for (int j = 0; j + 1 < c.n; j++)
    if (c.t[j] < c.t[j + 1] - eps) { m = j + 1; break; }
  if (m < 0) {
    // no gap inside the list: everything farther is beyond the window, is THAT a gap?
    if (!(c.t[c.n - 1] < hw1_window(c.tmin, eps) - eps)) return false;
    m = c.n;
  }
  for (int i = 1; i < m; i++) {                       // S in index order
    const float t = c.t[i]; const int id = c.idx[i];
    int j = i;
    while (j > 0 && c.idx[j - 1] > id) { c.t[j] = c.t[j - 1]; c.idx[j] = c.idx[j - 1]; j--; }
    c.t[j] = t; c.idx[j] = id;
  }
  float best = RT_INF;
  int bi = -1;
  for (int i = 0; i < m; i++)
    if (c.t[i] < best - eps && c.t[i] > 0.0f) { best = c.t[i]; bi = c.idx[i]; }
  *t_out = best; *idx_out = bi;
  return true;
}

template <bool ANY>
__global__ void __launch_bounds__(RT_BLOCK) k_trace_hw1_fast(DScene S, Pool P, ShadowQueue Q, const WaveInfo* waves, int d,
                                                             Counters* C) {
  const uint32_t begin = ANY ? 0u : waves[d].begin;
  const uint32_t count = ANY ? C->shadow_count : waves[d].count;
  const float eps = S.isect_eps;
  TravCounters tc = {0, 0, 0, 0};
  for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < count; k += gridDim.x * blockDim.x) {
    const uint32_t slot = begin + k;
    const float4 ro = ANY ? Q.o[k] : P.ray_o[slot], rd = ANY ? Q.d[k] : P.ray_d[slot];
    const v3 o = mk3(ro.x, ro.y, ro.z), dir = mk3(rd.x, rd.y, rd.z);
    const float limit = ANY ? ro.w : RT_INF;
    const FastRay fr = make_fast_ray(o, dir, eps);
    Hw1Cands cand;
    cand.n = 0; cand.tmin = RT_INF; cand.overflow = false;
    bool occluded = false;
    uint32_t stack[RT_FSTACK];
    float stack_t[RT_FSTACK];
    int sp = 0;
    uint32_t cur = S.hw1_fast_root;
    float cur_t = 0.0f;
    while (!occluded) {
      if (cur == B200RT_REF_NONE) {
        if (sp == 0) break;
        sp--; cur = stack[sp]; cur_t = stack_t[sp];
      }
      const float bound = ANY ? limit : (cand.n ? hw1_window(cand.tmin, eps) : RT_INF);
      if (cur_t > bound + fabsf(bound) * 1e-4f) { cur = B200RT_REF_NONE; continue; }
      const uint32_t kind = B200RT_REF_KIND(cur), idx = B200RT_REF_INDEX(cur);
      cur = B200RT_REF_NONE;
      if (kind == B200RT_KIND_FNODE) {
        tc.nodes++;
        const float4 n0 = __ldg(S.fast_nodes + 4 * (size_t)idx), n1 = __ldg(S.fast_nodes + 4 * (size_t)idx + 1);
        const float4 n2 = __ldg(S.fast_nodes + 4 * (size_t)idx + 2), n3 = __ldg(S.fast_nodes + 4 * (size_t)idx + 3);
        float e0, e1;
        const bool h0 = box_loose(fr, n0.x, n0.y, n0.z, n0.w, n1.x, n1.y, eps, &e0);
        const bool h1 = box_loose(fr, n1.z, n1.w, n2.x, n2.y, n2.z, n2.w, eps, &e1);
        const uint32_t c0 = __float_as_uint(n3.x), c1 = __float_as_uint(n3.y);
        if (h0 && h1) {
          const bool swap = e1 < e0;
          cur = swap ? c1 : c0; cur_t = swap ? e1 : e0;
          if (sp < RT_FSTACK) { stack[sp] = swap ? c0 : c1; stack_t[sp] = swap ? e0 : e1; sp++; }
          else { cand.overflow = true; }
        } else if (h0) { cur = c0; cur_t = e0; }
        else if (h1) { cur = c1; cur_t = e1; }
      } else if (kind >= B200RT_KIND_LEAF1 && kind <= B200RT_KIND_LEAF4) {
        const int n = (int)(kind - B200RT_KIND_LEAF1) + 1;
        for (int i = 0; i < n; i++) {
          const size_t fi = (size_t)idx + i;
          tc.tris++;
          const float4 a = __ldg(S.fast_tris + 3 * fi), b = __ldg(S.fast_tris + 3 * fi + 1), c = __ldg(S.fast_tris + 3 * fi + 2);
          float t, be, ga;
          if (tri_test<1>(mk3(a.x, a.y, a.z), mk3(a.w, b.x, b.y), mk3(b.z, b.w, c.x), o, dir, eps, &t, &be, &ga) && t > 0.0f) {
            if (ANY) { if (t < limit) { occluded = true; break; } }
            else hw1_add(cand, t, __ldg(S.fast_tri_ref + fi), eps);
          }
        }
      }
    }
    for (int i = 0; i < S.n_spheres && !occluded; i++) {
      float t;
      tc.spheres++;
      if (sphere_test(S.spheres + i, o, dir, 0.f, eps, &t) && t > 0.0f) {
        if (ANY) { if (t < limit) occluded = true; }
        else hw1_add(cand, t, S.n_tris + i, eps);
      }
    }
    if (ANY) {
      Q.vis[__float_as_uint(rd.w)] = occluded ? 0 : 1;
      continue;
    }
    float best;
    int bi;
    if (!hw1_resolve(cand, eps, &best, &bi)) {
      // the reference's own loop (HW1/Scene.cpp:35-54) for this one ray
      best = RT_INF; bi = -1;
      for (int i = 0; i < S.n_tris; i++) {
        const float4 a = __ldg(S.tris + 3 * (size_t)i), b = __ldg(S.tris + 3 * (size_t)i + 1), c = __ldg(S.tris + 3 * (size_t)i + 2);
        float t, be, ga;
        if (tri_test<1>(mk3(a.x, a.y, a.z), mk3(a.w, b.x, b.y), mk3(b.z, b.w, c.x), o, dir, eps, &t, &be, &ga) &&
            t < best - eps && t > 0.0f) { best = t; bi = i; }
      }
      for (int i = 0; i < S.n_spheres; i++) {
        float t;
        if (sphere_test(S.spheres + i, o, dir, 0.f, eps, &t) && t < best - eps && t > 0.0f) { best = t; bi = S.n_tris + i; }
      }
      atomicAdd(&C->hw1_fallbacks, 1u);
    }
    float bb = 0.f, bg = 0.f;
    uint32_t best_ref = B200RT_REF_NONE;
    int best_tri = -1;
    if (bi >= 0 && bi < S.n_tris) {
      const float4 a = __ldg(S.tris + 3 * (size_t)bi), b = __ldg(S.tris + 3 * (size_t)bi + 1), c = __ldg(S.tris + 3 * (size_t)bi + 2);
      float t;
      tri_test<1>(mk3(a.x, a.y, a.z), mk3(a.w, b.x, b.y), mk3(b.z, b.w, c.x), o, dir, eps, &t, &bb, &bg);
      best_ref = B200RT_REF(B200RT_KIND_TRI, bi); best_tri = bi;
    } else if (bi >= S.n_tris) {
      best_ref = B200RT_REF(B200RT_KIND_SPHERE, bi - S.n_tris);
    }
    P.hit_a[slot] = make_float4(best, bb, bg, __uint_as_float(best_ref));
    P.hit_b[slot] = make_int2(bi, best_tri);
  }
  flush_counters(C, tc, ANY);
}
