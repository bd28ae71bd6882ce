// TEST INFRASTRUCTURE ONLY — the DEVICE's leaf functions (csrc/rt_traverse.cuh: box gates, tri_test, sphere_test,
// sphere_normal) compiled for the HOST, so that the reference's known-answer vectors (tests/golden/leaf_vectors.bin) can
// be put through the very source the kernels are built from on a box without a GPU.  What this checks is the order of
// operations written in that source; that the device executes the same IEEE operations is the job of the build flags
// (-fmad=false, default -prec-div / -prec-sqrt; here -ffp-contract=off) and of the on-device tests.  Nothing in the
// product links this file: tests/test_leaf_vectors.py compiles it into build/libdevice_leaf_host.so with g++.
#include <cuda_runtime.h>   // float4, make_float4 ... (host-compatible header)

#include <cmath>
#include <cstdint>
#include <cstring>
#undef __noinline__
#define __noinline__
template <class T> static inline T __ldg(const T* p) { return *p; }
template <class T> static inline T __ldcs(const T* p) { return *p; }
static inline float __int_as_float(int i) { float f; std::memcpy(&f, &i, 4); return f; }
static inline float __uint_as_float(unsigned i) { float f; std::memcpy(&f, &i, 4); return f; }
static inline unsigned __float_as_uint(float f) { unsigned i; std::memcpy(&i, &f, 4); return i; }
static inline int __float_as_int(float f) { int i; std::memcpy(&i, &f, 4); return i; }
static inline float __fmaf_rn(float a, float b, float c) { return std::fmaf(a, b, c); }

#include "../ceng795_b200/csrc/rt_traverse.cuh"

extern "C" {
// out2 = {exact gate decision, filtered gate decision}
void hostleaf_box(const float* mn, const float* mx, const float* o, const float* d, float eps, int* out2) {
  const v3 O = mk3(o[0], o[1], o[2]), D = mk3(d[0], d[1], d[2]);
  float entry = 0.0f;
  out2[0] = box_gate_exact(O, D, mn[0], mn[1], mn[2], mx[0], mx[1], mx[2], eps, &entry) ? 1 : 0;
  const RayCtx r = make_ray(O, D, eps);
  out2[1] = box_gate(r, mn[0], mn[1], mn[2], mx[0], mx[1], mx[2], eps, &entry) ? 1 : 0;
}
// 1 / 0: the filter's own decision, -1: it would fall back to the divisions (slow rays always do)
int hostleaf_box_filter(const float* mn, const float* mx, const float* o, const float* d, float eps) {
  const RayCtx r = make_ray(mk3(o[0], o[1], o[2]), mk3(d[0], d[1], d[2]), eps);
  float entry = 0.0f;
  return r.slow ? -1 : box_gate_filter(r, mn[0], mn[1], mn[2], mx[0], mx[1], mx[2], &entry);
}
// Adversarial boxes for the filtered gate: rays aimed exactly at corners, edges and faces of the box (the decision hangs
// on the last bit), origins up to `far` box sizes away, on the box, inside it; directions with tiny and zero components;
// `flat_share` of the boxes are flat on one or two axes (axis-aligned walls, the second level of the filter).
// Returns the number of cases in which the filter's decision differs from the exact gate's; stats[0..2] = cases,
// accepted by the exact gate, decided by the filter without the exact fallback.
static uint32_t lcg(uint32_t& s) { s = s * 1664525u + 1013904223u; return s; }
static float urand(uint32_t& s) { return (float)(lcg(s) >> 8) * (1.0f / 16777216.0f); }
int hostleaf_box_fuzz(int n, uint32_t seed, float far, float eps, float flat_share, long long* stats) {
  uint32_t s = seed * 2654435761u + 12345u;
  int bad = 0;
  long long pass = 0, fast = 0;
  for (int k = 0; k < n; k++) {
    float mn[3], mx[3], o[3], d[3], aim[3];
    const float scale = std::exp((urand(s) - 0.5f) * 12.0f);
    for (int a = 0; a < 3; a++) {
      const float c = (urand(s) - 0.5f) * 20.0f * scale, h = (0.01f + urand(s)) * scale;
      mn[a] = c - h; mx[a] = c + h;
      const int where = (int)(lcg(s) >> 30);   // 0: min plane, 1: max plane, 2/3: somewhere inside (or just outside)
      aim[a] = where == 0 ? mn[a] : where == 1 ? mx[a] : mn[a] + (urand(s) * 1.2f - 0.1f) * (mx[a] - mn[a]);
    }
    float ext[3];
    for (int a = 0; a < 3; a++) ext[a] = mx[a] - mn[a];
    if (urand(s) < flat_share) {
      const int a = (int)(lcg(s) % 3u);
      mx[a] = mn[a]; aim[a] = mn[a];
      if ((lcg(s) >> 29) == 0) { const int b = (a + 1 + (int)(lcg(s) & 1u)) % 3; mx[b] = mn[b]; aim[b] = mn[b]; }
    }
    const int kind = (int)(lcg(s) >> 29);
    for (int a = 0; a < 3; a++) {
      const float dist = (kind == 0 ? 0.0f : kind == 1 ? urand(s) : urand(s) * far) * ext[a];
      o[a] = aim[a] + (urand(s) < 0.5f ? -dist : dist);
      if (kind == 2) o[a] = aim[a];            // origin ON the aimed point for this axis
    }
    for (int a = 0; a < 3; a++) d[a] = aim[a] - o[a];
    if ((lcg(s) >> 28) == 0) d[lcg(s) % 3] = 0.0f;
    if ((lcg(s) >> 28) == 1) d[lcg(s) % 3] = (urand(s) - 0.5f) * 4e-6f;
    if ((lcg(s) >> 27) == 2) for (int a = 0; a < 3; a++) d[a] = urand(s) - 0.5f;   // unrelated direction: mostly misses
    if ((lcg(s) >> 28) == 3) for (int a = 0; a < 3; a++) d[a] = -d[a];            // box behind the origin
    const v3 O = mk3(o[0], o[1], o[2]), D = mk3(d[0], d[1], d[2]);
    float e0, e1;
    const bool want = box_gate_exact(O, D, mn[0], mn[1], mn[2], mx[0], mx[1], mx[2], eps, &e0);
    const RayCtx r = make_ray(O, D, eps);
    const bool g1 = box_gate(r, mn[0], mn[1], mn[2], mx[0], mx[1], mx[2], eps, &e1);
    bad += (g1 != want);
    pass += want;
    if (!r.slow && box_gate_filter(r, mn[0], mn[1], mn[2], mx[0], mx[1], mx[2], &e1) >= 0) fast++;   // decided without the divisions
  }
  if (stats) { stats[0] = n; stats[1] = pass; stats[2] = fast; }
  return bad;
}
// g9 = v0, a1 = v0 - v1, a2 = v0 - v2; mode 1 / 2 / 7; out3 = {t, beta, gamma}
int hostleaf_tri(const float* g9, const float* o, const float* d, int mode, float eps, float* out3) {
  const v3 v0 = mk3(g9[0], g9[1], g9[2]), a1 = mk3(g9[3], g9[4], g9[5]), a2 = mk3(g9[6], g9[7], g9[8]);
  const v3 O = mk3(o[0], o[1], o[2]), D = mk3(d[0], d[1], d[2]);
  switch (mode) {
    case 1: return tri_test<1>(v0, a1, a2, O, D, eps, out3, out3 + 1, out3 + 2) ? 1 : 0;
    case 2: return tri_test<2>(v0, a1, a2, O, D, eps, out3, out3 + 1, out3 + 2) ? 1 : 0;
    default: return tri_test<7>(v0, a1, a2, O, D, eps, out3, out3 + 1, out3 + 2) ? 1 : 0;
  }
}
// out4 = {t, normal.xyz} (normal by the HW3 rule: the normal matrix only when the sphere is transformed)
int hostleaf_sphere(const b200rt_sphere* s, const float* o, const float* d, float time, float eps, float* out4) {
  const v3 O = mk3(o[0], o[1], o[2]), D = mk3(d[0], d[1], d[2]);
  float t;
  if (!sphere_test(s, O, D, time, eps, &t)) return 0;
  const v3 n = sphere_normal(s, O, D, time, t);
  out4[0] = t; out4[1] = n.x; out4[2] = n.y; out4[3] = n.z;
  return 1;
}
}
