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
