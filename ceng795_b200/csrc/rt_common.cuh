// rt_common.cuh — shared device types for the B200 rendering core.
//
// Arithmetic policy: this translation unit is compiled with -fmad=false and the
// default -prec-div=true / -prec-sqrt=true, so `a*b+c`, `/` and sqrtf are plain
// IEEE binary32 operations in source order — the same values the reference's
// x86-64 build produces (no FMA contraction there either).
// Where the reference's unqualified pow/exp/log bind to the double overloads
// the kernels call the double functions too.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/b200rt.h"

#define RT_INF __int_as_float(0x7f800000)

struct v3 {
  float x, y, z;
};
__host__ __device__ __forceinline__ v3 mk3(float x, float y, float z) { v3 r; r.x = x; r.y = y; r.z = z; return r; }
__host__ __device__ __forceinline__ v3 ld3(const float* p) { return mk3(p[0], p[1], p[2]); }
__device__ __forceinline__ v3 operator+(v3 a, v3 b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ v3 operator-(v3 a, v3 b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ v3 operator-(v3 a) { return mk3(-a.x, -a.y, -a.z); }
__device__ __forceinline__ v3 operator*(v3 a, v3 b) { return mk3(a.x * b.x, a.y * b.y, a.z * b.z); }
__device__ __forceinline__ v3 operator*(v3 a, float s) { return mk3(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ v3 operator*(float s, v3 a) { return mk3(a.x * s, a.y * s, a.z * s); }
// Vector3::operator/(float): three IEEE divisions.  The shading kernels divide in ~100 places; inlined, those
// divisions alone were 24 KB of k_shade7_b (profiles/r01f: the shade kernels stall on instruction fetch), so they share
// one out-of-line copy.  The triangle test keeps the inline form (div3_inl).
__device__ __forceinline__ v3 div3_inl(v3 a, float s) { return mk3(a.x / s, a.y / s, a.z / s); }
__device__ __noinline__ v3 div3_call(v3 a, float s) { return mk3(a.x / s, a.y / s, a.z / s); }
__device__ __forceinline__ v3 operator/(v3 a, float s) { return div3_call(a, s); }
// Vector3::dot — (x*x') + (y*y') + (z*z') left to right (HW3/include/Vector3.h:107-109)
__device__ __forceinline__ float dot(v3 a, v3 b) { return (a.x * b.x) + (a.y * b.y) + (a.z * b.z); }
__device__ __forceinline__ v3 cross(v3 a, v3 b) {
  return mk3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
__device__ __forceinline__ float length(v3 a) { return sqrtf(a.x * a.x + a.y * a.y + a.z * a.z); }
__device__ __noinline__ v3 normalize(v3 a) { return div3_inl(a, length(a)); }
__device__ __forceinline__ bool is_zero(v3 a) { return a.x == 0.0f && a.y == 0.0f && a.z == 0.0f; }

// pow(x, k) for the small integer exponents the reference spells pow(x, 2.0) .. pow(x, 5.0): products in double.
// glibc's pow is within 0.52 ulp of the exact power; these products (x is a float widened to double, so x*x is exact
// and x^3, x^4 are correctly rounded) are within 1.5 ulp of it, and every caller rounds the expression to float, so a
// result differs from the reference's in about one call in 2^27 — the same class as CUDA's pow against glibc's, at a
// hundredth of the instructions (the double pow was 42 % of k_shade7_a's issued instructions, profiles/r01f).
__device__ __forceinline__ double pow2d(double x) { return x * x; }
__device__ __forceinline__ double pow3d(double x) { return (x * x) * x; }
__device__ __forceinline__ double pow4d(double x) { const double x2 = x * x; return x2 * x2; }
__device__ __forceinline__ double pow5d(double x) { const double x2 = x * x; return (x2 * x2) * x; }
// pow(x, y), x >= 0 (clamped cosines), y a material exponent: whole exponents up to 4096 by square-and-multiply
// (<= 24 roundings of 2^-53: the float result differs from the reference's in about one call in 10^7 at worst), the rest through
// the double pow.
__device__ __noinline__ double pow_ref7(double x, double y) {
  const int n = (int)y;
  if ((double)n == y && n >= 0 && n <= 4096) {
    double r = 1.0, b = x;
    for (int k = n; k > 0; k >>= 1) {
      if (k & 1) r *= b;
      b *= b;
    }
    return r;
  }
  return pow(x, y);
}

// ---- counter-based RNG (specification shared with oracle/oracle.c) ----------
__host__ __device__ __forceinline__ uint32_t rng_mix(uint32_t a, uint32_t b) {
  uint32_t h = a * 0x9E3779B1u + b;
  h ^= h >> 16; h *= 0x85EBCA6Bu; h ^= h >> 13; h *= 0xC2B2AE35u; h ^= h >> 16;
  return h;
}
__host__ __device__ __forceinline__ float rng_u01(uint32_t key, uint32_t dim) {
  return (float)(rng_mix(key, dim ^ 0xA511E9B3u) >> 8) * (1.0f / 16777216.0f);
}

// ---- device scene -----------------------------------------------------------
struct DScene {
  int semantics, max_depth;
  int refill_closest, refill_any;  // idle lanes before a traversal warp takes new rays (32 = whole-warp batches)
  int node_vote;                   // lanes that must hold an inner node for the warp to keep descending (rt_fast.cuh)
  int wide_node_vote;   // the same threshold for k_trace_wide (its steps are longer: swept separately)
  v3 background, ambient;
  float shadow_eps, isect_eps;
  uint32_t root;
  float root_box[6];
  const float4* nodes;      // 4 x float4 per inner node (b200rt_bvh_node verbatim)
  // top-level objects flattened with the reference's innermost gate of each: 2 x float4 per object
  // {min.xyz, as_float(ref)}, {max.xyz, as_float(1 if the object has a gate, 0 if it is entered unconditionally)}
  const float4* top_objs;
  uint32_t hw1_fast_root;   // HW1: SAH tree over ALL triangles (rt_hw1_fast.cuh), or NONE
  int n_top;                // 0: walk the reference's top-level tree instead (more than RT_TOP_LINEAR_MAX objects)
  const float4* tris;       // 3 x float4 per triangle: v0|a1.x, a1.yz|a2.xy, a2.z|n
  const float4* fast_nodes; // SAH tree (csrc/fast_bvh.h), same 64-byte records
  const float4* fast_tris;  // triangles again, in SAH leaf order
  const int* fast_tri_ref;  // leaf-order position -> triangle index (DFS order)
  const float* tri_gate;    // 6 floats per triangle: the reference's innermost gate box
  // 8-wide tree with quantised child boxes (csrc/wide_bvh.h), walked by k_trace_wide
  const float4* wide_nodes; // 5 x float4 per node
  const float4* wide_tris;  // triangles in the wide leaves' order
  const int* wide_tri_ref;  // wide-order position -> triangle index (DFS order)
  const int* wide_root;     // per mesh: index of its root node, -1 none
  int use_wide;             // 0: k_trace_fast (binary tree, top-level tree walk) — more than RT_TOP_LINEAR_MAX objects, device-built tree
  const int2* tri_info;     // material, flags
  const int* tri_vidx;      // 3 per triangle
  const float4* vnormal;    // per vertex
  const b200rt_mesh* meshes;
  const b200rt_instance* inst;
  const b200rt_sphere* spheres;
  const b200rt_triprim* triprims;
  const b200rt_material* mats;
  const b200rt_light* lights;
  int n_lights, n_tris, n_spheres, n_inst;
  // HW4..HW7
  const float* vert_pos;           // 3 per vertex
  const float* tri_uv;             // 6 per triangle
  const int* tri_tex;              // 1 per triangle
  const b200rt_texture* textures;
  const float* texels;
  const b200rt_brdf* brdfs;
  const float* light_cdf;
  const int* light_tri;
  int background_texture, env_light, integrator, importance_sampling;
  int* lit_count;   // parity hook: per pixel x light, shadow tests that came out lit (HW5-HW7 shading); NULL = off
  b200rt_camera cam;
};

// ---- vertex pool (wavefront state) -----------------------------------------
// One slot per path vertex. Wave d occupies the contiguous slot range
// [wave[d].begin, wave[d].begin + wave[d].count); children are appended to the
// next wave with warp-aggregated atomics.
#define RT_META_DEPTH(m) ((m)&0xFFu)
#define RT_META_INMEDIUM 0x100u
#define RT_META_PRIMARY 0x200u
#define RT_META_LIGHTHIT 0x400u   /* HW7 Ray::light_hit */

struct WaveInfo {
  uint32_t begin, count;
};

struct Pool {
  uint32_t capacity;
  float4* ray_o;     // o.xyz, time
  float4* ray_d;     // d.xyz, as_float(meta)
  uint32_t* key;     // RNG stream of the vertex
  uint32_t* pixel;   // pixel index (wave 0: also sample id in the high bits via sample[])
  uint32_t* sample;  // stratum index of the camera sample (wave 0 only)
  unsigned long long* omask;  // k_top_mask: top-level objects whose reference gate the ray passes (8-wide path)
  float4* hit_a;     // t, beta, gamma, as_float(top_ref)
  int2* hit_b;       // top_rank, tri
  float4* nrm;       // shading normal xyz, as_float(material)
  float4* rec0;      // local colour rgb, fresnel r
  float4* rec1;      // km rgb, as_float(child_mirror)
  float4* rec2;      // k  rgb, as_float(child_reflect)
  uint32_t* rec3;    // child_transmit
  float4* final_c;   // combined colour (after the backward pass)
  // HW7
  float4* kd;        // texture-blended diffuse constant rgb, as_float(flags: bit0 replace_all, bit1 emissive)
  float4* rec4;      // path tracing: weight of the hemisphere-sample child rgb, as_float(child_diffuse)
};

struct ShadowQueue {
  uint32_t capacity;
  float4* o;         // o.xyz, t_limit
  float4* d;         // d.xyz, as_float(result slot)
  float* time;
  unsigned long long* omask;  // as Pool::omask, per queue entry
  uint8_t* vis;      // per (vertex slot * n_lights + light): 1 = lit, 0 = occluded / inactive
  float4* lv;        // HW7, same index: {vector to the light sample, pdf} drawn by shade_a, reused by shade_b
};

struct Counters {
  unsigned long long rays_primary, rays_shadow, rays_secondary;
  unsigned long long nodes_visited, tris_tested, spheres_tested, instances_entered;
  unsigned long long sh_nodes_visited, sh_tris_tested, sh_spheres_tested, sh_instances_entered;
  unsigned int overflow;   // pool / queue overflow flag
  unsigned int shadow_count;
  unsigned int work_extend, work_shadow;  // persistent-kernel work cursors
  unsigned int hw1_fallbacks;  // HW1 rays that took the brute-force loop (rt_hw1_fast.cuh)
};
