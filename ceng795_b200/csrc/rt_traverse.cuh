// rt_traverse.cuh — BVH traversal and primitive intersection.
//
// Semantics to reproduce (HW3/src/Bounding_volume_hierarchy.cpp:31-55): a
// primitive is *reachable* iff every INNER / INSTANCE box on its root path
// passes Bounding_box::intersect; the result is the reachable primitive with
// the smallest positive t, ties going to the earliest in DFS order.  The
// reference visits every reachable node; this traversal visits near children
// first and prunes boxes whose entry distance exceeds the best t so far.  The
// winner is chosen by the key (t, top-level DFS rank, triangle DFS index), so
// visiting order and pruning cannot change it.
#pragma once
#include "m4_invert.h"
#include "rt_common.cuh"

#define RT_STACK 96

struct HitRec {
  float t, beta, gamma;
  uint32_t top_ref;  // ref of the top-level object (or TRI for HW1)
  int top_rank, tri;
};

struct TravCounters {
  unsigned int nodes, tris, spheres, insts;
};

// Bounding_box::intersect — HW3/src/bounding_box.cpp:17-37.  Returns the
// reference's accept decision (`!(bbox_t < 0 || bbox_t == inf)`) with exactly
// the reference's arithmetic; *entry gets max(tmin, 0) for ordering/pruning.
__device__ __noinline__ bool box_gate_exact(v3 o, v3 d, float mnx, float mny, float mnz, float mxx, float mxy,
                                            float mxz, float eps, float* entry) {
  float tmin = -RT_INF, tmax = RT_INF;
  if (!(fabsf(d.x) < eps)) {
    float a = (mnx - o.x) / d.x, b = (mxx - o.x) / d.x;
    if (d.x < 0) { float t = a; a = b; b = t; }
    if (a > tmin) tmin = a;
    if (b < tmax) tmax = b;
    if (tmin > tmax) return false;
  }
  if (!(fabsf(d.y) < eps)) {
    float a = (mny - o.y) / d.y, b = (mxy - o.y) / d.y;
    if (d.y < 0) { float t = a; a = b; b = t; }
    if (a > tmin) tmin = a;
    if (b < tmax) tmax = b;
    if (tmin > tmax) return false;
  }
  if (!(fabsf(d.z) < eps)) {
    float a = (mnz - o.z) / d.z, b = (mxz - o.z) / d.z;
    if (d.z < 0) { float t = a; a = b; b = t; }
    if (a > tmin) tmin = a;
    if (b < tmax) tmax = b;
    if (tmin > tmax) return false;
  }
  const float bt = (tmin > 0.0f) ? tmin : tmax;
  if (bt < 0.0f || bt == RT_INF) return false;
  *entry = fmaxf(tmin, 0.0f);
  return true;
}

// Ray with the per-ray constants of the filtered box test.
struct RayCtx {
  v3 o, d, inv;
  bool slow;  // some |d_i| < eps (axis skipped by the reference): always take the exact path
};
__device__ __forceinline__ RayCtx make_ray(v3 o, v3 d, float eps) {
  RayCtx r;
  r.o = o; r.d = d;
  r.slow = (fabsf(d.x) < eps) || (fabsf(d.y) < eps) || (fabsf(d.z) < eps);
  r.inv = mk3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
  return r;
}

// Filtered gate: decides with (b - o) * (1/d) when the outcome is clear by a
// margin that covers the rounding difference to the reference's (b - o) / d
// (|diff| <= 2^-24 + 2^-23 relative, margin 5e-7), and falls back to the exact
// division-based test otherwise.  The accept/reject DECISION is therefore always
// the reference's; only the instruction count changes.
//   1 accept, 0 reject, -1 not clear.
// Two levels.  The first compares the entry (largest near plane) with the exit (smallest far plane).  It can never
// clear a FLAT box — an axis-aligned wall or light quad, mn == mx on one axis, gives near == far on that axis, so entry
// and exit coincide for every ray that meets the wall's plane inside the box — and the Cornell-style scenes of HW7 /
// PPM gate a dozen of those per ray (profiles/r02p: the gate pre-pass took 11 % of cfg4, most of it in the divisions
// of the fallback).  The second level uses what the reference's loop really decides: it rejects iff some near_i
// exceeds some far_j, and near_i <= far_i always holds (the same subtraction and division, rounded monotonically, then
// ordered by the sign of d), so only the six CROSS-axis pairs matter, each with a margin of its own size.
#ifndef RT_GATE_LEVELS
#define RT_GATE_LEVELS 2   // 1: entry against exit only (A/B aid, tools/build_variant.sh)
#endif
__device__ __forceinline__ int box_gate_filter(const RayCtx& r, float mnx, float mny, float mnz, float mxx, float mxy,
                                               float mxz, float* entry) {
  const float ax = (mnx - r.o.x) * r.inv.x, bx = (mxx - r.o.x) * r.inv.x;
  const float ay = (mny - r.o.y) * r.inv.y, by = (mxy - r.o.y) * r.inv.y;
  const float az = (mnz - r.o.z) * r.inv.z, bz = (mxz - r.o.z) * r.inv.z;
  const float lx = fminf(ax, bx), ly = fminf(ay, by), lz = fminf(az, bz);
  const float hx = fmaxf(ax, bx), hy = fmaxf(ay, by), hz = fmaxf(az, bz);
  const float tn = fmaxf(fmaxf(lx, ly), lz);
  const float tf = fminf(fminf(hx, hy), hz);
  const float m = 5e-7f * fmaxf(fabsf(tn), fabsf(tf));
  const bool range_ok = tf > 1e-30f && tf < 1e30f && tn > -1e30f;   // bbox_t is positive and finite
  if (tn + m < tf && range_ok) { *entry = fmaxf(tn, 0.0f); return 1; }
  if (tn > tf + m || tf < -1e-30f) return 0;
#if RT_GATE_LEVELS > 1
  const float Lx = fmaxf(ly, lz), Ly = fmaxf(lx, lz), Lz = fmaxf(lx, ly);   // largest near plane of the OTHER two axes
  const float qx = 5e-7f * fmaxf(fabsf(Lx), fabsf(hx)) + 1e-30f;
  const float qy = 5e-7f * fmaxf(fabsf(Ly), fabsf(hy)) + 1e-30f;
  const float qz = 5e-7f * fmaxf(fabsf(Lz), fabsf(hz)) + 1e-30f;
  if (Lx + qx < hx && Ly + qy < hy && Lz + qz < hz && range_ok) { *entry = fmaxf(tn, 0.0f); return 1; }
  if (Lx > hx + qx || Ly > hy + qy || Lz > hz + qz) return 0;
#endif
  return -1;
}
__device__ __forceinline__ bool box_gate(const RayCtx& r, float mnx, float mny, float mnz, float mxx, float mxy,
                                         float mxz, float eps, float* entry) {
  if (!r.slow) {
    const int g = box_gate_filter(r, mnx, mny, mnz, mxx, mxy, mxz, entry);
    if (g >= 0) return g != 0;
  }
  return box_gate_exact(r.o, r.d, mnx, mny, mnz, mxx, mxy, mxz, eps, entry);
}

// Mesh_triangle::determinant — HW3/include/Mesh_triangle.h:42-47
__device__ __forceinline__ float det3(v3 c1, v3 c2, v3 c3) {
  return c1.x * (c2.y * c3.z - c3.y * c2.z) + c2.x * (c3.y * c1.z - c1.y * c3.z) +
         c3.x * (c1.y * c2.z - c2.y * c1.z);
}

// Cramer's-rule triangle test.  MODE 1: HW1 band (HW1/Triangle.cpp:29-47),
// 2: exact [0,1], t>0 (HW3/src/Mesh_triangle.cpp:85-97), 7: HW7/PPM mesh band.
template <int MODE>
__device__ __forceinline__ bool tri_test(v3 v0, v3 a1, v3 a2, v3 o, v3 d, float eps, float* t_out, float* beta_out,
                                         float* gamma_out) {
  const float det_a = det3(a1, a2, d);
  if (det_a == 0.0f) return false;
  const v3 b = div3_inl(v0 - o, det_a);
  const float beta = det3(b, a2, d);
  if (MODE == 1) { if (beta < 0.0f - eps || beta > 1.0f + eps) return false; }
  else if (MODE == 2) { if (beta < 0.0f || beta > 1.0f) return false; }
  else { if (beta < -eps) return false; }
  const float gamma = det3(a1, b, d);
  if (MODE == 1) { if (gamma < 0.0f - eps || beta + gamma > 1.0f + eps) return false; }
  else if (MODE == 2) { if (gamma < 0.0f || beta + gamma > 1.0f) return false; }
  else { if (gamma < -eps || beta + gamma > 1.0f + eps) return false; }
  const float t = det3(a1, a2, b);
  if (MODE == 2) { if (!(t > 0.0f)) return false; }
  else { if (!(t > -eps)) return false; }
  *t_out = t; *beta_out = beta; *gamma_out = gamma;
  return true;
}

// Matrix4x4::multiply — HW3/src/Matrix4x4.cpp:29-50, rows of a 3x4 / 3x3 block.
__device__ __forceinline__ v3 xf_point(const float* m, v3 p) {
  v3 r;
  r.x = m[3];  r.x += m[0] * p.x; r.x += m[1] * p.y; r.x += m[2] * p.z;
  r.y = m[7];  r.y += m[4] * p.x; r.y += m[5] * p.y; r.y += m[6] * p.z;
  r.z = m[11]; r.z += m[8] * p.x; r.z += m[9] * p.y; r.z += m[10] * p.z;
  return r;
}
__device__ __forceinline__ v3 xf_vector4(const float* m, v3 p) {
  v3 r;
  r.x = 0.0f; r.x += m[0] * p.x; r.x += m[1] * p.y; r.x += m[2] * p.z;
  r.y = 0.0f; r.y += m[4] * p.x; r.y += m[5] * p.y; r.y += m[6] * p.z;
  r.z = 0.0f; r.z += m[8] * p.x; r.z += m[9] * p.y; r.z += m[10] * p.z;
  return r;
}
__device__ __forceinline__ v3 xf_vector3(const float* m, v3 p) {
  v3 r;
  r.x = 0.0f; r.x += m[0] * p.x; r.x += m[1] * p.y; r.x += m[2] * p.z;
  r.y = 0.0f; r.y += m[3] * p.x; r.y += m[4] * p.y; r.y += m[5] * p.z;
  r.z = 0.0f; r.z += m[6] * p.x; r.z += m[7] * p.y; r.z += m[8] * p.z;
  return r;
}

// Matrix4x4::invert_matrix — HW3/src/Matrix4x4.cpp:97-233: the one copy shared with the host builder (m4_invert.h)
__device__ __noinline__ bool m4_invert(const float* m, float* out) { return b200rt_m4_invert(m, out); }

// Motion blur: Translation(time*velocity) * M, then Arbitrary_transformation
// (HW3/include/Mesh.h:76-84).  inv16 = full inverse, row-major.
__device__ __noinline__ void moving_inverse(const float* fwd, const float* vel, float time, float* inv16) {
  const float dx = vel[0] * time, dy = vel[1] * time, dz = vel[2] * time;  // ray.time * velocity
  float tr[16] = {1, 0, 0, dx, 0, 1, 0, dy, 0, 0, 1, dz, 0, 0, 0, 1};
  float prod[16];
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 4; j++) {
      float s = 0.0f;
      for (int k = 0; k < 4; k++) s += tr[i * 4 + k] * fwd[k * 4 + j];
      prod[i * 4 + j] = s;
    }
  if (!m4_invert(prod, inv16)) {
    for (int i = 0; i < 16; i++) inv16[i] = (i % 5 == 0) ? 1.0f : 0.0f;
  }
}

// Sphere::intersect — HW3/include/Sphere.h:44-155, HW1/Sphere.h:20-49.
// Returns true even when t < 0, like the reference (the caller filters t > 0).
// The local-space ray used is returned for the normal computation.
__device__ __forceinline__ bool sphere_test(const b200rt_sphere* s, v3 o, v3 d, float time, float eps, float* t_out) {
  v3 lo = o, ldir = d;
  if (s->flags & B200RT_XF_MOVING) {
    float inv16[16];
    moving_inverse(s->fwd, s->velocity, time, inv16);
    lo = xf_point(inv16, o);
    ldir = xf_vector4(inv16, d);
  } else if (!(s->flags & B200RT_XF_IDENTITY)) {
    lo = xf_point(s->inv, o);
    ldir = xf_vector4(s->inv, d);
  }
  const v3 c = ld3(s->center);
  const v3 cto = lo - c;
  const float a = dot(ldir, ldir);
  const float b = 2 * dot(ldir, cto);
  const float cc = dot(cto, cto) - s->radius * s->radius;
  const float disc = b * b - 4 * a * cc;
  if (disc < -eps) return false;
  float t;
  if (disc < eps) {
    t = -b / (2 * a);
  } else {
    const float sq = sqrtf(disc);
    const float t1 = (-b + sq) / (2 * a);
    const float t2 = (-b - sq) / (2 * a);
    t = (t2 < 0.0f) ? t1 : t2;
  }
  *t_out = t;
  return true;
}

// Sphere normal at parameter t (same branches as above).
__device__ __forceinline__ v3 sphere_normal(const b200rt_sphere* s, v3 o, v3 d, float time, float t) {
  const v3 c = ld3(s->center);
  if (s->flags & B200RT_XF_MOVING) {
    float inv16[16];
    moving_inverse(s->fwd, s->velocity, time, inv16);
    const v3 lo = xf_point(inv16, o), ldir = xf_vector4(inv16, d);
    const v3 n = normalize((lo + (ldir * t)) - c);
    // normal matrix = transpose(inverse): row i of it is column i of inv16
    v3 r;
    r.x = 0.0f; r.x += inv16[0] * n.x; r.x += inv16[4] * n.y; r.x += inv16[8] * n.z;
    r.y = 0.0f; r.y += inv16[1] * n.x; r.y += inv16[5] * n.y; r.y += inv16[9] * n.z;
    r.z = 0.0f; r.z += inv16[2] * n.x; r.z += inv16[6] * n.y; r.z += inv16[10] * n.z;
    return normalize(r);
  }
  if (!(s->flags & B200RT_XF_IDENTITY)) {
    const v3 lo = xf_point(s->inv, o), ldir = xf_vector4(s->inv, d);
    const v3 n = normalize((lo + (ldir * t)) - c);
    return normalize(xf_vector3(s->nrm, n));
  }
  return normalize((o + (d * t)) - c);
}

__device__ __forceinline__ bool better(float t, int rank, int tri, const HitRec& b) {
  return t < b.t || (t == b.t && (rank < b.top_rank || (rank == b.top_rank && tri < b.tri)));
}

#define RT_SENTINEL 0xE0000000u /* kind 14: leave the current instance */

// Closest-hit (ANY=false) or any-hit in (0, tlimit) (ANY=true) for the
// BVH-based trees (HW2..HW7, PPM).  TRI_MODE selects the mesh-triangle accept
// rule.  EXACT=true disables t-pruning (visits exactly the reference's node set).
template <bool ANY, bool EXACT, int TRI_MODE>
__device__ __forceinline__ bool traverse(const DScene& S, v3 o, v3 d, float time, float tlimit, HitRec& best,
                                         TravCounters& cnt) {
  uint32_t stack[RT_STACK];
  float stack_t[RT_STACK];
  int sp = 0;
  const float eps = S.isect_eps;
  best.t = RT_INF; best.beta = 0.f; best.gamma = 0.f; best.top_ref = B200RT_REF_NONE; best.top_rank = 0x7fffffff;
  best.tri = 0x7fffffff;
  bool found = false;

  const RayCtx wr = make_ray(o, d, eps);
  RayCtx cr = wr;               // ray in the current space
  v3 co = o, cd = d;
  int cur_rank = 0x7fffffff;    // rank of the instance being traversed
  uint32_t cur_top = B200RT_REF_NONE;
  const bool hw7 = S.semantics >= B200RT_SEM_HW7;  // HW4: band triangle test (TRI_MODE 7) without culling
  bool culling = hw7;           // HW7/PPM: back-face culling, off inside refractive instances

  // root: a gate for INNER / INSTANCE, unconditional for primitives
  {
    const uint32_t k = B200RT_REF_KIND(S.root);
    if (k == B200RT_KIND_NONE) return false;
    float e = 0.0f;
    bool ok = true;
    if (k == B200RT_KIND_INNER || k == B200RT_KIND_INSTANCE)
      ok = box_gate(cr, S.root_box[0], S.root_box[1], S.root_box[2], S.root_box[3], S.root_box[4],
                    S.root_box[5], eps, &e);
    if (!ok) return false;
    stack[sp] = S.root; stack_t[sp] = e; sp++;
  }

  while (sp > 0) {
    sp--;
    const uint32_t ref = stack[sp];
    const float te = stack_t[sp];
    const uint32_t kind = B200RT_REF_KIND(ref);
    const uint32_t idx = B200RT_REF_INDEX(ref);
    if (kind == 14u) {  // leave instance
      co = o; cd = d; cr = wr; cur_rank = 0x7fffffff; cur_top = B200RT_REF_NONE; culling = hw7;
      continue;
    }
    if (!EXACT) {
      const float bound = ANY ? tlimit : best.t;
      if (te > bound + fabsf(bound) * 1e-4f) continue;
    }
    if (kind == B200RT_KIND_INNER) {
      cnt.nodes++;
      const float4 n0 = __ldg(S.nodes + 4 * (size_t)idx);
      const float4 n1 = __ldg(S.nodes + 4 * (size_t)idx + 1);
      const float4 n2 = __ldg(S.nodes + 4 * (size_t)idx + 2);
      const float4 n3 = __ldg(S.nodes + 4 * (size_t)idx + 3);
      uint32_t child[2] = {__float_as_uint(n3.x), __float_as_uint(n3.y)};
      float entry[2] = {0.f, 0.f};
      bool push[2] = {false, false};
#pragma unroll
      for (int c = 0; c < 2; c++) {
        const uint32_t ck = B200RT_REF_KIND(child[c]);
        const uint32_t ci = B200RT_REF_INDEX(child[c]);
        if (ck == B200RT_KIND_INNER || ck == B200RT_KIND_INSTANCE) {
          bool ok;
          if (c == 0) ok = box_gate(cr, n0.x, n0.y, n0.z, n0.w, n1.x, n1.y, eps, &entry[0]);
          else ok = box_gate(cr, n1.z, n1.w, n2.x, n2.y, n2.z, n2.w, eps, &entry[1]);
          push[c] = ok;
        } else if (ck == B200RT_KIND_LIGHTINST) {
          push[c] = true; entry[c] = 0.f;   // Light_mesh has no gate of its own
        } else if (ck == B200RT_KIND_TRI) {
          cnt.tris++;
          const float4 t0 = __ldg(S.tris + 3 * (size_t)ci);
          const float4 t1 = __ldg(S.tris + 3 * (size_t)ci + 1);
          const float4 t2 = __ldg(S.tris + 3 * (size_t)ci + 2);
          float t, be, ga;
          if (!(TRI_MODE == 7 && culling && dot(cd, mk3(t2.y, t2.z, t2.w)) > 0.0f) &&
              tri_test<TRI_MODE>(mk3(t0.x, t0.y, t0.z), mk3(t0.w, t1.x, t1.y), mk3(t1.z, t1.w, t2.x), co, cd, eps,
                                 &t, &be, &ga) && t > 0.0f) {
            if (ANY) { if (t < tlimit) return true; }
            else if (better(t, cur_rank, (int)ci, best)) {
              best.t = t; best.beta = be; best.gamma = ga; best.top_ref = cur_top; best.top_rank = cur_rank;
              best.tri = (int)ci; found = true;
            }
          }
        } else if (ck == B200RT_KIND_SPHERE) {
          cnt.spheres++;
          const b200rt_sphere* s = S.spheres + ci;
          float t;
          if (sphere_test(s, co, cd, time, eps, &t) && t > 0.0f) {
            if (ANY) { if (t < tlimit) return true; }
            else if (better(t, s->rank, -1, best)) {
              best.t = t; best.beta = 0.f; best.gamma = 0.f; best.top_ref = child[c]; best.top_rank = s->rank;
              best.tri = -1; found = true;
            }
          }
        } else if (ck == B200RT_KIND_TRIPRIM) {
          cnt.tris++;
          const b200rt_triprim* p = S.triprims + ci;
          v3 lo = co, ldir = cd;
          if (!(p->flags & B200RT_XF_IDENTITY)) { lo = xf_point(p->inv, co); ldir = xf_vector4(p->inv, cd); }
          float t, be, ga;
          if (!(TRI_MODE == 7 && culling && dot(ldir, ld3(p->normal)) > 0.0f) &&
              tri_test<2>(ld3(p->v0), ld3(p->a1), ld3(p->a2), lo, ldir, eps, &t, &be, &ga) && t > 0.0f) {
            if (ANY) { if (t < tlimit) return true; }
            else if (better(t, p->rank, -1, best)) {
              best.t = t; best.beta = be; best.gamma = ga; best.top_ref = child[c]; best.top_rank = p->rank;
              best.tri = -1; found = true;
            }
          }
        }
      }
      // push far first so that the near child is popped next
      if (push[0] && push[1]) {
        const int far_c = (entry[0] > entry[1]) ? 0 : 1;
        if (sp + 2 > RT_STACK) { continue; }
        stack[sp] = child[far_c]; stack_t[sp] = entry[far_c]; sp++;
        stack[sp] = child[1 - far_c]; stack_t[sp] = entry[1 - far_c]; sp++;
      } else if (push[0] || push[1]) {
        const int c = push[0] ? 0 : 1;
        if (sp + 1 > RT_STACK) { continue; }
        stack[sp] = child[c]; stack_t[sp] = entry[c]; sp++;
      }
    } else if (kind == B200RT_KIND_INSTANCE || kind == B200RT_KIND_LIGHTINST) {
      // Mesh_instance::intersect — HW3/include/Mesh.h:55-98 (gate already passed)
      const b200rt_instance* in = S.inst + idx;
      const b200rt_mesh* m = S.meshes + in->mesh;
      const uint32_t mk = B200RT_REF_KIND(m->root);
      if (mk == B200RT_KIND_NONE) continue;
      cnt.insts++;
      if (!(in->flags & B200RT_INST_DIRECT)) {
        if (in->flags & B200RT_INST_MOVING) {
          float inv16[16];
          moving_inverse(in->fwd, in->velocity, time, inv16);
          co = xf_point(inv16, o); cd = xf_vector4(inv16, d);
        } else {
          co = xf_point(in->inv, o); cd = xf_vector4(in->inv, d);
        }
      }
      cr = make_ray(co, cd, eps);
      cur_rank = in->rank; cur_top = ref;
      culling = hw7 && !(in->flags & B200RT_INST_REFRACTIVE);
      if (sp + 2 > RT_STACK) { co = o; cd = d; cr = wr; cur_rank = 0x7fffffff; cur_top = B200RT_REF_NONE; culling = hw7; continue; }
      stack[sp] = RT_SENTINEL; stack_t[sp] = 0.f; sp++;
      if (mk == B200RT_KIND_INNER) {
        float e;
        if (box_gate(cr, m->root_box[0], m->root_box[1], m->root_box[2], m->root_box[3], m->root_box[4],
                     m->root_box[5], eps, &e)) {
          stack[sp] = m->root; stack_t[sp] = e; sp++;
        }
      } else {
        stack[sp] = m->root; stack_t[sp] = 0.f; sp++;  // single-triangle mesh: handled below
      }
    } else if (kind == B200RT_KIND_TRI) {
      // a mesh whose BVH is a single triangle (create_bvh size==1)
      cnt.tris++;
      const float4 t0 = __ldg(S.tris + 3 * (size_t)idx);
      const float4 t1 = __ldg(S.tris + 3 * (size_t)idx + 1);
      const float4 t2 = __ldg(S.tris + 3 * (size_t)idx + 2);
      float t, be, ga;
      if (!(TRI_MODE == 7 && culling && dot(cd, mk3(t2.y, t2.z, t2.w)) > 0.0f) &&
          tri_test<TRI_MODE>(mk3(t0.x, t0.y, t0.z), mk3(t0.w, t1.x, t1.y), mk3(t1.z, t1.w, t2.x), co, cd, eps, &t,
                             &be, &ga) && t > 0.0f) {
        if (ANY) { if (t < tlimit) return true; }
        else if (better(t, cur_rank, (int)idx, best)) {
          best.t = t; best.beta = be; best.gamma = ga; best.top_ref = cur_top; best.top_rank = cur_rank;
          best.tri = (int)idx; found = true;
        }
      }
    } else if (kind == B200RT_KIND_SPHERE) {
      cnt.spheres++;
      const b200rt_sphere* s = S.spheres + idx;
      float t;
      const bool lone = !ANY && ref == S.root;  // root is the sphere itself: no parent filters t > 0 (Appendix B.10)
      if (sphere_test(s, co, cd, time, eps, &t) && (t > 0.0f || lone)) {
        if (ANY) { if (t < tlimit) return true; }
        else if (better(t, s->rank, -1, best)) {
          best.t = t; best.top_ref = ref; best.top_rank = s->rank; best.tri = -1; found = true;
        }
      }
    } else if (kind == B200RT_KIND_TRIPRIM) {
      cnt.tris++;
      const b200rt_triprim* p = S.triprims + idx;
      v3 lo = co, ldir = cd;
      if (!(p->flags & B200RT_XF_IDENTITY)) { lo = xf_point(p->inv, co); ldir = xf_vector4(p->inv, cd); }
      float t, be, ga;
      if (!(TRI_MODE == 7 && culling && dot(ldir, ld3(p->normal)) > 0.0f) &&
          tri_test<2>(ld3(p->v0), ld3(p->a1), ld3(p->a2), lo, ldir, eps, &t, &be, &ga) && t > 0.0f) {
        if (ANY) { if (t < tlimit) return true; }
        else if (better(t, p->rank, -1, best)) {
          best.t = t; best.beta = be; best.gamma = ga; best.top_ref = ref; best.top_rank = p->rank; best.tri = -1;
          found = true;
        }
      }
    }
  }
  return found;
}
