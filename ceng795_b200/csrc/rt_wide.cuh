// rt_wide.cuh — the default traversal: persistent warps, a flat top-level object list staged in shared memory, and
// inside a mesh the 8-wide tree with quantised child boxes of csrc/wide_bvh.h, walked with a stack of node GROUPS that
// lives in shared memory (no per-thread local arrays at all).
//
// Why the answer is still the reference's.  In the reference a mesh triangle k is reachable iff every
// BVH node on its root path passes Bounding_box::intersect (HW7/src/Bounding_volume_hierarchy.cpp:31-35).  Those boxes
// are nested and the slab test is monotone under IEEE rounding, so the chain passes iff the innermost box does; that
// box is stored per triangle (`tri_gate`).  A triangle found through OUR tree is admitted iff the reference's triangle
// test passes (same code, same object-space ray) and box_gate(tri_gate[k]) passes with the reference's own
// arithmetic.  Our boxes only need to be conservative: they are grown on the host, quantised outwards, and tested
// with slack.  The winner is the minimum of the key (t, top-level DFS rank, triangle DFS index), so neither the
// visiting order nor pruning can change it.
//
// Node step (one 80-byte fetch = five float4): the 8 child boxes are tested in the node's quantisation frame
// (t = q * (2^e / d) + (p - o) / d, one FMA per plane; q reaches the FMA through a byte-permute, not a conversion:
// wide_q2f) and leave one bit per hit SLOT.  Once per node, not per child, the inner slots among them are permuted by
// the ray's octant — bit 24 + (slot ^ octant), so "highest bit first" is front to back without sorting — and become the
// lane's current group {child base, unvisited inner slots}; the previous group is pushed only when it still has
// unvisited members, so the stack depth is the tree depth (<= 16), 8 bytes per level.  The leaf slots are kept as
// {node, slots}; which triangles they hold (unary count and offset in the node's meta bytes) is decoded when the lane
// gets to test them — at once for a closest-hit ray, from a short queue in shared memory for a shadow ray.
//
// Everything a lane does is written as functions of a WideLane so that tests/wide_host.cpp can run the SAME source
// on the host (g++) against the oracle; only the warp scheduling below (#ifdef __CUDACC__) is device-only.
#pragma once
#include "rt_traverse.cuh"
#include "wide_bvh.h"   // the node encoding (RT_WIDE_BF16) is shared with the host builder

#define RT_WNODE_F4 (RT_WIDE_BF16 ? 8 : 5)   // float4 per node

#define RT_WSTACK 16        // shared-memory stack entries per thread (node groups); scenes whose wide tree is deeper use k_trace_fast
#ifndef RT_WIDE_NODE_VOTE
#define RT_WIDE_NODE_VOTE 8   // lanes that must still want a node step for phase 1 to go on (swept: profiles/r02_tuning.txt)
#endif
#ifndef RT_WIDE_TQ_CLOSEST
#define RT_WIDE_TQ_CLOSEST 0   // the same for closest-hit rays (costs pruning: a queued leaf cannot shrink best.t yet)
#endif
#ifndef RT_WIDE_TQ
#define RT_WIDE_TQ 4   // leaf steps a shadow ray may queue before it has to test them (0 = test at once; swept)
#endif
#ifndef RT_TOP_LINEAR_MAX
#define RT_TOP_LINEAR_MAX 40  // top-level objects up to which phase 0 scans a flat gate list (one bit each in the lane's object mask)
#endif

struct WideRay {
  v3 o, d, inv;
  uint32_t oct_inv;   // bit 2/1/0 set: the ray travels towards +x/+y/+z
};

// 1/d for OUR boxes: a component below 1e-18 (zero, denormal) is replaced by +-1e-18 so that no product overflows or
// turns into a NaN.  Such a ray moves less than 1e-18 * t along that axis, far below the growth of the boxes.
__device__ __forceinline__ float wide_safe_inv(float d) {
  const float a = fabsf(d);
  return copysignf(1.0f / (a < 1e-18f ? 1e-18f : a), d);
}
__device__ __forceinline__ WideRay make_wide_ray(v3 o, v3 d) {
  WideRay r;
  r.o = o; r.d = d;
  r.inv = mk3(wide_safe_inv(d.x), wide_safe_inv(d.y), wide_safe_inv(d.z));
  r.oct_inv = (r.inv.x < 0.0f ? 0u : 4u) | (r.inv.y < 0.0f ? 0u : 2u) | (r.inv.z < 0.0f ? 0u : 1u);
  return r;
}

struct WideLane {
  WideRay cr;                   // the ray in the space of the object being traversed
  uint2 G;                      // node group: x = child_base, y = hit bits 31..24 | imask 7..0
  uint2 T;                      // triangle group: x = tri_base, y = hit bits 23..0
  unsigned long long objmask;   // top-level objects whose gate passed and that have not been entered yet
  HitRec best;
  bool found;
  int cur_rank;
  uint32_t cur_top;
  bool culling;                 // HW7/PPM back-face culling, off inside refractive instances
};

__device__ __forceinline__ void wide_reset(WideLane& L, bool hw7) {
  L.G = make_uint2(0u, 0u); L.T = make_uint2(0u, 0u); L.objmask = 0ull;
  L.best.t = RT_INF; L.best.beta = 0.f; L.best.gamma = 0.f; L.best.top_ref = B200RT_REF_NONE;
  L.best.top_rank = 0x7fffffff; L.best.tri = 0x7fffffff;
  L.found = false; L.cur_rank = 0x7fffffff; L.cur_top = B200RT_REF_NONE; L.culling = hw7;
}

// ---- phase 0: the reference's gates of the top-level objects (BVH::intersect HW3/src/Bounding_volume_hierarchy.cpp:
// 31-55, Mesh_instance's own gate HW3/include/Mesh.h:59-62), one bit per object that has to be entered
__device__ __forceinline__ unsigned long long wide_top_mask(const float4* top, int n_top, const RayCtx& wr,
                                                                     float eps, TravCounters& tc) {
  unsigned long long m = 0ull;
  for (int k = 0; k < n_top; k++) {
    const float4 b0 = top[2 * k], b1 = top[2 * k + 1];
    float e = 0.0f;
    tc.nodes++;
    if (__float_as_uint(b1.w) == 0u || box_gate(wr, b0.x, b0.y, b0.z, b1.x, b1.y, b1.z, eps, &e)) m |= 1ull << k;
  }
  return m;
}

// Where the ray enters the gate box of a top-level object (0 for an object without a gate): ordering and pruning
// only, the decision was taken by box_gate.  Axes the reference skips (|d| < eps) are skipped here too.
__device__ __forceinline__ float wide_top_entry(const float4 b0, const float4 b1, const RayCtx& wr, float eps) {
  if (__float_as_uint(b1.w) == 0u) return 0.0f;
  float tn = 0.0f;
  if (!(fabsf(wr.d.x) < eps)) tn = fmaxf(tn, fminf((b0.x - wr.o.x) * wr.inv.x, (b1.x - wr.o.x) * wr.inv.x));
  if (!(fabsf(wr.d.y) < eps)) tn = fmaxf(tn, fminf((b0.y - wr.o.y) * wr.inv.y, (b1.y - wr.o.y) * wr.inv.y));
  if (!(fabsf(wr.d.z) < eps)) tn = fmaxf(tn, fminf((b0.z - wr.o.z) * wr.inv.z, (b1.z - wr.o.z) * wr.inv.z));
  return tn;
}

// The pending object the ray enters first; objects that start beyond the best hit are dropped.  -1: none left.
__device__ __forceinline__ int wide_select(const float4* top, unsigned long long& mask, const RayCtx& wr,
                                                    float eps, float bound_slack) {
  int best_k = -1;
  float best_e = RT_INF;
  unsigned long long m = mask;
  while (m != 0ull) {
    const int k = __ffsll((long long)m) - 1;
    m &= m - 1ull;
    const float e = wide_top_entry(top[2 * k], top[2 * k + 1], wr, eps);
    if (e > bound_slack) { mask &= ~(1ull << k); continue; }
    if (best_k < 0 || e < best_e) { best_k = k; best_e = e; }
  }
  if (best_k >= 0) mask &= ~(1ull << best_k);
  return best_k;
}

// ---- one node of the wide tree ---------------------------------------------------------------------------------
// 32768 + (byte k of w), as a float, in one byte-permute: the quantised plane dropped into mantissa bits 8..15 of 2^15.
// (float)(byte) would be an I2F, which shares the quarter-rate conversion / special-function pipe with the divisions of
// the triangle test: 48 of them per node step kept that pipe 57 % busy and made it the limiter of a node step
// (profiles/r02f: 8 cycles per warp instruction there against 2 on the FMA and ALU pipes).  The 2^15 is taken out
// again by the constant term of the plane's FMA (c = b - 32768 a), at a cost of a/1024 in precision — a thousandth of
// a quantisation step, far inside the growth of the boxes.
__device__ __forceinline__ float wide_q2f(uint32_t w, int k) {
#ifdef __CUDA_ARCH__
  return __uint_as_float(__byte_perm(0x47000000u, w, 0x3040u | ((uint32_t)k << 4)));   // selector as an immediate
#else
  return __uint_as_float(0x47000000u | (((w >> (8 * k)) & 0xffu) << 8));
#endif
}

// Planes converted the other way, with a plain (float)(byte) on the conversion pipe: after the byte-permutes the ALU pipe
// is the busy one (62 %, profiles/r02h) and the conversion pipe idles at 14 %, so a share of the 48 conversions of a
// node step goes back there.  0: none, 1: the z planes (16), 2: z planes and far y planes (24).  Swept on the device.
#ifndef RT_WIDE_I2F_PLANES
#define RT_WIDE_I2F_PLANES 1
#endif
__device__ __forceinline__ float wide_q2i(uint32_t w, int k) { return (float)((w >> (8 * k)) & 0xffu); }

// Slots (bit s = slot s) of the node {n0, n2..n4} whose child box ray r enters no later than bound_slack.  An empty slot
// has an inside-out box and does not hit; if rounding ever let one through, it would name no inner child (imask bit 0)
// and no triangle (meta 0), so nothing depends on it.  Written so that the per-child work is conversions, FMAs,
// three-input min / max, one compare and one predicated OR: the ALU and conversion pipes are the busy ones in this
// kernel (profiles/r02f), so what can be done once per node (telling inner from leaf slots, front-to-back order,
// triangle bits) is done once per node, outside this loop.
__device__ __forceinline__ uint32_t wide_node_hits(const float4 n0, const float4 n2, const float4 n3, const float4 n4,
                                                            const WideRay& r, float bound_slack) {
  const uint32_t ei = __float_as_uint(n0.w);
  // not an accept/reject decision of the reference: FMA allowed
  const float ax = __uint_as_float((ei & 0xffu) << 23) * r.inv.x;
  const float ay = __uint_as_float(((ei >> 8) & 0xffu) << 23) * r.inv.y;
  const float az = __uint_as_float(((ei >> 16) & 0xffu) << 23) * r.inv.z;
  // plane t = q * a + b = (32768 + q) * a + c; the far planes carry the slack of the comparison (1 + 4e-6) as well
  const float by = (n0.y - r.o.y) * r.inv.y, bz = (n0.z - r.o.z) * r.inv.z;
  const float cx = __fmaf_rn(-32768.0f, ax, (n0.x - r.o.x) * r.inv.x);
  const float cy = __fmaf_rn(-32768.0f, ay, by);
  const float cz = __fmaf_rn(-32768.0f, az, bz);
  const bool px = !(r.inv.x < 0.0f), py = !(r.inv.y < 0.0f), pz = !(r.inv.z < 0.0f);
  uint32_t hits = 0u;
#pragma unroll
  for (int h = 0; h < 2; h++) {
    const uint32_t lox = __float_as_uint(h ? n2.y : n2.x), loy = __float_as_uint(h ? n2.w : n2.z);
    const uint32_t loz = __float_as_uint(h ? n3.y : n3.x), hix = __float_as_uint(h ? n3.w : n3.z);
    const uint32_t hiy = __float_as_uint(h ? n4.y : n4.x), hiz = __float_as_uint(h ? n4.w : n4.z);
    const uint32_t nx4 = px ? lox : hix, fx4 = px ? hix : lox;
    const uint32_t ny4 = py ? loy : hiy, fy4 = py ? hiy : loy;
    const uint32_t nz4 = pz ? loz : hiz, fz4 = pz ? hiz : loz;
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const float tnx = __fmaf_rn(wide_q2f(nx4, k), ax, cx);
      const float tny = __fmaf_rn(wide_q2f(ny4, k), ay, cy);
      const float tnz = RT_WIDE_I2F_PLANES >= 1 ? __fmaf_rn(wide_q2i(nz4, k), az, bz) : __fmaf_rn(wide_q2f(nz4, k), az, cz);
      const float tfx = __fmaf_rn(wide_q2f(fx4, k), ax, cx);
      const float tfy = RT_WIDE_I2F_PLANES >= 2 ? __fmaf_rn(wide_q2i(fy4, k), ay, by) : __fmaf_rn(wide_q2f(fy4, k), ay, cy);
      const float tfz = RT_WIDE_I2F_PLANES >= 1 ? __fmaf_rn(wide_q2i(fz4, k), az, bz) : __fmaf_rn(wide_q2f(fz4, k), az, cz);
      const float tn = fmaxf(fmaxf(tnx, tny), fmaxf(tnz, 0.0f));
      const float tf = fminf(fminf(tfx, tfy), tfz);
      // slack of the comparison: relative to tf (a box behind the origin, tf < 0, misses either way)
      if (tn <= fminf(tf * 1.000004f, bound_slack)) hits |= 1u << (4 * h + k);   // one compare, one predicated OR
    }
  }
  return hits;
}

// Bit s of x moves to bit s ^ oct (oct < 8): the order in which a ray with direction octant `oct` meets the slots.
__device__ __forceinline__ uint32_t wide_octant_permute(uint32_t x, uint32_t oct) {
  if (oct & 1u) x = ((x & 0x55u) << 1) | ((x >> 1) & 0x55u);
  if (oct & 2u) x = ((x & 0x33u) << 2) | ((x >> 2) & 0x33u);
  if (oct & 4u) x = ((x & 0x0fu) << 4) | (x >> 4);
  return x;
}

#if RT_WIDE_BF16
// The same answer for the bfloat16 encoding (wide_bvh.h).  A slot's word per axis is (hi << 16) | lo, both offsets from
// the node's origin.  `w * 65536` is lo as a float, exactly; `w * 1` is hi with the lo bits left in its low mantissa,
// which can only enlarge the box.  Which of the two is the near plane depends on the sign of the direction, and that
// choice is the multiplier: one IMAD on the FMA pipe per plane, no select, no shift, nothing on the ALU pipe.
__device__ __forceinline__ uint32_t wide_node_hits_bf16(const float4 n0, const float4* pl, const WideRay& r, float bound_slack) {
  const float bx = (n0.x - r.o.x) * r.inv.x, by = (n0.y - r.o.y) * r.inv.y, bz = (n0.z - r.o.z) * r.inv.z;
  // 65536 where the ray travels towards + on that axis (near = lo), else 1 (near = hi); the far multiplier is the other
  const uint32_t ox = (r.oct_inv >> 2) & 1u, oy = (r.oct_inv >> 1) & 1u, oz = r.oct_inv & 1u;
  const uint32_t mnx = 1u + 65535u * ox, mny = 1u + 65535u * oy, mnz = 1u + 65535u * oz;
  const uint32_t mfx = 65536u - 65535u * ox, mfy = 65536u - 65535u * oy, mfz = 65536u - 65535u * oz;
  const float fsx = r.inv.x * 1.000004f, fsy = r.inv.y * 1.000004f, fsz = r.inv.z * 1.000004f;   // far planes carry the slack
  const float fbx = bx * 1.000004f, fby = by * 1.000004f, fbz = bz * 1.000004f;
  uint32_t hits = 0u;
#pragma unroll
  for (int h = 0; h < 2; h++) {
    const float4 X = __ldg(pl + h), Y = __ldg(pl + 2 + h), Z = __ldg(pl + 4 + h);
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const uint32_t wx = __float_as_uint(k == 0 ? X.x : k == 1 ? X.y : k == 2 ? X.z : X.w);
      const uint32_t wy = __float_as_uint(k == 0 ? Y.x : k == 1 ? Y.y : k == 2 ? Y.z : Y.w);
      const uint32_t wz = __float_as_uint(k == 0 ? Z.x : k == 1 ? Z.y : k == 2 ? Z.z : Z.w);
      const float tnx = __fmaf_rn(__uint_as_float(wx * mnx), r.inv.x, bx);
      const float tny = __fmaf_rn(__uint_as_float(wy * mny), r.inv.y, by);
      const float tnz = __fmaf_rn(__uint_as_float(wz * mnz), r.inv.z, bz);
      const float tfx = __fmaf_rn(__uint_as_float(wx * mfx), fsx, fbx);
      const float tfy = __fmaf_rn(__uint_as_float(wy * mfy), fsy, fby);
      const float tfz = __fmaf_rn(__uint_as_float(wz * mfz), fsz, fbz);
      const float tn = fmaxf(fmaxf(tnx, tny), fmaxf(tnz, 0.0f));
      const float tf = fminf(fminf(tfx, tfy), tfz);
      if (tn <= fminf(tf, bound_slack)) hits |= 1u << (4 * h + k);
    }
  }
  return hits;
}
#endif

// Takes the nearest unvisited child of the lane's node group, tests its 8 children and makes them the new groups.
// Returns true when the old group still has unvisited children: the caller pushes `rest`.
__device__ __forceinline__ bool wide_node_step(const float4* wide_nodes, WideLane& L, float bound_slack,
                                                        uint2& rest, TravCounters& tc) {
  const uint32_t bit = 31u - (uint32_t)__clz((int)L.G.y);
  const uint32_t slot = (bit - 24u) ^ L.cr.oct_inv;
  const uint32_t rel = (uint32_t)__popc(L.G.y & 0xffu & ((1u << slot) - 1u));
  const size_t idx = (size_t)L.G.x + rel;
  rest = make_uint2(L.G.x, L.G.y & ~(1u << bit));
  const bool more = (rest.y & 0xff000000u) != 0u;
  tc.nodes++;
  const float4 n0 = __ldg(wide_nodes + RT_WNODE_F4 * idx);
  const float child_base = __ldg(reinterpret_cast<const float*>(wide_nodes + RT_WNODE_F4 * idx + 1));
#if RT_WIDE_BF16
  const uint32_t hits = wide_node_hits_bf16(n0, wide_nodes + RT_WNODE_F4 * idx + 2, L.cr, bound_slack);
#else
  const float4 n2 = __ldg(wide_nodes + 5 * idx + 2), n3 = __ldg(wide_nodes + 5 * idx + 3);
  const float4 n4 = __ldg(wide_nodes + 5 * idx + 4);
  const uint32_t hits = wide_node_hits(n0, n2, n3, n4, L.cr, bound_slack);
#endif
  const uint32_t imask = __float_as_uint(n0.w) >> 24;
  // inner slots: front to back for this ray is "highest bit first" after the octant permutation
  L.G = make_uint2(__float_as_uint(child_base), (wide_octant_permute(hits & imask, L.cr.oct_inv) << 24) | imask);
  // leaf slots: which triangles they hold is looked up when (and if) the lane gets to test them (wide_triangles)
  L.T = make_uint2((uint32_t)idx, hits & ~imask);
  return more;
}

// The lane's pending triangles (all of them): the reference's test, then the reference's innermost gate.
// Returns true when an any-hit ray has found its occluder.
template <bool ANY, int TRI_MODE>
__device__ __forceinline__ bool wide_triangles(const DScene& S, WideLane& L, float tlimit, float eps,
                                                        TravCounters& tc) {
  // L.T = {node, hit leaf slots}: the node's second float4 says where its triangles start and, per slot, how many
  // (unary 1 / 3 / 7 in the top bits of the meta byte) at which offset (low 5 bits): up to 24 triangles per node
  const float4 n1 = __ldg(S.wide_nodes + RT_WNODE_F4 * (size_t)L.T.x + 1);
  const uint32_t tri_base = __float_as_uint(n1.y);
  uint32_t slots = L.T.y, bits = 0u;
  L.T.y = 0u;
  while (slots != 0u) {
    const uint32_t s_ = (uint32_t)__ffs((int)slots) - 1u;
    slots &= slots - 1u;
    const uint32_t meta = (__float_as_uint(s_ < 4u ? n1.z : n1.w) >> (8u * (s_ & 3u))) & 0xffu;
    bits |= (meta >> 5) << (meta & 31u);
  }
  while (bits != 0u) {
    const uint32_t b = (uint32_t)__ffs((int)bits) - 1u;
    bits &= bits - 1u;
    const size_t fi = (size_t)tri_base + b;
    tc.tris++;
    const float4 t0 = __ldg(S.wide_tris + 3 * fi), t1 = __ldg(S.wide_tris + 3 * fi + 1), t2 = __ldg(S.wide_tris + 3 * fi + 2);
    float t, be, ga;
    if (!(TRI_MODE == 7 && L.culling && dot(L.cr.d, mk3(t2.y, t2.z, t2.w)) > 0.0f) &&
        tri_test<TRI_MODE>(mk3(t0.x, t0.y, t0.z), mk3(t0.w, t1.x, t1.y), mk3(t1.z, t1.w, t2.x), L.cr.o, L.cr.d, eps, &t,
                           &be, &ga) && t > 0.0f && (ANY ? (t < tlimit) : (t <= L.best.t))) {
      const int tri = __ldg(S.wide_tri_ref + fi);
      if (ANY || better(t, L.cur_rank, tri, L.best)) {
        const float2* g = reinterpret_cast<const float2*>(S.tri_gate + 6 * (size_t)tri);
        const float2 g0 = __ldg(g), g1 = __ldg(g + 1), g2 = __ldg(g + 2);
        RayCtx c;
        c.o = L.cr.o; c.d = L.cr.d;
        c.slow = (fabsf(c.d.x) < eps) || (fabsf(c.d.y) < eps) || (fabsf(c.d.z) < eps);
        c.inv = mk3(1.0f / c.d.x, 1.0f / c.d.y, 1.0f / c.d.z);   // the gate's own reciprocal (IEEE division), not the clamped one
        float e;
        if (box_gate(c, g0.x, g0.y, g1.x, g1.y, g2.x, g2.y, eps, &e)) {
          if (ANY) { L.found = true; return true; }
          L.best.t = t; L.best.beta = be; L.best.gamma = ga; L.best.top_ref = L.cur_top; L.best.top_rank = L.cur_rank;
          L.best.tri = tri; L.found = true;
        }
      }
    }
  }
  return false;
}

// Enters top-level object `ref` (its gate has passed): an instance sets up the object-space ray and the root group,
// a sphere / stand-alone triangle is tested on the spot.  Returns true when an any-hit ray has found its occluder.
template <bool ANY, int TRI_MODE>
__device__ __forceinline__ bool wide_enter(const DScene& S, WideLane& L, uint32_t ref, v3 o, v3 dir, float time,
                                                    float tlimit, float eps, bool hw7, TravCounters& tc) {
  const uint32_t kind = B200RT_REF_KIND(ref), idx = B200RT_REF_INDEX(ref);
  if (kind == B200RT_KIND_INSTANCE || kind == B200RT_KIND_LIGHTINST) {
    // Mesh_instance::intersect — HW3/include/Mesh.h:55-98 (gate already passed; a Light_mesh has none)
    const b200rt_instance* in = S.inst + idx;
    const b200rt_mesh* m = S.meshes + in->mesh;
    const int wroot = S.wide_root[in->mesh];
    if (wroot < 0 && B200RT_REF_KIND(m->root) != B200RT_KIND_TRI) return false;   // a mesh without triangles
    tc.insts++;
    v3 co = o, cd = dir;
    if (!(in->flags & B200RT_INST_DIRECT)) {
      if (in->flags & B200RT_INST_MOVING) {
        float inv16[16];
        moving_inverse(in->fwd, in->velocity, time, inv16);
        co = xf_point(inv16, o); cd = xf_vector4(inv16, dir);
      } else {
        co = xf_point(in->inv, o); cd = xf_vector4(in->inv, dir);
      }
    }
    L.cr = make_wide_ray(co, cd);
    L.cur_rank = in->rank; L.cur_top = ref;
    L.culling = hw7 && !(in->flags & B200RT_INST_REFRACTIVE);
    if (wroot >= 0) {
      L.G = make_uint2((uint32_t)wroot, (1u << (24u + L.cr.oct_inv)) | 1u);   // "slot 0 of a group whose base is the root"
      return false;
    }
    // a mesh whose BVH is a single triangle (create_bvh size==1): no gate at all
    const uint32_t ti = B200RT_REF_INDEX(m->root);
    tc.tris++;
    const float4 t0 = __ldg(S.tris + 3 * (size_t)ti), t1 = __ldg(S.tris + 3 * (size_t)ti + 1), t2 = __ldg(S.tris + 3 * (size_t)ti + 2);
    float t, be, ga;
    if (!(TRI_MODE == 7 && L.culling && dot(cd, mk3(t2.y, t2.z, t2.w)) > 0.0f) &&
        tri_test<TRI_MODE>(mk3(t0.x, t0.y, t0.z), mk3(t0.w, t1.x, t1.y), mk3(t1.z, t1.w, t2.x), co, cd, eps, &t, &be, &ga) &&
        t > 0.0f) {
      if (ANY) { if (t < tlimit) { L.found = true; return true; } }
      else if (better(t, L.cur_rank, (int)ti, L.best)) {
        L.best.t = t; L.best.beta = be; L.best.gamma = ga; L.best.top_ref = ref; L.best.top_rank = L.cur_rank;
        L.best.tri = (int)ti; L.found = true;
      }
    }
    return false;
  }
  if (kind == B200RT_KIND_SPHERE) {
    tc.spheres++;
    const b200rt_sphere* s = S.spheres + idx;
    float t;
    // Sphere::intersect returns true even for t < 0 and only a BVH parent filters `t > 0`: a scene whose root IS the
    // sphere (create_bvh size==1) reports the negative hit
    const bool lone = !ANY && ref == S.root;
    if (sphere_test(s, o, dir, time, eps, &t) && (t > 0.0f || lone)) {
      if (ANY) { if (t < tlimit) { L.found = true; return true; } }
      else if (better(t, s->rank, -1, L.best)) {
        L.best.t = t; L.best.beta = 0.f; L.best.gamma = 0.f; L.best.top_ref = ref; L.best.top_rank = s->rank;
        L.best.tri = -1; L.found = true;
      }
    }
    return false;
  }
  if (kind == B200RT_KIND_TRIPRIM) {
    tc.tris++;
    const b200rt_triprim* p = S.triprims + idx;
    v3 lo = o, ldir = dir;
    if (!(p->flags & B200RT_XF_IDENTITY)) { lo = xf_point(p->inv, o); ldir = xf_vector4(p->inv, dir); }
    float t, be, ga;
    if (!(TRI_MODE == 7 && hw7 && dot(ldir, ld3(p->normal)) > 0.0f) &&   // top-level objects are culled (HW7)
        tri_test<2>(ld3(p->v0), ld3(p->a1), ld3(p->a2), lo, ldir, eps, &t, &be, &ga) && t > 0.0f) {
      if (ANY) { if (t < tlimit) { L.found = true; return true; } }
      else if (better(t, p->rank, -1, L.best)) {
        L.best.t = t; L.best.beta = be; L.best.gamma = ga; L.best.top_ref = ref; L.best.top_rank = p->rank;
        L.best.tri = -1; L.found = true;
      }
    }
  }
  return false;
}

// One ray, start to finish, with the stack in a plain array: what a lane of k_trace_wide computes, without the warp
// scheduling.  Used by tests/wide_host.cpp (the host twin) — the kernel below is the product path.
template <bool ANY, int TRI_MODE>
__device__ __forceinline__ bool wide_trace_one(const DScene& S, const float4* top, v3 o, v3 d, float time,
                                                        float tlimit, HitRec& out, TravCounters& tc, int* max_sp) {
  const float eps = S.isect_eps;
  const bool hw7 = S.semantics >= B200RT_SEM_HW7;
  WideLane L;
  wide_reset(L, hw7);
  L.cr = make_wide_ray(o, d);
  const RayCtx wr = make_ray(o, d, eps);
  L.objmask = wide_top_mask(top, S.n_top, wr, eps, tc);
  uint2 stack[64];
  int sp = 0;
  for (;;) {
    const float bound = ANY ? tlimit : L.best.t;
    const float bs = bound + fabsf(bound) * 1e-4f;
    if (L.T.y != 0u) {
      if (wide_triangles<ANY, TRI_MODE>(S, L, tlimit, eps, tc)) break;
    } else if (L.G.y & 0xff000000u) {
      uint2 rest;
      if (wide_node_step(S.wide_nodes, L, bs, rest, tc)) { stack[sp++] = rest; if (max_sp && sp > *max_sp) *max_sp = sp; }
    } else if (sp > 0) {
      L.G = stack[--sp];
    } else {
      const int k = wide_select(top, L.objmask, wr, eps, bs);
      if (k < 0) break;
      if (wide_enter<ANY, TRI_MODE>(S, L, __float_as_uint(top[2 * k].w), o, d, time, tlimit, eps, hw7, tc)) break;
    }
  }
  out = L.best;
  return L.found;
}

#ifdef __CUDACC__
#ifndef RT_WIDE_MIN_BLOCKS
#define RT_WIDE_MIN_BLOCKS 8  // default resident CTAs of 128 threads per SM the register budget is held to (64 registers; swept: profiles/r02_tuning.txt)
#endif

// MINB: resident CTAs per SM the register budget is held to (4: 128 registers, 6: 80, 8: 64) — instantiated for all
// three so that one run on the device can sweep it (B200RT_WIDE_BLOCKS), profiles/r02*_tuning.txt.
// Phase 0 as a pass of its own: the reference's top-level gates of every ray of the wave, 32 lanes wide, one 64-bit
// mask per ray.  The traversal kernel then starts a ray by loading 8 bytes, so topping a warp up costs next to
// nothing and idle lanes are refilled early (S.refill_*) instead of waiting until the warp has almost drained.
template <bool ANY>
__global__ void __launch_bounds__(256) k_top_mask(DScene S, Pool P, ShadowQueue Q, const WaveInfo* waves, int d, Counters* C) {
  __shared__ float4 s_top[2 * RT_TOP_LINEAR_MAX];
  for (int i = threadIdx.x; i < 2 * S.n_top; i += blockDim.x) s_top[i] = __ldg(S.top_objs + i);
  __syncthreads();
  const uint32_t begin = ANY ? 0u : waves[d].begin;
  const uint32_t count = ANY ? min(C->shadow_count, Q.capacity) : waves[d].count;
  const float eps = S.isect_eps;
  TravCounters tc = {0, 0, 0, 0};
  for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < count; k += gridDim.x * blockDim.x) {
    const float4 ro = ANY ? Q.o[k] : P.ray_o[begin + k];
    const float4 rd = ANY ? Q.d[k] : P.ray_d[begin + k];
    const RayCtx wr = make_ray(mk3(ro.x, ro.y, ro.z), mk3(rd.x, rd.y, rd.z), eps);
    const unsigned long long m = wide_top_mask(s_top, S.n_top, wr, eps, tc);
    if (ANY) Q.omask[k] = m; else P.omask[begin + k] = m;
  }
  flush_counters(C, tc, ANY);
}

template <bool ANY, int TRI_MODE, int MINB>
__global__ void __launch_bounds__(RT_BLOCK, MINB) k_trace_wide(DScene S, Pool P, ShadowQueue Q,
                                                               const WaveInfo* waves, int d, Counters* C) {
  __shared__ uint2 s_stack[RT_WSTACK][RT_BLOCK];
  __shared__ float4 s_top[2 * RT_TOP_LINEAR_MAX];
#if RT_WIDE_TQ
  // any-hit rays only: leaf steps are QUEUED (up to RT_WIDE_TQ per lane) while the lane goes on descending, and tested
  // together in phase 2 — no pruning is lost (a shadow ray has no closest hit to shrink towards), an occluded ray
  // merely learns it a little later
  __shared__ uint2 s_tq[RT_WIDE_TQ][RT_BLOCK];
#endif
  int tqn = 0;
  for (int i = threadIdx.x; i < 2 * S.n_top; i += RT_BLOCK) s_top[i] = __ldg(S.top_objs + i);
  __syncthreads();

  const uint32_t begin = ANY ? 0u : waves[d].begin;
  const uint32_t count = ANY ? min(C->shadow_count, Q.capacity) : waves[d].count;
  unsigned int* cursor = ANY ? &C->work_shadow : &C->work_extend;
  const float eps = S.isect_eps;
  // Topping a warp up before it has drained pays when ray costs inside a batch differ a lot (bounce waves, one sample
  // per pixel); it costs ~10 % when they do not: the strata of one pixel travel together and finish together
  const int refill_idle = ANY ? S.refill_any : S.refill_closest;
  const unsigned lane = lane_id();
  const unsigned tid = threadIdx.x;
  const bool hw7 = S.semantics >= B200RT_SEM_HW7;  // HW4 has the band triangle test (TRI_MODE 7) but no culling
  TravCounters tc = {0, 0, 0, 0};
  if (ANY && blockIdx.x == 0 && tid == 0) atomicAdd(&C->rays_shadow, (unsigned long long)count);

  WideLane L;
  wide_reset(L, hw7);
  L.cr = make_wide_ray(mk3(0, 0, 0), mk3(1, 1, 1));
  int sp = 0;
  bool has_ray = false, exhausted = false;
  uint32_t out_slot = 0, src = 0;
  float time = 0.f, tlimit = RT_INF;

  for (;;) {
    // ---- refill idle lanes from the wave's work cursor ----------------------
    const unsigned need = __ballot_sync(0xffffffffu, !has_ray);
    if (need != 0u && !exhausted) {
      uint32_t base = 0;
      const int leader = __ffs(need) - 1;
      if ((int)lane == leader) base = atomicAdd(cursor, (unsigned)__popc(need));
      base = __shfl_sync(0xffffffffu, base, leader);
      if (base + (uint32_t)__popc(need) >= count) exhausted = true;
      if (!has_ray) {
        const uint32_t k = base + __popc(need & ((1u << lane) - 1u));
        if (k < count) {
          float4 ro, rd;
          if (ANY) {
            src = k;
            ro = Q.o[k]; rd = Q.d[k];
            time = Q.time[k]; tlimit = ro.w;
            out_slot = __float_as_uint(rd.w);
          } else {
            out_slot = begin + k; src = out_slot;
            ro = P.ray_o[out_slot]; rd = P.ray_d[out_slot];
            time = ro.w; tlimit = RT_INF;
          }
          wide_reset(L, hw7);
          sp = 0; tqn = 0;
          has_ray = true;
          L.objmask = ANY ? Q.omask[k] : P.omask[out_slot];   // phase 0 was done by k_top_mask
        }
      }
    }
    if (!__any_sync(0xffffffffu, has_ray)) break;

    for (;;) {
      // ---- phase 1 (warp-synchronous): node steps, for as long as enough lanes hold an unvisited inner child.  A
      // lane with triangles to test, or with nothing left in its object, waits — but only until fewer than
      // S.wide_node_vote lanes are still descending: then everybody goes through phase 2.
      for (;;) {
        // a group without unvisited children: take the next one off the stack right here, a lane should not have to
        // wait for phase 2 just to pop
        if (has_ray && L.T.y == 0u && (L.G.y & 0xff000000u) == 0u && sp > 0) { sp--; L.G = s_stack[sp][tid]; }
        const bool want_node = has_ray && L.T.y == 0u && (L.G.y & 0xff000000u) != 0u;
        const unsigned mn = __ballot_sync(0xffffffffu, want_node);
        if (mn == 0u) break;
        if (__popc(mn) < S.wide_node_vote && __any_sync(0xffffffffu, has_ray && !want_node)) break;
        if (want_node) {
          const float bound = ANY ? tlimit : L.best.t;
          uint2 rest;
          if (wide_node_step(S.wide_nodes, L, bound + fabsf(bound) * 1e-4f, rest, tc)) {
            if (sp < RT_WSTACK) { s_stack[sp][tid] = rest; sp++; }
            else atomicExch(&C->overflow, 2u);
          }
#if RT_WIDE_TQ
          if ((ANY || RT_WIDE_TQ_CLOSEST) && L.T.y != 0u && tqn < (ANY ? RT_WIDE_TQ : RT_WIDE_TQ_CLOSEST)) { s_tq[tqn][tid] = L.T; tqn++; L.T.y = 0u; }
#endif
        }
      }
      // ---- phase 2: triangles; then, with nothing pending in the group, the stack or the next top-level object
      if (has_ray && (tqn > 0 || !(L.T.y == 0u && (L.G.y & 0xff000000u) != 0u))) {
        bool finished = false;
        if (L.T.y != 0u) finished = wide_triangles<ANY, TRI_MODE>(S, L, tlimit, eps, tc);
#if RT_WIDE_TQ
        while (!finished && tqn > 0) {
          tqn--;
          L.T = s_tq[tqn][tid];
          finished = wide_triangles<ANY, TRI_MODE>(S, L, tlimit, eps, tc);
        }
#endif
        if (!finished && (L.G.y & 0xff000000u) == 0u) {
          if (sp > 0) {
            sp--;
            L.G = s_stack[sp][tid];
          } else {
            const float4 ro = ANY ? Q.o[src] : P.ray_o[src];
            const float4 rd = ANY ? Q.d[src] : P.ray_d[src];
            const v3 o = mk3(ro.x, ro.y, ro.z), dir = mk3(rd.x, rd.y, rd.z);
            const RayCtx wr = make_ray(o, dir, eps);
            const float bound = ANY ? tlimit : L.best.t;
            const int k = wide_select(s_top, L.objmask, wr, eps, bound + fabsf(bound) * 1e-4f);
            if (k < 0) finished = true;
            else finished = wide_enter<ANY, TRI_MODE>(S, L, __float_as_uint(s_top[2 * k].w), o, dir, time, tlimit, eps, hw7, tc);
          }
        }
        if (finished) {
          if (ANY) {
            Q.vis[out_slot] = L.found ? 0 : 1;
          } else {
            P.hit_a[out_slot] = make_float4(L.found ? L.best.t : RT_INF, L.best.beta, L.best.gamma,
                                            __uint_as_float(L.found ? L.best.top_ref : B200RT_REF_NONE));
            P.hit_b[out_slot] = make_int2(L.found ? L.best.top_rank : -1, L.found ? L.best.tri : -1);
          }
          has_ray = false;
        }
      }
      const unsigned act = __ballot_sync(0xffffffffu, has_ray);
      if (act == 0u) break;
      if (!exhausted && __popc(~act) >= refill_idle) break;
    }
  }
  flush_counters(C, tc, ANY);
}
#endif  // __CUDACC__
