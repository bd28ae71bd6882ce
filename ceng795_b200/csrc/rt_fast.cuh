// rt_fast.cuh — the default traversal: persistent warps with per-lane ray
// replacement, walking the reference's (small) top-level tree with its exact
// gates and, inside a mesh, the SAH "fast" BVH (csrc/fast_bvh.h).
//
// Why the answer is still the reference's.  In the reference a
// mesh triangle k is reachable iff every BVH node on its root path passes
// Bounding_box::intersect (HW7/src/Bounding_volume_hierarchy.cpp:31-35).  All
// those boxes are nested (Bounding_box::expand is min/max, bounding_box.cpp:6-17)
// and the slab test is monotone under IEEE rounding — a superset box yields a
// superset [tmin,tmax] at every stage — so the whole chain passes iff the
// innermost box does.  That box (the reference node holding k) is stored per
// triangle in `tri_gate`; a triangle found through the SAH tree is admitted iff
// `box_gate(tri_gate[k])` passes, with the reference's own division-based
// arithmetic.  The SAH boxes themselves only need to be conservative (they are
// grown, and tested with slack), because every candidate is re-decided by the
// reference's triangle test + gate.  The winner is chosen by the key
// (t, top-level DFS rank, triangle DFS index), independent of visiting order.
#pragma once
#include "rt_traverse.cuh"

#define RT_FSTACK 64
#ifndef RT_REFILL_MIN_IDLE
// lanes idle before a warp takes new rays.  32 = whole-warp batches.  With the flat phase 0 (uniform cost, no
// divergence) topping a warp up once it is down to its last few rays pays: 26 (swept in profiles/r01g_tuning.txt)
#define RT_REFILL_MIN_IDLE 26
#endif
#ifndef RT_TOP_LINEAR_MAX
#define RT_TOP_LINEAR_MAX 40  // top-level objects up to which phase 0 scans a flat gate list instead of the tree
#endif
#ifndef RT_NODE_VOTE
#define RT_NODE_VOTE 4  // see phase 1 of k_trace_fast (swept in profiles/r01g_tuning.txt)
#endif
#ifndef RT_FAST_MIN_BLOCKS
#define RT_FAST_MIN_BLOCKS 8  // resident CTAs per SM the register budget is held to (64 regs; measured sweet spot, profiles/r01b)
#endif

// Ray in the space being traversed, with the constants of the two box tests.
struct FastRay {
  v3 o, d, inv, oi;  // oi = o * inv (for the fused loose test)
  bool slow;         // some |d_i| < eps: the reference skips that axis; take the generic path
};
__device__ __forceinline__ FastRay make_fast_ray(v3 o, v3 d, float eps) {
  FastRay r;
  r.o = o; r.d = d;
  r.slow = (fabsf(d.x) < eps) || (fabsf(d.y) < eps) || (fabsf(d.z) < eps);
  r.inv = mk3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
  r.oi = mk3(o.x * r.inv.x, o.y * r.inv.y, o.z * r.inv.z);
  return r;
}
__device__ __forceinline__ RayCtx as_ctx(const FastRay& r) {
  RayCtx c;
  c.o = r.o; c.d = r.d; c.inv = r.inv; c.slow = r.slow;
  return c;
}

// Conservative slab test for SAH boxes (grown on the host): may accept a box
// the exact test would reject, never the reverse by more than the growth.  Not
// an accept/reject decision of the reference, so FMA is allowed here.
__device__ __forceinline__ bool box_loose(const FastRay& r, float mnx, float mny, float mnz, float mxx, float mxy,
                                          float mxz, float eps, float* entry) {
  if (!r.slow) {
    const float ax = __fmaf_rn(mnx, r.inv.x, -r.oi.x), bx = __fmaf_rn(mxx, r.inv.x, -r.oi.x);
    const float ay = __fmaf_rn(mny, r.inv.y, -r.oi.y), by = __fmaf_rn(mxy, r.inv.y, -r.oi.y);
    const float az = __fmaf_rn(mnz, r.inv.z, -r.oi.z), bz = __fmaf_rn(mxz, r.inv.z, -r.oi.z);
    const float tn = fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fminf(az, bz));
    const float tf = fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fmaxf(az, bz));
    const float tfa = __fmaf_rn(2e-6f, fmaxf(fabsf(tn), fabsf(tf)), tf);
    *entry = fmaxf(tn, 0.0f);
    return tn <= tfa && tfa >= 0.0f;
  }
  // A direction component below the reference's axis-skip threshold.  The reference's OWN gates treat the
  // box as infinite on such an axis (and tri_gate reproduces that), but this tree is ours: a candidate still
  // has to pass through its triangle geometrically, so a true slab test on every axis stays conservative —
  // and keeps an axis-parallel ray from visiting the whole mesh (it does in the reference).
  float tn = -RT_INF, tf = RT_INF;
  const float o[3] = {r.o.x, r.o.y, r.o.z}, d[3] = {r.d.x, r.d.y, r.d.z};
  const float mn[3] = {mnx, mny, mnz}, mx[3] = {mxx, mxy, mxz};
#pragma unroll
  for (int i = 0; i < 3; i++) {
    if (d[i] == 0.0f) {
      if (o[i] < mn[i] || o[i] > mx[i]) return false;
    } else {
      const float a = (mn[i] - o[i]) / d[i], b = (mx[i] - o[i]) / d[i];
      tn = fmaxf(tn, fminf(a, b));
      tf = fminf(tf, fmaxf(a, b));
    }
  }
  const float tfa = tf + 2e-6f * fmaxf(fabsf(tn), fabsf(tf));
  *entry = fmaxf(tn, 0.0f);
  return !(tn > tfa) && !(tfa < 0.0f);
}

template <bool ANY, int TRI_MODE>
__global__ void __launch_bounds__(RT_BLOCK, RT_FAST_MIN_BLOCKS) k_trace_fast(DScene S, Pool P, ShadowQueue Q,
                                                                             const WaveInfo* waves, int d,
                                                                             Counters* C) {
  const uint32_t begin = ANY ? 0u : waves[d].begin;
  const uint32_t count = ANY ? C->shadow_count : waves[d].count;
  unsigned int* cursor = ANY ? &C->work_shadow : &C->work_extend;
  const float eps = S.isect_eps;
  // Topping a warp up before it has drained pays when ray costs inside a batch differ a lot (bounce waves, one
  // sample per pixel); it costs ~10 % when they do not: the strata of one pixel travel together, finish together,
  // and a partial refill then leaves the warp running 26 wide (cfg3, profiles/r01g_tuning.txt).
  const int refill_idle = (d == 0 && S.cam.samples_sqrt > 1) ? 32 : (ANY ? S.refill_any : S.refill_closest);
  const unsigned lane = lane_id();
  TravCounters tc = {0, 0, 0, 0};

  // stack[0, n_obj)  : the ray's gated top-level objects, nearest first — a queue consumed through obj_next
  // stack[n_obj, sp) : SAH nodes pending inside the object being traversed — a stack
  // (while the top-level tree is expanded, its pending nodes live at the far end, stack[tp, RT_FSTACK))
  uint32_t stack[RT_FSTACK];
  float stack_t[RT_FSTACK];
  int sp = 0, n_obj = 0, obj_next = 0;
  bool has_ray = false, exhausted = false;
  uint32_t out_slot = 0, src = 0;
  float time = 0.f, tlimit = RT_INF;
  FastRay cr;
  cr.o = mk3(0, 0, 0); cr.d = cr.o; cr.inv = cr.o; cr.oi = cr.o; cr.slow = false;
  HitRec best;
  best.t = RT_INF; best.beta = 0.f; best.gamma = 0.f; best.top_ref = B200RT_REF_NONE; best.top_rank = 0x7fffffff;
  best.tri = 0x7fffffff;
  bool found = false;
  int cur_rank = 0x7fffffff;
  uint32_t cur_top = B200RT_REF_NONE;
  uint32_t cur = B200RT_REF_NONE;  // node to process next without going through the stack
  float cur_t = 0.f;
  const bool hw7 = S.semantics >= B200RT_SEM_HW7;  // HW4 has the band triangle test (TRI_MODE 7) but no culling
  bool culling = hw7;              // HW7/PPM back-face culling, off inside refractive instances
#ifdef RT_DEBUG_RAY_COST
  unsigned int ray_nodes0 = 0, ray_tris0 = 0;
#endif

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
          best.t = RT_INF; best.beta = 0.f; best.gamma = 0.f; best.top_ref = B200RT_REF_NONE;
          best.top_rank = 0x7fffffff; best.tri = 0x7fffffff;
          found = false;
          cur_rank = 0x7fffffff; cur_top = B200RT_REF_NONE; culling = hw7;
          cur = B200RT_REF_NONE;
          has_ray = true;
#ifdef RT_DEBUG_RAY_COST
          ray_nodes0 = tc.nodes; ray_tris0 = tc.tris;
#endif
          // ---- phase 0: expand the reference's top-level tree with its exact gates (BVH::intersect,
          // HW3/src/Bounding_volume_hierarchy.cpp:31-55; Mesh_instance's own gate HW3/include/Mesh.h:59-62)
          // into the queue of objects this ray has to enter, nearest first
          const RayCtx wr = make_ray(mk3(ro.x, ro.y, ro.z), mk3(rd.x, rd.y, rd.z), eps);
          n_obj = 0; obj_next = 0;
          int tp = RT_FSTACK;
          if (S.n_top == 0) {
            const uint32_t rk = B200RT_REF_KIND(S.root);
            if (rk != B200RT_KIND_NONE) {
              float e = 0.0f;
              bool ok = true;
              if (rk == B200RT_KIND_INNER || rk == B200RT_KIND_INSTANCE)
                ok = box_gate(wr, S.root_box[0], S.root_box[1], S.root_box[2], S.root_box[3], S.root_box[4],
                              S.root_box[5], eps, &e);
              if (ok) { tp--; stack[tp] = S.root; stack_t[tp] = e; }
            }
          } else {
            // flat list: every lane tests the same object's innermost gate in lock step (uniform loads), then
            // inserts it into its queue by entry distance
            for (int k2 = 0; k2 < S.n_top; k2++) {
              const float4 b0 = __ldg(S.top_objs + 2 * k2), b1 = __ldg(S.top_objs + 2 * k2 + 1);
              float e = 0.0f;
              tc.nodes++;
              if (__float_as_uint(b1.w) != 0u && !box_gate(wr, b0.x, b0.y, b0.z, b1.x, b1.y, b1.z, eps, &e)) continue;
              int at = n_obj;
              while (at > 0 && stack_t[at - 1] > e) { stack[at] = stack[at - 1]; stack_t[at] = stack_t[at - 1]; at--; }
              stack[at] = __float_as_uint(b0.w); stack_t[at] = e;
              n_obj++;
            }
          }
          while (tp < RT_FSTACK) {
            const uint32_t ref = stack[tp];
            const float te = stack_t[tp];
            tp++;
            if (B200RT_REF_KIND(ref) != B200RT_KIND_INNER) {
              if (n_obj < tp) { stack[n_obj] = ref; stack_t[n_obj] = te; n_obj++; }
              else atomicExch(&C->overflow, 2u);
              continue;
            }
            tc.nodes++;
            const uint32_t idx = B200RT_REF_INDEX(ref);
            const float4 n0 = __ldg(S.nodes + 4 * (size_t)idx);
            const float4 n1 = __ldg(S.nodes + 4 * (size_t)idx + 1);
            const float4 n2 = __ldg(S.nodes + 4 * (size_t)idx + 2);
            const float4 n3 = __ldg(S.nodes + 4 * (size_t)idx + 3);
            const uint32_t child[2] = {__float_as_uint(n3.x), __float_as_uint(n3.y)};
            float entry[2] = {0.f, 0.f};
            bool push[2] = {false, false};
#pragma unroll
            for (int c = 0; c < 2; c++) {
              const uint32_t ck = B200RT_REF_KIND(child[c]);
              if (ck == B200RT_KIND_INNER || ck == B200RT_KIND_INSTANCE) {
                if (c == 0) push[0] = box_gate(wr, n0.x, n0.y, n0.z, n0.w, n1.x, n1.y, eps, &entry[0]);
                else push[1] = box_gate(wr, n1.z, n1.w, n2.x, n2.y, n2.z, n2.w, eps, &entry[1]);
              } else if (ck != B200RT_KIND_NONE) {
                push[c] = true;  // primitives and Light_mesh have no gate of their own
              }
            }
            const int far_c = (push[0] && push[1] && entry[0] > entry[1]) ? 0 : 1;
#pragma unroll
            for (int k2 = 0; k2 < 2; k2++) {
              const int c = k2 == 0 ? far_c : 1 - far_c;  // far first: the near child is popped next
              if (push[c]) {
                if (tp - 1 > n_obj) { tp--; stack[tp] = child[c]; stack_t[tp] = entry[c]; }
                else atomicExch(&C->overflow, 2u);
              }
            }
          }
          sp = n_obj;
        }
      }
    }
    if (!__any_sync(0xffffffffu, has_ray)) break;

    for (;;) {
      // ---- phase 1 (warp-synchronous): inner nodes of the SAH tree, one step per lane per iteration, for as long as
      // enough lanes hold one.  A lane that reaches a leaf, or runs out of nodes in its object, waits — but only
      // until fewer than S.node_vote lanes are still descending: then everybody goes through phase 2 (the few
      // descending lanes idle there instead).  node_vote = 1 is the plain while-while loop; incoherent waves
      // (shadow / bounce rays) ran it at 2-3 active lanes (profiles/r01g).
      for (;;) {
        bool want_node = false;
        if (has_ray) {
          for (;;) {
            if (cur == B200RT_REF_NONE) {
              if (sp <= n_obj) break;
              sp--; cur = stack[sp]; cur_t = stack_t[sp];
            }
            const float bound = ANY ? tlimit : best.t;
            if (cur_t > bound + fabsf(bound) * 1e-4f) { cur = B200RT_REF_NONE; continue; }  // pruned
            break;
          }
          want_node = cur != B200RT_REF_NONE && B200RT_REF_KIND(cur) == B200RT_KIND_FNODE;
        }
        const unsigned mn = __ballot_sync(0xffffffffu, want_node);
        if (mn == 0u) break;
        if (__popc(mn) < S.node_vote && __any_sync(0xffffffffu, has_ray && !want_node)) break;
        if (want_node) {
          const uint32_t idx = B200RT_REF_INDEX(cur);
          cur = B200RT_REF_NONE;
          tc.nodes++;
          const float4 n0 = __ldg(S.fast_nodes + 4 * (size_t)idx);
          const float4 n1 = __ldg(S.fast_nodes + 4 * (size_t)idx + 1);
          const float4 n2 = __ldg(S.fast_nodes + 4 * (size_t)idx + 2);
          const float4 n3 = __ldg(S.fast_nodes + 4 * (size_t)idx + 3);
          float e0, e1;
          const bool h0 = box_loose(cr, n0.x, n0.y, n0.z, n0.w, n1.x, n1.y, eps, &e0);
          const bool h1 = box_loose(cr, n1.z, n1.w, n2.x, n2.y, n2.z, n2.w, eps, &e1);
          const uint32_t c0 = __float_as_uint(n3.x), c1 = __float_as_uint(n3.y);
          if (h0 && h1) {
            const bool swap = e1 < e0;
            cur = swap ? c1 : c0; cur_t = swap ? e1 : e0;
            if (sp < RT_FSTACK) { stack[sp] = swap ? c0 : c1; stack_t[sp] = swap ? e0 : e1; sp++; }
            else atomicExch(&C->overflow, 2u);
          } else if (h0) { cur = c0; cur_t = e0; }
          else if (h1) { cur = c1; cur_t = e1; }
        }
      }
      if (has_ray && !(cur != B200RT_REF_NONE && B200RT_REF_KIND(cur) == B200RT_KIND_FNODE)) {
        bool finished = false;
        // ---- phase 2a: nothing pending in the current object -> enter the next one of the queue
        if (cur == B200RT_REF_NONE) {
          if (obj_next >= n_obj) {
            finished = true;
          } else {
            const uint32_t ref = stack[obj_next];
            const float te = stack_t[obj_next];
            obj_next++;
            const uint32_t kind = B200RT_REF_KIND(ref);
            const uint32_t idx = B200RT_REF_INDEX(ref);
            const float bound = ANY ? tlimit : best.t;
            if (te > bound + fabsf(bound) * 1e-4f) {
              // pruned: the object's gate box starts beyond the best hit
            } else if (kind == B200RT_KIND_INSTANCE || kind == B200RT_KIND_LIGHTINST) {
              // Mesh_instance::intersect — HW3/include/Mesh.h:55-98 (gate already passed; a Light_mesh has none)
              const b200rt_instance* in = S.inst + idx;
              const b200rt_mesh* m = S.meshes + in->mesh;
              const uint32_t froot = (m->fast_root != B200RT_REF_NONE) ? m->fast_root : m->root;
              if (B200RT_REF_KIND(froot) != B200RT_KIND_NONE) {
                tc.insts++;
                const float4 ro = ANY ? Q.o[src] : P.ray_o[src];
                const float4 rd = ANY ? Q.d[src] : P.ray_d[src];
                const v3 o = mk3(ro.x, ro.y, ro.z), dir = mk3(rd.x, rd.y, rd.z);
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
                cr = make_fast_ray(co, cd, eps);
                cur_rank = in->rank; cur_top = ref;
                culling = hw7 && !(in->flags & B200RT_INST_REFRACTIVE);
                cur = froot; cur_t = 0.f;
              }
            } else if (kind == B200RT_KIND_SPHERE) {
              tc.spheres++;
              const b200rt_sphere* s = S.spheres + idx;
              const float4 ro = ANY ? Q.o[src] : P.ray_o[src];
              const float4 rd = ANY ? Q.d[src] : P.ray_d[src];
              float t;
              // Sphere::intersect returns true even for t < 0 and only a BVH parent filters `t > 0`: a scene whose
              // root IS the sphere (create_bvh size==1) reports the negative hit
              const bool lone = !ANY && ref == S.root;
              if (sphere_test(s, mk3(ro.x, ro.y, ro.z), mk3(rd.x, rd.y, rd.z), time, eps, &t) && (t > 0.0f || lone)) {
                if (ANY) { if (t < tlimit) { found = true; finished = true; } }
                else if (better(t, s->rank, -1, best)) {
                  best.t = t; best.beta = 0.f; best.gamma = 0.f; best.top_ref = ref; best.top_rank = s->rank;
                  best.tri = -1; found = true;
                }
              }
            } else if (kind == B200RT_KIND_TRIPRIM) {
              tc.tris++;
              const b200rt_triprim* p = S.triprims + idx;
              const float4 ro = ANY ? Q.o[src] : P.ray_o[src];
              const float4 rd = ANY ? Q.d[src] : P.ray_d[src];
              v3 lo = mk3(ro.x, ro.y, ro.z), ldir = mk3(rd.x, rd.y, rd.z);
              if (!(p->flags & B200RT_XF_IDENTITY)) { const v3 wo = lo, wd = ldir; lo = xf_point(p->inv, wo); ldir = xf_vector4(p->inv, wd); }
              float t, be, ga;
              if (!(TRI_MODE == 7 && hw7 && dot(ldir, ld3(p->normal)) > 0.0f) &&   // top-level objects are culled (HW7)
                  tri_test<2>(ld3(p->v0), ld3(p->a1), ld3(p->a2), lo, ldir, eps, &t, &be, &ga) && t > 0.0f) {
                if (ANY) { if (t < tlimit) { found = true; finished = true; } }
                else if (better(t, p->rank, -1, best)) {
                  best.t = t; best.beta = be; best.gamma = ga; best.top_ref = ref; best.top_rank = p->rank;
                  best.tri = -1; found = true;
                }
              }
            }
          }
        }
        // ---- phase 2b: a leaf of the SAH tree (or the lone triangle of a one-triangle mesh)
        if (cur != B200RT_REF_NONE && B200RT_REF_KIND(cur) != B200RT_KIND_FNODE) {
          const uint32_t kind = B200RT_REF_KIND(cur);
          const uint32_t idx = B200RT_REF_INDEX(cur);
          cur = B200RT_REF_NONE;
          if (kind >= B200RT_KIND_LEAF1 && kind <= B200RT_KIND_LEAF4) {
            const int n = (int)(kind - B200RT_KIND_LEAF1) + 1;
            for (int i = 0; i < n; i++) {
              const size_t fi = (size_t)idx + i;
              tc.tris++;
              const float4 t0 = __ldg(S.fast_tris + 3 * fi);
              const float4 t1 = __ldg(S.fast_tris + 3 * fi + 1);
              const float4 t2 = __ldg(S.fast_tris + 3 * fi + 2);
              float t, be, ga;
              if (!(TRI_MODE == 7 && culling && dot(cr.d, mk3(t2.y, t2.z, t2.w)) > 0.0f) &&
                  tri_test<TRI_MODE>(mk3(t0.x, t0.y, t0.z), mk3(t0.w, t1.x, t1.y), mk3(t1.z, t1.w, t2.x), cr.o, cr.d,
                                     eps, &t, &be, &ga) && t > 0.0f && (ANY ? (t < tlimit) : (t <= best.t))) {
                const int tri = __ldg(S.fast_tri_ref + fi);
                if (ANY || better(t, cur_rank, tri, best)) {
                  // the reference's innermost gate for this triangle
                  const float2* g = reinterpret_cast<const float2*>(S.tri_gate + 6 * (size_t)tri);
                  const float2 g0 = __ldg(g), g1 = __ldg(g + 1), g2 = __ldg(g + 2);
                  float e;
                  if (box_gate(as_ctx(cr), g0.x, g0.y, g1.x, g1.y, g2.x, g2.y, eps, &e)) {
                    if (ANY) { found = true; finished = true; break; }
                    best.t = t; best.beta = be; best.gamma = ga; best.top_ref = cur_top; best.top_rank = cur_rank;
                    best.tri = tri; found = true;
                  }
                }
              }
            }
          } else if (kind == B200RT_KIND_TRI) {
            // a mesh whose BVH is a single triangle (create_bvh size==1): no gate at all
            tc.tris++;
            const float4 t0 = __ldg(S.tris + 3 * (size_t)idx);
            const float4 t1 = __ldg(S.tris + 3 * (size_t)idx + 1);
            const float4 t2 = __ldg(S.tris + 3 * (size_t)idx + 2);
            float t, be, ga;
            if (!(TRI_MODE == 7 && culling && dot(cr.d, mk3(t2.y, t2.z, t2.w)) > 0.0f) &&
                tri_test<TRI_MODE>(mk3(t0.x, t0.y, t0.z), mk3(t0.w, t1.x, t1.y), mk3(t1.z, t1.w, t2.x), cr.o, cr.d, eps,
                                   &t, &be, &ga) && t > 0.0f) {
              if (ANY) { if (t < tlimit) { found = true; finished = true; } }
              else if (better(t, cur_rank, (int)idx, best)) {
                best.t = t; best.beta = be; best.gamma = ga; best.top_ref = cur_top; best.top_rank = cur_rank;
                best.tri = (int)idx; found = true;
              }
            }
          }
        }
        if (finished) {
          if (ANY) {
            Q.vis[out_slot] = found ? 0 : 1;
          } else {
            P.hit_a[out_slot] = make_float4(found ? best.t : RT_INF, best.beta, best.gamma,
                                            __uint_as_float(found ? best.top_ref : B200RT_REF_NONE));
            P.hit_b[out_slot] = make_int2(found ? best.top_rank : -1, found ? best.tri : -1);
#ifdef RT_DEBUG_RAY_COST
            P.rec3[out_slot] = (tc.nodes - ray_nodes0) | ((tc.tris - ray_tris0) << 16);
#endif
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
