// rt_kernels.cuh — wavefront stages: generate / extend / shade_a / shadow /
// shade_b / combine (SURVEY.md §2c kernel table, §8 a1-a2, a9-a11, a15).
//
// The reference evaluates `trace_ray` recursively.  Here the recursion tree is
// evaluated level by level: wave d holds every path vertex at recursion depth
// d in a contiguous range of the vertex pool; shade_b appends the children to
// wave d+1; after the last wave a backward `combine` pass folds the children's
// colours into their parents in exactly the reference's operation order
//   color = local; color += km*mirror; color += k*(r*reflect + (1-r)*transmit)
// so the float result does not depend on the wavefront structure.
#pragma once
#include "rt_traverse.cuh"

#define RT_BLOCK 128

__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31u; }

// Warp-aggregated append: one atomicAdd per warp (ballot + popc), returns the
// element index for participating lanes.
__device__ __forceinline__ uint32_t warp_append(unsigned int* counter, bool want) {
  const unsigned mask = __ballot_sync(0xffffffffu, want);
  if (mask == 0) return 0;
  const int leader = __ffs(mask) - 1;
  uint32_t base = 0;
  if ((int)lane_id() == leader) base = atomicAdd(counter, (unsigned)__popc(mask));
  base = __shfl_sync(0xffffffffu, base, leader);
  return base + __popc(mask & ((1u << lane_id()) - 1u));
}

struct RenderParams {
  // rows rendered: groups of row_group consecutive rows starting at row_begin, row_begin + row_stride, ... (< row_end);
  // row_group 1 is the reference's own interleave `j += height_increase` (HW7/src/Scene.cpp:28)
  int row_begin, row_end, row_stride, row_group;
  int row_index0;   // index (among the rendered rows) of the first row of this batch
  int n_rows;
  int sample_begin, sample_count;      // strata [begin, begin+count) of samples_sqrt^2
  uint32_t seed_key;
  uint32_t n_paths;
};

// ---- camera: Camera::calculate_ray_at (HW1/Camera.h:24-32, HW3/include/Camera.h:35-47)
__device__ __forceinline__ void camera_ray(const DScene& S, float x, float y, bool dof, float ex, float ey, v3* o,
                                           v3* d) {
  const b200rt_camera& c = S.cam;
  const v3 e = ld3(c.e), u = ld3(c.u), v = ld3(c.v);
  if (S.semantics == B200RT_SEM_HW1) {
    const float su = (c.right - c.left) * (x + 0.5f) / (float)c.width;
    const float sv = (c.top - c.bottom) * (y + 0.5f) / (float)c.height;
    const v3 s = (ld3(c.top_left) + (u * su)) - (v * sv);
    *o = e; *d = s - e;
    return;
  }
  // HW4..HW7 callers pass the absolute sample position (HW7/include/Camera.h:66-84)
  const v3 s = (S.semantics >= B200RT_SEM_HW4)
                   ? (ld3(c.top_left) + (ld3(c.su) * x)) - (ld3(c.sv) * y)
                   : (ld3(c.top_left) + (ld3(c.su) * (x + 0.5f))) - (ld3(c.sv) * (y + 0.5f));
  if (dof) {
    const v3 ne = e + (((u * ex) + (v * ey)) * c.aperture) * 0.5f;
    *o = ne; *d = normalize(s - ne);
  } else {
    *o = e; *d = normalize(s - e);
  }
}

__global__ void k_generate(DScene S, Pool P, RenderParams R, WaveInfo* waves, Counters* C) {
  const uint32_t stride = gridDim.x * blockDim.x;
  const int W = S.cam.width;
  const int ns = S.cam.samples_sqrt;
  const bool tiled = (W % 8 == 0) && (R.n_rows % 4 == 0);
  for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < R.n_paths; k += stride) {
    const uint32_t pl = k / (uint32_t)R.sample_count;
    const uint32_t sidx = R.sample_begin + (k - pl * (uint32_t)R.sample_count);
    uint32_t rowi = pl / (uint32_t)W, coli = pl % (uint32_t)W;
    if (tiled) {
      // 8x4 pixel tiles: the 32 rays of a warp start out spatially coherent
      const uint32_t tile = pl >> 5, in = pl & 31u;
      const uint32_t tiles_per_row = (uint32_t)W >> 3;
      rowi = (tile / tiles_per_row) * 4u + (in >> 3);
      coli = (tile % tiles_per_row) * 8u + (in & 7u);
    }
    const int ri = R.row_index0 + (int)rowi;
    const int row = R.row_begin + (ri / R.row_group) * R.row_stride + (ri % R.row_group);
    const int col = (int)coli;
    const uint32_t pix = (uint32_t)row * (uint32_t)W + (uint32_t)col;
    const uint32_t key = rng_mix(rng_mix(R.seed_key, pix), ns == 1 ? 0u : sidx);
    v3 o, d;
    float time = 0.0f;
    const bool abs_pos = S.semantics >= B200RT_SEM_HW4;
    if (ns == 1) {
      if (abs_pos) camera_ray(S, col + 0.5f, row + 0.5f, false, 0.f, 0.f, &o, &d);
      else camera_ray(S, (float)col, (float)row, false, 0.f, 0.f, &o, &d);
    } else {
      // HW3/src/Scene.cpp:55-78 with the clock-seeded engines replaced by the shared stream
      const int x = (int)(sidx / (uint32_t)ns), y = (int)(sidx % (uint32_t)ns);
      const float ex = rng_u01(key, 4), ey = rng_u01(key, 5);
      const float sx = ((float)x + ex) / (float)ns, sy = ((float)y + ey) / (float)ns;
      time = rng_u01(key, 6);
      const float px = abs_pos ? ((float)col + sx) : (((float)col + sx) - 0.5f);
      const float py = abs_pos ? ((float)row + sy) : (((float)row + sy) - 0.5f);
      if (S.cam.aperture == 0.0f) {
        camera_ray(S, px, py, false, 0.f, 0.f, &o, &d);
      } else {
        const float dx = rng_u01(key, 7) * 2.0f - 1.0f, dy = rng_u01(key, 8) * 2.0f - 1.0f;
        camera_ray(S, px, py, true, dx, dy, &o, &d);
      }
    }
    P.ray_o[k] = make_float4(o.x, o.y, o.z, time);
    P.ray_d[k] = make_float4(d.x, d.y, d.z, __uint_as_float(RT_META_PRIMARY));
    P.key[k] = key;
    P.pixel[k] = pix;
    P.sample[k] = sidx;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    waves[0].begin = 0; waves[0].count = R.n_paths;
    atomicAdd(&C->rays_primary, (unsigned long long)R.n_paths);
  }
}

__global__ void k_next_wave(WaveInfo* waves, int d, Counters* C) {
  waves[d + 1].begin = waves[d].begin + waves[d].count;
  waves[d + 1].count = 0;
  C->work_extend = 0; C->work_shadow = 0; C->shadow_count = 0;
}

__device__ __forceinline__ void flush_counters(Counters* C, const TravCounters& c, bool shadow = false) {
  unsigned int n = c.nodes, t = c.tris, s = c.spheres, i = c.insts;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    n += __shfl_xor_sync(0xffffffffu, n, o); t += __shfl_xor_sync(0xffffffffu, t, o);
    s += __shfl_xor_sync(0xffffffffu, s, o); i += __shfl_xor_sync(0xffffffffu, i, o);
  }
  if (lane_id() == 0) {
    unsigned long long* base = shadow ? &C->sh_nodes_visited : &C->nodes_visited;
    atomicAdd(base + 0, (unsigned long long)n); atomicAdd(base + 1, (unsigned long long)t);
    atomicAdd(base + 2, (unsigned long long)s); atomicAdd(base + 3, (unsigned long long)i);
  }
}

#include "rt_fast.cuh"
#include "rt_wide.cuh"

// ---- extend (exact mode): closest hit, persistent warps pulling 32-ray batches
template <bool EXACT, int TRI_MODE>
__global__ void __launch_bounds__(RT_BLOCK) k_extend(DScene S, Pool P, const WaveInfo* waves, int d, Counters* C) {
  const uint32_t begin = waves[d].begin, count = waves[d].count;
  TravCounters tc = {0, 0, 0, 0};
  for (;;) {
    uint32_t base = 0;
    if (lane_id() == 0) base = atomicAdd(&C->work_extend, 32u);
    base = __shfl_sync(0xffffffffu, base, 0);
    if (base >= count) break;
    const uint32_t k = base + lane_id();
    if (k < count) {
      const uint32_t slot = begin + k;
      const float4 ro = P.ray_o[slot], rd = P.ray_d[slot];
      HitRec h;
      const bool f = traverse<false, EXACT, TRI_MODE>(S, mk3(ro.x, ro.y, ro.z), mk3(rd.x, rd.y, rd.z), ro.w, RT_INF,
                                                      h, tc);
      P.hit_a[slot] = make_float4(f ? h.t : RT_INF, h.beta, h.gamma, __uint_as_float(f ? h.top_ref : B200RT_REF_NONE));
      P.hit_b[slot] = make_int2(f ? h.top_rank : -1, f ? h.tri : -1);
    }
  }
  flush_counters(C, tc);
}

// ---- HW1: brute force in file order (HW1/Scene.cpp:33-55), triangles staged
// through shared memory in tiles shared by the whole block -------------------
#define HW1_TILE 256
__global__ void __launch_bounds__(RT_BLOCK) k_extend_hw1(DScene S, Pool P, const WaveInfo* waves, int d, Counters* C) {
  __shared__ float4 tile[HW1_TILE * 3];
  const uint32_t begin = waves[d].begin, count = waves[d].count;
  const float eps = S.isect_eps;
  TravCounters tc = {0, 0, 0, 0};
  for (uint32_t base = blockIdx.x * blockDim.x; base < count; base += gridDim.x * blockDim.x) {
    const uint32_t k = base + threadIdx.x;
    const bool act = k < count;
    const uint32_t slot = begin + (act ? k : 0);
    const float4 ro = P.ray_o[slot], rd = P.ray_d[slot];
    const v3 o = mk3(ro.x, ro.y, ro.z), dir = mk3(rd.x, rd.y, rd.z);
    float best = RT_INF, bb = 0.f, bg = 0.f;
    int best_tri = -1, best_rank = -1;
    uint32_t best_ref = B200RT_REF_NONE;
    for (int t0 = 0; t0 < S.n_tris; t0 += HW1_TILE) {
      const int nt = min(HW1_TILE, S.n_tris - t0);
      __syncthreads();
      for (int i = threadIdx.x; i < nt * 3; i += blockDim.x) tile[i] = __ldg(S.tris + 3 * (size_t)t0 + i);
      __syncthreads();
      if (act) {
        for (int i = 0; i < nt; i++) {
          const float4 a = tile[3 * i], b = tile[3 * i + 1], c = tile[3 * i + 2];
          float t, be, ga;
          tc.tris++;
          if (tri_test<1>(mk3(a.x, a.y, a.z), mk3(a.w, b.x, b.y), mk3(b.z, b.w, c.x), o, dir, eps, &t, &be, &ga) &&
              t < best - eps && t > 0.0f) {
            best = t; bb = be; bg = ga; best_tri = t0 + i; best_rank = t0 + i;
            best_ref = B200RT_REF(B200RT_KIND_TRI, t0 + i);
          }
        }
      }
    }
    if (act) {
      for (int i = 0; i < S.n_spheres; i++) {
        float t;
        tc.spheres++;
        if (sphere_test(S.spheres + i, o, dir, 0.f, eps, &t) && t < best - eps && t > 0.0f) {
          best = t; best_tri = -1; best_rank = S.n_tris + i; best_ref = B200RT_REF(B200RT_KIND_SPHERE, i);
        }
      }
      P.hit_a[slot] = make_float4(best, bb, bg, __uint_as_float(best_ref));
      P.hit_b[slot] = make_int2(best_rank, best_tri);
    }
  }
  flush_counters(C, tc);
}

// HW1 shadow predicate: any primitive with 0 < t < limit (HW1/Scene.cpp:72-101)
__global__ void __launch_bounds__(RT_BLOCK) k_shadow_hw1(DScene S, ShadowQueue Q, Counters* C) {
  __shared__ float4 tile[HW1_TILE * 3];
  const uint32_t count = C->shadow_count;
  const float eps = S.isect_eps;
  for (uint32_t base = blockIdx.x * blockDim.x; base < count; base += gridDim.x * blockDim.x) {
    const uint32_t k = base + threadIdx.x;
    const bool act = k < count;
    const float4 ro = Q.o[act ? k : 0], rd = Q.d[act ? k : 0];
    const v3 o = mk3(ro.x, ro.y, ro.z), dir = mk3(rd.x, rd.y, rd.z);
    const float limit = ro.w;
    bool occluded = false;
    for (int t0 = 0; t0 < S.n_tris; t0 += HW1_TILE) {
      const int nt = min(HW1_TILE, S.n_tris - t0);
      __syncthreads();
      for (int i = threadIdx.x; i < nt * 3; i += blockDim.x) tile[i] = __ldg(S.tris + 3 * (size_t)t0 + i);
      __syncthreads();
      if (act && !occluded) {
        for (int i = 0; i < nt; i++) {
          const float4 a = tile[3 * i], b = tile[3 * i + 1], c = tile[3 * i + 2];
          float t, be, ga;
          if (tri_test<1>(mk3(a.x, a.y, a.z), mk3(a.w, b.x, b.y), mk3(b.z, b.w, c.x), o, dir, eps, &t, &be, &ga) &&
              t < limit && t > 0.0f) { occluded = true; break; }
        }
      }
    }
    if (act) {
      for (int i = 0; i < S.n_spheres && !occluded; i++) {
        float t;
        if (sphere_test(S.spheres + i, o, dir, 0.f, eps, &t) && t < limit && t > 0.0f) occluded = true;
      }
      Q.vis[__float_as_uint(rd.w)] = occluded ? 0 : 1;
    }
  }
}

// ---- shadow: any-hit in (0, limit) — the reference runs a full closest-hit and
// tests `0 < t < dist - eps` (HW3/src/Scene.cpp:152-157); any-hit is the same predicate.
template <bool EXACT, int TRI_MODE>
__global__ void __launch_bounds__(RT_BLOCK) k_shadow(DScene S, ShadowQueue Q, Counters* C) {
  const uint32_t count = C->shadow_count;
  TravCounters tc = {0, 0, 0, 0};
  for (;;) {
    uint32_t base = 0;
    if (lane_id() == 0) base = atomicAdd(&C->work_shadow, 32u);
    base = __shfl_sync(0xffffffffu, base, 0);
    if (base >= count) break;
    const uint32_t k = base + lane_id();
    if (k < count) {
      const float4 ro = Q.o[k], rd = Q.d[k];
      HitRec h;
      const bool occ = traverse<true, EXACT, TRI_MODE>(S, mk3(ro.x, ro.y, ro.z), mk3(rd.x, rd.y, rd.z), Q.time[k],
                                                       ro.w, h, tc);
      Q.vis[__float_as_uint(rd.w)] = occ ? 0 : 1;
    }
  }
  flush_counters(C, tc, true);
}

// ---- surface reconstruction ---------------------------------------------------
__device__ __forceinline__ v3 hit_normal(const DScene& S, uint32_t top_ref, int tri, float beta, float gamma, v3 o,
                                         v3 d, float time, float t, int* material) {
  const uint32_t kind = B200RT_REF_KIND(top_ref), idx = B200RT_REF_INDEX(top_ref);
  if (kind == B200RT_KIND_INSTANCE || kind == B200RT_KIND_TRI) {
    const int2 info = S.tri_info[tri];
    v3 n;
    if (info.y & 1) {  // smooth — HW3/src/Mesh_triangle.cpp:101-106
      const int* vi = S.tri_vidx + 3 * (size_t)tri;
      const float4 n0 = S.vnormal[vi[0]], n1 = S.vnormal[vi[1]], n2 = S.vnormal[vi[2]];
      n = normalize(((mk3(n0.x, n0.y, n0.z) * ((1 - beta) - gamma)) + (mk3(n1.x, n1.y, n1.z) * beta)) +
                    (mk3(n2.x, n2.y, n2.z) * gamma));
    } else {
      const float4 c = S.tris[3 * (size_t)tri + 2];
      n = mk3(c.y, c.z, c.w);
    }
    if (kind == B200RT_KIND_TRI) { *material = info.x; return n; }
    const b200rt_instance* in = S.inst + idx;
    if (in->flags & B200RT_INST_DIRECT) { *material = info.x; return n; }
    *material = in->material;
    if (in->flags & B200RT_INST_MOVING) {
      float inv16[16];
      moving_inverse(in->fwd, in->velocity, time, inv16);
      v3 r;
      r.x = 0.0f; r.x += inv16[0] * n.x; r.x += inv16[4] * n.y; r.x += inv16[8] * n.z;
      r.y = 0.0f; r.y += inv16[1] * n.x; r.y += inv16[5] * n.y; r.y += inv16[9] * n.z;
      r.z = 0.0f; r.z += inv16[2] * n.x; r.z += inv16[6] * n.y; r.z += inv16[10] * n.z;
      return normalize(r);
    }
    return normalize(xf_vector3(in->nrm, n));
  }
  if (kind == B200RT_KIND_SPHERE) {
    const b200rt_sphere* s = S.spheres + idx;
    *material = s->material;
    return sphere_normal(s, o, d, time, t);
  }
  const b200rt_triprim* p = S.triprims + idx;
  *material = p->material;
  if (p->flags & B200RT_XF_IDENTITY) return ld3(p->normal);
  return normalize(xf_vector3(p->nrm, ld3(p->normal)));
}

__device__ __forceinline__ v3 light_position(const b200rt_light& L, uint32_t key, int area_index) {
  v3 lp = ld3(L.position);
  if (L.kind == B200RT_LIGHT_AREA) {
    // HW3/src/Scene.cpp:184-191 (clock-seeded engine -> shared stream, dims 16+2k, 17+2k)
    const float e1 = rng_u01(key, 16 + 2 * area_index), e2 = rng_u01(key, 17 + 2 * area_index);
    lp = (lp + (ld3(L.edge1) * e1)) + (ld3(L.edge2) * e2);
  }
  return lp;
}

// ---- shade_a: surface record + shadow-ray emission ----------------------------
#ifndef RT_SHADE3_MIN_BLOCKS
#define RT_SHADE3_MIN_BLOCKS 8  // resident CTAs the HW1-HW3 shade kernels are held to (64 registers): they wait on memory (profiles/r01l)
#endif
__global__ void __launch_bounds__(RT_BLOCK, RT_SHADE3_MIN_BLOCKS) k_shade_a(DScene S, Pool P, ShadowQueue Q, const WaveInfo* waves, int d,
                                                      Counters* C) {
  const uint32_t begin = waves[d].begin, count = waves[d].count;
  const uint32_t stride = gridDim.x * blockDim.x;
  const uint32_t rounds = (count + stride - 1) / stride;
  const bool hw1 = S.semantics == B200RT_SEM_HW1;
  for (uint32_t r = 0; r < rounds; r++) {
    const uint32_t k = r * stride + blockIdx.x * blockDim.x + threadIdx.x;
    const bool act = k < count;
    const uint32_t slot = begin + (act ? k : 0);
    const float4 ha = P.hit_a[slot];
    const uint32_t top_ref = __float_as_uint(ha.w);
    const bool hit = act && B200RT_REF_KIND(top_ref) != B200RT_KIND_NONE;
    v3 p = mk3(0, 0, 0);
    float time = 0.f;
    uint32_t key = 0;
    bool lights_on = false;
    bool has_d = false, has_s = false;
    if (hit) {
      const float4 ro = P.ray_o[slot], rd = P.ray_d[slot];
      const v3 o = mk3(ro.x, ro.y, ro.z), dir = mk3(rd.x, rd.y, rd.z);
      const uint32_t meta = __float_as_uint(rd.w);
      time = ro.w;
      int mat;
      const v3 n = hit_normal(S, top_ref, P.hit_b[slot].y, ha.y, ha.z, o, dir, time, ha.x, &mat);
      P.nrm[slot] = make_float4(n.x, n.y, n.z, __int_as_float(mat));
      p = o + (dir * ha.x);
      key = P.key[slot];
      const b200rt_material& m = S.mats[mat];
      has_d = !is_zero(ld3(m.diffuse)); has_s = !is_zero(ld3(m.specular));
      lights_on = hw1 || !(meta & RT_META_INMEDIUM);
      if (!hw1 && !has_d && !has_s) lights_on = false;  // HW3/src/Scene.cpp:141-145
    }
    int area_index = 0;
    for (int li = 0; li < S.n_lights; li++) {
      const b200rt_light& L = S.lights[li];
      bool want = lights_on;
      v3 wi = mk3(0, 0, 0);
      float ldist = 0.f;
      if (want) {
        const v3 lvec = light_position(L, key, area_index) - p;
        wi = normalize(lvec);
        ldist = length(lvec);
      }
      if (L.kind == B200RT_LIGHT_AREA) area_index++;
      const uint32_t qi = warp_append(&C->shadow_count, want);
      if (want) {
        if (qi < Q.capacity) {
          const v3 so = p + (wi * S.shadow_eps);
          const uint32_t vslot = slot * (uint32_t)S.n_lights + (uint32_t)li;
          Q.o[qi] = make_float4(so.x, so.y, so.z, ldist - S.shadow_eps);
          Q.d[qi] = make_float4(wi.x, wi.y, wi.z, __uint_as_float(vslot));
          Q.time[qi] = time;
          Q.vis[vslot] = 0;
        } else {
          C->overflow = 1;
        }
      }
    }
  }
}

__global__ void k_clamp_shadow_count(ShadowQueue Q, Counters* C) {
  atomicAdd(&C->rays_shadow, (unsigned long long)min(C->shadow_count, Q.capacity));
  if (C->shadow_count > Q.capacity) C->shadow_count = Q.capacity;
}

// Scene::refract_ray — HW3/src/Scene.cpp:100-114
__device__ __forceinline__ bool refract_ray(v3 du, v3 n, float index, v3* out) {
  const float n_ratio = 1 / index;
  const float cos_theta = dot(-du, n);
  const float delta = 1 - (n_ratio) * (n_ratio) * (1 - (cos_theta * cos_theta));
  if (delta < 0.0f) return false;
  *out = normalize(((du + (n * cos_theta)) * n_ratio) - (n * sqrtf(delta)));
  return true;
}

// ---- shade_b: Scene::trace_ray body (HW1/Scene.cpp:56-123, HW3/src/Scene.cpp:125-311)
__global__ void __launch_bounds__(RT_BLOCK, RT_SHADE3_MIN_BLOCKS) k_shade_b(DScene S, Pool P, ShadowQueue Q, WaveInfo* waves, int d,
                                                      Counters* C) {
  const uint32_t begin = waves[d].begin, count = waves[d].count;
  const uint32_t next_begin = waves[d + 1].begin;
  const uint32_t stride = gridDim.x * blockDim.x;
  const uint32_t rounds = (count + stride - 1) / stride;
  const bool hw1 = S.semantics == B200RT_SEM_HW1;
  for (uint32_t r = 0; r < rounds; r++) {
    const uint32_t k = r * stride + blockIdx.x * blockDim.x + threadIdx.x;
    const bool act = k < count;
    const uint32_t slot = begin + (act ? k : 0);
    const float4 ha = P.hit_a[slot];
    const float4 ro = P.ray_o[slot], rd = P.ray_d[slot];
    const uint32_t meta = __float_as_uint(rd.w);
    const uint32_t depth = RT_META_DEPTH(meta);
    const uint32_t top_ref = __float_as_uint(ha.w);
    const bool hit = act && B200RT_REF_KIND(top_ref) != B200RT_KIND_NONE;
    v3 color = mk3(0, 0, 0);
    v3 km = mk3(0, 0, 0), kk = mk3(0, 0, 0);
    float fres = 0.f;
    bool want_m = false, want_r = false, want_t = false;
    v3 dir_m = mk3(0, 0, 0), dir_r = mk3(0, 0, 0), dir_t = mk3(0, 0, 0);
    bool med_r = false, med_t = false;
    v3 p = mk3(0, 0, 0);
    const float time = ro.w;
    const uint32_t key = act ? P.key[slot] : 0u;
    if (act && !hit) {
      // "primary" is decided by the recursion budget (HW3/src/Scene.cpp:125-128)
      if (depth == 0) color = S.background;
    }
    if (hit) {
      const v3 o = mk3(ro.x, ro.y, ro.z), dir = mk3(rd.x, rd.y, rd.z);
      const float4 nn = P.nrm[slot];
      const v3 n = mk3(nn.x, nn.y, nn.z);
      const b200rt_material& m = S.mats[__float_as_int(nn.w)];
      p = o + (dir * ha.x);
      const v3 w0 = normalize(o - p);
      const v3 kd = ld3(m.diffuse), ks = ld3(m.specular);
      const bool has_d = !is_zero(kd), has_s = !is_zero(ks);
      if (hw1 || !(meta & RT_META_INMEDIUM)) {
        color = color + (ld3(m.ambient) * S.ambient);
        int area_index = 0;
        for (int li = 0; li < S.n_lights; li++) {
          const b200rt_light& L = S.lights[li];
          const int ai = area_index;
          if (L.kind == B200RT_LIGHT_AREA) area_index++;
          if (!hw1 && !has_d && !has_s) continue;
          if (!Q.vis[slot * (uint32_t)S.n_lights + (uint32_t)li]) continue;
          const v3 lvec = light_position(L, key, ai) - p;
          const v3 wi = normalize(lvec);
          const float ldist = length(lvec);
          const float d2 = ldist * ldist;
          const v3 inten = ld3(L.intensity);
          if (L.kind == B200RT_LIGHT_AREA) {
            const float icos = dot(-wi, ld3(L.normal));  // HW3/src/Scene.cpp:208
            if (has_d) {
              const float dc = dot(n, wi);
              color = color + ((((kd * inten) * icos) * dc) / d2);
            }
            if (has_s) {
              const float sc = fmaxf(0.0f, dot(n, normalize(w0 + wi)));
              color = color + ((((ks * inten) * icos) * (float)pow_ref7((double)sc, (double)m.phong_exponent)) / d2);
            }
          } else {
            if (hw1 || has_d) {
              const float dc = dot(n, wi);
              color = color + (((kd * inten) * dc) / d2);
            }
            if (hw1 || has_s) {
              const float sc = fmaxf(0.0f, dot(n, normalize(w0 + wi)));
              color = color + (((ks * inten) * (float)pow_ref7((double)sc, (double)m.phong_exponent)) / d2);
            }
          }
        }
      }
      const bool budget = (int)depth < S.max_depth;  // `current_recursion_depth > 0`
      km = ld3(m.mirror);
      if (!is_zero(km) && budget) {
        const v3 wr = normalize((n * (2 * dot(n, w0))) - w0);
        dir_m = wr;
        if (!hw1 && m.roughness != 0.0f) {  // glossy — HW3/src/Scene.cpp:236-260
          v3 rp;
          if (wr.x < wr.y && wr.x < wr.z) rp = mk3(1.0f, wr.y, wr.z);
          else if (wr.y < wr.z && wr.y < wr.x) rp = mk3(wr.x, 1.0f, wr.z);
          else rp = mk3(wr.x, wr.y, 1.0f);
          const v3 u = normalize(cross(rp, wr));
          const v3 v = normalize(cross(u, wr));
          const float eu = rng_u01(key, 2) - 0.5f, ev = rng_u01(key, 3) - 0.5f;
          dir_m = normalize(wr + (((u * eu) + (v * ev)) * m.roughness));
        }
        want_m = true;
      }
      const v3 kt = ld3(m.transparency);
      if (!hw1 && !is_zero(kt) && budget) {  // HW3/src/Scene.cpp:264-310
        const v3 wr = normalize((n * (2 * dot(n, w0))) - w0);
        const v3 dn = normalize(dir);
        const float ni = m.refraction_index;
        float cos_theta = 0.0f;
        bool tir = false, entering;
        if (dot(dn, n) < 0.0f) {
          refract_ray(dn, n, ni, &dir_t);
          cos_theta = dot(-dn, n);
          kk = mk3(1.0f, 1.0f, 1.0f);
          entering = true;
        } else {
          const double ht = (double)ha.x;
          const double sgn = (S.semantics == B200RT_SEM_HW2) ? -1.0 : 1.0;  // HW2/Scene.cpp:166-168
          kk = mk3((float)exp(sgn * log((double)kt.x) * ht), (float)exp(sgn * log((double)kt.y) * ht),
                   (float)exp(sgn * log((double)kt.z) * ht));
          entering = false;
          if (refract_ray(dn, -n, 1.0f / ni, &dir_t)) cos_theta = dot(dir_t, n);
          else tir = true;
        }
        dir_r = wr;
        want_r = true;
        if (tir) {
          med_r = true;
          fres = -1.0f;  // marks total internal reflection for combine
        } else {
          const float r0 = ((ni - 1) * (ni - 1)) / ((ni + 1) * (ni + 1));
          fres = (float)((double)r0 + (double)(1 - r0) * pow5d((double)(1.0f - cos_theta)));
          med_r = !entering;
          med_t = entering;
          want_t = true;
        }
      }
    }
    // children -> next wave
    uint32_t cm = 0xffffffffu, cr = 0xffffffffu, ct = 0xffffffffu;
#pragma unroll
    for (int c = 0; c < 3; c++) {
      const bool want = (c == 0) ? want_m : (c == 1 ? want_r : want_t);
      const uint32_t ci = warp_append(&waves[d + 1].count, want);
      if (want) {
        const uint32_t cs = next_begin + ci;
        if (cs < P.capacity) {
          const v3 cd = (c == 0) ? dir_m : (c == 1 ? dir_r : dir_t);
          const v3 co = p + (cd * S.shadow_eps);
          const bool med = (c == 0) ? false : (c == 1 ? med_r : med_t);
          const uint32_t cmeta = (depth + 1) | (med ? RT_META_INMEDIUM : 0u);
          P.ray_o[cs] = make_float4(co.x, co.y, co.z, time);
          P.ray_d[cs] = make_float4(cd.x, cd.y, cd.z, __uint_as_float(cmeta));
          P.key[cs] = rng_mix(key, 0x101u + (uint32_t)c);
          P.pixel[cs] = P.pixel[slot];
          if (c == 0) cm = cs; else if (c == 1) cr = cs; else ct = cs;
        } else {
          C->overflow = 1;
        }
      }
    }
    if (act) {
      P.rec0[slot] = make_float4(color.x, color.y, color.z, fres);
      P.rec1[slot] = make_float4(km.x, km.y, km.z, __uint_as_float(cm));
      P.rec2[slot] = make_float4(kk.x, kk.y, kk.z, __uint_as_float(cr));
      P.rec3[slot] = ct;
    }
  }
}

// Closes wave d+1 (clamped to the pool) and opens wave d+2 behind it: what k_next_wave(d+1) did as a launch of its own.
__global__ void k_finish_wave(WaveInfo* waves, int d, Pool P, Counters* C) {
  const uint32_t room = P.capacity > waves[d + 1].begin ? P.capacity - waves[d + 1].begin : 0;
  if (waves[d + 1].count > room) { waves[d + 1].count = room; C->overflow = 1; }
  atomicAdd(&C->rays_secondary, (unsigned long long)waves[d + 1].count);
  waves[d + 2].begin = waves[d + 1].begin + waves[d + 1].count;
  waves[d + 2].count = 0;
  C->work_extend = 0; C->work_shadow = 0; C->shadow_count = 0;
}

// gaussian_filter — HW3/src/Scene.cpp:15-17 (double exp, float divisor)
__device__ __forceinline__ float gaussian_filter(float x, float y, float sigma) {
  return (float)(exp((double)(-(x * x + y * y) / (2 * sigma * sigma))) /
                 (double)(float)(2 * 3.14159265358979323846 * (double)sigma));
}
// HW2/Scene.cpp:12-14: the same without the float cast of the divisor
__device__ __forceinline__ float gaussian_filter_hw2(float x, float y, float sigma) {
  return (float)(exp((double)(-(x * x + y * y) / (2 * sigma * sigma))) / (2 * 3.14159265358979323846 * (double)sigma));
}

// ---- combine: fold children into parents, deepest wave first -------------------
__global__ void k_combine(DScene S, Pool P, const WaveInfo* waves, int d, float4* accum) {
  const uint32_t begin = waves[d].begin, count = waves[d].count;
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < count; k += stride) {
    const uint32_t slot = begin + k;
    const float4 r0 = P.rec0[slot], r1 = P.rec1[slot], r2 = P.rec2[slot];
    const uint32_t cm = __float_as_uint(r1.w), cr = __float_as_uint(r2.w), ct = P.rec3[slot];
    v3 color = mk3(r0.x, r0.y, r0.z);
    if (cm != 0xffffffffu) {
      const float4 c = P.final_c[cm];
      color = color + (mk3(r1.x, r1.y, r1.z) * mk3(c.x, c.y, c.z));
    }
    if (cr != 0xffffffffu) {
      const float4 a = P.final_c[cr];
      const v3 kk = mk3(r2.x, r2.y, r2.z);
      if (ct != 0xffffffffu) {
        const float4 b = P.final_c[ct];
        const float r = r0.w;
        color = color + (kk * ((mk3(a.x, a.y, a.z) * r) + (mk3(b.x, b.y, b.z) * (1 - r))));
      } else {
        color = color + (kk * mk3(a.x, a.y, a.z));
      }
    }
    if (d > 0) {
      P.final_c[slot] = make_float4(color.x, color.y, color.z, 0.f);
    } else {
      const uint32_t pix = P.pixel[slot];
      const int ns = S.cam.samples_sqrt;
      if (ns == 1 && S.semantics == B200RT_SEM_HW4) {
        // HW4 splats the 3x3 Gaussian (sigma 1/2) even at one sample, with integer offsets (HW4/src/Scene.cpp:26-49)
        const int W = S.cam.width, H = S.cam.height;
        const int i = (int)(pix % (uint32_t)W), j = (int)(pix / (uint32_t)W);
        for (int aj = j - 1; aj < j + 2; aj++) {
          if (aj < 0 || aj >= H) continue;
          for (int ai = i - 1; ai < i + 2; ai++) {
            if (ai < 0 || ai >= W) continue;
            const float w = gaussian_filter((float)(i - ai), (float)(j - aj), 1.5f / 3.0f);
            // Pixel::add_color(color, w): one 16-byte vector atomic (sm_90+) instead of four scalar ones
            atomicAdd(accum + (size_t)aj * W + ai, make_float4(color.x * w, color.y * w, color.z * w, w));
          }
        }
      } else if (ns == 1) {
        // Pixel::add_color(color, 1.0f) on a fresh pixel — HW3/src/Scene.cpp:31-34
        accum[pix] = make_float4(color.x * 1.0f, color.y * 1.0f, color.z * 1.0f, 1.0f);
      } else {
        // 3x3 Gaussian splat sigma=1/3 — HW3/src/Scene.cpp:80-93
        const int W = S.cam.width, H = S.cam.height;
        const int i = (int)(pix % (uint32_t)W), j = (int)(pix / (uint32_t)W);
        const uint32_t sidx = P.sample[slot], key = P.key[slot];
        const int x = (int)(sidx / (uint32_t)ns), y = (int)(sidx % (uint32_t)ns);
        const float sx = ((float)x + rng_u01(key, 4)) / (float)ns, sy = ((float)y + rng_u01(key, 5)) / (float)ns;
        for (int aj = j - 1; aj < j + 2; aj++) {
          if (aj < 0 || aj >= H) continue;
          for (int ai = i - 1; ai < i + 2; ai++) {
            if (ai < 0 || ai >= W) continue;
            const float fx = ((float)i + sx) - ((float)ai + 0.5f);
            const float fy = ((float)j + sy) - ((float)aj + 0.5f);
            const float w = S.semantics == B200RT_SEM_HW2
                                ? gaussian_filter_hw2(fx, fy, 1.0f / 3.0f)
                                : gaussian_filter(fx, fy, S.semantics == B200RT_SEM_HW4 ? 1.5f / 3.0f : 1.0f / 3.0f);
            // Pixel::add_color(color, w): one 16-byte vector atomic (sm_90+) instead of four scalar ones
            atomicAdd(accum + (size_t)aj * W + ai, make_float4(color.x * w, color.y * w, color.z * w, w));
          }
        }
      }
    }
  }
}

__global__ void k_export_ids(Pool P, uint32_t n, int32_t* ids, float* t) {
  for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
    const float4 ha = P.hit_a[k];
    const int2 hb = P.hit_b[k];
    const uint32_t pix = P.pixel[k];
    ids[2 * (size_t)pix] = hb.x; ids[2 * (size_t)pix + 1] = hb.y;
    t[pix] = ha.x;
  }
}

__global__ void k_import_rays(Pool P, const float* rays, uint32_t n, WaveInfo* waves) {
  for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
    const float* r = rays + 8 * (size_t)k;
    P.ray_o[k] = make_float4(r[0], r[1], r[2], r[3]);
    P.ray_d[k] = make_float4(r[4], r[5], r[6], __uint_as_float(RT_META_PRIMARY));
    P.pixel[k] = k;
    P.key[k] = 0; P.sample[k] = 0;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) { waves[0].begin = 0; waves[0].count = n; }
}
