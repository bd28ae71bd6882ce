// rt_shade7.cuh — HW7 semantics of the shade stages: textures and Perlin noise
// (HW7/src/Texture.cpp, Perlin_noise.cpp), five BRDFs (HW7/src/*_BRDF.cpp), seven
// light kinds (HW7/src/*_light.cpp, Light_mesh.cpp, Light_sphere.cpp), Whitted
// ray tracing (Scene::trace_ray, HW7/src/Scene.cpp:415-447) and the branching
// path tracer (Scene::trace_path, :312-414) — SURVEY.md §8 rows a9-a14.
//
// Float/double discipline follows the reference's overload resolution: e.g.
// the unqualified `pow(cos, phong_exponent)` in Scene.cpp is the DOUBLE pow,
// `std::pow(float, float)` in the BRDF files is powf, `Vector3 / M_PI` divides
// by (float)M_PI (Vector3::operator/(float)).
#pragma once
#include "rt_kernels.cuh"

#define RT_PI 3.14159265358979323846
// The HW7 shade kernels stall on instruction fetch (profiles/r01i: no_instructions 45-59 % of stall samples, 90 KB of
// code walked by 16 warps per SM).  Warps of a block that walk the same stretch of code together share its cache lines,
// so the per-light loops are block-uniform and start with a barrier.
#define RT_SHADE_SYNC 1
#ifndef RT_SHADE7_BLOCK
#define RT_SHADE7_BLOCK 256  // threads per block of k_shade7_a/b: more warps behind each barrier
#endif


// Second measure against the same stall, and against the 15 of 32 lanes the two kernels keep busy (profiles/r01i): a
// block takes the wave in tiles of RT_SHADE7_BLOCK * RT_SHADE7_ROUNDS slots and walks each tile grouped by CLASS — by
// the kind of surface in shade_a (stand-alone triangle, sphere, mesh instance: the branches of surface7), by material
// in shade_b (BRDF model, mirror, glass) — so that the 32 vertices of a warp, and the warps of a block, mostly run the
// same branch.  Only the assignment of slots to threads changes: a vertex is still shaded from its own slot alone, so
// every value written is the one it was before; the children of a tile merely enter the next wave class by class.
// The grouping is a stable counting sort in shared memory (position = vertices of lower classes + vertices of the same
// class in earlier warp-rounds + same-class lanes below this one), deterministic from run to run.  0 rounds: off.
// Register budget of k_shade7_a/b: left to the compiler (128 / 127 registers, 2 resident blocks).  Holding it to 3 or 4
// blocks (-DRT_SHADE7_MIN_BLOCKS=3: 80 registers, 4: 64) was measured and is no faster (profiles/r02_tuning.txt 18).
#ifdef RT_SHADE7_MIN_BLOCKS
#define RT_SHADE7_BOUNDS __launch_bounds__(RT_SHADE7_BLOCK, RT_SHADE7_MIN_BLOCKS)
#else
#define RT_SHADE7_BOUNDS __launch_bounds__(RT_SHADE7_BLOCK)
#endif
#ifndef RT_SHADE7_ROUNDS
#define RT_SHADE7_ROUNDS 8   // swept on cfg4: 0 / 2 / 4 / 8 rounds = 842 / 759 / 731 / 720 ms per 64 strata (profiles/r02_tuning.txt)
#endif
#define RT_SHADE7_TILE (RT_SHADE7_BLOCK * (RT_SHADE7_ROUNDS > 0 ? RT_SHADE7_ROUNDS : 1))
#ifndef RT_SHADE7_CLASSES
#define RT_SHADE7_CLASSES 64   // a multiple of 32
#endif
#define RT_SHADE7_SEGS ((RT_SHADE7_BLOCK / 32) * (RT_SHADE7_ROUNDS > 0 ? RT_SHADE7_ROUNDS : 1))
struct TileOrder7 {
  unsigned short cnt[RT_SHADE7_CLASSES][RT_SHADE7_SEGS];   // vertices of a class in one warp-round, then their offset in the class
  unsigned short first[RT_SHADE7_CLASSES];                  // vertices of a class, then the first position of the class
  unsigned short perm[RT_SHADE7_TILE];                      // position in the walk -> vertex of the tile
};
// Fills T.perm[0 .. n_valid-1]; cls(i) = class of vertex i of the tile.  All threads of the block call it; position p of
// the walk is taken by thread p % BLOCK in round p / BLOCK (tile_vertex7), and the block meets at a barrier before it
// orders the next tile.
template <class ClassOf>
__device__ __forceinline__ void tile_order7(TileOrder7& T, uint32_t n_valid, ClassOf cls) {
#if RT_SHADE7_ROUNDS > 0
  const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  for (uint32_t x = tid; x < RT_SHADE7_CLASSES * RT_SHADE7_SEGS; x += RT_SHADE7_BLOCK) (&T.cnt[0][0])[x] = 0;
  __syncthreads();
  uint32_t c[RT_SHADE7_ROUNDS], rank[RT_SHADE7_ROUNDS];
#pragma unroll
  for (int j = 0; j < RT_SHADE7_ROUNDS; j++) {
    const uint32_t i = (uint32_t)j * RT_SHADE7_BLOCK + tid;
    c[j] = (i < n_valid) ? (cls(i) & (RT_SHADE7_CLASSES - 1)) : RT_SHADE7_CLASSES;   // beyond the wave: in no class
    const uint32_t same = __match_any_sync(0xffffffffu, c[j]);
    rank[j] = __popc(same & ((1u << lane) - 1u));
    if (rank[j] == 0 && c[j] < RT_SHADE7_CLASSES) T.cnt[c[j]][j * (RT_SHADE7_BLOCK / 32) + warp] = (unsigned short)__popc(same);
  }
  __syncthreads();
  if (tid < RT_SHADE7_CLASSES) {
    uint32_t run = 0;
    for (int sgm = 0; sgm < RT_SHADE7_SEGS; sgm++) { const uint32_t v = T.cnt[tid][sgm]; T.cnt[tid][sgm] = (unsigned short)run; run += v; }
    T.first[tid] = (unsigned short)run;
  }
  __syncthreads();
  if (warp == 0) {   // exclusive scan of the class sizes, CLASSES / 32 consecutive classes per lane
    constexpr int PER = RT_SHADE7_CLASSES / 32;
    uint32_t sz[PER], mine = 0;
#pragma unroll
    for (int q = 0; q < PER; q++) { sz[q] = T.first[PER * lane + q]; mine += sz[q]; }
    uint32_t incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o); if ((int)lane >= o) incl += v; }
    uint32_t at = incl - mine;
#pragma unroll
    for (int q = 0; q < PER; q++) { T.first[PER * lane + q] = (unsigned short)at; at += sz[q]; }
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < RT_SHADE7_ROUNDS; j++)
    if (c[j] < RT_SHADE7_CLASSES)
      T.perm[T.first[c[j]] + T.cnt[c[j]][j * (RT_SHADE7_BLOCK / 32) + warp] + rank[j]] = (unsigned short)((uint32_t)j * RT_SHADE7_BLOCK + tid);
  __syncthreads();
#endif
}
__device__ __forceinline__ uint32_t tile_vertex7(const TileOrder7& T, uint32_t position) {
#if RT_SHADE7_ROUNDS > 0
  return T.perm[position];
#else
  return position;
#endif
}
// Class of a closest hit for shade_a: kind of the top-level object first (the branch of surface7), then the object.
__device__ __forceinline__ uint32_t surf_class7(uint32_t top_ref) {
  const uint32_t kind = B200RT_REF_KIND(top_ref);
  if (kind == B200RT_KIND_NONE) return RT_SHADE7_CLASSES - 1;
  return (((kind & 7u) * (RT_SHADE7_CLASSES / 8)) | (B200RT_REF_INDEX(top_ref) & (RT_SHADE7_CLASSES / 8 - 1))) % (RT_SHADE7_CLASSES - 1);
}

__device__ __forceinline__ v3 operator/(v3 a, v3 b) { return mk3(a.x / b.x, a.y / b.y, a.z / b.z); }

// ---- textures ----------------------------------------------------------------
__device__ __forceinline__ float texel7(const DScene& S, const b200rt_texture& T, long long i) {
  const long long n = (long long)T.width * T.height * 3;
  if (i < 0 || i >= n) return 0.0f;  // the reference reads past the image here (UB); defined as 0
  return __ldg(S.texels + T.texel_offset + i);
}
// Texture::get_color_at — HW7/src/Texture.cpp:66-143
__device__ __noinline__ v3 tex_color7(const DScene& S, const b200rt_texture& T, float u, float v) {
  const int W = T.width, H = T.height;
  if (T.appearance == B200RT_TEX_CLAMP) {
    u = fmaxf(0.0f, fminf(1.0f, u));
    v = fmaxf(0.0f, fminf(1.0f, v));
  } else {
    u = fmaxf(0.0f, fminf(1.0f, u - floorf(u)));
    v = fmaxf(0.0f, fminf(1.0f, v - floorf(v)));
  }
  u *= W;
  if (u >= W) u--;
  v *= H;
  if (v >= H) v--;
  float c[3];
  if (T.interp == B200RT_TEX_NEAREST) {
    const long long coord = 3LL * W * (unsigned int)v + 3LL * (unsigned int)u;
    return mk3(texel7(S, T, coord), texel7(S, T, coord + 1), texel7(S, T, coord + 2));
  }
  if (T.appearance == B200RT_TEX_REPEAT) {
    const unsigned int p = ((unsigned int)u) % (unsigned)W, pn = (p + 1) % (unsigned)W;
    const unsigned int q = ((unsigned int)v) % (unsigned)H, qn = (q + 1) % (unsigned)H;
    const float dx = u - p, dy = v - q;
    const long long c_p_q = 3LL * (W * q + p), c_pn_q = 3LL * (W * q + pn), c_p_qn = 3LL * (W * qn + p),
                    c_pn_qn = 3LL * (W * qn + pn);
#pragma unroll
    for (int k = 0; k < 3; k++) {
      float x = texel7(S, T, c_p_q + k) * (1 - dx) * (1 - dy);
      x += texel7(S, T, c_pn_q + k) * (dx) * (1 - dy);
      x += texel7(S, T, c_pn_qn + k) * (dx) * (dy);
      x += texel7(S, T, c_p_qn + k) * (1 - dx) * (dy);
      c[k] = x;
    }
    return mk3(c[0], c[1], c[2]);
  }
  const unsigned int p = (unsigned int)u, pn = p + 1;
  const unsigned int q = (unsigned int)v, qn = (unsigned int)(v + 1);
  const float dx = u - p, dy = v - q;
  const long long c_p_q = 3LL * ((long long)W * q + p), c_pn_q = 3LL * ((long long)W * q + pn),
                  c_p_qn = 3LL * ((long long)W * qn + p), c_pn_qn = 3LL * ((long long)W * qn + pn);
#pragma unroll
  for (int k = 0; k < 3; k++) {
    float x = texel7(S, T, c_p_q + k) * (1 - dx) * (1 - dy);
    if (pn < (unsigned)W) x += texel7(S, T, c_pn_q + k) * (dx) * (1 - dy);
    if (qn < (unsigned)H) x += texel7(S, T, c_pn_qn + k) * (dx) * (dy);  // crossed guards
    if (pn < (unsigned)W && qn < (unsigned)H) x += texel7(S, T, c_p_qn + k) * (1 - dx) * (dy);
    c[k] = x;
  }
  return mk3(c[0], c[1], c[2]);
}
// Texture::blend_color — HW7/src/Texture.cpp:53-65
__device__ __forceinline__ v3 tex_blend7(const b200rt_texture& T, v3 tc, v3 kd) {
  if (T.decal == B200RT_TEX_REPLACE_KD) return tc;
  if (T.decal == B200RT_TEX_BLEND_KD) return (tc + kd) / 2.0f;
  return mk3(0, 0, 0);
}
// Perlin_noise — HW7/src/Perlin_noise.cpp:23-62
__device__ __forceinline__ float perlin_weight7(float x) {
  const double a = fabs((double)x);
  return (float)(-6 * pow5d(a) + 15 * pow4d(a) + -10 * pow3d(a) + 1);
}
__device__ __noinline__ float perlin_value7(const b200rt_texture& T, v3 pos) {
  const float G[16][3] = {{1, 1, 0}, {-1, 1, 0}, {1, -1, 0}, {-1, -1, 0}, {1, 0, 1}, {-1, 0, 1}, {1, 0, -1}, {-1, 0, -1},
                          {0, 1, 1}, {0, -1, 1}, {0, 1, -1}, {0, -1, -1}, {1, 1, 0}, {-1, 1, 0}, {0, -1, 1}, {0, -1, -1}};
  const v3 sp = pos * T.scaling;
  float value = 0.0f;
  const int x = (int)floorf(sp.x), y = (int)floorf(sp.y), z = (int)floorf(sp.z);
  for (int i = x; i <= x + 1; i++)
    for (int j = y; j <= y + 1; j++)
      for (int k = z; k <= z + 1; k++) {
        const v3 vv = mk3(sp.x - i, sp.y - j, sp.z - k);
        int a = i % 16, b = j % 16, c = k % 16;
        if (a < 0) a += 16;
        if (b < 0) b += 16;
        if (c < 0) c += 16;
        const float* g = G[T.perm[(a + T.perm[(b + T.perm[c % 16]) % 16]) % 16]];
        const v3 e = mk3(g[0], g[1], g[2]);
        const float wx = perlin_weight7(vv.x), wy = perlin_weight7(vv.y), wz = perlin_weight7(vv.z);
        value += dot(e, vv) * wx * wy * wz;
      }
  if (T.appearance == B200RT_TEX_VEIN) return fabsf(value);
  return (value + 1.0f) / 2.0f;
}
// Texture::bump_normal (Perlin) — HW7/src/Texture.cpp:145-165
__device__ __forceinline__ v3 bump_perlin7(const b200rt_texture& T, v3 normal, v3 pos) {
  const float pv = perlin_value7(T, pos);
  const float epsilon = 0.001f;
  const float dx = (perlin_value7(T, pos + mk3(epsilon, 0, 0)) - pv) / epsilon;
  const float dy = (perlin_value7(T, pos + mk3(0, epsilon, 0)) - pv) / epsilon;
  const float dz = (perlin_value7(T, pos + mk3(0, 0, epsilon)) - pv) / epsilon;
  const v3 g = mk3(dx, dy, dz) * T.bump_multiplier;
  const v3 g_ii = normal * dot(g, normal);
  return normal - (g - g_ii);
}
// Texture::bump_normal (image) — HW7/src/Texture.cpp:167-181
__device__ __forceinline__ v3 bump_image7(const DScene& S, const b200rt_texture& T, v3 dp_du, v3 dp_dv, v3 normal,
                                          float u, float v) {
  const v3 c0 = tex_color7(S, T, u, v);
  const v3 cu = tex_color7(S, T, u + 1.0f / T.width, v);
  const v3 cv = tex_color7(S, T, u, v + 1.0f / T.height);
  const float v0 = (c0.x + c0.y + c0.z) / 3.0f, vu = (cu.x + cu.y + cu.z) / 3.0f, vv = (cv.x + cv.y + cv.z) / 3.0f;
  const float dd_du = (vu - v0) * T.bump_multiplier, dd_dv = (vv - v0) * T.bump_multiplier;
  const v3 dq_du = dp_du + (normal * dd_du);
  const v3 dq_dv = dp_dv + (normal * dd_dv);
  return normalize(cross(dq_dv, dq_du));
}

// ---- surface of the closest hit (the texture / normal part of the intersect functions) ----
struct Surf7 {
  v3 n;
  float u, v, perlin;
  int material, texture, light;
};
__device__ __noinline__ Surf7 surface7(const DScene& S, uint32_t top_ref, int tri, float beta, float gamma, v3 o, v3 d,
                                       float time, float t) {
  Surf7 sf;
  sf.u = -1; sf.v = -1; sf.perlin = -1; sf.texture = -1; sf.light = -1; sf.material = 0;
  sf.n = mk3(0, 0, 0);
  const uint32_t kind = B200RT_REF_KIND(top_ref), idx = B200RT_REF_INDEX(top_ref);
  if (kind == B200RT_KIND_INSTANCE || kind == B200RT_KIND_LIGHTINST) {
    // Mesh_triangle::intersect — HW7/src/Mesh_triangle.cpp:89-157, then Mesh_instance::intersect HW7/include/Mesh.h:83-114
    const b200rt_instance* in = S.inst + idx;
    float inv16[16];
    const bool moving = (in->flags & B200RT_INST_MOVING) != 0;
    if (moving) moving_inverse(in->fwd, in->velocity, time, inv16);
    const int2 info = S.tri_info[tri];
    v3 n;
    if (info.y & 1) {
      const int* vi = S.tri_vidx + 3 * (size_t)tri;
      const float4 n0 = S.vnormal[vi[0]], n1 = S.vnormal[vi[1]], n2 = S.vnormal[vi[2]];
      n = normalize(((mk3(n0.x, n0.y, n0.z) * ((1 - beta) - gamma)) + (mk3(n1.x, n1.y, n1.z) * beta)) +
                    (mk3(n2.x, n2.y, n2.z) * gamma));
    } else {
      const float4 c = S.tris[3 * (size_t)tri + 2];
      n = mk3(c.y, c.z, c.w);
    }
    const int tex = S.tri_tex[tri];
    if (tex != -1) {
      const b200rt_texture& T = S.textures[tex];
      if (T.kind == B200RT_TEX_PERLIN) {
        const v3 lo = moving ? xf_point(inv16, o) : xf_point(in->inv, o);
        const v3 ldir = moving ? xf_vector4(inv16, d) : xf_vector4(in->inv, d);
        const v3 lp = lo + (ldir * t);
        sf.perlin = perlin_value7(T, lp);
        if (T.is_bump) n = bump_perlin7(T, n, lp);
      } else {
        const float* uv = S.tri_uv + 6 * (size_t)tri;
        const float u = uv[0] + beta * (uv[2] - uv[0]) + gamma * (uv[4] - uv[0]);
        const float v = uv[1] + beta * (uv[3] - uv[1]) + gamma * (uv[5] - uv[1]);
        sf.u = u; sf.v = v;
        if (T.is_bump) {
          const float ub_ua = uv[2] - uv[0], uc_ua = uv[4] - uv[0], vb_va = uv[3] - uv[1], vc_va = uv[5] - uv[1];
          const int* vi = S.tri_vidx + 3 * (size_t)tri;
          const v3 p0 = ld3(S.vert_pos + 3 * (size_t)vi[0]), p1 = ld3(S.vert_pos + 3 * (size_t)vi[1]),
                   p2 = ld3(S.vert_pos + 3 * (size_t)vi[2]);
          const v3 pb_pa = p1 - p0, pc_pa = p2 - p0;
          float ic = (ub_ua * vc_va) - (vb_va * uc_ua);
          if (ic == 0.0f) ic = 0.000001f;
          ic = 1.0f / ic;
          const v3 dp_du = ((pb_pa * vc_va) - (pc_pa * vb_va)) * ic;
          const v3 dp_dv = ((pc_pa * ub_ua) - (pb_pa * uc_ua)) * ic;
          n = bump_image7(S, T, dp_du, dp_dv, n, u, v);
        }
      }
    }
    if (moving) {
      v3 r;
      r.x = 0.0f; r.x += inv16[0] * n.x; r.x += inv16[4] * n.y; r.x += inv16[8] * n.z;
      r.y = 0.0f; r.y += inv16[1] * n.x; r.y += inv16[5] * n.y; r.y += inv16[9] * n.z;
      r.z = 0.0f; r.z += inv16[2] * n.x; r.z += inv16[6] * n.y; r.z += inv16[10] * n.z;
      sf.n = normalize(r);
    } else {
      sf.n = normalize(xf_vector3(in->nrm, n));
    }
    sf.material = in->material;
    sf.texture = in->texture;
    sf.light = (in->flags & B200RT_INST_LIGHT) ? in->light : -1;
    return sf;
  }
  if (kind == B200RT_KIND_SPHERE) {
    // Sphere::intersect — HW7/src/Sphere.cpp:32-103
    const b200rt_sphere* s = S.spheres + idx;
    const v3 c = ld3(s->center);
    float inv16[16];
    const bool moving = (s->flags & B200RT_XF_MOVING) != 0;
    v3 lo = o, ldir = d;
    if (moving) {
      moving_inverse(s->fwd, s->velocity, time, inv16);
      lo = xf_point(inv16, o); ldir = xf_vector4(inv16, d);
    } else if (!(s->flags & B200RT_XF_IDENTITY)) {
      lo = xf_point(s->inv, o); ldir = xf_vector4(s->inv, d);
    }
    const v3 lp = lo + (ldir * t);
    const v3 lc = lp - c;
    v3 n = normalize(lc);
    if (s->texture != -1) {
      sf.texture = s->texture;
      const b200rt_texture& T = S.textures[s->texture];
      // HW7/src/Sphere.cpp:107-108 clamps and calls std::acos; HW4-HW6 call the unqualified (double) acos on the raw
      // ratio (HW4/src/Sphere.cpp:101, HW6/src/Sphere.cpp:106)
      const float theta = (S.semantics >= B200RT_SEM_HW7) ? acosf(fmaxf(-1.0f, fminf(1.0f, lc.y / s->radius)))
                                                          : (float)acos((double)(lc.y / s->radius));
      const float phi = (float)atan2((double)lc.z, (double)lc.x);  // unqualified atan2: the double function
      const float u = (float)((RT_PI - (double)phi) / (2 * RT_PI));
      const float v = (float)((double)theta / RT_PI);
      sf.u = u; sf.v = v;
      if (T.kind == B200RT_TEX_PERLIN) {
        sf.perlin = perlin_value7(T, lp);
        if (T.is_bump) n = bump_perlin7(T, n, lp);
      } else if (T.is_bump) {
        const float th = (float)(RT_PI * (double)v);
        const float ph = (float)(RT_PI * (double)(1 - 2 * u));
        const v3 dp_du = mk3((float)((double)(lc.z * 2) * RT_PI), 0, (float)((double)(lc.x * -2) * RT_PI));
        const v3 dp_dv = mk3((float)((double)lc.y * RT_PI * cos((double)ph)),
                             (float)((double)(-s->radius) * RT_PI * sin((double)th)),
                             (float)((double)lc.y * RT_PI * sin((double)ph)));
        n = bump_image7(S, T, dp_du, dp_dv, n, u, v);
      }
    }
    if (moving) {
      v3 r;
      r.x = 0.0f; r.x += inv16[0] * n.x; r.x += inv16[4] * n.y; r.x += inv16[8] * n.z;
      r.y = 0.0f; r.y += inv16[1] * n.x; r.y += inv16[5] * n.y; r.y += inv16[9] * n.z;
      r.z = 0.0f; r.z += inv16[2] * n.x; r.z += inv16[6] * n.y; r.z += inv16[10] * n.z;
      sf.n = normalize(r);
    } else {
      sf.n = normalize(xf_vector3(s->nrm, n));  // always applied: a second normalise even for identity
    }
    sf.material = s->material;
    sf.light = (s->flags & B200RT_XF_LIGHT) ? s->light : -1;
    return sf;
  }
  if (kind == B200RT_KIND_TRIPRIM) {
    const b200rt_triprim* p = S.triprims + idx;
    sf.material = p->material;
    sf.u = 0; sf.v = 0; sf.perlin = 0;
    sf.n = (p->flags & B200RT_XF_IDENTITY) ? ld3(p->normal) : normalize(xf_vector3(p->nrm, ld3(p->normal)));
    return sf;
  }
  return sf;
}

// Scene::calculate_diffuse_constant — HW7/src/Scene.cpp:231-258
__device__ __forceinline__ bool diffuse_constant7(const DScene& S, const Surf7& sf, v3* out) {
  const b200rt_material& m = S.mats[sf.material];
  *out = ld3(m.diffuse);
  if (sf.texture < 0) return false;
  const b200rt_texture& T = S.textures[sf.texture];
  if (T.decal == B200RT_TEX_REPLACE_ALL) {
    *out = (T.kind == B200RT_TEX_PERLIN) ? mk3(0, 0, 0) : tex_color7(S, T, sf.u, sf.v);
    return true;
  }
  if (T.kind == B200RT_TEX_PERLIN) *out = tex_blend7(T, mk3(sf.perlin, sf.perlin, sf.perlin), *out);
  else *out = tex_blend7(T, tex_color7(S, T, sf.u, sf.v) / T.normalizer, *out);
  return false;
}

// BRDF::get_reflectance x5 — HW7/src/*_BRDF.cpp
__device__ __noinline__ v3 brdf_eval7(const b200rt_brdf& B, v3 normal, v3 kd, v3 ks, v3 w_i, v3 w_o) {
  const float cti = dot(normal, w_i);
  if (cti > 1.0f || cti <= 0.0f) return mk3(0, 0, 0);
  const float e = B.exponent;
  switch (B.kind) {
    case B200RT_BRDF_PHONG: {
      const v3 rwi = (normal * (2 * cti)) - w_i;
      const float car = fmaxf(0.0f, dot(w_o, rwi));
      return ((ks * powf(car, e)) / cti) + kd;
    }
    case B200RT_BRDF_MODIFIED_PHONG: {
      const v3 rwi = (normal * (2 * cti)) - w_i;
      const float car = fmaxf(0.0f, dot(w_o, rwi));
      v3 k_d = kd, k_s = ks * powf(car, e);
      if (B.normalized) {
        k_s = (k_s * (e + 2.0f)) / (float)(2.0f * RT_PI);
        k_d = k_d / (float)RT_PI;
      }
      return k_s + k_d;
    }
    case B200RT_BRDF_BLINN_PHONG: {
      const v3 wh = normalize(w_i + w_o);
      const float cah = dot(wh, normal);
      return ((ks * powf(cah, e)) / cti) + kd;
    }
    case B200RT_BRDF_MODIFIED_BLINN_PHONG: {
      const v3 wh = normalize(w_i + w_o);
      const float cah = dot(wh, normal);
      v3 k_d = kd, k_s = ks * powf(cah, e);
      if (B.normalized) {
        k_d = k_d / (float)RT_PI;
        k_s = (k_s * (e + 8.0f)) / (float)(8.0f * RT_PI);
      }
      return k_s + k_d;
    }
    default: {  // Torrance-Sparrow
      const v3 wh = normalize(w_i + w_o);
      const float ca = dot(wh, normal);
      const float d_alpha = (float)((double)((e + 2) * powf(ca, e)) / (2.0f * RT_PI));
      const float cto = dot(w_o, normal);
      const float cb = dot(w_o, wh);
      float g = fminf(2 * ca * cto / cb, 2 * ca * cti / cb);
      g = fminf(1.0f, g);
      const float ri = B.refractive_index;
      const float r0 = (float)(pow2d((double)(ri - 1.0f)) / pow2d((double)(ri + 1.0f)));
      const float f_beta = (float)((double)r0 + (double)(1 - r0) * pow5d((double)(1 - cb)));
      const v3 k_d = kd / (float)RT_PI;
      const v3 k_s = (((ks * d_alpha) * f_beta) * g) / (4 * cti * cto);
      return k_d + k_s;
    }
  }
}

// w*cos(theta) + v*sin(theta)*cos(phi) + u*sin(theta)*sin(phi), normalised (trace_path, env light, sphere light)
__device__ __noinline__ v3 onb_dir7(v3 w, float theta, float phi) {
  const v3 u = normalize((w.x != 0.0f || w.y != 0.0f) ? mk3(-w.y, w.x, 0.0f) : mk3(0.0f, 1.0f, 0.0f));
  const v3 v = cross(w, u);
  return normalize(((w * cosf(theta)) + ((v * sinf(theta)) * cosf(phi))) + ((u * sinf(theta)) * sinf(phi)));
}

// Light::direction_and_distance.  RNG: dims 32+4*li+{0,1,2} of the vertex key (shared with oracle.c).
__device__ __noinline__ v3 light_dir7(const DScene& S, int li, v3 p, v3 normal, uint32_t key, float* distance,
                                      float* prob) {
  const b200rt_light& L = S.lights[li];
  const uint32_t dim = 32u + 4u * (uint32_t)li;
  *prob = 0.0f;
  switch (L.kind) {
    case B200RT_LIGHT_AREA: {
      const float e1 = rng_u01(key, dim), e2 = rng_u01(key, dim + 1);
      const v3 pos = (ld3(L.position) + (ld3(L.edge1) * e1)) + (ld3(L.edge2) * e2);
      const v3 dir = pos - p;
      *distance = length(dir);
      return dir;
    }
    case B200RT_LIGHT_DIRECTIONAL:
      *distance = 3.402823466e+38f;
      return -ld3(L.edge1);
    case B200RT_LIGHT_ENV: {
      const float e1 = rng_u01(key, dim), e2 = rng_u01(key, dim + 1);
      const float phi = (float)(2 * RT_PI * (double)e1);
      const float theta = acosf(e2);
      *distance = 3.402823466e+38f;
      *prob = (float)(1.0f / (2.0f * RT_PI));
      return onb_dir7(normal, theta, phi);
    }
    case B200RT_LIGHT_MESH: {
      const float e0 = rng_u01(key, dim), e1 = sqrtf(rng_u01(key, dim + 1)), e2 = rng_u01(key, dim + 2);
      int tri = -1;
      for (int i = 0; i < L.cdf_count; i++)
        if (e0 <= S.light_cdf[L.cdf_begin + i]) { tri = S.light_tri[L.cdf_begin + i]; break; }
      if (tri < 0) tri = S.light_tri[L.cdf_begin + L.cdf_count - 1];
      const int* vi = S.tri_vidx + 3 * (size_t)tri;
      const v3 p0 = ld3(S.vert_pos + 3 * (size_t)vi[0]), p1 = ld3(S.vert_pos + 3 * (size_t)vi[1]),
               p2 = ld3(S.vert_pos + 3 * (size_t)vi[2]);
      const v3 q = (p1 * (1.0f - e2)) + (p2 * e2);
      const v3 po = (p0 * (1.0f - e1)) + (q * e1);
      const b200rt_instance* in = S.inst + L.object;
      const v3 pw = xf_point(in->fwd, po);
      const v3 dir = pw - p;
      *distance = length(dir);
      const float4 c = S.tris[3 * (size_t)tri + 2];
      const float cti = fmaxf(0.001f, -dot(normalize(dir), mk3(c.y, c.z, c.w)));
      *prob = *distance * *distance / (L.total_area * cti);
      return dir;
    }
    case B200RT_LIGHT_SPHERE: {
      const float e1 = rng_u01(key, dim), e2 = rng_u01(key, dim + 1);
      const b200rt_sphere* s = S.spheres + L.object;
      const v3 ps = xf_point(s->inv, p);
      v3 w = ld3(s->center) - ps;
      const float dd = length(w);
      w = normalize(w);
      const float stm = fmaxf(-1.0f, fminf(1.0f, s->radius / dd));
      const float tmax = asinf(stm);
      const float ctm = cosf(tmax);
      const float phi = (float)(2 * RT_PI * (double)e1);
      const float theta = acosf(fmaxf(-1.0f, fminf(1.0f, 1 - e2 + e2 * ctm)));
      const v3 lo = onb_dir7(w, theta, phi);
      const v3 lw = normalize(xf_vector4(s->fwd, lo));
      float t = RT_INF;
      sphere_test(s, p + (lw * S.shadow_eps), lw, 0.0f, S.isect_eps, &t);
      *distance = t;
      *prob = (float)(1 / (2 * RT_PI * (double)(1 - ctm)));
      return lw;
    }
    default: {  // point, spot
      const v3 dir = ld3(L.position) - p;
      *distance = length(dir);
      return dir;
    }
  }
}
// Light::incoming_radiance
__device__ __noinline__ v3 light_radiance7(const DScene& S, int li, v3 to_light, float prob) {
  const b200rt_light& L = S.lights[li];
  const float x = to_light.x, y = to_light.y, z = to_light.z;
  switch (L.kind) {
    case B200RT_LIGHT_POINT: return ld3(L.intensity) / (x * x + y * y + z * z);
    case B200RT_LIGHT_AREA: {
      const v3 rw = -normalize(to_light);
      return (ld3(L.intensity) * dot(rw, ld3(L.normal))) / (x * x + y * y + z * z);
    }
    case B200RT_LIGHT_SPOT: {
      const v3 rw = -normalize(to_light);
      const float ct = dot(rw, ld3(L.edge1));
      const float ccov = L.edge2[0], cfall = L.edge2[1];
      if (ct > cfall) return ld3(L.intensity) / (x * x + y * y + z * z);
      if (ct < ccov) return mk3(0, 0, 0);
      const float c = (ct - ccov) / (cfall - ccov);
      return (ld3(L.intensity) * (float)pow4d((double)c)) / (x * x + y * y + z * z);
    }
    case B200RT_LIGHT_DIRECTIONAL: return ld3(L.intensity);
    case B200RT_LIGHT_ENV: {
      const b200rt_texture& T = S.textures[L.object];
      const float theta = acosf(fminf(1.0f, fmaxf(-1.0f, y)));
      const float phi = atan2f(z, x);
      float u = (float)((RT_PI - (double)phi) / (2.0f * RT_PI));
      float v = (float)((double)theta / RT_PI);
      u = fmaxf(0.0f, fminf(1.0f, u));
      v = fmaxf(0.0f, fminf(1.0f, v));
      u *= T.width;
      if (u >= T.width) u--;
      v *= T.height;
      if (v >= T.height) v--;
      const long long coord = 3LL * T.width * (unsigned int)v + 3LL * (unsigned int)u;
      return mk3(texel7(S, T, coord), texel7(S, T, coord + 1), texel7(S, T, coord + 2)) / prob;
    }
    default: return ld3(L.intensity) / prob;  // mesh, sphere
  }
}

// ---- shade_a (HW7): surface record, emissive / replace_all short cuts, shadow-ray emission ----
#define RT_KDF_REPLACE_ALL 1u
#define RT_KDF_EMISSIVE 2u
__global__ void RT_SHADE7_BOUNDS k_shade7_a(DScene S, Pool P, ShadowQueue Q, const WaveInfo* waves, int d,
                                                       Counters* C) {
  const uint32_t begin = waves[d].begin, count = waves[d].count;
  __shared__ TileOrder7 s_order;
  for (uint32_t base = blockIdx.x * RT_SHADE7_TILE; base < count; base += gridDim.x * RT_SHADE7_TILE) {
   const uint32_t n_valid = min((uint32_t)RT_SHADE7_TILE, count - base);
   tile_order7(s_order, n_valid, [&](uint32_t i) { return surf_class7(__float_as_uint(P.hit_a[begin + base + i].w)); });
#pragma unroll 1
   for (int rj = 0; rj < (RT_SHADE7_ROUNDS > 0 ? RT_SHADE7_ROUNDS : 1); rj++) {
    const uint32_t position = (uint32_t)rj * RT_SHADE7_BLOCK + threadIdx.x;
    const bool act = position < n_valid;
    const uint32_t slot = begin + base + (act ? tile_vertex7(s_order, position) : 0);
    const float4 ha = P.hit_a[slot];
    const uint32_t top_ref = __float_as_uint(ha.w);
    const bool hit = act && B200RT_REF_KIND(top_ref) != B200RT_KIND_NONE;
    v3 p = mk3(0, 0, 0), n = mk3(0, 0, 0);
    float time = 0.f;
    uint32_t key = 0;
    bool lights_on = false;
    if (hit) {
      const float4 ro = P.ray_o[slot], rd = P.ray_d[slot];
      const v3 o = mk3(ro.x, ro.y, ro.z), dir = mk3(rd.x, rd.y, rd.z);
      const uint32_t meta = __float_as_uint(rd.w);
      time = ro.w;
      const Surf7 sf = surface7(S, top_ref, P.hit_b[slot].y, ha.y, ha.z, o, dir, time, ha.x);
      n = sf.n;
      P.nrm[slot] = make_float4(n.x, n.y, n.z, __int_as_float(sf.material));
      uint32_t flags = 0;
      v3 kd = mk3(0, 0, 0);
      if (sf.light >= 0) {
        flags = RT_KDF_EMISSIVE;
        kd = ld3(S.lights[sf.light].intensity);  // radiance of the emissive shape
      } else {
        if (diffuse_constant7(S, sf, &kd)) flags = RT_KDF_REPLACE_ALL;
      }
      P.kd[slot] = make_float4(kd.x, kd.y, kd.z, __uint_as_float(flags));
      p = o + (dir * ha.x);
      key = P.key[slot];
      // trace_ray returns early for emissive shapes and replace_all decals; trace_path only for emissive shapes
      const bool early = (flags & RT_KDF_EMISSIVE) || (!S.integrator && (flags & RT_KDF_REPLACE_ALL));
      lights_on = !early && !(meta & RT_META_INMEDIUM);
    }
    for (int li = 0; li < S.n_lights; li++) {
#ifdef RT_SHADE_SYNC
      __syncthreads();
#endif
      bool want = false;
      v3 wi = mk3(0, 0, 0);
      float ldist = 0.f;
      const uint32_t vslot = slot * (uint32_t)S.n_lights + (uint32_t)li;
      if (lights_on) {
        float prob;
        const v3 lvec = light_dir7(S, li, p, n, key, &ldist, &prob);
        wi = normalize(lvec);
        want = S.semantics < B200RT_SEM_HW7 || dot(n, wi) >= 0.0f;  // HW7/src/Scene.cpp:281; HW5/HW6 have no gate
        if (want) Q.lv[vslot] = make_float4(lvec.x, lvec.y, lvec.z, prob);  // shade_b reuses the sample
      }
      if (act) Q.vis[vslot] = 0;
      const uint32_t qi = warp_append(&C->shadow_count, want);
      if (want) {
        if (qi < Q.capacity) {
          const v3 so = p + (S.shadow_eps * wi);
          Q.o[qi] = make_float4(so.x, so.y, so.z, ldist - S.shadow_eps);
          Q.d[qi] = make_float4(wi.x, wi.y, wi.z, __uint_as_float(vslot));
          Q.time[qi] = time;
        } else {
          C->overflow = 1;
        }
      }
    }
   }
   __syncthreads();   // s_order is rewritten for the next tile
  }
}

// ---- HW4: HW3's shading loop with the textured diffuse colour (HW4/src/Scene.cpp:129-357) ----------
// shade_a: surface record through surface7 (HW4's Mesh_triangle / Sphere already are HW7's, minus culling),
// diffuse colour / replace_all, shadow rays with HW3's rules (no n.w_i gate, lights skipped when kd = ks = 0).
__global__ void __launch_bounds__(RT_BLOCK) k_shade4_a(DScene S, Pool P, ShadowQueue Q, const WaveInfo* waves, int d,
                                                       Counters* C) {
  const uint32_t begin = waves[d].begin, count = waves[d].count;
  const uint32_t stride = gridDim.x * blockDim.x;
  const uint32_t rounds = (count + stride - 1) / stride;
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
    if (hit) {
      const float4 ro = P.ray_o[slot], rd = P.ray_d[slot];
      const v3 o = mk3(ro.x, ro.y, ro.z), dir = mk3(rd.x, rd.y, rd.z);
      const uint32_t meta = __float_as_uint(rd.w);
      time = ro.w;
      const Surf7 sf = surface7(S, top_ref, P.hit_b[slot].y, ha.y, ha.z, o, dir, time, ha.x);
      P.nrm[slot] = make_float4(sf.n.x, sf.n.y, sf.n.z, __int_as_float(sf.material));
      v3 kd;
      const bool replace_all = diffuse_constant7(S, sf, &kd);
      P.kd[slot] = make_float4(kd.x, kd.y, kd.z, __uint_as_float(replace_all ? RT_KDF_REPLACE_ALL : 0u));
      p = o + (dir * ha.x);
      key = P.key[slot];
      const b200rt_material& m = S.mats[sf.material];
      lights_on = !replace_all && !(meta & RT_META_INMEDIUM) && !(is_zero(ld3(m.diffuse)) && is_zero(ld3(m.specular)));
    }
    int area_index = 0;
    for (int li = 0; li < S.n_lights; li++) {
      const b200rt_light& L = S.lights[li];
      v3 wi = mk3(0, 0, 0);
      float ldist = 0.f;
      if (lights_on) {
        const v3 lvec = light_position(L, key, area_index) - p;
        wi = normalize(lvec);
        ldist = length(lvec);
      }
      if (L.kind == B200RT_LIGHT_AREA) area_index++;
      const uint32_t vslot = slot * (uint32_t)S.n_lights + (uint32_t)li;
      if (act) Q.vis[vslot] = 0;
      const uint32_t qi = warp_append(&C->shadow_count, lights_on);
      if (lights_on) {
        if (qi < Q.capacity) {
          const v3 so = p + (S.shadow_eps * wi);
          Q.o[qi] = make_float4(so.x, so.y, so.z, ldist - S.shadow_eps);
          Q.d[qi] = make_float4(wi.x, wi.y, wi.z, __uint_as_float(vslot));
          Q.time[qi] = time;
        } else {
          C->overflow = 1;
        }
      }
    }
  }
}

__global__ void __launch_bounds__(RT_BLOCK) k_shade4_b(DScene S, Pool P, ShadowQueue Q, WaveInfo* waves, int d,
                                                       Counters* C) {
  const uint32_t begin = waves[d].begin, count = waves[d].count;
  const uint32_t next_begin = waves[d + 1].begin;
  const uint32_t stride = gridDim.x * blockDim.x;
  const uint32_t rounds = (count + stride - 1) / stride;
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
    const v3 o = mk3(ro.x, ro.y, ro.z), dir = mk3(rd.x, rd.y, rd.z);
    if (act && !hit && (meta & RT_META_PRIMARY)) {
      // HW4/src/Scene.cpp:135-144: primary by ray type; background texture at (x / width, y / height)
      if (S.background_texture >= 0) {
        const uint32_t pix = P.pixel[slot];
        const int W = S.cam.width, ns = S.cam.samples_sqrt;
        const int i = (int)(pix % (uint32_t)W), j = (int)(pix / (uint32_t)W);
        float x, y;
        if (ns == 1) { x = i + 0.5f; y = j + 0.5f; }
        else {
          const uint32_t sidx = P.sample[slot];
          const int sx = (int)(sidx / (uint32_t)ns), sy = (int)(sidx % (uint32_t)ns);
          x = i + (sx + rng_u01(key, 4)) / ns; y = j + (sy + rng_u01(key, 5)) / ns;
        }
        color = tex_color7(S, S.textures[S.background_texture], x / S.cam.width, y / S.cam.height);
      } else {
        color = S.background;
      }
    }
    if (hit) {
      const float4 nn = P.nrm[slot];
      const v3 n = mk3(nn.x, nn.y, nn.z);
      const b200rt_material& m = S.mats[__float_as_int(nn.w)];
      const float4 kdf = P.kd[slot];
      const v3 kd = mk3(kdf.x, kdf.y, kdf.z);
      p = o + (dir * ha.x);
      const v3 w0 = normalize(o - p);
      if (__float_as_uint(kdf.w) & RT_KDF_REPLACE_ALL) {
        color = kd;   // the texel; no lights, no mirror, no refraction (everything sits inside `if (!is_replace_all)`)
      } else {
        const v3 ks = ld3(m.specular);
        const bool has_d = !is_zero(ld3(m.diffuse)), has_s = !is_zero(ks);
        if (!(meta & RT_META_INMEDIUM)) {
          color = color + (ld3(m.ambient) * S.ambient);
          int area_index = 0;
          for (int li = 0; li < S.n_lights; li++) {
            const b200rt_light& L = S.lights[li];
            const int ai = area_index;
            if (L.kind == B200RT_LIGHT_AREA) area_index++;
            if (!has_d && !has_s) continue;
            if (!Q.vis[slot * (uint32_t)S.n_lights + (uint32_t)li]) continue;
            const v3 lvec = light_position(L, key, ai) - p;
            const v3 wi = normalize(lvec);
            const float ldist = length(lvec);
            const float d2 = ldist * ldist;
            const v3 inten = ld3(L.intensity);
            if (L.kind == B200RT_LIGHT_AREA) {
              const float icos = dot(-wi, ld3(L.normal));
              if (has_d) color = color + ((((kd * inten) * icos) * dot(n, wi)) / d2);
              if (has_s) {
                const float sc = fmaxf(0.0f, dot(n, normalize(w0 + wi)));
                color = color + ((((ks * inten) * icos) * (float)pow_ref7((double)sc, (double)m.phong_exponent)) / d2);
              }
            } else {
              if (has_d) color = color + (((kd * inten) * dot(n, wi)) / d2);
              if (has_s) {
                const float sc = fmaxf(0.0f, dot(n, normalize(w0 + wi)));
                color = color + (((ks * inten) * (float)pow_ref7((double)sc, (double)m.phong_exponent)) / d2);
              }
            }
          }
        }
        const bool budget = (int)depth < S.max_depth;
        km = ld3(m.mirror);
        if (!is_zero(km) && budget) {
          const v3 wr = normalize((n * (2 * dot(n, w0))) - w0);
          dir_m = wr;
          if (m.roughness != 0.0f) {
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
        if (!is_zero(kt) && budget) {
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
            kk = mk3((float)exp(log((double)kt.x) * ht), (float)exp(log((double)kt.y) * ht),
                     (float)exp(log((double)kt.z) * ht));
            entering = false;
            if (refract_ray(dn, -n, 1.0f / ni, &dir_t)) cos_theta = dot(dir_t, n);
            else tir = true;
          }
          dir_r = wr;
          want_r = true;
          if (tir) {
            med_r = true;
            fres = -1.0f;
          } else {
            const float r0 = ((ni - 1) * (ni - 1)) / ((ni + 1) * (ni + 1));
            fres = (float)((double)r0 + (double)(1 - r0) * pow5d((double)(1.0f - cos_theta)));
            med_r = !entering;
            med_t = entering;
            want_t = true;
          }
        }
      }
    }
    uint32_t cm = 0xffffffffu, cr = 0xffffffffu, ct = 0xffffffffu;
#pragma unroll
    for (int c = 0; c < 3; c++) {
      const bool want = (c == 0) ? want_m : (c == 1 ? want_r : want_t);
      const uint32_t ci = warp_append(&waves[d + 1].count, want);
      if (want) {
        const uint32_t cs = next_begin + ci;
        if (cs < P.capacity) {
          const v3 cdir = (c == 0) ? dir_m : (c == 1 ? dir_r : dir_t);
          const v3 co = p + (cdir * S.shadow_eps);
          const bool med = (c == 0) ? false : (c == 1 ? med_r : med_t);
          const uint32_t cmeta = (depth + 1) | (med ? RT_META_INMEDIUM : 0u);
          P.ray_o[cs] = make_float4(co.x, co.y, co.z, time);
          P.ray_d[cs] = make_float4(cdir.x, cdir.y, cdir.z, __uint_as_float(cmeta));
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

// ---- shade_b (HW7): Scene::send_ray / trace_ray / trace_path bodies -------------------------
__global__ void RT_SHADE7_BOUNDS k_shade7_b(DScene S, Pool P, ShadowQueue Q, WaveInfo* waves, int d,
                                                       Counters* C) {
  const uint32_t begin = waves[d].begin, count = waves[d].count;
  const uint32_t next_begin = waves[d + 1].begin;
  const bool pt = S.integrator != 0;
  __shared__ TileOrder7 s_order;
  for (uint32_t base = blockIdx.x * RT_SHADE7_TILE; base < count; base += gridDim.x * RT_SHADE7_TILE) {
   const uint32_t n_valid = min((uint32_t)RT_SHADE7_TILE, count - base);
   // by material (written by shade_a next to the normal); vertices that hit nothing go last
   tile_order7(s_order, n_valid,
               [&](uint32_t i) {
                 const uint32_t sl = begin + base + i;
                 const uint32_t top_ref = __float_as_uint(P.hit_a[sl].w);
                 if (B200RT_REF_KIND(top_ref) == B200RT_KIND_NONE) return (uint32_t)(RT_SHADE7_CLASSES - 1);
#ifdef RT_SHADE7_B_BY_OBJECT
                 return surf_class7(top_ref);
#endif
                 return (uint32_t)__float_as_int(P.nrm[sl].w) % (RT_SHADE7_CLASSES - 1);
               });
#pragma unroll 1
   for (int rj = 0; rj < (RT_SHADE7_ROUNDS > 0 ? RT_SHADE7_ROUNDS : 1); rj++) {
    const uint32_t position = (uint32_t)rj * RT_SHADE7_BLOCK + threadIdx.x;
    const bool act = position < n_valid;
    const uint32_t slot = begin + base + (act ? tile_vertex7(s_order, position) : 0);
    const float4 ha = P.hit_a[slot];
    const float4 ro = P.ray_o[slot], rd = P.ray_d[slot];
    const uint32_t meta = __float_as_uint(rd.w);
    const uint32_t depth = RT_META_DEPTH(meta);
    const uint32_t top_ref = __float_as_uint(ha.w);
    const bool hit = act && B200RT_REF_KIND(top_ref) != B200RT_KIND_NONE;
    v3 color = mk3(0, 0, 0);
    v3 km = mk3(0, 0, 0), kk = mk3(0, 0, 0), wdiff = mk3(0, 0, 0);
    float fres = 0.f;
    bool want_m = false, want_r = false, want_t = false, want_d = false;
    v3 dir_m = mk3(0, 0, 0), dir_r = mk3(0, 0, 0), dir_t = mk3(0, 0, 0), dir_d = mk3(0, 0, 0);
    bool med_r = false, med_t = false;
    bool lh_m = false, lh_rt = false, lh_d = false;  // light_hit of the children
    v3 p = mk3(0, 0, 0);
    const float time = ro.w;
    const uint32_t key = act ? P.key[slot] : 0u;
    const v3 o = mk3(ro.x, ro.y, ro.z), dir = mk3(rd.x, rd.y, rd.z);
    if (act && !hit) {
      // Scene::send_ray miss branch — HW7/src/Scene.cpp:170-183
      if (meta & RT_META_PRIMARY) {
        if (S.background_texture >= 0) {
          const uint32_t pix = P.pixel[slot];
          const int W = S.cam.width, ns = S.cam.samples_sqrt;
          const int i = (int)(pix % (uint32_t)W), j = (int)(pix / (uint32_t)W);
          float x, y;
          if (ns == 1) { x = i + 0.5f; y = j + 0.5f; }
          else {
            const uint32_t sidx = P.sample[slot];
            const int sx = (int)(sidx / (uint32_t)ns), sy = (int)(sidx % (uint32_t)ns);
            x = i + (sx + rng_u01(key, 4)) / ns; y = j + (sy + rng_u01(key, 5)) / ns;
          }
          color = tex_color7(S, S.textures[S.background_texture], x / S.cam.width, y / S.cam.height);
        } else if (S.env_light >= 0) {
          color = light_radiance7(S, S.env_light, dir, 1.0f);
        } else {
          color = S.background;
        }
      } else if (S.env_light >= 0) {
        color = light_radiance7(S, S.env_light, dir, 1.0f);
      }
    }
    // Surface state of a hit vertex, hoisted so that the light loop below runs in lock step for the whole block
    v3 n = mk3(0, 0, 0), kd = mk3(0, 0, 0), ks = mk3(0, 0, 0), w_o = mk3(0, 0, 0);
    int mat = 0;
    uint32_t flags = 0;
    const bool in_medium = (meta & RT_META_INMEDIUM) != 0;
    const bool light_hit = (meta & RT_META_LIGHTHIT) != 0;
    if (hit) {
      const float4 nn = P.nrm[slot];
      n = mk3(nn.x, nn.y, nn.z);
      mat = __float_as_int(nn.w);
      const float4 kdf = P.kd[slot];
      kd = mk3(kdf.x, kdf.y, kdf.z);
      flags = __float_as_uint(kdf.w);
      p = o + (dir * ha.x);
      w_o = normalize(o - p);
      ks = ld3(S.mats[mat].specular);
    }
    // trace_ray / trace_path go on past the emissive and replace_all short cuts only for these vertices
    const bool shaded = hit && !(flags & RT_KDF_EMISSIVE) && !(!pt && (flags & RT_KDF_REPLACE_ALL));
    // HW7 sums the lights on their own and adds the sum to ambient (HW7/src/Scene.cpp:260-311, 434-437); HW5/HW6
    // add every term straight onto the colour, which starts as the ambient term (HW6/src/Scene.cpp:169-207)
    const bool onto_colour = S.semantics < B200RT_SEM_HW7;
    v3 direct = (shaded && onto_colour) ? (color + (ld3(S.mats[mat].ambient) * S.ambient)) : mk3(0, 0, 0);
    // calculate_diffuse_and_specular_radiance — HW7/src/Scene.cpp:260-311
    for (int li = 0; li < S.n_lights; li++) {
#ifdef RT_SHADE_SYNC
      __syncthreads();
#endif
      const uint32_t vslot = slot * (uint32_t)S.n_lights + (uint32_t)li;
      if (!(shaded && !in_medium) || !Q.vis[vslot]) continue;
      if (S.lit_count) atomicAdd(S.lit_count + (size_t)P.pixel[slot] * S.n_lights + li, 1);   // parity hook (b200rt_debug_lit_counts)
      const b200rt_material& m = S.mats[mat];
      const float4 ls = Q.lv[vslot];  // light_dir7's sample, drawn once in shade_a
      const v3 lvec = mk3(ls.x, ls.y, ls.z);
      const float prob = ls.w;
      const v3 w_i = normalize(lvec);
      const v3 Li = light_radiance7(S, li, lvec, prob);
      const float cti = fmaxf(0.0f, dot(n, w_i));
      if (m.brdf != -1) {
        direct = direct + ((Li * cti) * brdf_eval7(S.brdfs[m.brdf], n, kd, ks, w_i, w_o));
      } else {
        direct = direct + ((kd * Li) * cti);
        const float sct = fmaxf(0.0f, dot(n, normalize(w_o + w_i)));
        direct = direct + ((ks * Li) * (float)pow_ref7((double)sct, (double)m.phong_exponent));
      }
    }
#ifdef RT_SHADE_SYNC
    __syncthreads();
#endif
    if (hit) {
      const b200rt_material& m = S.mats[mat];
      if (flags & RT_KDF_EMISSIVE) {
        color = (pt && light_hit) ? mk3(0, 0, 0) : kd;
      } else if (!pt && (flags & RT_KDF_REPLACE_ALL)) {
        color = kd;
      } else {
        bool copy_light_hit = light_hit;
        if (pt) {
          if (!is_zero(direct)) { copy_light_hit = true; color = color + direct; }
        } else if (!in_medium) {
          if (onto_colour) {
            color = direct;
          } else {
            color = color + (ld3(m.ambient) * S.ambient);
            color = color + direct;
          }
        }
        const bool budget = (int)depth < S.max_depth;
        if (pt && !in_medium && budget) {
          // hemisphere sample — HW7/src/Scene.cpp:341-399
          const float e1 = rng_u01(key, 9), e2 = rng_u01(key, 10);
          const float phi = (float)(2 * RT_PI * (double)e1);
          const float theta = S.importance_sampling ? asinf(sqrtf(e2)) : acosf(e2);
          const v3 w_i = onb_dir7(n, theta, phi);
          const float cti = fmaxf(0.0f, dot(n, w_i));
          const float prob = S.importance_sampling ? (float)((double)cti / RT_PI) : (float)(1.0f / (2 * RT_PI));
          dir_d = w_i;
          want_d = true;  // the reference traces the sample ray even when probability == 0
          lh_d = light_hit;
          if (prob != 0.0f) {
            if (m.brdf != -1) {
              wdiff = (brdf_eval7(S.brdfs[m.brdf], n, kd, ks, w_i, w_o) * cti) / prob;
            } else {
              const float sct = fmaxf(0.0f, dot(n, normalize(w_o + w_i)));
              wdiff = ((kd * cti) / prob) + ((ks * (float)pow_ref7((double)sct, (double)m.phong_exponent)) / prob);
            }
          }
        }
        km = ld3(m.mirror);
        if (!is_zero(km) && budget) {
          // Scene::reflect_ray — HW7/src/Scene.cpp:191-229
          const v3 wr = normalize((n * (2 * dot(n, w_o))) - w_o);
          dir_m = wr;
          if (m.roughness != 0.0f) {
            v3 rp;
            if (wr.x < wr.y && wr.x < wr.z) rp = mk3(1.0f, wr.y, wr.z);
            else if (wr.y < wr.z && wr.y < wr.x) rp = mk3(wr.x, 1.0f, wr.z);
            else rp = mk3(wr.x, wr.y, 1.0f);
            const v3 u = normalize(cross(rp, wr));
            const v3 v = normalize(cross(u, wr));
            const float eu = rng_u01(key, 2) - 0.5f, ev = rng_u01(key, 3) - 0.5f;
            dir_m = normalize(wr + (((u * eu) + (v * ev)) * m.roughness));
            lh_m = copy_light_hit;  // only the glossy branch propagates light_hit
          }
          want_m = true;
        }
        const v3 kt = ld3(m.transparency);
        if (!is_zero(kt) && budget) {
          // Scene::refract_ray — HW7/src/Scene.cpp:108-166
          const v3 wr = normalize((n * (2 * dot(n, w_o))) - w_o);
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
            kk = mk3((float)exp(log((double)kt.x) * ht), (float)exp(log((double)kt.y) * ht),
                     (float)exp(log((double)kt.z) * ht));
            entering = false;
            if (refract_ray(dn, -n, 1.0f / ni, &dir_t)) cos_theta = dot(dir_t, n);
            else tir = true;
          }
          dir_r = wr;
          want_r = true;
          lh_rt = copy_light_hit;
          if (tir) {
            med_r = true;
            fres = -1.0f;
          } else {
            const float r0 = ((ni - 1) * (ni - 1)) / ((ni + 1) * (ni + 1));
            fres = (float)((double)r0 + (double)(1 - r0) * pow5d((double)(1.0f - cos_theta)));
            med_r = !entering;
            med_t = entering;
            want_t = true;
          }
        }
      }
    }
    // children -> next wave (order: diffuse sample, mirror, reflect, transmit)
#ifdef RT_SHADE_SYNC
    __syncthreads();
#endif
    uint32_t cd = 0xffffffffu, cm = 0xffffffffu, cr = 0xffffffffu, ct = 0xffffffffu;
#pragma unroll
    for (int c = 0; c < 4; c++) {
      const bool want = (c == 0) ? want_d : (c == 1 ? want_m : (c == 2 ? want_r : want_t));
      const uint32_t ci = warp_append(&waves[d + 1].count, want);
      if (want) {
        const uint32_t cs = next_begin + ci;
        if (cs < P.capacity) {
          const v3 cdir = (c == 0) ? dir_d : (c == 1 ? dir_m : (c == 2 ? dir_r : dir_t));
          const v3 co = p + (cdir * S.shadow_eps);
          const bool med = (c == 0) ? ((meta & RT_META_INMEDIUM) != 0) : (c == 1 ? false : (c == 2 ? med_r : med_t));
          const bool lh = (c == 0) ? lh_d : (c == 1 ? lh_m : lh_rt);
          const uint32_t cmeta = (depth + 1) | (med ? RT_META_INMEDIUM : 0u) | (lh ? RT_META_LIGHTHIT : 0u);
          P.ray_o[cs] = make_float4(co.x, co.y, co.z, time);
          P.ray_d[cs] = make_float4(cdir.x, cdir.y, cdir.z, __uint_as_float(cmeta));
          const uint32_t sub = (c == 0) ? 0x104u : (0x100u + (uint32_t)c);
          P.key[cs] = rng_mix(key, sub);
          P.pixel[cs] = P.pixel[slot];
          if (c == 0) cd = cs; else if (c == 1) cm = cs; else if (c == 2) cr = cs; else ct = cs;
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
      P.rec4[slot] = make_float4(wdiff.x, wdiff.y, wdiff.z, __uint_as_float(cd));
    }
   }
   __syncthreads();   // s_order is rewritten for the next tile
  }
}

// ---- combine (HW7): radiance = local + w_d*L_d + km*L_m + k*(r*L_r + (1-r)*L_t) ------------
__global__ void k_combine7(DScene S, Pool P, const WaveInfo* waves, int d, float4* accum) {
  const uint32_t begin = waves[d].begin, count = waves[d].count;
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < count; k += stride) {
    const uint32_t slot = begin + k;
    const float4 r0 = P.rec0[slot], r1 = P.rec1[slot], r2 = P.rec2[slot], r4 = P.rec4[slot];
    const uint32_t cm = __float_as_uint(r1.w), cr = __float_as_uint(r2.w), ct = P.rec3[slot],
                   cd = __float_as_uint(r4.w);
    v3 color = mk3(r0.x, r0.y, r0.z);
    if (cd != 0xffffffffu) {
      const float4 c = P.final_c[cd];
      color = color + (mk3(r4.x, r4.y, r4.z) * mk3(c.x, c.y, c.z));
    }
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
      if (S.cam.samples_sqrt == 1) {
        accum[pix] = make_float4(color.x * 1.0f, color.y * 1.0f, color.z * 1.0f, 1.0f);
      } else if (S.semantics < B200RT_SEM_HW7) {
        // HW5/HW6: 3x3 Gaussian splat, sigma 1/2, around the absolute sample position — HW6/src/Scene.cpp:84-97
        const int W = S.cam.width, H = S.cam.height, ns = S.cam.samples_sqrt;
        const int i = (int)(pix % (uint32_t)W), j = (int)(pix / (uint32_t)W);
        const uint32_t sidx = P.sample[slot], key = P.key[slot];
        const int x = (int)(sidx / (uint32_t)ns), y = (int)(sidx % (uint32_t)ns);
        const float sx = ((float)x + rng_u01(key, 4)) / (float)ns, sy = ((float)y + rng_u01(key, 5)) / (float)ns;
        for (int aj = j - 1; aj < j + 2; aj++) {
          if (aj < 0 || aj >= H) continue;
          for (int ai = i - 1; ai < i + 2; ai++) {
            if (ai < 0 || ai >= W) continue;
            const float w = gaussian_filter(((float)i + sx) - ((float)ai + 0.5f), ((float)j + sy) - ((float)aj + 0.5f),
                                            1.5f / 3.0f);
            // Pixel::add_color(color, w): one 16-byte vector atomic (sm_90+) instead of four scalar ones
            atomicAdd(accum + (size_t)aj * W + ai, make_float4(color.x * w, color.y * w, color.z * w, w));
          }
        }
      } else {
        // box filter: Pixel::add_color(color, 1.0f) per sample — HW7/src/Scene.cpp:80
        atomicAdd(accum + pix, make_float4(color.x, color.y, color.z, 1.0f));
      }
    }
  }
}
