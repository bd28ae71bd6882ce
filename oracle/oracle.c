/*
 * oracle.c — TEST INFRASTRUCTURE ONLY (never linked into or called from the
 * product path; only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load it).
 *
 * A plain-C, CPU restatement of the ceng795 hot path: ray generation, the
 * reference's recursive *unpruned* BVH traversal, Cramer's-rule triangle and
 * quadratic sphere intersection, instancing, and the Whitted shading loop, in
 * the reference's operation order and IEEE binary32 (build with
 * -ffp-contract=off).  Every function cites the reference lines it follows.
 * It consumes the same flattened B200SceneDesc the CUDA core consumes.
 *
 * Pinning: the reference ships no tests or golden vectors, so
 * this restatement is pinned against the reference ITSELF: oracle/build_ref.sh
 * compiles the untouched trees into oracle/_ref/, and tests/test_oracle_vs_ref.py
 * checks that this file reproduces their PNGs byte for byte on deterministic
 * scenes; tests/golden/ holds small reference-rendered fixtures (with the
 * generating script) for boxes where /root/reference is absent.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../include/b200rt.h"

typedef struct { float x, y, z; } v3;
static inline v3 V(float x, float y, float z) { v3 r = {x, y, z}; return r; }
static inline v3 ld(const float* p) { return V(p[0], p[1], p[2]); }
static inline v3 add(v3 a, v3 b) { return V(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline v3 sub(v3 a, v3 b) { return V(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline v3 mulv(v3 a, v3 b) { return V(a.x * b.x, a.y * b.y, a.z * b.z); }
static inline v3 muls(v3 a, float s) { return V(a.x * s, a.y * s, a.z * s); }
static inline v3 divs(v3 a, float s) { return V(a.x / s, a.y / s, a.z / s); }
static inline v3 neg(v3 a) { return V(-a.x, -a.y, -a.z); }
static inline float dot(v3 a, v3 b) { return (a.x * b.x) + (a.y * b.y) + (a.z * b.z); }
static inline v3 cross(v3 a, v3 b) {
  return V(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
static inline float len(v3 a) { return sqrtf(a.x * a.x + a.y * a.y + a.z * a.z); }
static inline v3 normalize(v3 a) { return divs(a, len(a)); }
static inline int is_zero(v3 a) { return a.x == 0.0f && a.y == 0.0f && a.z == 0.0f; }
static inline float comp(v3 a, int i) { return i == 0 ? a.x : (i == 1 ? a.y : a.z); }

typedef struct {
  v3 o, d;
  float time;
  int in_medium;
  int is_shadow;
  int light_hit;   /* HW7 Ray::light_hit */
  int primary;     /* HW7 ray_type == r_primary */
  float bg_u, bg_v;
} Ray;

typedef struct {
  float t;
  int top_rank;   /* DFS rank of the top-level object, -1 none */
  int tri;        /* triangle index (mesh DFS order) or -1      */
  uint32_t shape; /* ref of the object that owns the material   */
  int material;
  v3 normal;
  int valid;      /* shape != NULL */
  /* HW4..HW7 */
  float u, v, perlin;
  int texture;    /* texture id of the shape that owns the hit (instance / sphere), -1 none */
  int light;      /* is_light_object: index into lights[] (radiance), else -1 */
} Hit;

typedef struct OracleCounters {
  uint64_t rays_primary, rays_shadow, rays_secondary;
  uint64_t nodes_visited, tris_tested, spheres_tested, instances_entered;
} OracleCounters;

typedef struct {
  const B200SceneDesc* d;
  OracleCounters* c;
} Ctx;
/* Parity hook: per pixel x light, how many shadow tests of direct7 came out "lit" (the device counts the same thing,
 * b200rt_debug_lit_counts).  The pixel being rendered is thread-local; NULL buffer = off. */
static int32_t* g_lit_count = NULL;
static _Thread_local int tl_pixel = -1;
void oracle_set_lit_counts(int32_t* buf) { g_lit_count = buf; }
static inline int sem_hw7(const B200SceneDesc* d) { return d->semantics >= B200RT_SEM_HW7; }
/* HW5 and HW6 shade like HW7's ray-tracing mode (unified light list, clamped cosine, BRDF objects from HW6 on:
 * HW6/src/Scene.cpp:119-301) but without back-face culling, without the n.w_i >= 0 gate in front of the shadow ray,
 * and with the sigma-1/2 Gaussian splat on the multi-sample path (HW6/src/Scene.cpp:18-101). */
static inline int sem_hw5(const B200SceneDesc* d) { return d->semantics >= B200RT_SEM_HW5; }
/* HW4 already has the textured Mesh_triangle / Sphere of HW7 (band acceptance, per-call sphere transform), but no culling */
static inline int sem_hw4(const B200SceneDesc* d) { return d->semantics >= B200RT_SEM_HW4; }

/* Matrix4x4::multiply — HW3/src/Matrix4x4.cpp:29-50 (3x4 rows) */
static v3 xf_point(const float* m, v3 p) {
  float r[3];
  const float q[3] = {p.x, p.y, p.z};
  for (int i = 0; i < 3; i++) {
    r[i] = m[i * 4 + 3];
    for (int j = 0; j < 3; j++) r[i] += m[i * 4 + j] * q[j];
  }
  return V(r[0], r[1], r[2]);
}
static v3 xf_vector(const float* m, int stride, v3 p) {
  float r[3];
  const float q[3] = {p.x, p.y, p.z};
  for (int i = 0; i < 3; i++) {
    r[i] = 0;
    for (int j = 0; j < 3; j++) r[i] += m[i * stride + j] * q[j];
  }
  return V(r[0], r[1], r[2]);
}

/* 4x4 helpers for the motion-blur branch (HW3/include/Mesh.h:76-95) */
static void m4_mul(const float* a, const float* b, float* r) {
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 4; j++) {
      float s = 0.0f;
      for (int k = 0; k < 4; k++) s += a[i * 4 + k] * b[k * 4 + j];
      r[i * 4 + j] = s;
    }
}
/* Matrix4x4::invert_matrix — HW3/src/Matrix4x4.cpp:97-233 */
static int m4_invert(const float* m, float* out) {
  float inv[16], det;
  inv[0] = m[5] * m[10] * m[15] - m[5] * m[11] * m[14] - m[9] * m[6] * m[15] +
           m[9] * m[7] * m[14] + m[13] * m[6] * m[11] - m[13] * m[7] * m[10];
  inv[4] = -m[4] * m[10] * m[15] + m[4] * m[11] * m[14] + m[8] * m[6] * m[15] -
           m[8] * m[7] * m[14] - m[12] * m[6] * m[11] + m[12] * m[7] * m[10];
  inv[8] = m[4] * m[9] * m[15] - m[4] * m[11] * m[13] - m[8] * m[5] * m[15] +
           m[8] * m[7] * m[13] + m[12] * m[5] * m[11] - m[12] * m[7] * m[9];
  inv[12] = -m[4] * m[9] * m[14] + m[4] * m[10] * m[13] + m[8] * m[5] * m[14] -
            m[8] * m[6] * m[13] - m[12] * m[5] * m[10] + m[12] * m[6] * m[9];
  inv[1] = -m[1] * m[10] * m[15] + m[1] * m[11] * m[14] + m[9] * m[2] * m[15] -
           m[9] * m[3] * m[14] - m[13] * m[2] * m[11] + m[13] * m[3] * m[10];
  inv[5] = m[0] * m[10] * m[15] - m[0] * m[11] * m[14] - m[8] * m[2] * m[15] +
           m[8] * m[3] * m[14] + m[12] * m[2] * m[11] - m[12] * m[3] * m[10];
  inv[9] = -m[0] * m[9] * m[15] + m[0] * m[11] * m[13] + m[8] * m[1] * m[15] -
           m[8] * m[3] * m[13] - m[12] * m[1] * m[11] + m[12] * m[3] * m[9];
  inv[13] = m[0] * m[9] * m[14] - m[0] * m[10] * m[13] - m[8] * m[1] * m[14] +
            m[8] * m[2] * m[13] + m[12] * m[1] * m[10] - m[12] * m[2] * m[9];
  inv[2] = m[1] * m[6] * m[15] - m[1] * m[7] * m[14] - m[5] * m[2] * m[15] +
           m[5] * m[3] * m[14] + m[13] * m[2] * m[7] - m[13] * m[3] * m[6];
  inv[6] = -m[0] * m[6] * m[15] + m[0] * m[7] * m[14] + m[4] * m[2] * m[15] -
           m[4] * m[3] * m[14] - m[12] * m[2] * m[7] + m[12] * m[3] * m[6];
  inv[10] = m[0] * m[5] * m[15] - m[0] * m[7] * m[13] - m[4] * m[1] * m[15] +
            m[4] * m[3] * m[13] + m[12] * m[1] * m[7] - m[12] * m[3] * m[5];
  inv[14] = -m[0] * m[5] * m[14] + m[0] * m[6] * m[13] + m[4] * m[1] * m[14] -
            m[4] * m[2] * m[13] - m[12] * m[1] * m[6] + m[12] * m[2] * m[5];
  inv[3] = -m[1] * m[6] * m[11] + m[1] * m[7] * m[10] + m[5] * m[2] * m[11] -
           m[5] * m[3] * m[10] - m[9] * m[2] * m[7] + m[9] * m[3] * m[6];
  inv[7] = m[0] * m[6] * m[11] - m[0] * m[7] * m[10] - m[4] * m[2] * m[11] +
           m[4] * m[3] * m[10] + m[8] * m[2] * m[7] - m[8] * m[3] * m[6];
  inv[11] = -m[0] * m[5] * m[11] + m[0] * m[7] * m[9] + m[4] * m[1] * m[11] -
            m[4] * m[3] * m[9] - m[8] * m[1] * m[7] + m[8] * m[3] * m[5];
  inv[15] = m[0] * m[5] * m[10] - m[0] * m[6] * m[9] - m[4] * m[1] * m[10] +
            m[4] * m[2] * m[9] + m[8] * m[1] * m[6] - m[8] * m[2] * m[5];
  det = m[0] * inv[0] + m[1] * inv[4] + m[2] * inv[8] + m[3] * inv[12];
  if (det == 0.0f) return 0;
  det = 1.0f / det;
  for (int i = 0; i < 16; i++) out[i] = inv[i] * det;
  return 1;
}
/* Translation(time*velocity) * M, inverse and normal matrix per ray. */
static void moving_xf(const float* fwd, const float* vel, float time, float* inv16, float* nrm16) {
  float tr[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
  const v3 dp = muls(ld(vel), time); /* ray.time * velocity */
  tr[3] = dp.x; tr[7] = dp.y; tr[11] = dp.z;
  float prod[16];
  m4_mul(tr, fwd, prod);
  if (!m4_invert(prod, inv16)) { /* reference exit(-1)s; keep going with identity */
    memset(inv16, 0, 64);
    inv16[0] = inv16[5] = inv16[10] = inv16[15] = 1.0f;
  }
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 4; j++) nrm16[i * 4 + j] = inv16[j * 4 + i];
}

/* Bounding_box::intersect — HW3/src/bounding_box.cpp:17-37 */
static float box_intersect(const float* mn, const float* mx, const Ray* ray, float eps) {
  float tmin = -INFINITY, tmax = INFINITY;
  for (int i = 0; i < 3; i++) {
    const float di = comp(ray->d, i), oi = comp(ray->o, i);
    if (fabsf(di) < eps) continue;
    float ti_min = (mn[i] - oi) / di;
    float ti_max = (mx[i] - oi) / di;
    if (di < 0) { float tmp = ti_min; ti_min = ti_max; ti_max = tmp; }
    if (ti_min > tmin) tmin = ti_min;
    if (ti_max < tmax) tmax = ti_max;
    if (tmin > tmax) return INFINITY;
  }
  return (tmin > 0.0f) ? tmin : tmax;
}

/* Mesh_triangle::determinant — HW3/include/Mesh_triangle.h:42-47 */
static inline float det3(v3 c1, v3 c2, v3 c3) {
  return c1.x * (c2.y * c3.z - c3.y * c2.z) + c2.x * (c3.y * c1.z - c1.y * c3.z) +
         c3.x * (c1.y * c2.z - c2.y * c1.z);
}

/* Cramer's-rule core shared by Mesh_triangle/Triangle of every tree.  mode:
 *   1 = HW1 band (HW1/Triangle.cpp:29-47), 2 = exact [0,1] t>0
 *   (HW3/src/Mesh_triangle.cpp:85-97), 7 = HW7/PPM mesh band (no beta>1 test). */
static int tri_core(const float* g, v3 o, v3 d, int mode, float eps, float* t_out, float* beta_out,
                    float* gamma_out) {
  const v3 v0 = ld(g), a1 = ld(g + 3), a2 = ld(g + 6);
  const float det_a = det3(a1, a2, d);
  if (det_a == 0.0f) return 0;
  const v3 b = divs(sub(v0, o), det_a);
  const float beta = det3(b, a2, d);
  if (mode == 1) {
    if (beta < 0.0f - eps || beta > 1.0f + eps) return 0;
  } else if (mode == 2) {
    if (beta < 0.0f || beta > 1.0f) return 0;
  } else {
    if (beta < -eps) return 0;
  }
  const float gamma = det3(a1, b, d);
  if (mode == 1) {
    if (gamma < 0.0f - eps || beta + gamma > 1.0f + eps) return 0;
  } else if (mode == 2) {
    if (gamma < 0.0f || beta + gamma > 1.0f) return 0;
  } else {
    if (gamma < -eps || beta + gamma > 1.0f + eps) return 0;
  }
  const float t = det3(a1, a2, b);
  if (mode == 2) {
    if (!(t > 0.0f)) return 0;
  } else {
    if (!(t > -eps)) return 0;
  }
  *t_out = t; *beta_out = beta; *gamma_out = gamma;
  return 1;
}

static int shape_intersect(const Ctx* cx, uint32_t ref, const float* box, const Ray* ray, Hit* hit, int culling);
static void tex_surface_tri(const B200SceneDesc* d, int tri, float beta, float gamma, v3 local_p, Hit* hit);
static void tex_surface_sphere(const B200SceneDesc* d, const b200rt_sphere* s, v3 local_coords, v3 local_p, Hit* hit);

/* Mesh_triangle::intersect — HW3/src/Mesh_triangle.cpp:75-114, HW7/src/Mesh_triangle.cpp:56-157 */
static int mesh_tri_intersect(const Ctx* cx, int tri, const Ray* ray, Hit* hit, int culling) {
  const B200SceneDesc* d = cx->d;
  cx->c->tris_tested++;
  const int hw7 = sem_hw7(d), hw4 = sem_hw4(d);
  if (hw7 && culling && dot(ray->d, ld(d->tri_geom + 12 * (size_t)tri + 9)) > 0.0f) return 0;
  float t, beta, gamma;
  /* HW4/src/Mesh_triangle.cpp:64-76 already uses the +-eps band of HW7/src/Mesh_triangle.cpp:75-88 */
  if (!tri_core(d->tri_geom + 12 * (size_t)tri, ray->o, ray->d, hw4 ? 7 : 2, d->isect_eps, &t, &beta, &gamma))
    return 0;
  hit->t = t;
  hit->tri = tri;
  hit->shape = B200RT_REF(B200RT_KIND_TRI, tri);
  hit->material = d->tri_info[2 * (size_t)tri];
  hit->valid = 1;
  hit->light = -1;
  hit->texture = -1;
  if (d->tri_info[2 * (size_t)tri + 1] & 1) {
    const int32_t* vi = d->tri_vidx + 3 * (size_t)tri;
    const v3 n0 = ld(d->vert_normal + 3 * (size_t)vi[0]);
    const v3 n1 = ld(d->vert_normal + 3 * (size_t)vi[1]);
    const v3 n2 = ld(d->vert_normal + 3 * (size_t)vi[2]);
    hit->normal = normalize(add(add(muls(n0, 1 - beta - gamma), muls(n1, beta)), muls(n2, gamma)));
  } else {
    hit->normal = ld(d->tri_geom + 12 * (size_t)tri + 9);
  }
  if (hw4) {
    hit->u = -1; hit->v = -1; hit->perlin = -1;
    if (d->tri_tex[tri] != -1) tex_surface_tri(d, tri, beta, gamma, add(ray->o, muls(ray->d, t)), hit);
  }
  return 1;
}

/* Triangle::intersect (stand-alone primitive) — HW3/src/Triangle.cpp:51-109 */
static int triprim_intersect(const Ctx* cx, int idx, const Ray* ray, Hit* hit, int culling) {
  const b200rt_triprim* p = cx->d->triprims + idx;
  cx->c->tris_tested++;
  float g[12];
  memcpy(g, p->v0, 12); memcpy(g + 3, p->a1, 12); memcpy(g + 6, p->a2, 12);
  v3 o = ray->o, dd = ray->d;
  const int ident = (p->flags & B200RT_XF_IDENTITY) != 0;
  if (!ident) {
    o = xf_point(p->inv, ray->o);
    dd = xf_vector(p->inv, 4, ray->d);
  }
  /* HW7/src/Triangle.cpp:69-71,106-108: culling against the object-space normal */
  if (sem_hw7(cx->d) && culling && dot(dd, ld(p->normal)) > 0) return 0;
  float t, beta, gamma;
  if (!tri_core(g, o, dd, 2, cx->d->isect_eps, &t, &beta, &gamma)) return 0;
  hit->t = t;
  hit->tri = -1;
  hit->light = -1; hit->texture = -1; hit->u = 0; hit->v = 0; hit->perlin = 0;
  hit->shape = B200RT_REF(B200RT_KIND_TRIPRIM, idx);
  hit->material = p->material;
  hit->top_rank = p->rank;
  hit->valid = 1;
  hit->normal = ident ? ld(p->normal) : normalize(xf_vector(p->nrm, 3, ld(p->normal)));
  return 1;
}

/* Sphere::intersect — HW3/include/Sphere.h:44-155 (three branches) and
 * HW1/Sphere.h:20-49.  Returns 1 even when t < 0, like the reference. */
static int sphere_intersect(const Ctx* cx, int idx, const Ray* ray, Hit* hit) {
  const b200rt_sphere* s = cx->d->spheres + idx;
  const float eps = cx->d->isect_eps;
  cx->c->spheres_tested++;
  const v3 center = ld(s->center);
  const float radius = s->radius;
  v3 o = ray->o, dd = ray->d;
  float inv16[16], nrm16[16];
  const float* nrm = NULL;
  int nstride = 3;
  const int moving = (s->flags & B200RT_XF_MOVING) != 0;
  const int ident = (s->flags & B200RT_XF_IDENTITY) != 0;
  if (moving) {
    moving_xf(s->fwd, s->velocity, ray->time, inv16, nrm16);
    o = xf_point(inv16, ray->o);
    dd = xf_vector(inv16, 4, ray->d);
    nrm = nrm16; nstride = 4;
  } else if (!ident) {
    o = xf_point(s->inv, ray->o);
    dd = xf_vector(s->inv, 4, ray->d);
    nrm = s->nrm; nstride = 3;
  }
  const v3 cto = sub(o, center);
  const float a = dot(dd, dd);
  const float b = 2 * dot(dd, cto);
  const float c = dot(cto, cto) - radius * radius;
  const float disc = b * b - 4 * a * c;
  float t;
  if (disc < -eps) return 0;
  if (disc < eps) {
    t = -b / (2 * a);
  } else {
    const float sq = sqrtf(disc);
    const float t1 = (-b + sq) / (2 * a);
    const float t2 = (-b - sq) / (2 * a);
    t = (t2 < .0) ? t1 : t2;
  }
  hit->t = t;
  hit->light = -1; hit->texture = -1; hit->u = -1; hit->v = -1; hit->perlin = -1;
  const v3 local_p = add(o, muls(dd, t));
  const v3 local_c = sub(local_p, center);
  v3 n = normalize(local_c);
  if (sem_hw4(cx->d)) {
    /* HW4/src/Sphere.cpp:31-100, HW7/src/Sphere.cpp:65-97: the normal matrix is ALWAYS applied (a second normalise even for identity) */
    hit->normal = n;
    if (s->texture != -1) {
      hit->texture = s->texture;
      tex_surface_sphere(cx->d, s, local_c, local_p, hit);
      n = hit->normal;
    }
    if (nrm) n = normalize(xf_vector(nrm, nstride, n));
    else n = normalize(xf_vector(s->nrm, 3, n));
    if (s->flags & B200RT_XF_LIGHT) hit->light = s->light;   /* Light_sphere::intersect */
  } else if (nrm) {
    n = normalize(xf_vector(nrm, nstride, n));
  }
  hit->normal = n;
  hit->tri = -1;
  hit->shape = B200RT_REF(B200RT_KIND_SPHERE, idx);
  hit->material = s->material;
  hit->top_rank = s->rank;
  hit->valid = 1;
  return 1;
}

/* Mesh_instance::intersect — HW3/include/Mesh.h:55-98, HW7/include/Mesh.h:65-115 (HW2: Mesh::intersect
 * HW2/Mesh.h:13-15; HW7 Light_mesh::intersect HW7/include/Light_mesh.h:52-68) */
static int instance_intersect(const Ctx* cx, int idx, const float* box, const Ray* ray, Hit* hit, int culling) {
  const B200SceneDesc* d = cx->d;
  const b200rt_instance* in = d->instances + idx;
  const b200rt_mesh* m = d->meshes + in->mesh;
  if (in->flags & B200RT_INST_DIRECT) {
    if (!shape_intersect(cx, m->root, m->root_box, ray, hit, culling)) return 0;
    hit->top_rank = in->rank;
    return 1;
  }
  if (!(in->flags & B200RT_INST_LIGHT)) {
    const float bt = box_intersect(box, box + 3, ray, d->isect_eps);
    if (bt < 0.0f || bt == INFINITY) return 0;
  }
  if (in->flags & B200RT_INST_REFRACTIVE) culling = 0;
  cx->c->instances_entered++;
  Ray local = *ray;
  float inv16[16], nrm16[16];
  const float* nrm = in->nrm;
  int nstride = 3;
  if (in->flags & B200RT_INST_MOVING) {
    moving_xf(in->fwd, in->velocity, ray->time, inv16, nrm16);
    local.o = xf_point(inv16, ray->o);
    local.d = xf_vector(inv16, 4, ray->d);
    nrm = nrm16; nstride = 4;
  } else {
    local.o = xf_point(in->inv, ray->o);
    local.d = xf_vector(in->inv, 4, ray->d);
  }
  if (m->root == B200RT_REF_NONE) return 0;
  if (!shape_intersect(cx, m->root, m->root_box, &local, hit, culling)) return 0;
  hit->normal = normalize(xf_vector(nrm, nstride, hit->normal));
  hit->shape = B200RT_REF(B200RT_KIND_INSTANCE, idx);
  hit->material = in->material;
  hit->texture = in->texture;
  hit->light = (in->flags & B200RT_INST_LIGHT) ? in->light : -1;
  hit->top_rank = in->rank;
  return 1;
}

/* BVH::intersect — HW3/src/Bounding_volume_hierarchy.cpp:31-55 */
static int bvh_intersect(const Ctx* cx, uint32_t node, const float* box, const Ray* ray, Hit* hit, int culling) {
  const b200rt_bvh_node* n = cx->d->nodes + node;
  const float bt = box_intersect(box, box + 3, ray, cx->d->isect_eps);
  if (bt < 0.0f || bt == INFINITY) return 0;
  cx->c->nodes_visited++; /* nodes whose gate passed = 64-byte node records a flattened traversal fetches */
  int any = 0;
  Hit lh; memset(&lh, 0, sizeof(lh)); lh.t = INFINITY; lh.top_rank = -1; lh.tri = -1;
  if (shape_intersect(cx, n->left, n->lmin, ray, &lh, culling) && lh.t > 0.0f && lh.t < hit->t) {
    *hit = lh; any = 1;
  }
  Hit rh; memset(&rh, 0, sizeof(rh)); rh.t = INFINITY; rh.top_rank = -1; rh.tri = -1;
  if (shape_intersect(cx, n->right, n->rmin, ray, &rh, culling) && rh.t > 0.0f && rh.t < hit->t) {
    *hit = rh; any = 1;
  }
  return any;
}

static int shape_intersect(const Ctx* cx, uint32_t ref, const float* box, const Ray* ray, Hit* hit, int culling) {
  const uint32_t idx = B200RT_REF_INDEX(ref);
  switch (B200RT_REF_KIND(ref)) {
    case B200RT_KIND_INNER: return bvh_intersect(cx, idx, box, ray, hit, culling);
    case B200RT_KIND_TRI: return mesh_tri_intersect(cx, (int)idx, ray, hit, culling);
    case B200RT_KIND_SPHERE: return sphere_intersect(cx, (int)idx, ray, hit);
    case B200RT_KIND_TRIPRIM: return triprim_intersect(cx, (int)idx, ray, hit, culling);
    case B200RT_KIND_LIGHTINST:
    case B200RT_KIND_INSTANCE: return instance_intersect(cx, (int)idx, box, ray, hit, culling);
    default: return 0;
  }
}

/* Scene-level closest hit: `hit_data.t = inf; bvh->intersect(ray, hit_data)`
 * (HW3/src/Scene.cpp:121-124) or the HW1 brute-force scan (HW1/Scene.cpp:33-55). */
static int scene_intersect(const Ctx* cx, const Ray* ray, Hit* hit) {
  const B200SceneDesc* d = cx->d;
  memset(hit, 0, sizeof(*hit));
  hit->t = INFINITY; hit->top_rank = -1; hit->tri = -1;
  if (d->semantics == B200RT_SEM_HW1) {
    const float eps = d->isect_eps;
    for (int i = 0; i < d->n_tris; i++) {
      float t, beta, gamma;
      cx->c->tris_tested++;
      if (!tri_core(d->tri_geom + 12 * (size_t)i, ray->o, ray->d, 1, eps, &t, &beta, &gamma)) continue;
      if (t < hit->t - eps && t > 0.0f) {
        hit->t = t; hit->tri = i; hit->top_rank = i; hit->valid = 1;
        hit->shape = B200RT_REF(B200RT_KIND_TRI, i);
        hit->material = d->tri_info[2 * (size_t)i];
        hit->normal = ld(d->tri_geom + 12 * (size_t)i + 9);
      }
    }
    for (int i = 0; i < d->n_spheres; i++) {
      Hit h; memset(&h, 0, sizeof(h)); h.t = INFINITY;
      if (!sphere_intersect(cx, i, ray, &h)) continue;
      if (h.t < hit->t - eps && h.t > 0.0f) {
        *hit = h; hit->top_rank = d->n_tris + i; hit->tri = -1;
      }
    }
    return hit->valid;
  }
  if (d->root == B200RT_REF_NONE) return 0;
  hit->light = -1; hit->texture = -1;
  /* HW7: `bvh->intersect(ray, hit_data, true)` — culling on for every ray type (HW7/src/Scene.cpp:170,286) */
  return shape_intersect(cx, d->root, d->root_box, ray, hit, 1);
}

/* ------------------------------------------------------------------------ */
/* Counter-based RNG shared (as a specification) with the CUDA core: the    */
/* reference seeds everything from the clock / random_device, so matched    */
/* sampling needs an injectable stream.                     */
/* ------------------------------------------------------------------------ */
static inline uint32_t rng_mix(uint32_t a, uint32_t b) {
  uint32_t h = a * 0x9E3779B1u + b;
  h ^= h >> 16; h *= 0x85EBCA6Bu; h ^= h >> 13; h *= 0xC2B2AE35u; h ^= h >> 16;
  return h;
}
static inline float rng_u01(uint32_t key, uint32_t dim) {
  return (float)(rng_mix(key, dim ^ 0xA511E9B3u) >> 8) * (1.0f / 16777216.0f);
}

typedef struct {
  uint32_t key;   /* stream of this path vertex */
  int pixel;
} PathState;

/* Scene::refract_ray — HW3/src/Scene.cpp:100-114 */
static int refract_ray(v3 du, v3 n, float index, v3* out) {
  const float n_ratio = 1 / index;
  const float cos_theta = dot(neg(du), n);
  const float delta = 1 - (n_ratio) * (n_ratio) * (1 - (cos_theta * cos_theta));
  if (delta < 0.0f) return 0;
  *out = normalize(sub(muls(add(du, muls(n, cos_theta)), n_ratio), muls(n, sqrtf(delta))));
  return 1;
}

/* Scene::trace_ray — HW1/Scene.cpp:28-123 and HW3/src/Scene.cpp:116-312. */
static int diffuse_constant7(const B200SceneDesc* d, const Hit* hit, v3* out);
/* ---- the reference's own random streams, for the bit-exact pin of the photon pass -------------------------------
 * PPM draws from three function-local `thread_local static std::mt19937 generator(rd())` engines through
 * `std::uniform_real_distribution<float>(0, 1)`: Point_light::generate_photon (PPM/src/Point_light.cpp:10-14),
 * Scene::sample_hemisphere (PPM/src/Scene.cpp:17-21), Scene::photon_trace (:107-109).  oracle/ref_driver_ppm.cpp
 * runs the tree single-threaded with random_device interposed to a constant; with mt mode on, the restatement draws
 * from the same three streams in the same program order instead of its counter-based generator.
 * mt19937: the standard recurrence.  uniform_real_distribution<float> in libstdc++ (bits/random.tcc,
 * generate_canonical<float, 24>): ONE 32-bit draw, float(draw) / 2^32, and nextafter(1, 0) if that rounds to 1. */
typedef struct { uint32_t x[624]; int i; } Mt19937;
static void mt_seed(Mt19937* m, uint32_t seed) {
  m->x[0] = seed;
  for (int k = 1; k < 624; k++) m->x[k] = 1812433253u * (m->x[k - 1] ^ (m->x[k - 1] >> 30)) + (uint32_t)k;
  m->i = 624;
}
static uint32_t mt_next(Mt19937* m) {
  if (m->i >= 624) {
    for (int k = 0; k < 624; k++) {
      const uint32_t y = (m->x[k] & 0x80000000u) | (m->x[(k + 1) % 624] & 0x7fffffffu);
      m->x[k] = m->x[(k + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
    }
    m->i = 0;
  }
  uint32_t y = m->x[m->i++];
  y ^= y >> 11; y ^= (y << 7) & 0x9d2c5680u; y ^= (y << 15) & 0xefc60000u; y ^= y >> 18;
  return y;
}
static float mt_u01(Mt19937* m) {
  float r = (float)mt_next(m) / 4294967296.0f;
  if (r >= 1.0f) r = nextafterf(1.0f, 0.0f);
  return r;
}
static struct { int on; Mt19937 light, hemi, trace, eye; float* state_out; int64_t state_cap; } g_mt;   /* test infrastructure: single-threaded use only */
#define PPM_DRAW(stream, key, dim) (g_mt.on ? mt_u01(&g_mt.stream) : rng_u01((key), (dim)))
/* state_out (optional): 5 floats per hit point {flux.rgb, radius_squared, n} written by the next oracle_ppm call */
/* HW7 (and HW3-HW6, whose sites are a subset): one engine per draw site, all seeded from random_device —
 * Scene::render_image (HW7/src/Scene.cpp:36-61: jitter x, y, [lens x, y], time), Area_light.cpp:18-22,
 * Light_mesh.cpp:7-12, Light_sphere.cpp:9-13, Spherical_directional_light.cpp:18-22, the glossy perturbation
 * (Scene.cpp:216-220) and the hemisphere sample of trace_path (Scene.cpp:347-351).  oracle/ref_driver_hdr.cpp with
 * REF_SINGLE_THREAD=1 runs the tree on one thread with random_device interposed. */
static struct { int on; Mt19937 render, area, mesh, sphere, env, glossy, hemi; } g_mt7;
#define DRAW7(stream, key, dim) (g_mt7.on ? mt_u01(&g_mt7.stream) : rng_u01((key), (dim)))
/* Clock-seeded engines (HW2-HW4 everywhere; the glossy engines of HW5/HW6): `engine.seed(system_clock::now()...)` on a
 * FRESH engine per pixel (jitter / lens / time: three engines with the same seed, HW3/src/Scene.cpp:40-53) or per call
 * (area light :184-189, glossy u and v :247-256).  The drivers interpose std::chrono::system_clock::now() too
 * (REF_SEED), which turns every such engine into "the first draws of engine(seed)".  HW2/HW3 use
 * default_random_engine = minstd_rand0, HW4-HW6 mt19937. */
typedef struct { int is_mt; uint64_t x; Mt19937 mt; } ClockEngine;
static struct { int on; uint32_t seed; } g_clk;
static void ce_seed(ClockEngine* e, int is_mt) {
  e->is_mt = is_mt;
  if (is_mt) { mt_seed(&e->mt, g_clk.seed); return; }
  e->x = g_clk.seed % 2147483647u;        /* linear_congruential_engine::seed: 0 becomes 1 */
  if (e->x == 0) e->x = 1;
}
static float ce_u01(ClockEngine* e) {
  if (e->is_mt) return mt_u01(&e->mt);
  /* minstd_rand0: x <- 16807 x mod (2^31 - 1), range [1, 2^31 - 2]; generate_canonical<float, 24>: one draw,
   * float(x - min) / float(range) with range = 2147483646 (rounds to 2^31 as a float) */
  e->x = (e->x * 16807u) % 2147483647u;
  float r = (float)(e->x - 1u) / (float)2147483646.0L;
  if (r >= 1.0f) r = nextafterf(1.0f, 0.0f);
  return r;
}
void oracle_clock_mode(int on, uint32_t seed) { g_clk.on = on; g_clk.seed = seed; }

void oracle_hw7_mt_mode(int on, uint32_t seed) {
  g_mt7.on = on;
  if (on) {
    mt_seed(&g_mt7.render, seed); mt_seed(&g_mt7.area, seed); mt_seed(&g_mt7.mesh, seed); mt_seed(&g_mt7.sphere, seed);
    mt_seed(&g_mt7.env, seed); mt_seed(&g_mt7.glossy, seed); mt_seed(&g_mt7.hemi, seed);
  }
}
void oracle_ppm_mt_mode(int on, uint32_t seed, float* state_out, int64_t state_cap) {
  g_mt.on = on; g_mt.state_out = on ? state_out : NULL; g_mt.state_cap = state_cap;
  if (on) { mt_seed(&g_mt.light, seed); mt_seed(&g_mt.hemi, seed); mt_seed(&g_mt.trace, seed); mt_seed(&g_mt.eye, seed); }
}

static v3 tex_color(const B200SceneDesc* d, const b200rt_texture* T, float u, float v);
static v3 trace_ray(const Ctx* cx, const Ray* ray, int depth, uint32_t key, int count_as) {
  const B200SceneDesc* d = cx->d;
  const int hw1 = d->semantics == B200RT_SEM_HW1;
  const int hw4 = d->semantics == B200RT_SEM_HW4;   /* HW4/src/Scene.cpp:129-357 */
  v3 color = V(0, 0, 0);
  Hit hit;
  if (count_as == 0) cx->c->rays_primary++; else cx->c->rays_secondary++;
  const int found = scene_intersect(cx, ray, &hit);
  if (!found) {
    if (hw4) {
      if (ray->primary) {
        if (d->background_texture >= 0) return tex_color(d, d->textures + d->background_texture, ray->bg_u, ray->bg_v);
        return ld(d->background);
      }
      return color;
    }
    if (d->max_depth == depth) return ld(d->background);
    return color;
  }
  const v3 p = add(ray->o, muls(ray->d, hit.t));
  const b200rt_material* m = d->materials + hit.material;
  const v3 w0 = normalize(sub(ray->o, p));
  const v3 n = hit.normal;
  const float seps = d->shadow_eps;
  v3 kd = ld(m->diffuse);
  const v3 ks = ld(m->specular);
  const v3 kd_mat = kd;   /* has_diffuse looks at the MATERIAL's diffuse, the term uses the textured one */
  if (hw4) {
    v3 dc;
    if (diffuse_constant7(d, &hit, &dc)) return dc;   /* replace_all: the texel, nothing else (HW4/src/Scene.cpp:157-170) */
    kd = dc;
  }
  if (hw1 || !ray->in_medium) {
    color = add(color, mulv(ld(m->ambient), ld(d->ambient)));
    int rdim = 0;
    for (int li = 0; li < d->n_lights; li++) {
      const b200rt_light* L = d->lights + li;
      const int has_d = !is_zero(kd_mat), has_s = !is_zero(ks);
      if (!hw1 && !has_d && !has_s) continue;
      v3 lpos = ld(L->position);
      if (L->kind == B200RT_LIGHT_AREA) {
        /* HW3/src/Scene.cpp:184-191; the clock-seeded engine is replaced by the shared stream */
        float e1, e2;
        if (g_clk.on) {   /* a fresh clock-seeded engine per area light per shading point */
          ClockEngine eng; ce_seed(&eng, sem_hw4(d));
          e1 = ce_u01(&eng); e2 = ce_u01(&eng);
        } else {
          e1 = rng_u01(key, 16 + 2 * rdim); e2 = rng_u01(key, 17 + 2 * rdim);
        }
        rdim++;
        lpos = add(add(lpos, muls(ld(L->edge1), e1)), muls(ld(L->edge2), e2));
      }
      const v3 lvec = sub(lpos, p);
      const v3 wi = normalize(lvec);
      const float ldist = len(lvec);
      Ray sr; memset(&sr, 0, sizeof(sr));
      sr.o = add(p, muls(wi, seps)); sr.d = wi; sr.time = ray->time; sr.is_shadow = 1;
      cx->c->rays_shadow++;
      Hit sh;
      scene_intersect(cx, &sr, &sh);
      if (hw1) {
        /* HW1 tests every primitive separately: any 0 < t < dist-eps (HW1/Scene.cpp:72-101).
         * scene_intersect's HW1 scan keeps a running best with the -eps rule, which is not the
         * same predicate, so redo it literally. */
        int in_shadow = 0;
        for (int i = 0; i < d->n_tris && !in_shadow; i++) {
          float t, be, ga;
          if (tri_core(d->tri_geom + 12 * (size_t)i, sr.o, sr.d, 1, d->isect_eps, &t, &be, &ga) &&
              t < (ldist - seps) && t > 0.0f)
            in_shadow = 1;
        }
        for (int i = 0; i < d->n_spheres && !in_shadow; i++) {
          Hit h; memset(&h, 0, sizeof(h)); h.t = INFINITY;
          OracleCounters dummy; memset(&dummy, 0, sizeof(dummy));
          Ctx c2 = {d, &dummy};
          if (sphere_intersect(&c2, i, &sr, &h) && h.t < (ldist - seps) && h.t > 0.0f) in_shadow = 1;
        }
        if (in_shadow) continue;
      } else {
        if (sh.t < (ldist - seps) && sh.t > 0.0f) continue;
      }
      const float d2 = ldist * ldist;
      v3 inten = ld(L->intensity);
      if (L->kind == B200RT_LIGHT_AREA) {
        const float icos = dot(neg(wi), ld(L->normal));  /* HW3/src/Scene.cpp:208 */
        if (has_d) {
          const float dc = dot(n, wi);
          color = add(color, divs(muls(muls(mulv(kd, inten), icos), dc), d2));
        }
        if (has_s) {
          const float sc = fmaxf(0.0f, dot(n, normalize(add(w0, wi))));
          color = add(color, divs(muls(muls(mulv(ks, inten), icos), (float)pow((double)sc, (double)m->phong_exponent)), d2));
        }
      } else {
        if (hw1 || has_d) {
          const float dc = dot(n, wi);
          color = add(color, divs(muls(mulv(kd, inten), dc), d2));
        }
        if (hw1 || has_s) {
          const float sc = fmaxf(0.0f, dot(n, normalize(add(w0, wi))));
          color = add(color, divs(muls(mulv(ks, inten), (float)pow((double)sc, (double)m->phong_exponent)), d2));
        }
      }
    }
  }
  const v3 km = ld(m->mirror);
  if (!is_zero(km) && depth > 0) {
    const v3 wr = normalize(sub(muls(n, 2 * dot(n, w0)), w0));
    v3 dir = wr;
    if (!hw1 && m->roughness != 0.0f) {
      /* glossy — HW3/src/Scene.cpp:236-260 */
      v3 rp;
      if (wr.x < wr.y && wr.x < wr.z) rp = V(1.0f, wr.y, wr.z);
      else if (wr.y < wr.z && wr.y < wr.x) rp = V(wr.x, 1.0f, wr.z);
      else rp = V(wr.x, wr.y, 1.0f);
      const v3 u = normalize(cross(rp, wr));
      const v3 v = normalize(cross(u, wr));
      float eu = DRAW7(glossy, key, 2) - 0.5f;
    float ev = DRAW7(glossy, key, 3) - 0.5f;
    if (g_clk.on && !sem_hw7(d)) {   /* fresh clock-seeded engines for u and for v: the same first draw twice */
      ClockEngine gu; ce_seed(&gu, sem_hw4(d)); eu = ce_u01(&gu) - 0.5f;
      ClockEngine gv; ce_seed(&gv, sem_hw4(d)); ev = ce_u01(&gv) - 0.5f;
    }
      dir = normalize(add(wr, muls(add(muls(u, eu), muls(v, ev)), m->roughness)));
    }
    Ray mr; memset(&mr, 0, sizeof(mr));
    mr.o = add(p, muls(dir, seps)); mr.d = dir; mr.time = ray->time;
    color = add(color, mulv(km, trace_ray(cx, &mr, depth - 1, rng_mix(key, 0x101u), 1)));
  }
  const v3 kt = ld(m->transparency);
  if (!hw1 && !is_zero(kt) && depth > 0) {
    /* dielectric — HW3/src/Scene.cpp:264-310 (HW2 flips the Beer sign, HW2/Scene.cpp:166-168) */
    const v3 wr = normalize(sub(muls(n, 2 * dot(n, w0)), w0));
    v3 tdir = V(0, 0, 0);
    float cos_theta = 0.0f;
    v3 k = V(0, 0, 0);
    const v3 dn = normalize(ray->d);
    const float ni = m->refraction_index;
    int tir = 0, entering;
    if (dot(dn, n) < 0.0f) {
      refract_ray(dn, n, ni, &tdir);
      cos_theta = dot(neg(dn), n);
      k = V(1.0f, 1.0f, 1.0f);
      entering = 1;
    } else {
      const float ht = hit.t;
      /* unqualified exp/log resolve to the DOUBLE overloads in the reference TU */
      if (d->semantics == B200RT_SEM_HW2) {
        k = V((float)exp(-log((double)kt.x) * (double)ht), (float)exp(-log((double)kt.y) * (double)ht),
              (float)exp(-log((double)kt.z) * (double)ht));
      } else {
        k = V((float)exp(log((double)kt.x) * (double)ht), (float)exp(log((double)kt.y) * (double)ht),
              (float)exp(log((double)kt.z) * (double)ht));
      }
      entering = 0;
      if (refract_ray(dn, neg(n), 1.0f / ni, &tdir)) cos_theta = dot(tdir, n);
      else tir = 1;
    }
    Ray rr; memset(&rr, 0, sizeof(rr));
    rr.o = add(p, muls(wr, seps)); rr.d = wr; rr.time = ray->time;
    if (tir) {
      rr.in_medium = 1;
      color = add(color, mulv(k, trace_ray(cx, &rr, depth - 1, rng_mix(key, 0x102u), 1)));
    } else {
      const float r0 = ((ni - 1) * (ni - 1)) / ((ni + 1) * (ni + 1));
      /* `pow(float,int)` is the double overload: the whole sum is evaluated in double */
      const float r = (float)((double)r0 + (double)(1 - r0) * pow((double)(1.0f - cos_theta), 5.0));
      rr.in_medium = !entering;
      Ray tr; memset(&tr, 0, sizeof(tr));
      tr.o = add(p, muls(tdir, seps)); tr.d = tdir; tr.time = ray->time; tr.in_medium = entering;
      const v3 cr = trace_ray(cx, &rr, depth - 1, rng_mix(key, 0x102u), 1);
      const v3 ct = trace_ray(cx, &tr, depth - 1, rng_mix(key, 0x103u), 1);
      color = add(color, mulv(k, add(muls(cr, r), muls(ct, 1 - r))));
    }
  }
  return color;
}

/* ======================================================================== */
/* HW4..HW7: textures, Perlin noise, BRDFs, light sources, path tracing       */
/* ======================================================================== */
static inline float texel(const B200SceneDesc* d, const b200rt_texture* T, long i) {
  const long n = (long)T->width * T->height * 3;
  if (i < 0 || i >= n) return 0.0f; /* the reference reads past the image here (UB); defined as 0 */
  return d->texels[T->texel_offset + i];
}
/* Texture::get_color_at — HW7/src/Texture.cpp:66-143 */
static v3 tex_color(const B200SceneDesc* d, const b200rt_texture* T, float u, float v) {
  const int W = T->width, H = T->height;
  if (T->appearance == B200RT_TEX_CLAMP) {
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
  v3 c;
  if (T->interp == B200RT_TEX_NEAREST) {
    const long coord = 3L * W * (unsigned int)v + 3L * (unsigned int)u;
    return V(texel(d, T, coord), texel(d, T, coord + 1), texel(d, T, coord + 2));
  }
  if (T->appearance == B200RT_TEX_REPEAT) {
    const unsigned int p = ((unsigned int)u) % (unsigned)W, pn = (p + 1) % (unsigned)W;
    const unsigned int q = ((unsigned int)v) % (unsigned)H, qn = (q + 1) % (unsigned)H;
    const float dx = u - p, dy = v - q;
    const long c_p_q = 3L * (W * q + p), c_pn_q = 3L * (W * q + pn), c_p_qn = 3L * (W * qn + p),
               c_pn_qn = 3L * (W * qn + pn);
    float* cc = &c.x;
    for (int k = 0; k < 3; k++) {
      float x = texel(d, T, c_p_q + k) * (1 - dx) * (1 - dy);
      x += texel(d, T, c_pn_q + k) * (dx) * (1 - dy);
      x += texel(d, T, c_pn_qn + k) * (dx) * (dy);
      x += texel(d, T, c_p_qn + k) * (1 - dx) * (dy);
      if (k == 0) c.x = x; else if (k == 1) c.y = x; else c.z = x;
    }
    (void)cc;
    return c;
  }
  {
    const unsigned int p = (unsigned int)u, pn = p + 1;
    const unsigned int q = (unsigned int)v, qn = (unsigned int)(v + 1);
    const float dx = u - p, dy = v - q;
    const long c_p_q = 3L * ((long)W * q + p), c_pn_q = 3L * ((long)W * q + pn), c_p_qn = 3L * ((long)W * qn + p),
               c_pn_qn = 3L * ((long)W * qn + pn);
    for (int k = 0; k < 3; k++) {
      float x = texel(d, T, c_p_q + k) * (1 - dx) * (1 - dy);
      if (pn < (unsigned)W) x += texel(d, T, c_pn_q + k) * (dx) * (1 - dy);
      if (qn < (unsigned)H) x += texel(d, T, c_pn_qn + k) * (dx) * (dy);       /* crossed guards: Appendix B.4 */
      if (pn < (unsigned)W && qn < (unsigned)H) x += texel(d, T, c_p_qn + k) * (1 - dx) * (dy);
      if (k == 0) c.x = x; else if (k == 1) c.y = x; else c.z = x;
    }
    return c;
  }
}
/* Texture::blend_color — HW7/src/Texture.cpp:53-65 */
static v3 tex_blend(const b200rt_texture* T, v3 tc, v3 kd) {
  if (T->decal == B200RT_TEX_REPLACE_KD) return tc;
  if (T->decal == B200RT_TEX_BLEND_KD) return divs(add(tc, kd), 2.0f);
  return V(0, 0, 0);
}
/* Perlin_noise — HW7/src/Perlin_noise.cpp:23-62 */
static const float kPerlinG[16][3] = {{1, 1, 0}, {-1, 1, 0}, {1, -1, 0}, {-1, -1, 0}, {1, 0, 1}, {-1, 0, 1},
                                      {1, 0, -1}, {-1, 0, -1}, {0, 1, 1}, {0, -1, 1}, {0, 1, -1}, {0, -1, -1},
                                      {1, 1, 0}, {-1, 1, 0}, {0, -1, 1}, {0, -1, -1}};
static inline float perlin_weight(float x) {
  const double a = fabs((double)x);
  return (float)(-6 * pow(a, 5) + 15 * pow(a, 4) + -10 * pow(a, 3) + 1);
}
static float perlin_value(const b200rt_texture* T, v3 pos) {
  const v3 sp = muls(pos, T->scaling);
  float value = 0.0f;
  const int x = (int)floorf(sp.x), y = (int)floorf(sp.y), z = (int)floorf(sp.z);
  for (int i = x; i <= x + 1; i++)
    for (int j = y; j <= y + 1; j++)
      for (int k = z; k <= z + 1; k++) {
        const v3 vv = V(sp.x - i, sp.y - j, sp.z - k);
        int a = i % 16, b = j % 16, c = k % 16;
        if (a < 0) a += 16;
        if (b < 0) b += 16;
        if (c < 0) c += 16;
        const float* g = kPerlinG[T->perm[(a + T->perm[(b + T->perm[c % 16]) % 16]) % 16]];
        const v3 e = V(g[0], g[1], g[2]);
        const float wx = perlin_weight(vv.x), wy = perlin_weight(vv.y), wz = perlin_weight(vv.z);
        value += dot(e, vv) * wx * wy * wz;
      }
  if (T->appearance == B200RT_TEX_VEIN) return fabsf(value);
  return (value + 1.0f) / 2.0f;
}
/* Texture::bump_normal (Perlin) — HW7/src/Texture.cpp:145-165 */
static v3 bump_perlin(const b200rt_texture* T, v3 normal, v3 pos) {
  const float pv = perlin_value(T, pos);
  const float epsilon = 0.001;
  const float dx = (perlin_value(T, add(pos, V(epsilon, 0, 0))) - pv) / epsilon;
  const float dy = (perlin_value(T, add(pos, V(0, epsilon, 0))) - pv) / epsilon;
  const float dz = (perlin_value(T, add(pos, V(0, 0, epsilon))) - pv) / epsilon;
  const v3 g = muls(V(dx, dy, dz), T->bump_multiplier);
  const v3 g_ii = muls(normal, dot(g, normal));
  return sub(normal, sub(g, g_ii));
}
/* Texture::bump_normal (image) — HW7/src/Texture.cpp:167-181 */
static v3 bump_image(const B200SceneDesc* d, const b200rt_texture* T, v3 dp_du, v3 dp_dv, v3 normal, float u, float v) {
  const v3 c0 = tex_color(d, T, u, v);
  const v3 cu = tex_color(d, T, u + 1.0f / T->width, v);
  const v3 cv = tex_color(d, T, u, v + 1.0f / T->height);
  const float v0 = (c0.x + c0.y + c0.z) / 3.0f, vu = (cu.x + cu.y + cu.z) / 3.0f, vv = (cv.x + cv.y + cv.z) / 3.0f;
  const float dd_du = (vu - v0) * T->bump_multiplier, dd_dv = (vv - v0) * T->bump_multiplier;
  const v3 dq_du = add(dp_du, muls(normal, dd_du));
  const v3 dq_dv = add(dp_dv, muls(normal, dd_dv));
  return normalize(cross(dq_dv, dq_du));
}
/* texture part of Mesh_triangle::intersect — HW7/src/Mesh_triangle.cpp:107-149 */
static void tex_surface_tri(const B200SceneDesc* d, int tri, float beta, float gamma, v3 local_p, Hit* hit) {
  const b200rt_texture* T = d->textures + d->tri_tex[tri];
  if (T->kind == B200RT_TEX_PERLIN) {
    hit->perlin = perlin_value(T, local_p);
    if (T->is_bump) hit->normal = bump_perlin(T, hit->normal, local_p);
    return;
  }
  const float* uv = d->tri_uv + 6 * (size_t)tri;
  const float u = uv[0] + beta * (uv[2] - uv[0]) + gamma * (uv[4] - uv[0]);
  const float v = uv[1] + beta * (uv[3] - uv[1]) + gamma * (uv[5] - uv[1]);
  hit->u = u; hit->v = v;
  if (T->is_bump) {
    const float ub_ua = uv[2] - uv[0], uc_ua = uv[4] - uv[0], vb_va = uv[3] - uv[1], vc_va = uv[5] - uv[1];
    const float* g = d->tri_geom + 12 * (size_t)tri;
    /* p1 - p0 and p2 - p0 from the vertex positions themselves */
    const int32_t* vi = d->tri_vidx + 3 * (size_t)tri;
    const v3 p0 = ld(d->vert_pos + 3 * (size_t)vi[0]), p1 = ld(d->vert_pos + 3 * (size_t)vi[1]),
             p2 = ld(d->vert_pos + 3 * (size_t)vi[2]);
    (void)g;
    const v3 pb_pa = sub(p1, p0), pc_pa = sub(p2, p0);
    float ic = (ub_ua * vc_va) - (vb_va * uc_ua);
    if (ic == 0.0f) ic = 0.000001f;
    ic = 1.0f / ic;
    const v3 dp_du = muls(sub(muls(pb_pa, vc_va), muls(pc_pa, vb_va)), ic);
    const v3 dp_dv = muls(sub(muls(pc_pa, ub_ua), muls(pb_pa, uc_ua)), ic);
    hit->normal = bump_image(d, T, dp_du, dp_dv, hit->normal, u, v);
  }
}
/* texture part of Sphere::intersect + get_uv — HW7/src/Sphere.cpp:70-112 */
static void tex_surface_sphere(const B200SceneDesc* d, const b200rt_sphere* s, v3 lc, v3 local_p, Hit* hit) {
  const b200rt_texture* T = d->textures + s->texture;
  /* HW7/src/Sphere.cpp:107-108 clamps; HW4-HW6 call the unqualified (double) acos on the raw ratio (HW4/src/Sphere.cpp:101, HW6/src/Sphere.cpp:106) */
  const float theta = sem_hw7(d) ? acosf(fmaxf(-1.0f, fminf(1.0f, lc.y / s->radius))) : (float)acos((double)(lc.y / s->radius));
  const float phi = (float)atan2((double)lc.z, (double)lc.x); /* unqualified atan2: the double function */
  const float u = (float)((M_PI - (double)phi) / (2 * M_PI));
  const float v = (float)((double)theta / M_PI);
  hit->u = u; hit->v = v;
  if (T->kind == B200RT_TEX_PERLIN) {
    hit->perlin = perlin_value(T, local_p);
    if (T->is_bump) hit->normal = bump_perlin(T, hit->normal, local_p);
  } else if (T->is_bump) {
    const float th = (float)(M_PI * (double)v);
    const float ph = (float)(M_PI * (double)(1 - 2 * u));
    const v3 dp_du = V((float)((double)(lc.z * 2) * M_PI), 0, (float)((double)(lc.x * -2) * M_PI));
    const v3 dp_dv = V((float)((double)lc.y * M_PI * cos((double)ph)), (float)((double)(-s->radius) * M_PI * sin((double)th)),
                       (float)((double)lc.y * M_PI * sin((double)ph)));
    hit->normal = bump_image(d, T, dp_du, dp_dv, hit->normal, u, v);
  }
}

/* BRDF::get_reflectance ×5 — HW7/src/*_BRDF.cpp */
static v3 brdf_eval(const b200rt_brdf* B, v3 normal, v3 kd, v3 ks, v3 w_i, v3 w_o) {
  const float cti = dot(normal, w_i);
  if (cti > 1.0f || cti <= 0.0f) return V(0, 0, 0);
  const float e = B->exponent;
  switch (B->kind) {
    case B200RT_BRDF_PHONG: {
      const v3 rwi = sub(muls(normal, 2 * cti), w_i);
      const float car = fmaxf(0.0f, dot(w_o, rwi));
      return add(divs(muls(ks, powf(car, e)), cti), kd);
    }
    case B200RT_BRDF_MODIFIED_PHONG: {
      const v3 rwi = sub(muls(normal, 2 * cti), w_i);
      const float car = fmaxf(0.0f, dot(w_o, rwi));
      v3 k_d = kd, k_s = muls(ks, powf(car, e));
      if (B->normalized) {
        /* Vector3::operator/(float): the double divisor is converted to float first */
        k_s = divs(muls(k_s, e + 2.0f), (float)(2.0f * M_PI));
        k_d = divs(k_d, (float)M_PI);
      }
      return add(k_s, k_d);
    }
    case B200RT_BRDF_BLINN_PHONG: {
      const v3 wh = normalize(add(w_i, w_o));
      const float cah = dot(wh, normal);
      return add(divs(muls(ks, powf(cah, e)), cti), kd);
    }
    case B200RT_BRDF_MODIFIED_BLINN_PHONG: {
      const v3 wh = normalize(add(w_i, w_o));
      const float cah = dot(wh, normal);
      v3 k_d = kd, k_s = muls(ks, powf(cah, e));
      if (B->normalized) {
        k_d = divs(k_d, (float)M_PI);
        k_s = divs(muls(k_s, e + 8.0f), (float)(8.0f * M_PI));
      }
      return add(k_s, k_d);
    }
    default: { /* Torrance–Sparrow */
      const v3 wh = normalize(add(w_i, w_o));
      const float ca = dot(wh, normal);
      const float d_alpha = (float)((double)((e + 2) * powf(ca, e)) / (2.0f * M_PI));
      const float cto = dot(w_o, normal);
      const float cb = dot(w_o, wh);
      float g = fminf(2 * ca * cto / cb, 2 * ca * cti / cb);
      g = fminf(1.0f, g);
      const float ri = B->refractive_index;
      const float r0 = (float)(pow((double)(ri - 1.0f), 2) / pow((double)(ri + 1.0f), 2));
      const float f_beta = (float)((double)r0 + (double)(1 - r0) * pow((double)(1 - cb), 5));
      const v3 k_d = divs(kd, (float)M_PI);
      const v3 k_s = divs(muls(muls(muls(ks, d_alpha), f_beta), g), 4 * cti * cto);
      return add(k_d, k_s);
    }
  }
}

/* ONB sample used by trace_path, the env light and the sphere light: w*cos(theta) + v*sin*cos(phi) + u*sin*sin(phi) */
static v3 onb_dir(v3 w, float theta, float phi) {
  const v3 u = normalize((w.x != 0.0f || w.y != 0.0f) ? V(-w.y, w.x, 0.0f) : V(0.0f, 1.0f, 0.0f));
  const v3 v = cross(w, u);
  return normalize(add(add(muls(w, cosf(theta)), muls(muls(v, sinf(theta)), cosf(phi))), muls(muls(u, sinf(theta)), sinf(phi))));
}

static int sphere_intersect(const Ctx* cx, int idx, const Ray* ray, Hit* hit);

/* Light::direction_and_distance — HW7/src/{Point,Area,Spot,Directional,Spherical_directional}_light.cpp,
 * Light_mesh.cpp:3-49, Light_sphere.cpp:5-55.  RNG: dims 32+4*li+{0,1,2} of the vertex key. */
static v3 light_dir(const Ctx* cx, int li, v3 p, v3 normal, uint32_t key, float* distance, float* prob) {
  const B200SceneDesc* d = cx->d;
  const b200rt_light* L = d->lights + li;
  const uint32_t dim = 32u + 4u * (uint32_t)li;
  *prob = 0.0f;
  switch (L->kind) {
    case B200RT_LIGHT_AREA: {
      const float e1 = DRAW7(area, key, dim);
      const float e2 = DRAW7(area, key, dim + 1);
      const v3 pos = add(add(ld(L->position), muls(ld(L->edge1), e1)), muls(ld(L->edge2), e2));
      const v3 dir = sub(pos, p);
      *distance = len(dir);
      return dir;
    }
    case B200RT_LIGHT_DIRECTIONAL:
      *distance = 3.402823466e+38f;
      return neg(ld(L->edge1));
    case B200RT_LIGHT_ENV: {
      const float e1 = DRAW7(env, key, dim);
      const float e2 = DRAW7(env, key, dim + 1);
      const float phi = (float)(2 * M_PI * (double)e1);
      const float theta = acosf(e2);
      *distance = 3.402823466e+38f;
      *prob = (float)(1.0f / (2.0f * M_PI));
      return onb_dir(normal, theta, phi);
    }
    case B200RT_LIGHT_MESH: {
      const float e0 = DRAW7(mesh, key, dim);
      const float e1 = sqrtf(DRAW7(mesh, key, dim + 1));
      const float e2 = DRAW7(mesh, key, dim + 2);
      int tri = -1;
      for (int i = 0; i < L->cdf_count; i++)
        if (e0 <= d->light_cdf[L->cdf_begin + i]) { tri = d->light_tri[L->cdf_begin + i]; break; }
      if (tri < 0) tri = d->light_tri[L->cdf_begin + L->cdf_count - 1]; /* reference dereferences NULL here */
      const int32_t* vi = d->tri_vidx + 3 * (size_t)tri;
      const v3 p0 = ld(d->vert_pos + 3 * (size_t)vi[0]), p1 = ld(d->vert_pos + 3 * (size_t)vi[1]),
               p2 = ld(d->vert_pos + 3 * (size_t)vi[2]);
      const v3 q = add(muls(p1, 1.0f - e2), muls(p2, e2));
      const v3 po = add(muls(p0, 1.0f - e1), muls(q, e1));
      const b200rt_instance* in = d->instances + L->object;
      const v3 pw = xf_point(in->fwd, po);
      const v3 dir = sub(pw, p);
      *distance = len(dir);
      const float cti = fmaxf(0.001f, -dot(normalize(dir), ld(d->tri_geom + 12 * (size_t)tri + 9)));
      *prob = *distance * *distance / (L->total_area * cti);
      return dir;
    }
    case B200RT_LIGHT_SPHERE: {
      const float e1 = DRAW7(sphere, key, dim);
      const float e2 = DRAW7(sphere, key, dim + 1);
      const b200rt_sphere* s = d->spheres + L->object;
      const v3 ps = xf_point(s->inv, p);
      v3 w = sub(ld(s->center), ps);
      const float dd = len(w);
      w = normalize(w);
      const float stm = fmaxf(-1.0f, fminf(1.0f, s->radius / dd));
      const float tmax = asinf(stm);
      const float ctm = cosf(tmax);
      const float phi = (float)(2 * M_PI * (double)e1);
      const float theta = acosf(fmaxf(-1.0f, fminf(1.0f, 1 - e2 + e2 * ctm)));
      const v3 lo = onb_dir(w, theta, phi);
      const v3 lw = normalize(xf_vector(s->fwd, 4, lo));
      Ray r; memset(&r, 0, sizeof(r));
      r.o = add(p, muls(lw, d->shadow_eps)); r.d = lw;
      Hit h; memset(&h, 0, sizeof(h)); h.t = INFINITY;
      OracleCounters dummy; memset(&dummy, 0, sizeof(dummy));
      Ctx c2 = {d, &dummy};
      sphere_intersect(&c2, L->object, &r, &h);
      *distance = h.t;
      *prob = (float)(1 / (2 * M_PI * (double)(1 - ctm)));
      return lw;
    }
    default: { /* point, spot */
      const v3 dir = sub(ld(L->position), p);
      *distance = len(dir);
      return dir;
    }
  }
}
/* Light::incoming_radiance */
static v3 light_radiance(const B200SceneDesc* d, int li, v3 to_light, float prob) {
  const b200rt_light* L = d->lights + li;
  const float x = to_light.x, y = to_light.y, z = to_light.z;
  switch (L->kind) {
    case B200RT_LIGHT_POINT: return divs(ld(L->intensity), x * x + y * y + z * z);
    case B200RT_LIGHT_AREA: {
      const v3 rw = neg(normalize(to_light));
      return divs(muls(ld(L->intensity), dot(rw, ld(L->normal))), x * x + y * y + z * z);
    }
    case B200RT_LIGHT_SPOT: {
      const v3 rw = neg(normalize(to_light));
      const float ct = dot(rw, ld(L->edge1));
      const float ccov = L->edge2[0], cfall = L->edge2[1];
      if (ct > cfall) return divs(ld(L->intensity), x * x + y * y + z * z);
      if (ct < ccov) return V(0, 0, 0);
      const float c = (ct - ccov) / (cfall - ccov);
      return divs(muls(ld(L->intensity), (float)pow((double)c, 4)), x * x + y * y + z * z);
    }
    case B200RT_LIGHT_DIRECTIONAL: return ld(L->intensity);
    case B200RT_LIGHT_ENV: {
      const b200rt_texture* T = d->textures + L->object;
      const float theta = acosf(fminf(1.0f, fmaxf(-1.0f, y))); /* clamp(-1,1,cos) with the swapped-argument helper */
      const float phi = atan2f(z, x);
      float u = (float)((M_PI - (double)phi) / (2.0f * M_PI));
      float v = (float)((double)theta / M_PI);
      u = fmaxf(0.0f, fminf(1.0f, u));
      v = fmaxf(0.0f, fminf(1.0f, v));
      u *= T->width;
      if (u >= T->width) u--;
      v *= T->height;
      if (v >= T->height) v--;
      const long coord = 3L * T->width * (unsigned int)v + 3L * (unsigned int)u;
      return divs(V(texel(d, T, coord), texel(d, T, coord + 1), texel(d, T, coord + 2)), prob);
    }
    default: return divs(ld(L->intensity), prob); /* mesh, sphere */
  }
}

static v3 send_ray7(const Ctx* cx, const Ray* ray, int level, uint32_t key, int count_as);

/* Scene::calculate_diffuse_constant — HW7/src/Scene.cpp:231-258 */
static int diffuse_constant7(const B200SceneDesc* d, const Hit* hit, v3* out) {
  const b200rt_material* m = d->materials + hit->material;
  *out = ld(m->diffuse);
  if (hit->texture < 0) return 0;
  const b200rt_texture* T = d->textures + hit->texture;
  if (T->decal == B200RT_TEX_REPLACE_ALL) {
    *out = (T->kind == B200RT_TEX_PERLIN) ? V(0, 0, 0) : tex_color(d, T, hit->u, hit->v);
    return 1;
  }
  if (T->kind == B200RT_TEX_PERLIN) *out = tex_blend(T, V(hit->perlin, hit->perlin, hit->perlin), *out);
  else *out = tex_blend(T, divs(tex_color(d, T, hit->u, hit->v), T->normalizer), *out);
  return 0;
}
/* Scene::calculate_diffuse_and_specular_radiance — HW7/src/Scene.cpp:260-311 */
static v3 direct7(const Ctx* cx, const Ray* ray, const Hit* hit, v3 kd, uint32_t key, v3 start) {
  const B200SceneDesc* d = cx->d;
  v3 radiance = start;   /* HW7 sums the lights on their own (0); HW5/HW6 add them straight onto the colour */
  const v3 p = add(ray->o, muls(ray->d, hit->t));
  const b200rt_material* m = d->materials + hit->material;
  const v3 w_o = normalize(sub(ray->o, p));
  const v3 n = hit->normal;
  const b200rt_brdf* B = (m->brdf != -1) ? d->brdfs + m->brdf : NULL;
  for (int li = 0; li < d->n_lights; li++) {
    float ldist, prob;
    const v3 lvec = light_dir(cx, li, p, n, key, &ldist, &prob);
    const v3 w_i = normalize(lvec);
    if (!sem_hw7(d) || dot(n, w_i) >= 0.0f) {   /* the gate is HW7's (HW7/src/Scene.cpp:281); HW6 has none */
      Ray sr; memset(&sr, 0, sizeof(sr));
      sr.o = add(p, muls(w_i, d->shadow_eps)); sr.d = w_i; sr.time = ray->time; sr.is_shadow = 1;
      cx->c->rays_shadow++;
      Hit sh;
      scene_intersect(cx, &sr, &sh);
      if (sh.t < (ldist - d->shadow_eps) && sh.t > 0.0f) continue;
      if (g_lit_count && tl_pixel >= 0) __atomic_fetch_add(&g_lit_count[(size_t)tl_pixel * d->n_lights + li], 1, __ATOMIC_RELAXED);
      const v3 Li = light_radiance(d, li, lvec, prob);
      const float cti = fmaxf(0.0f, dot(n, w_i));
      if (B) {
        radiance = add(radiance, mulv(muls(Li, cti), brdf_eval(B, n, kd, ld(m->specular), w_i, w_o)));
      } else {
        radiance = add(radiance, muls(mulv(kd, Li), cti));
        const float sct = fmaxf(0.0f, dot(n, normalize(add(w_o, w_i))));
        radiance = add(radiance, muls(mulv(ld(m->specular), Li), (float)pow((double)sct, (double)m->phong_exponent)));
      }
    }
  }
  return radiance;
}
/* Scene::reflect_ray — HW7/src/Scene.cpp:191-229 */
static v3 reflect7(const Ctx* cx, const Ray* ray, const Hit* hit, int level, uint32_t key) {
  const B200SceneDesc* d = cx->d;
  const b200rt_material* m = d->materials + hit->material;
  const v3 n = hit->normal;
  const v3 p = add(ray->o, muls(ray->d, hit->t));
  const v3 w_o = normalize(sub(ray->o, p));
  const v3 wr = normalize(sub(muls(n, 2 * dot(n, w_o)), w_o));
  Ray mr; memset(&mr, 0, sizeof(mr));
  mr.time = ray->time;
  if (m->roughness == 0.0f) {
    mr.o = add(p, muls(wr, d->shadow_eps)); mr.d = wr;   /* light_hit NOT propagated here */
  } else {
    v3 rp;
    if (wr.x < wr.y && wr.x < wr.z) rp = V(1.0f, wr.y, wr.z);
    else if (wr.y < wr.z && wr.y < wr.x) rp = V(wr.x, 1.0f, wr.z);
    else rp = V(wr.x, wr.y, 1.0f);
    const v3 u = normalize(cross(rp, wr));
    const v3 v = normalize(cross(u, wr));
    float eu = DRAW7(glossy, key, 2) - 0.5f;
    float ev = DRAW7(glossy, key, 3) - 0.5f;
    if (g_clk.on && !sem_hw7(d)) {   /* fresh clock-seeded engines for u and for v: the same first draw twice */
      ClockEngine gu; ce_seed(&gu, sem_hw4(d)); eu = ce_u01(&gu) - 0.5f;
      ClockEngine gv; ce_seed(&gv, sem_hw4(d)); ev = ce_u01(&gv) - 0.5f;
    }
    const v3 wrp = normalize(add(wr, muls(add(muls(u, eu), muls(v, ev)), m->roughness)));
    mr.o = add(p, muls(wrp, d->shadow_eps)); mr.d = wrp;
    mr.light_hit = ray->light_hit;
  }
  return send_ray7(cx, &mr, level + 1, rng_mix(key, 0x101u), 1);
}
/* Scene::refract_ray — HW7/src/Scene.cpp:108-166 */
static v3 refract7(const Ctx* cx, const Ray* ray, const Hit* hit, int level, uint32_t key) {
  const B200SceneDesc* d = cx->d;
  const b200rt_material* m = d->materials + hit->material;
  const v3 n = hit->normal;
  const v3 p = add(ray->o, muls(ray->d, hit->t));
  const v3 w_o = normalize(sub(ray->o, p));
  const v3 wr = normalize(sub(muls(n, 2 * dot(n, w_o)), w_o));
  v3 tdir = V(0, 0, 0), k = V(0, 0, 0);
  float cos_theta = 0.0f;
  const v3 dn = normalize(ray->d);
  const float ni = m->refraction_index;
  int tir = 0, entering;
  if (dot(dn, n) < 0.0f) {
    refract_ray(dn, n, ni, &tdir);
    cos_theta = dot(neg(dn), n);
    k = V(1.0f, 1.0f, 1.0f);
    entering = 1;
  } else {
    const v3 kt = ld(m->transparency);
    const float ht = hit->t;
    k = V((float)exp(log((double)kt.x) * (double)ht), (float)exp(log((double)kt.y) * (double)ht),
          (float)exp(log((double)kt.z) * (double)ht));
    entering = 0;
    if (refract_ray(dn, neg(n), 1.0f / ni, &tdir)) cos_theta = dot(tdir, n);
    else tir = 1;
  }
  Ray rr; memset(&rr, 0, sizeof(rr));
  rr.o = add(p, muls(wr, d->shadow_eps)); rr.d = wr; rr.time = ray->time; rr.light_hit = ray->light_hit;
  if (tir) {
    rr.in_medium = 1;
    return mulv(k, send_ray7(cx, &rr, level + 1, rng_mix(key, 0x102u), 1));
  }
  const float r0 = ((ni - 1) * (ni - 1)) / ((ni + 1) * (ni + 1));
  const float r = (float)((double)r0 + (double)(1 - r0) * pow((double)(1.0f - cos_theta), 5.0));
  rr.in_medium = !entering;
  Ray tr; memset(&tr, 0, sizeof(tr));
  tr.o = add(p, muls(tdir, d->shadow_eps)); tr.d = tdir; tr.time = ray->time; tr.in_medium = entering;
  tr.light_hit = ray->light_hit;
  /* `r * send_ray(reflection) + (1 - r) * send_ray(transmission)`: the order of the two calls is unspecified in C++;
   * the reference as built here (g++ 13 -O3) traces the TRANSMISSION first — visible only when the draws are replayed */
  v3 cr, ct;
  if (g_mt7.on) {
    ct = send_ray7(cx, &tr, level + 1, rng_mix(key, 0x103u), 1);
    cr = send_ray7(cx, &rr, level + 1, rng_mix(key, 0x102u), 1);
  } else {
    cr = send_ray7(cx, &rr, level + 1, rng_mix(key, 0x102u), 1);
    ct = send_ray7(cx, &tr, level + 1, rng_mix(key, 0x103u), 1);
  }
  return mulv(k, add(muls(cr, r), muls(ct, 1 - r)));
}
/* Scene::trace_ray — HW7/src/Scene.cpp:415-447 */
static v3 trace_ray7(const Ctx* cx, const Ray* ray, const Hit* hit, int level, uint32_t key) {
  const B200SceneDesc* d = cx->d;
  v3 radiance = V(0, 0, 0);
  if (hit->light >= 0) return ld(d->lights[hit->light].intensity);
  const b200rt_material* m = d->materials + hit->material;
  v3 kd;
  if (diffuse_constant7(d, hit, &kd)) return kd;
  if (!ray->in_medium) {
    radiance = add(radiance, mulv(ld(m->ambient), ld(d->ambient)));
    if (sem_hw7(d)) radiance = add(radiance, direct7(cx, ray, hit, kd, key, V(0, 0, 0)));
    else radiance = direct7(cx, ray, hit, kd, key, radiance);   /* color += ... per light, HW6/src/Scene.cpp:196-207 */
  }
  if (!is_zero(ld(m->mirror)) && level < d->max_depth)
    radiance = add(radiance, mulv(ld(m->mirror), reflect7(cx, ray, hit, level, key)));
  if (!is_zero(ld(m->transparency)) && level < d->max_depth)
    radiance = add(radiance, refract7(cx, ray, hit, level, key));
  return radiance;
}
/* Scene::trace_path — HW7/src/Scene.cpp:312-414 */
static v3 trace_path7(const Ctx* cx, const Ray* ray, const Hit* hit, int level, uint32_t key) {
  const B200SceneDesc* d = cx->d;
  v3 radiance = V(0, 0, 0);
  if (hit->light >= 0) return ray->light_hit ? radiance : ld(d->lights[hit->light].intensity);
  Ray ray_copy = *ray;
  v3 kd, direct = V(0, 0, 0);
  diffuse_constant7(d, hit, &kd);
  if (!ray->in_medium) direct = direct7(cx, ray, hit, kd, key, V(0, 0, 0));
  if (!is_zero(direct)) { ray_copy.light_hit = 1; radiance = add(radiance, direct); }
  const b200rt_material* m = d->materials + hit->material;
  const b200rt_brdf* B = (m->brdf != -1) ? d->brdfs + m->brdf : NULL;
  if (!ray->in_medium && level < d->max_depth) {
    const v3 p = add(ray->o, muls(ray->d, hit->t));
    const float e1 = DRAW7(hemi, key, 9);
    const float e2 = DRAW7(hemi, key, 10);
    const float phi = (float)(2 * M_PI * (double)e1);
    const float theta = d->importance_sampling ? asinf(sqrtf(e2)) : acosf(e2);
    const v3 n = hit->normal;
    const v3 w_i = onb_dir(n, theta, phi);
    const float cti = fmaxf(0.0f, dot(n, w_i));
    const v3 w_o = normalize(sub(ray->o, p));
    Ray sr; memset(&sr, 0, sizeof(sr));
    sr.o = add(p, muls(w_i, d->shadow_eps)); sr.d = w_i; sr.time = ray->time;
    sr.in_medium = ray->in_medium; sr.light_hit = ray->light_hit;
    const v3 Li = send_ray7(cx, &sr, level + 1, rng_mix(key, 0x104u), 1);
    const float prob = d->importance_sampling ? (float)((double)cti / M_PI) : (float)(1.0f / (2 * M_PI));
    if (prob != 0.0f) {
      if (B) {
        radiance = add(radiance, divs(mulv(muls(Li, cti), brdf_eval(B, n, kd, ld(m->specular), w_i, w_o)), prob));
      } else {
        radiance = add(radiance, divs(muls(mulv(kd, Li), cti), prob));
        const float sct = fmaxf(0.0f, dot(n, normalize(add(w_o, w_i))));
        radiance = add(radiance, divs(muls(mulv(ld(m->specular), Li), (float)pow((double)sct, (double)m->phong_exponent)), prob));
      }
    }
  }
  if (!is_zero(ld(m->mirror)) && level < d->max_depth)
    radiance = add(radiance, mulv(ld(m->mirror), reflect7(cx, &ray_copy, hit, level, key)));
  if (!is_zero(ld(m->transparency)) && level < d->max_depth)
    radiance = add(radiance, refract7(cx, &ray_copy, hit, level, key));
  return radiance;
}
/* Scene::send_ray — HW7/src/Scene.cpp:168-189 */
static v3 send_ray7(const Ctx* cx, const Ray* ray, int level, uint32_t key, int count_as) {
  const B200SceneDesc* d = cx->d;
  if (count_as == 0) cx->c->rays_primary++; else cx->c->rays_secondary++;
  Hit hit;
  if (!scene_intersect(cx, ray, &hit)) {
    if (ray->primary) {
      if (d->background_texture >= 0) return tex_color(d, d->textures + d->background_texture, ray->bg_u, ray->bg_v);
      if (d->env_light >= 0) return light_radiance(d, d->env_light, ray->d, 1.0f);
      return ld(d->background);
    }
    if (d->env_light >= 0) return light_radiance(d, d->env_light, ray->d, 1.0f);
    return V(0, 0, 0);
  }
  if (!d->integrator) return trace_ray7(cx, ray, &hit, level, key);
  return trace_path7(cx, ray, &hit, level, key);
}

/* Camera::calculate_ray_at — HW1/Camera.h:24-32, HW3/include/Camera.h:35-47 */
static Ray camera_ray(const B200SceneDesc* d, const b200rt_camera* c, float x, float y, int use_dof,
                      float ex, float ey, float time) {
  Ray r; memset(&r, 0, sizeof(r));
  r.time = time;
  const v3 e = ld(c->e), u = ld(c->u), v = ld(c->v);
  if (d->semantics == B200RT_SEM_HW1) {
    const float su = (c->right - c->left) * (x + 0.5f) / c->width;
    const float sv = (c->top - c->bottom) * (y + 0.5f) / c->height;
    const v3 s = sub(add(ld(c->top_left), muls(u, su)), muls(v, sv));
    r.o = e; r.d = sub(s, e);
    return r;
  }
  /* HW4..HW7 callers pass the absolute sample position (HW7/include/Camera.h:66-84): x_off = 0 */
  const float off = (d->semantics >= B200RT_SEM_HW4) ? 0.0f : 0.5f;
  const v3 s = (d->semantics >= B200RT_SEM_HW4)
                   ? sub(add(ld(c->top_left), muls(ld(c->su), x)), muls(ld(c->sv), y))
                   : sub(add(ld(c->top_left), muls(ld(c->su), x + off)), muls(ld(c->sv), y + off));
  r.primary = 1;
  r.bg_u = x / c->width; r.bg_v = y / c->height;
  if (use_dof) {
    const v3 ne = add(e, muls(muls(add(muls(u, ex), muls(v, ey)), c->aperture), 0.5f));
    r.o = ne; r.d = normalize(sub(s, ne));
  } else {
    r.o = e; r.d = normalize(sub(s, e));
  }
  return r;
}

/* gaussian_filter — HW3/src/Scene.cpp:15-17 divides by (float)(2 pi sigma); HW2/Scene.cpp:12-14 by the double
 * 2 * M_PI * sigma (no cast): the weights differ by an ulp now and then (found by the replay against the HW2 tree) */
static inline float gaussian_filter_hw2(float x, float y, float sigma) {
  return (float)(exp(-(x * x + y * y) / (2 * sigma * sigma)) / (2 * M_PI * sigma));
}
static inline float gaussian_filter(float x, float y, float sigma) {
  return (float)(exp(-(x * x + y * y) / (2 * sigma * sigma)) / (float)(2 * M_PI * sigma));
}

/* Scene::render_image — HW1/Scene.cpp:9-27, HW3/src/Scene.cpp:21-99.
 * rgbw: W*H*4 floats {sum color*w, sum w}. ids (2 per pixel) / tout may be NULL.
 * For HW1 `rgbw` receives the raw float colour with w=1; quantisation lives in
 * the host mirror (ceng795_b200/image.py) for both arms. */
int oracle_render(const B200SceneDesc* d, int camera_index, int row_begin, int row_end, uint64_t seed,
                  float* rgbw, int32_t* ids, float* tout, OracleCounters* counters) {
  if (!d || camera_index < 0 || camera_index >= d->n_cameras) return -1;
  const b200rt_camera* c = d->cameras + camera_index;
  const int W = c->width, H = c->height;
  if (row_end <= 0 || row_end > H) row_end = H;
  if (row_begin < 0) row_begin = 0;
  OracleCounters total; memset(&total, 0, sizeof(total));
  const int ns = c->samples_sqrt;
  const uint32_t s0 = rng_mix((uint32_t)seed, (uint32_t)(seed >> 32));
  /* tools/check_nonfinite_oracle.py: ORACLE_COLS=x0:x1 restricts the per-pixel loop to a column window */
  int col_begin = 0, col_end = W;
  int strata_begin = 0, strata_end = 0x7fffffff;   /* ORACLE_STRATA=a:b: only those strata of every pixel (debugging) */
  if (getenv("ORACLE_STRATA")) sscanf(getenv("ORACLE_STRATA"), "%d:%d", &strata_begin, &strata_end);
  if (getenv("ORACLE_COLS")) { sscanf(getenv("ORACLE_COLS"), "%d:%d", &col_begin, &col_end); if (col_end > W) col_end = W; }
#pragma omp parallel if (!g_mt7.on)
  {
    OracleCounters cnt; memset(&cnt, 0, sizeof(cnt));
    Ctx cx = {d, &cnt};
    const int hw7 = sem_hw7(d), hw5 = sem_hw5(d);
#pragma omp for collapse(2) schedule(dynamic, 64)
    for (int j = row_begin; j < row_end; j++) {
      for (int i = col_begin; i < col_end; i++) {
        const int pix = j * W + i;
        tl_pixel = pix;
        if (hw7 && ns > 1 && rgbw) {
          /* HW7/src/Scene.cpp:35-85: strata loop, box filter (GAUSSIAN_FILTER is off) */
          float acc[4] = {0, 0, 0, 0};
          for (int x = 0; x < ns; x++)
            for (int y = 0; y < ns; y++) {
              if (x * ns + y < strata_begin || x * ns + y >= strata_end) continue;
              const uint32_t key = rng_mix(rng_mix(s0, (uint32_t)pix), (uint32_t)(x * ns + y));
              /* draw order of HW7/src/Scene.cpp:47-61: jitter x, y, then (lens x, y), then time */
              const float ex = DRAW7(render, key, 4);
              const float ey = DRAW7(render, key, 5);
              const float sx = (x + ex) / ns, sy = (y + ey) / ns;
              Ray r;
              if (c->aperture == 0.0f) {
                const float time = DRAW7(render, key, 6);
                r = camera_ray(d, c, i + sx, j + sy, 0, 0, 0, time);
              } else {
                const float dx = DRAW7(render, key, 7) * 2.0f - 1.0f;
                const float dy = DRAW7(render, key, 8) * 2.0f - 1.0f;
                const float time = DRAW7(render, key, 6);
                r = camera_ray(d, c, i + sx, j + sy, 1, dx, dy, time);
              }
              const v3 col = send_ray7(&cx, &r, 0, key, 0);
              acc[0] += col.x * 1.0f; acc[1] += col.y * 1.0f; acc[2] += col.z * 1.0f; acc[3] += 1.0f;
            }
          memcpy(rgbw + 4 * (size_t)pix, acc, sizeof(acc));
          continue;
        }
        if (ns == 1) {
          Ray r = sem_hw4(d) ? camera_ray(d, c, i + 0.5f, j + 0.5f, 0, 0, 0, 0.0f)
                      : camera_ray(d, c, (float)i, (float)j, 0, 0, 0, 0.0f);
          if (ids || tout) {
            Hit h; OracleCounters dummy; memset(&dummy, 0, sizeof(dummy));
            Ctx c2 = {d, &dummy};
            const int f = scene_intersect(&c2, &r, &h);
            if (ids) { ids[2 * pix] = f ? h.top_rank : -1; ids[2 * pix + 1] = f ? h.tri : -1; }
            if (tout) tout[pix] = f ? h.t : INFINITY;
          }
          if (rgbw && d->semantics == B200RT_SEM_HW4) continue;   /* splatted below, sequentially */
          if (rgbw) {
            const uint32_t key = rng_mix(rng_mix(s0, (uint32_t)pix), 0u);
            const v3 col = hw5 ? send_ray7(&cx, &r, 0, key, 0) : trace_ray(&cx, &r, d->max_depth, key, 0);
            rgbw[4 * pix] = col.x * 1.0f; rgbw[4 * pix + 1] = col.y * 1.0f;
            rgbw[4 * pix + 2] = col.z * 1.0f; rgbw[4 * pix + 3] = 1.0f;
          }
        }
      }
    }
#pragma omp critical
    {
      total.rays_primary += cnt.rays_primary; total.rays_shadow += cnt.rays_shadow;
      total.rays_secondary += cnt.rays_secondary; total.nodes_visited += cnt.nodes_visited;
      total.tris_tested += cnt.tris_tested; total.spheres_tested += cnt.spheres_tested;
      total.instances_entered += cnt.instances_entered;
    }
  }
  if (ns == 1 && rgbw && d->semantics == B200RT_SEM_HW4) {
    /* HW4 splats the 3x3 Gaussian (sigma 1/2) even at one sample per pixel, with integer offsets
     * (APPLY_FILTER_SINGLE_SAMPLE, HW4/src/Scene.cpp:10,26-49) */
    OracleCounters cnt; memset(&cnt, 0, sizeof(cnt));
    Ctx cx = {d, &cnt};
    for (int j = row_begin; j < row_end; j++)
      for (int i = 0; i < W; i++) {
        const int pix = j * W + i;
        const uint32_t key = rng_mix(rng_mix(s0, (uint32_t)pix), 0u);
        Ray r = camera_ray(d, c, i + 0.5f, j + 0.5f, 0, 0, 0, 0.0f);
        const v3 col = trace_ray(&cx, &r, d->max_depth, key, 0);
        for (int aj = j - 1; aj < j + 2; aj++) {
          if (aj < 0 || aj >= H) continue;
          for (int ai = i - 1; ai < i + 2; ai++) {
            if (ai < 0 || ai >= W) continue;
            const float w = gaussian_filter((float)(i - ai), (float)(j - aj), 1.5f / 3.0f);
            float* px = rgbw + 4 * (size_t)(aj * W + ai);
            px[0] += col.x * w; px[1] += col.y * w; px[2] += col.z * w; px[3] += w;
          }
        }
      }
    total.rays_primary += cnt.rays_primary; total.rays_shadow += cnt.rays_shadow;
    total.rays_secondary += cnt.rays_secondary; total.nodes_visited += cnt.nodes_visited;
    total.tris_tested += cnt.tris_tested; total.spheres_tested += cnt.spheres_tested;
    total.instances_entered += cnt.instances_entered;
  }
  if (ns > 1 && rgbw && !sem_hw7(d)) {
    /* multi-sample path with the 3x3 Gaussian splat: sequential over pixels so
     * that the accumulation order is defined (the reference's is thread-racy). */
    OracleCounters cnt; memset(&cnt, 0, sizeof(cnt));
    Ctx cx = {d, &cnt};
    const int hw4ms = d->semantics >= B200RT_SEM_HW4;   /* HW4-HW6: sigma 1/2, absolute sample positions */
    const int hw5 = sem_hw5(d);
    ClockEngine ce_ms, ce_dof, ce_time;
    memset(&ce_ms, 0, sizeof(ce_ms)); memset(&ce_dof, 0, sizeof(ce_dof)); memset(&ce_time, 0, sizeof(ce_time));
    for (int j = row_begin; j < row_end; j++)
      for (int i = 0; i < W; i++) {
        const int pix = j * W + i;
        for (int x = 0; x < ns; x++)
          for (int y = 0; y < ns; y++) {
            const uint32_t key = rng_mix(rng_mix(s0, (uint32_t)pix), (uint32_t)(x * ns + y));
            /* HW5/HW6 draw in this order from one engine (HW6/src/Scene.cpp:60-77): jitter x, y, (lens x, y), time;
             * HW2-HW4 from three engines re-seeded from the clock at every pixel (HW3/src/Scene.cpp:40-53) */
            const int clk = g_clk.on && !hw5;
            if (clk && x == 0 && y == 0) { ce_seed(&ce_ms, sem_hw4(d)); ce_seed(&ce_dof, sem_hw4(d)); ce_seed(&ce_time, sem_hw4(d)); }
            const float ex = clk ? ce_u01(&ce_ms) : DRAW7(render, key, 4);
            const float ey = clk ? ce_u01(&ce_ms) : DRAW7(render, key, 5);
            const float sx = (x + ex) / ns, sy = (y + ey) / ns;
            Ray r;
            const float px_ = hw4ms ? (i + sx) : (i + sx - 0.5f), py_ = hw4ms ? (j + sy) : (j + sy - 0.5f);
            if (c->aperture == 0.0f) {
              const float time = clk ? ce_u01(&ce_time) : DRAW7(render, key, 6);
              r = camera_ray(d, c, px_, py_, 0, 0, 0, time);
            } else {
              const float dx = (clk ? ce_u01(&ce_dof) : DRAW7(render, key, 7)) * 2.0f - 1.0f;
              const float dy = (clk ? ce_u01(&ce_dof) : DRAW7(render, key, 8)) * 2.0f - 1.0f;
              const float time = clk ? ce_u01(&ce_time) : DRAW7(render, key, 6);
              r = camera_ray(d, c, px_, py_, 1, dx, dy, time);
            }
            const v3 col = hw5 ? send_ray7(&cx, &r, 0, key, 0) : trace_ray(&cx, &r, d->max_depth, key, 0);
            for (int aj = j - 1; aj < j + 2; aj++) {
              if (aj < 0 || aj >= H) continue;
              for (int ai = i - 1; ai < i + 2; ai++) {
                if (ai < 0 || ai >= W) continue;
                const float fx = (i + sx) - (ai + 0.5f);
                const float fy = (j + sy) - (aj + 0.5f);
                const float w = d->semantics == B200RT_SEM_HW2 ? gaussian_filter_hw2(fx, fy, 1.0f / 3.0f)
                                                               : gaussian_filter(fx, fy, hw4ms ? 1.5f / 3.0f : 1.0f / 3.0f);
                float* px = rgbw + 4 * (size_t)(aj * W + ai);
                px[0] += col.x * w; px[1] += col.y * w; px[2] += col.z * w; px[3] += w;
              }
            }
          }
      }
    total.rays_primary += cnt.rays_primary; total.rays_shadow += cnt.rays_shadow;
    total.rays_secondary += cnt.rays_secondary; total.nodes_visited += cnt.nodes_visited;
    total.tris_tested += cnt.tris_tested; total.spheres_tested += cnt.spheres_tested;
    total.instances_entered += cnt.instances_entered;
  }
  if (counters) *counters = total;
  return 0;
}

/* ======================================================================== */
/* PPM — progressive photon mapping (PPM/src/Scene.cpp:15-372, main.cpp:33-157) */
/* ======================================================================== */
typedef struct {
  v3 position, normal, w_o, attenuation, flux;
  float radius_squared, pixel_weight;
  uint32_t n;
  int material, pixel;
} OracleHitPoint;
typedef struct {
  OracleHitPoint* hp;
  int64_t n, cap;
  /* hash grid (CSR) */
  int64_t num_hash;
  int64_t* cell_start;
  int32_t* cell_items;
  v3 bbox_min;
  float hash_scale, initial_radius;
} OraclePpm;

static int ppm_material_type(const b200rt_material* m) { /* PPM/src/Material.cpp:55-61 */
  if (!is_zero(ld(m->mirror))) return 1;
  if (!is_zero(ld(m->transparency))) return 2;
  return 0;
}
static void ppm_add_hp(OraclePpm* P, const OracleHitPoint* h) {
  if (P->n == P->cap) {
    P->cap = P->cap ? 2 * P->cap : 4096;
    P->hp = (OracleHitPoint*)realloc(P->hp, sizeof(OracleHitPoint) * (size_t)P->cap);
  }
  P->hp[P->n++] = *h;
}
/* Scene::eye_trace — PPM/src/Scene.cpp:286-362 */
static void ppm_eye_trace(const Ctx* cx, OraclePpm* P, const Ray* ray, int depth, v3 att, int pixel, float weight) {
  const B200SceneDesc* d = cx->d;
  Hit hit;
  cx->c->rays_primary += (depth == 0);
  cx->c->rays_secondary += (depth != 0);
  if (!scene_intersect(cx, ray, &hit)) return;
  const v3 p = add(ray->o, muls(ray->d, hit.t));
  const v3 n = hit.normal;
  const b200rt_material* m = d->materials + hit.material;
  const int type = ppm_material_type(m);
  if (type == 0) {
    OracleHitPoint h; memset(&h, 0, sizeof(h));
    h.material = hit.material; h.attenuation = att; h.w_o = normalize(sub(ray->o, p)); h.normal = n; h.position = p;
    h.pixel = pixel; h.pixel_weight = weight;
    ppm_add_hp(P, &h);
    return;
  }
  if (depth >= d->max_depth) return;
  const v3 w_o = normalize(sub(ray->o, p));
  const v3 w_r = normalize(sub(muls(n, 2.0f * dot(n, w_o)), w_o));
  Ray rr; memset(&rr, 0, sizeof(rr));
  rr.o = add(p, muls(w_r, d->shadow_eps)); rr.d = w_r;
  if (type == 1) {
    ppm_eye_trace(cx, P, &rr, depth + 1, mulv(ld(m->mirror), att), pixel, weight);
    return;
  }
  const v3 nl = dot(n, ray->d) < 0.0f ? n : muls(n, -1);
  const int into = dot(n, nl) > 0.0f;
  const float air = 1.0f;
  const float nnt = into ? air / m->refraction_index : m->refraction_index / air;
  const float ddn = dot(ray->d, nl);
  const float cos2t = 1 - nnt * nnt * (1 - ddn * ddn);
  if (cos2t < 0.0f) {
    ppm_eye_trace(cx, P, &rr, depth + 1, mulv(ld(m->transparency), att), pixel, weight);
    return;
  }
  const v3 rdir = normalize(sub(muls(ray->d, nnt), muls(n, (into ? 1 : -1) * (ddn * nnt + sqrtf(cos2t)))));
  const float a = m->refraction_index - air, b = m->refraction_index + air;
  const float R0 = a * a / (b * b);
  const float ca = into ? -ddn : dot(rdir, n);
  const float c = 1 - ca;
  const float fres = R0 + (1 - R0) * c * c * c * c * c;
  Ray tr; memset(&tr, 0, sizeof(tr));
  tr.o = add(p, muls(rdir, d->shadow_eps)); tr.d = rdir;
  const v3 ac = mulv(ld(m->transparency), att);
  if (into) {
    ppm_eye_trace(cx, P, &rr, depth + 1, muls(att, fres), pixel, weight);
    ppm_eye_trace(cx, P, &tr, depth + 1, muls(ac, 1.0f - fres), pixel, weight);
  } else {
    ppm_eye_trace(cx, P, &tr, depth + 1, ac, pixel, weight);
  }
}
/* Scene::hash — PPM/include/Scene.h:65-68 (signed products wrap; cast to unsigned) */
static inline int64_t ppm_hash(const OraclePpm* P, int ix, int iy, int iz) {
  const uint32_t h = ((uint32_t)ix * 73856093u) ^ ((uint32_t)iy * 19349663u) ^ ((uint32_t)iz * 83492791u);
  return (int64_t)(h % (uint32_t)P->num_hash);
}
/* Scene::build_hash_grid — PPM/src/Scene.cpp:53-93 */
static void ppm_build_grid(OraclePpm* P, int width, int height) {
  v3 mn = V(INFINITY, INFINITY, INFINITY), mx = V(-INFINITY, -INFINITY, -INFINITY);
  for (int64_t i = 0; i < P->n; i++) {
    const v3 q = P->hp[i].position;
    mn = V(fminf(mn.x, q.x), fminf(mn.y, q.y), fminf(mn.z, q.z));
    mx = V(fmaxf(mx.x, q.x), fmaxf(mx.y, q.y), fmaxf(mx.z, q.z));
  }
  const v3 size = sub(mx, mn);
  const float r0 = ((size.x + size.y + size.z) / 3.0f) / ((width + height) / 2.0f) * 2.0f * 4.0f;
  P->initial_radius = r0;
  mn = V(INFINITY, INFINITY, INFINITY); mx = V(-INFINITY, -INFINITY, -INFINITY);
  P->num_hash = P->n;
  for (int64_t i = 0; i < P->n; i++) {
    OracleHitPoint* h = P->hp + i;
    h->radius_squared = r0 * r0; h->n = 0; h->flux = V(0, 0, 0);
    const v3 lo = V(h->position.x - r0, h->position.y - r0, h->position.z - r0);
    const v3 hi = V(h->position.x + r0, h->position.y + r0, h->position.z + r0);
    mn = V(fminf(mn.x, lo.x), fminf(mn.y, lo.y), fminf(mn.z, lo.z));
    mx = V(fmaxf(mx.x, lo.x), fmaxf(mx.y, lo.y), fmaxf(mx.z, lo.z));
    mn = V(fminf(mn.x, hi.x), fminf(mn.y, hi.y), fminf(mn.z, hi.z));
    mx = V(fmaxf(mx.x, hi.x), fmaxf(mx.y, hi.y), fmaxf(mx.z, hi.z));
  }
  P->bbox_min = mn;
  P->hash_scale = (float)(1.0 / ((double)r0 * 2.0));
  free(P->cell_start); free(P->cell_items);
  P->cell_start = (int64_t*)calloc((size_t)P->num_hash + 1, sizeof(int64_t));
  for (int pass = 0; pass < 2; pass++) {
    if (pass == 1) {
      int64_t acc = 0;
      for (int64_t c = 0; c <= P->num_hash; c++) { const int64_t k = P->cell_start[c]; P->cell_start[c] = acc; acc += k; }
      P->cell_items = (int32_t*)malloc(sizeof(int32_t) * (size_t)(acc ? acc : 1));
    }
    int64_t* fill = NULL;
    if (pass == 1) { fill = (int64_t*)malloc(sizeof(int64_t) * (size_t)(P->num_hash + 1)); memcpy(fill, P->cell_start, sizeof(int64_t) * (size_t)(P->num_hash + 1)); }
    for (int64_t i = 0; i < P->n; i++) {
      const OracleHitPoint* h = P->hp + i;
      const v3 bmin = muls(sub(V(h->position.x - r0, h->position.y - r0, h->position.z - r0), mn), P->hash_scale);
      const v3 bmax = muls(sub(V(h->position.x + r0, h->position.y + r0, h->position.z + r0), mn), P->hash_scale);
      for (int iz = abs((int)bmin.z); iz <= abs((int)bmax.z); iz++)
        for (int iy = abs((int)bmin.y); iy <= abs((int)bmax.y); iy++)
          for (int ix = abs((int)bmin.x); ix <= abs((int)bmax.x); ix++) {
            const int64_t hv = ppm_hash(P, ix, iy, iz);
            if (pass == 0) P->cell_start[hv]++; else P->cell_items[fill[hv]++] = (int32_t)i;
          }
    }
    free(fill);
  }
}
/* Scene::photon_trace — PPM/src/Scene.cpp:106-249; RNG dims of the photon key: emission 0,1; bounce b: 8b+{2,3} hemisphere,
 * 8b+4 roulette, 8b+5 glass roulette */
#define PPM_ALPHA 0.7f
static void ppm_photon_trace(const Ctx* cx, OraclePpm* P, Ray ray, v3 flux, uint32_t key, OracleCounters* pc) {
  const B200SceneDesc* d = cx->d;
  int depth = 0;
  for (int b = 0;; b++) {
    depth++;
    if (depth >= d->max_depth) return;
    Hit hit;
    pc->rays_shadow++; /* reused as "photon rays" */
    if (!scene_intersect(cx, &ray, &hit)) return;
    const v3 p = add(ray.o, muls(ray.d, hit.t));
    const v3 n = hit.normal;
    const v3 nl = dot(n, ray.d) < 0 ? n : muls(n, -1);
    const b200rt_material* m = d->materials + hit.material;
    const int type = ppm_material_type(m);
    const uint32_t dim = 8u * (uint32_t)b;
    if (type == 0) {
      if (P->num_hash > 0) {
        const v3 hh = muls(sub(p, P->bbox_min), P->hash_scale);
        const int ix = abs((int)hh.x), iy = abs((int)hh.y), iz = abs((int)hh.z);
        const int64_t cell = ppm_hash(P, ix, iy, iz);
        for (int64_t k = P->cell_start[cell]; k < P->cell_start[cell + 1]; k++) {
          OracleHitPoint* h = P->hp + P->cell_items[k];
          const v3 v = sub(h->position, p);
          pc->spheres_tested++; /* reused as "hit points scanned" */
          if ((dot(h->normal, n) > 1e-3f) && (dot(v, v) <= h->radius_squared)) {
            const float rr = (float)(((double)(h->n * PPM_ALPHA + PPM_ALPHA)) / ((double)(h->n * PPM_ALPHA) + 1.0));
            h->radius_squared = h->radius_squared * rr;
            h->n++;
            pc->instances_entered++; /* reused as "photons accepted" */
            v3 color = V(0, 0, 0);
            const v3 w_i = neg(normalize(ray.d));
            const b200rt_material* hm = d->materials + h->material;
            const float cti = dot(h->normal, w_i);
            if (!(cti > 1.0f || cti <= 0.0f)) {
              const float sct = fmaxf(0.0f, dot(h->normal, normalize(add(h->w_o, w_i))));
              color = mulv(add(ld(hm->diffuse), divs(muls(ld(hm->specular), powf(sct, hm->phong_exponent)), cti)), h->attenuation);
            }
            h->flux = muls(add(h->flux, mulv(color, flux)), rr);
          }
        }
      }
      /* sample_hemisphere(normal, d, probability) — cosine, PPM/src/Scene.cpp:15-44 */
      const float e1 = PPM_DRAW(hemi, key, dim + 2);
      const float e2 = PPM_DRAW(hemi, key, dim + 3);
      const float phi = (float)(2 * M_PI * (double)e1);
      const float theta = asinf(sqrtf(e2));
      const v3 dd = onb_dir(n, theta, phi);
      const float prob = (float)((double)fmaxf(0.0f, dot(n, dd)) / M_PI);
      v3 base = V(0, 0, 0);
      const v3 w_i = neg(normalize(ray.d));
      const float cti = fmaxf(0.0f, dot(n, w_i));
      if (!(cti > 1.0f || cti <= 0.0f)) {
        const float sct = fmaxf(0.0f, dot(nl, normalize(add(dd, w_i))));
        base = add(ld(m->diffuse), divs(muls(ld(m->specular), powf(sct, m->phong_exponent)), cti));
      }
      const float cto = fmaxf(0.0f, dot(n, dd));
      base = muls(base, cto);
      if (PPM_DRAW(trace, key, dim + 4) < prob) {
        ray.o = add(p, muls(dd, d->shadow_eps)); ray.d = dd;
        flux = divs(mulv(base, flux), prob);
        continue;
      }
      return;
    }
    const v3 w_o = normalize(sub(ray.o, p));
    const v3 w_r = normalize(sub(muls(n, 2.0f * dot(n, w_o)), w_o));
    if (type == 1) {
      ray.o = add(p, muls(w_r, d->shadow_eps)); ray.d = w_r;
      flux = mulv(ld(m->mirror), flux);
      continue;
    }
    const int into = dot(n, nl) > 0.0f;
    const float air = 1.0f;
    const float nnt = into ? air / m->refraction_index : m->refraction_index / air;
    const float ddn = dot(ray.d, nl);
    const float cos2t = 1 - nnt * nnt * (1 - ddn * ddn);
    if (cos2t < 0.0f) { ray.o = add(p, muls(w_r, d->shadow_eps)); ray.d = w_r; continue; }
    const v3 rdir = normalize(sub(muls(ray.d, nnt), muls(n, (into ? 1 : -1) * (ddn * nnt + sqrtf(cos2t)))));
    const float a = m->refraction_index - air, bb = m->refraction_index + air;
    const float R0 = a * a / (bb * bb);
    const float ca = into ? -ddn : dot(rdir, n);
    const float c = 1 - ca;
    const float fres = R0 + (1 - R0) * c * c * c * c * c;
    if (into && PPM_DRAW(trace, key, dim + 5) < fres) { ray.o = add(p, muls(w_r, d->shadow_eps)); ray.d = w_r; }
    else { ray.o = add(p, muls(rdir, d->shadow_eps)); ray.d = rdir; }
  }
}

/* The whole PPM frame: reset_hash_grid, eye_trace_lines, build_hash_grid, trace_n_photons (ONE sequential stream,
 * like the reference built without NDEBUG), density_estimation.  hp_out (optional): 16 floats per hit point
 * {position, normal, w_o, attenuation, material, pixel, pixel_weight, n}. */
/* Test yardstick: 1 = trace the same photons in REVERSE index order.  The order photons reach a hit point in is free
 * in the reference (its threads race for the mutex, PPM/src/Scene.cpp:135-169); two orders of one photon set bound
 * what "same photons, any serialisation order" can differ by. */
static int g_ppm_reverse = 0;
void oracle_ppm_photon_order(int reverse) { g_ppm_reverse = reverse; }
int oracle_ppm(const B200SceneDesc* d, int camera_index, uint64_t seed, int64_t n_photons, double total_num_of_photons,
               float* rgbw, float* hp_out, int64_t hp_cap, int64_t* n_hp, float* initial_radius, OracleCounters* counters) {
  if (!d || camera_index < 0 || camera_index >= d->n_cameras || d->n_lights < 1) return -1;
  const b200rt_camera* c = d->cameras + camera_index;
  const int W = c->width, H = c->height, ns = c->samples_sqrt;
  OracleCounters cnt; memset(&cnt, 0, sizeof(cnt));
  Ctx cx = {d, &cnt};
  OraclePpm P; memset(&P, 0, sizeof(P));
  const uint32_t s0 = rng_mix((uint32_t)seed, (uint32_t)(seed >> 32));
  for (int j = 0; j < H; j++)
    for (int i = 0; i < W; i++) {
      const int pix = j * W + i;
      if (ns == 1) {
        Ray r = camera_ray(d, c, i + 0.5f, j + 0.5f, 0, 0, 0, 0.0f);
        ppm_eye_trace(&cx, &P, &r, 0, V(1, 1, 1), pix, 1.0f);
      } else {
        for (int x = 0; x < ns; x++)
          for (int y = 0; y < ns; y++) {
            const uint32_t key = rng_mix(rng_mix(s0, (uint32_t)pix), (uint32_t)(x * ns + y));
            /* jitter of the multi-sample eye pass: one engine per eye_trace_lines call, PPM/src/Scene.cpp:264-274 */
            const float jx = PPM_DRAW(eye, key, 4);
            const float jy = PPM_DRAW(eye, key, 5);
            const float sx = (x + jx) / ns, sy = (y + jy) / ns;
            Ray r = camera_ray(d, c, i + sx, j + sy, 0, 0, 0, 0.0f);
            ppm_eye_trace(&cx, &P, &r, 0, V(1, 1, 1), pix, 1.0f);
          }
      }
    }
  ppm_build_grid(&P, W, H);
  if (initial_radius) *initial_radius = P.initial_radius;
  /* Point_light::generate_photon — PPM/src/Point_light.cpp:7-30 (theta = 2*pi*xi: Appendix B.6) */
  const b200rt_light* L = d->lights;
  OracleCounters pc; memset(&pc, 0, sizeof(pc));
  for (int64_t ii = 0; ii < n_photons; ii++) {
    const int64_t i = g_ppm_reverse ? n_photons - 1 - ii : ii;
    const uint32_t key = rng_mix(rng_mix(s0 ^ 0x50484F54u, (uint32_t)i), (uint32_t)((uint64_t)i >> 32));
    const v3 flux = muls(ld(L->intensity), (float)(M_PI * 4.0f));
    const float le1 = PPM_DRAW(light, key, 0);
    const float le2 = PPM_DRAW(light, key, 1);
    const float phi = (float)(2 * M_PI * (double)le1);
    const float theta = (float)(2 * M_PI * (double)le2);
    Ray r; memset(&r, 0, sizeof(r));
    r.d = onb_dir(V(0.0f, 1.0f, 0.0f), theta, phi);
    r.o = ld(L->position);
    ppm_photon_trace(&cx, &P, r, flux, key, &pc);
  }
  /* Scene::density_estimation — PPM/src/Scene.cpp:363-372 */
  if (rgbw) {
    memset(rgbw, 0, sizeof(float) * 4 * (size_t)W * H);
    for (int64_t i = 0; i < P.n; i++) {
      const OracleHitPoint* h = P.hp + i;
      const float k = (float)(1.0f / (M_PI * (double)h->radius_squared * total_num_of_photons));
      const v3 col = muls(h->flux, k);
      float* px = rgbw + 4 * (size_t)h->pixel;
      px[0] += col.x * h->pixel_weight; px[1] += col.y * h->pixel_weight; px[2] += col.z * h->pixel_weight;
      px[3] += h->pixel_weight;
    }
  }
  if (hp_out)
    for (int64_t i = 0; i < P.n && i < hp_cap; i++) {
      const OracleHitPoint* h = P.hp + i;
      float* o = hp_out + 16 * i;
      o[0] = h->position.x; o[1] = h->position.y; o[2] = h->position.z; o[3] = h->normal.x; o[4] = h->normal.y; o[5] = h->normal.z;
      o[6] = h->w_o.x; o[7] = h->w_o.y; o[8] = h->w_o.z; o[9] = h->attenuation.x; o[10] = h->attenuation.y; o[11] = h->attenuation.z;
      o[12] = (float)h->material; o[13] = (float)h->pixel; o[14] = h->pixel_weight; o[15] = (float)h->n;
    }
  if (g_mt.state_out)
    for (int64_t i = 0; i < P.n && i < g_mt.state_cap; i++) {
      const OracleHitPoint* h = P.hp + i;
      float* o = g_mt.state_out + 5 * i;
      o[0] = h->flux.x; o[1] = h->flux.y; o[2] = h->flux.z; o[3] = h->radius_squared; o[4] = (float)h->n;
    }
  if (n_hp) *n_hp = P.n;
  if (counters) {
    *counters = cnt;
    counters->rays_shadow = pc.rays_shadow;            /* photon rays        */
    counters->spheres_tested = pc.spheres_tested;      /* hit points scanned */
    counters->instances_entered = pc.instances_entered; /* photons accepted   */
  }
  free(P.hp); free(P.cell_start); free(P.cell_items);
  return 0;
}

/* Closest-hit on caller rays (8 floats per ray: o.xyz, time, d.xyz, unused). */
int oracle_trace_rays(const B200SceneDesc* d, const float* rays, int64_t n, int32_t* ids, float* tout,
                      OracleCounters* counters) {
  OracleCounters total; memset(&total, 0, sizeof(total));
#pragma omp parallel
  {
    OracleCounters cnt; memset(&cnt, 0, sizeof(cnt));
    Ctx cx = {d, &cnt};
#pragma omp for schedule(dynamic, 256)
    for (int64_t k = 0; k < n; k++) {
      Ray r; memset(&r, 0, sizeof(r));
      r.o = ld(rays + 8 * k); r.time = rays[8 * k + 3]; r.d = ld(rays + 8 * k + 4);
      Hit h;
      const int f = scene_intersect(&cx, &r, &h);
      if (ids) { ids[2 * k] = f ? h.top_rank : -1; ids[2 * k + 1] = f ? h.tri : -1; }
      if (tout) tout[k] = f ? h.t : INFINITY;
    }
#pragma omp critical
    {
      total.nodes_visited += cnt.nodes_visited; total.tris_tested += cnt.tris_tested;
      total.spheres_tested += cnt.spheres_tested; total.instances_entered += cnt.instances_entered;
    }
  }
  if (counters) *counters = total;
  return 0;
}

/* Leaf known-answer hooks (compared with oracle/_ref/ref_probe on seeded inputs). */
float oracle_box_intersect(const float* mn, const float* mx, const float* o, const float* dd, float eps) {
  Ray r; memset(&r, 0, sizeof(r)); r.o = ld(o); r.d = ld(dd);
  return box_intersect(mn, mx, &r, eps);
}
int oracle_tri_intersect(const float* g12, const float* o, const float* dd, int mode, float eps, float* out3) {
  return tri_core(g12, ld(o), ld(dd), mode, eps, out3, out3 + 1, out3 + 2);
}
/* Sphere::intersect on one sphere record: out5 = {t, normal.xyz, 0}.  `semantics` picks the tree's rules. */
int oracle_sphere_intersect(const b200rt_sphere* s, int semantics, float eps, const float* o, const float* dd, float time,
                            float* out4) {
  B200SceneDesc d; memset(&d, 0, sizeof(d));
  d.semantics = semantics; d.isect_eps = eps; d.spheres = s; d.n_spheres = 1;
  OracleCounters c; memset(&c, 0, sizeof(c));
  Ctx cx = {&d, &c};
  Ray r; memset(&r, 0, sizeof(r)); r.o = ld(o); r.d = ld(dd); r.time = time;
  Hit h; memset(&h, 0, sizeof(h));
  if (!sphere_intersect(&cx, 0, &r, &h)) return 0;
  out4[0] = h.t; out4[1] = h.normal.x; out4[2] = h.normal.y; out4[3] = h.normal.z;
  return 1;
}
/* Light::incoming_radiance for the kinds that need no texture (point, area, spot, directional, mesh, sphere). */
void oracle_light_radiance(const b200rt_light* L, const float* to_light, float prob, float* out3) {
  B200SceneDesc d; memset(&d, 0, sizeof(d));
  d.semantics = B200RT_SEM_HW7; d.lights = L; d.n_lights = 1;
  const v3 r = light_radiance(&d, 0, ld(to_light), prob);
  out3[0] = r.x; out3[1] = r.y; out3[2] = r.z;
}
/* BRDF::get_reflectance, Texture::get_color_at, Perlin_noise::get_value_at (tests/test_leaf_vectors.py) */
void oracle_brdf_eval(int kind, float exponent, int normalized, float refractive_index, const float* n, const float* kd,
                      const float* ks, const float* wi, const float* wo, float* out3) {
  b200rt_brdf B; memset(&B, 0, sizeof(B));
  B.kind = kind; B.exponent = exponent; B.normalized = normalized; B.refractive_index = refractive_index;
  const v3 f = brdf_eval(&B, ld(n), ld(kd), ld(ks), ld(wi), ld(wo));
  out3[0] = f.x; out3[1] = f.y; out3[2] = f.z;
}
void oracle_tex_color(const b200rt_texture* T, const float* texels, float u, float v, float* out3) {
  B200SceneDesc d; memset(&d, 0, sizeof(d));
  d.texels = texels;
  const v3 c = tex_color(&d, T, u, v);
  out3[0] = c.x; out3[1] = c.y; out3[2] = c.z;
}
float oracle_perlin_value(const b200rt_texture* T, const float* pos) { return perlin_value(T, ld(pos)); }
/* Deterministic half of the PPM frame (eye pass at 1 sample + hash grid) for tests/test_ppm.py against oracle/_ref/ppm_eye:
 * out5 = {r0^2, hash_scale, bbox_min.xyz}; cell_counts[num_hash] = size of every hash-grid cell. */
int oracle_ppm_grid(const B200SceneDesc* d, int camera_index, float* out5, int32_t* cell_counts, int64_t cap,
                    int64_t* num_hash) {
  if (!d || camera_index < 0 || camera_index >= d->n_cameras) return -1;
  const b200rt_camera* c = d->cameras + camera_index;
  if (c->samples_sqrt != 1) return -2;
  OracleCounters cnt; memset(&cnt, 0, sizeof(cnt));
  Ctx cx = {d, &cnt};
  OraclePpm P; memset(&P, 0, sizeof(P));
  for (int j = 0; j < c->height; j++)
    for (int i = 0; i < c->width; i++) {
      Ray r = camera_ray(d, c, i + 0.5f, j + 0.5f, 0, 0, 0, 0.0f);
      ppm_eye_trace(&cx, &P, &r, 0, V(1, 1, 1), j * c->width + i, 1.0f);
    }
  ppm_build_grid(&P, c->width, c->height);
  out5[0] = P.initial_radius * P.initial_radius; out5[1] = P.hash_scale;
  out5[2] = P.bbox_min.x; out5[3] = P.bbox_min.y; out5[4] = P.bbox_min.z;
  *num_hash = P.num_hash;
  for (int64_t k = 0; k < P.num_hash && k < cap; k++) cell_counts[k] = (int32_t)(P.cell_start[k + 1] - P.cell_start[k]);
  free(P.hp); free(P.cell_start); free(P.cell_items);
  return 0;
}
