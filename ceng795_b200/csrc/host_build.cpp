// host_build.cpp — host-side scene assembly for the B200 rendering core.
//
// This is the "host stays" half of the boundary written for a
// host that does not link the reference: it restates, operation for operation
// in IEEE binary32 (compile with -ffp-contract=off), the load-time arithmetic
// whose results the hot path consumes:
//   * Matrix4x4 / Transformation        HW3/src/Matrix4x4.cpp:29-233, HW3/src/Transformation.cpp:6-84
//   * Camera frame                      HW3/include/Camera.h:11-34, HW1/Camera.h:11-23
//   * Mesh_triangle / Triangle records  HW3/src/Mesh_triangle.cpp:4-37, HW3/src/Triangle.cpp:4-44
//   * vertex normals                    HW3/include/Vertex.h:13-22, HW3/src/Scene.cpp:690-697,872-877
//   * Bounding_box expand/apply_transform HW3/src/bounding_box.cpp:5-15,39-113
//   * BVH topology                      HW3/src/Bounding_volume_hierarchy.cpp:3-29
//   * object order of Scene::bvh        HW3/src/Scene.cpp:703-869
// The output is a B200SceneDesc (include/b200rt.h).  tests/test_host_build.py
// compares it against a dump made by walking the real reference objects
// (oracle/ref_probe.cpp built into oracle/_ref/).
#include "m4_invert.h"
#include <chrono>
#include <future>
#include <thread>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <string>
#include <system_error>
#include <utility>
#include <vector>

#include "../../include/b200rt.h"
#include "fast_bvh.h"
#include "api_guard.h"

namespace {

thread_local std::string g_host_error;

struct V3 {
  float x, y, z;
  V3() : x(0.f), y(0.f), z(0.f) {}
  V3(float a) : x(a), y(a), z(a) {}
  V3(float a, float b, float c) : x(a), y(b), z(c) {}
  float operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
  float& operator[](int i) { return i == 0 ? x : (i == 1 ? y : z); }
};
inline V3 operator+(V3 a, V3 b) { return V3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline V3 operator-(V3 a, V3 b) { return V3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline V3 operator-(V3 a) { return V3(-a.x, -a.y, -a.z); }
inline V3 operator*(V3 a, float s) { return V3(a.x * s, a.y * s, a.z * s); }
inline V3 operator*(float s, V3 a) { return a * s; }
inline V3 operator/(V3 a, float s) { return V3(a.x / s, a.y / s, a.z / s); }
inline bool operator==(V3 a, V3 b) { return a.x == b.x && a.y == b.y && a.z == b.z; }
inline bool operator!=(V3 a, V3 b) { return !(a == b); }
inline float dot(V3 a, V3 b) { return (a.x * b.x) + (a.y * b.y) + (a.z * b.z); }
inline V3 cross(V3 a, V3 b) {
  return V3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
inline float length(V3 a) { return std::sqrt(a.x * a.x + a.y * a.y + a.z * a.z); }
inline V3 normalize(V3 a) { return a / length(a); }
inline V3 load3(const float* p) { return V3(p[0], p[1], p[2]); }
inline void store3(float* p, V3 v) { p[0] = v.x; p[1] = v.y; p[2] = v.z; }

const float kInf = std::numeric_limits<float>::infinity();

// ---- Matrix4x4 -------------------------------------------------------------
struct M4 {
  float e[4][4];
  M4() { std::memset(e, 0, sizeof(e)); }
  static M4 identity() {
    M4 m;
    for (int i = 0; i < 4; i++) m.e[i][i] = 1.0f;
    return m;
  }
  static M4 from(const float* p) {
    M4 m;
    std::memcpy(m.e, p, sizeof(m.e));
    return m;
  }
  void to(float* p) const { std::memcpy(p, e, sizeof(e)); }
};
// Matrix4x4::operator* — HW3/src/Matrix4x4.cpp:52-62 (accumulates into 0).
M4 mul(const M4& a, const M4& b) {
  M4 r;
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 4; j++)
      for (int k = 0; k < 4; k++) r.e[i][j] += a.e[i][k] * b.e[k][j];
  return r;
}
M4 transpose(const M4& a) {
  M4 t;
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 4; j++) t.e[i][j] = a.e[j][i];
  return t;
}
// Matrix4x4::multiply — HW3/src/Matrix4x4.cpp:29-50.
V3 mul_point(const M4& m, V3 p) {
  V3 r;
  for (int i = 0; i < 3; i++) {
    r[i] = m.e[i][3];
    for (int j = 0; j < 3; j++) r[i] += m.e[i][j] * p[j];
  }
  return r;
}
bool is_identity(const M4& m) {
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 4; j++)
      if (m.e[i][j] != (i == j ? 1.0f : 0.0f)) return false;
  return true;
}
// Matrix4x4::invert_matrix — HW3/src/Matrix4x4.cpp:97-233: the one copy shared with the device (m4_invert.h)
bool invert(const M4& a, M4& out) {
  float m[16], inv[16];
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 4; j++) m[i * 4 + j] = a.e[i][j];
  if (!b200rt_m4_invert(m, inv)) return false;
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 4; j++) out.e[i][j] = inv[i * 4 + j];
  return true;
}

// ---- Bounding_box ----------------------------------------------------------
struct Box {
  V3 mn, mx, center;
  Box() : mn(kInf), mx(-kInf), center(0.f) {}
  Box(V3 a, V3 b) : mn(a), mx(b), center((b + a) / 2.0f) {}
  // Bounding_box::expand — HW3/src/bounding_box.cpp:5-15
  void expand(const Box& o) {
    mn.x = std::fmin(mn.x, o.mn.x);
    mn.y = std::fmin(mn.y, o.mn.y);
    mn.z = std::fmin(mn.z, o.mn.z);
    mx.x = std::fmax(mx.x, o.mx.x);
    mx.y = std::fmax(mx.y, o.mx.y);
    mx.z = std::fmax(mx.z, o.mx.z);
    center = (mx + mn) / 2.0f;
  }
};
// Bounding_box::apply_transform — HW3/src/bounding_box.cpp:39-113
Box apply_transform(const Box& b, const M4& m) {
  V3 c[8];
  int k = 0;
  for (int ix = 0; ix < 2; ix++)
    for (int iy = 0; iy < 2; iy++)
      for (int iz = 0; iz < 2; iz++)
        c[k++] = mul_point(m, V3(ix ? b.mx.x : b.mn.x, iy ? b.mx.y : b.mn.y, iz ? b.mx.z : b.mn.z));
  V3 mn = c[0], mx = c[0];
  for (int i = 1; i < 8; i++) {
    mn.x = std::fmin(mn.x, c[i].x);
    mn.y = std::fmin(mn.y, c[i].y);
    mn.z = std::fmin(mn.z, c[i].z);
    mx.x = std::fmax(mx.x, c[i].x);
    mx.y = std::fmax(mx.y, c[i].y);
    mx.z = std::fmax(mx.z, c[i].z);
  }
  return Box(mn, mx);
}
Box tri_box(V3 v0, V3 v1, V3 v2) {
  V3 mn = v0, mx = v0;
  mn.x = std::fmin(mn.x, v1.x); mn.y = std::fmin(mn.y, v1.y); mn.z = std::fmin(mn.z, v1.z);
  mx.x = std::fmax(mx.x, v1.x); mx.y = std::fmax(mx.y, v1.y); mx.z = std::fmax(mx.z, v1.z);
  mn.x = std::fmin(mn.x, v2.x); mn.y = std::fmin(mn.y, v2.y); mn.z = std::fmin(mn.z, v2.z);
  mx.x = std::fmax(mx.x, v2.x); mx.y = std::fmax(mx.y, v2.y); mx.z = std::fmax(mx.z, v2.z);
  return Box(mn, mx);
}
void motion_expand(Box& b, V3 vel) {  // HW3/include/Mesh.h:104-110, Sphere.h:30-36
  if (vel != V3(0.0f)) {
    V3 mn = b.mn, mx = b.mx;
    b.expand(Box(mn + vel, mx + vel));
    b.expand(Box(mn - vel, mx - vel));
  }
}

// ---- BVH in the reference's topology ---------------------------------------
struct Obj {
  uint32_t ref;
  Box box;
};
struct BuildOut {
  std::vector<b200rt_bvh_node>* nodes;
  std::vector<uint32_t>* leaf_order;  // refs in DFS order
  std::vector<Box>* gate;             // optional: per leaf (by ref index) the box of the node that holds it
};
void put_box(float* mn, float* mx, const Box& b) {
  store3(mn, b.mn);
  store3(mx, b.mx);
}
// BVH::BVH — HW3/src/Bounding_volume_hierarchy.cpp:3-29.  Returns the node's
// index; `box_out` receives the node's own box (stored by the parent).
//
// A subtree over n >= 2 objects always holds n-1 inner nodes and n leaves, so
// the depth-first slot of every node and leaf is known before the subtree is
// built: `node_pos` / `leaf_pos` are this subtree's first slots in the
// (pre-sized) output arrays.  That makes the two halves of a large range
// independent — they touch disjoint slices of `objs`, `nodes`, `leaf_order` and
// `gate` — and the left half is handed to another thread for the top
// `par_levels` levels.  The result is the same array, byte for byte, as the
// sequential recursion's.
constexpr int kParMinObjs = 1 << 14;
uint32_t build_bvh_at(std::vector<Obj>& objs, int start, int end, int dim, BuildOut& out, Box& box_out,
                      uint32_t node_pos, uint32_t leaf_pos, int par_levels) {
  Box box;
  for (int i = start; i < end; i++) box.expand(objs[i].box);
  const float center = box.center[dim];
  int mid = start;
  for (int i = start; i < end; i++) {
    if (objs[i].box.center[dim] < center) {
      std::swap(objs[i], objs[mid]);
      mid++;
    }
  }
  if (mid == start || mid == end) mid = start + ((end - start) / 2);
  const uint32_t me = node_pos;
  const int nl = mid - start, nr = end - mid;
  const uint32_t left_nodes = nl > 1 ? (uint32_t)(nl - 1) : 0u;
  b200rt_bvh_node n;
  std::memset(&n, 0, sizeof(n));
  Box lb, rb;
  std::future<uint32_t> left_task;
  if (nl == 1) {
    n.left = objs[start].ref;
    lb = objs[start].box;
    (*out.leaf_order)[leaf_pos] = objs[start].ref;
    if (out.gate) (*out.gate)[B200RT_REF_INDEX(objs[start].ref)] = box;
  } else if (par_levels > 0 && end - start >= kParMinObjs) {
    try {
      left_task = std::async(std::launch::async, [&objs, &out, &lb, start, mid, dim, me, leaf_pos, par_levels]() {
        return build_bvh_at(objs, start, mid, (dim + 1) % 3, out, lb, me + 1, leaf_pos, par_levels - 1);
      });
    } catch (const std::system_error&) {  // no thread to be had: same slots, built here
      uint32_t c = build_bvh_at(objs, start, mid, (dim + 1) % 3, out, lb, me + 1, leaf_pos, 0);
      n.left = B200RT_REF(B200RT_KIND_INNER, c);
    }
  } else {
    uint32_t c = build_bvh_at(objs, start, mid, (dim + 1) % 3, out, lb, me + 1, leaf_pos, 0);
    n.left = B200RT_REF(B200RT_KIND_INNER, c);
  }
  if (nr == 1) {
    n.right = objs[mid].ref;
    rb = objs[mid].box;
    (*out.leaf_order)[leaf_pos + (uint32_t)nl] = objs[mid].ref;
    if (out.gate) (*out.gate)[B200RT_REF_INDEX(objs[mid].ref)] = box;
  } else {
    uint32_t c = build_bvh_at(objs, mid, end, (dim + 1) % 3, out, rb, me + 1 + left_nodes, leaf_pos + (uint32_t)nl,
                              par_levels > 0 ? par_levels - 1 : 0);
    n.right = B200RT_REF(B200RT_KIND_INNER, c);
  }
  if (left_task.valid()) n.left = B200RT_REF(B200RT_KIND_INNER, left_task.get());
  put_box(n.lmin, n.lmax, lb);
  put_box(n.rmin, n.rmax, rb);
  (*out.nodes)[me] = n;
  box_out = box;
  return me;
}
int build_par_levels() {
  if (const char* e = std::getenv("B200RT_BUILD_THREADS")) {
    int t = std::atoi(e), l = 0;
    while ((1 << l) < t) l++;
    return l;
  }
  unsigned hc = std::thread::hardware_concurrency();
  int l = 0;
  while ((1u << l) < hc && l < 6) l++;
  return l;
}
// Appends the tree over objs[start,end) (end - start >= 2) to out.nodes / out.leaf_order.
uint32_t build_bvh(std::vector<Obj>& objs, int start, int end, int dim, BuildOut& out, Box& box_out) {
  const uint32_t node_pos = (uint32_t)out.nodes->size();
  const uint32_t leaf_pos = (uint32_t)out.leaf_order->size();
  out.nodes->resize((size_t)node_pos + (size_t)(end - start - 1));
  out.leaf_order->resize((size_t)leaf_pos + (size_t)(end - start));
  return build_bvh_at(objs, start, end, dim, out, box_out, node_pos, leaf_pos, build_par_levels());
}

struct MeshIn {
  int material, smooth;
  std::vector<int32_t> faces;
  M4 base_xf;
  V3 velocity;
  int texture = -1, texture_offset = 0;
  bool is_light = false;
  V3 radiance;
};
struct InstIn {
  int mesh, material;
  M4 xf;
  V3 velocity;
  bool direct;
  bool is_light = false;
};
struct TriIn {
  int material;
  int32_t idx[3];
  M4 xf;
};
struct SphIn {
  int material, center_vertex;
  float radius;
  M4 xf;
  V3 velocity;
  int texture = -1;
  bool is_light = false;
  V3 radiance;
};

void fill_xf(const M4& m, float inv12[12], float nrm9[9], float fwd16[16], bool& ok) {
  // Arbitrary_transformation — HW3/src/Transformation.cpp:73-84
  M4 inv;
  ok = invert(m, inv);
  if (!ok) return;
  M4 nrm = transpose(inv);
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 4; j++) inv12[i * 4 + j] = inv.e[i][j];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) nrm9[i * 3 + j] = nrm.e[i][j];
  if (fwd16) m.to(fwd16);
}

}  // namespace

struct b200rt_builder {
  int semantics;
  B200SceneDesc desc;
  std::vector<float> verts;
  std::vector<b200rt_camera> cameras;
  std::vector<b200rt_material> materials;
  std::vector<b200rt_light> lights;
  std::vector<MeshIn> meshes_in;
  std::vector<InstIn> inst_in;
  std::vector<InstIn> light_inst_in;   // LightMesh objects: after the instances (HW7/src/Scene.cpp:1246-1311)
  std::vector<SphIn> light_sph_in;     // LightSphere objects: after the triangles (HW7/src/Scene.cpp:1366-1424)
  std::vector<float> texcoords;
  std::vector<b200rt_texture> textures;
  std::vector<float> texels;
  std::vector<b200rt_brdf> brdfs;
  std::vector<float> tri_uv, light_cdf;
  std::vector<int32_t> tri_tex, light_tri;
  std::vector<TriIn> tri_in;
  std::vector<SphIn> sph_in;
  // flattened
  std::vector<b200rt_bvh_node> nodes;
  std::vector<float> tri_geom;
  std::vector<int32_t> tri_vidx, tri_info;
  std::vector<float> vert_normal;
  std::vector<b200rt_bvh_node> fast_nodes;
  std::vector<float> fast_tri_geom, tri_gate;
  std::vector<int32_t> fast_tri_ref;
  std::vector<b200rt_mesh> meshes;
  std::vector<b200rt_instance> instances;
  std::vector<b200rt_sphere> spheres;
  std::vector<b200rt_triprim> triprims;
  size_t n_mesh_nodes, n_user_lights;   // what b->nodes / b->lights held before the top-level phase (refit truncates to it)
  bool finished;
};

namespace {

// Second half of b200rt_builder_finish, and the whole of b200rt_builder_refit: the top-level objects with their
// transforms and world boxes, Scene::bvh over them, and the objects' ranks.  Mesh trees are not touched.
int finish_top(b200rt_builder* b) {
  const int nv = (int)(b->verts.size() / 3);
  auto vert = [&](int i) { return load3(&b->verts[3 * (size_t)i]); };
  auto check_idx = [&](int i) { return i >= 0 && i < nv; };
  const bool hw1 = b->semantics == B200RT_SEM_HW1;
  // ---- top-level objects in the reference's push order ----------------------
  std::vector<Obj> top;
  b->instances.clear(); b->spheres.clear(); b->triprims.clear();
  std::vector<InstIn> all_inst(b->inst_in);
  all_inst.insert(all_inst.end(), b->light_inst_in.begin(), b->light_inst_in.end());
  for (const InstIn& in : all_inst) {
    b200rt_instance fi;
    std::memset(&fi, 0, sizeof(fi));
    fi.light = -1;
    const b200rt_mesh& fm = b->meshes[in.mesh];
    Box mb(load3(fm.root_box), load3(fm.root_box + 3));
    Box ib;
    if (in.is_light) {
      // Light_mesh sits in Scene::bvh itself; its box is the OBJECT-space mesh box
      // and it has no gate of its own (HW7/include/Light_mesh.h:52-68).
      bool ok;
      fill_xf(in.xf, fi.inv, fi.nrm, fi.fwd, ok);
      if (!ok) {
        g_host_error = "NOT INVERTIBLE MATRIX (light mesh)";
        return B200RT_E_SINGULAR;
      }
      ib = mb;
      fi.flags |= B200RT_INST_LIGHT;
      // Light_mesh ctor: areas of the TRANSFORMED triangles in the (BVH-reordered) vector order
      const MeshIn& lm = b->meshes_in[in.mesh];
      b200rt_light L;
      std::memset(&L, 0, sizeof(L));
      L.kind = B200RT_LIGHT_MESH;
      store3(L.intensity, lm.radiance);
      L.object = (int)b->instances.size();
      L.cdf_begin = (int)b->light_cdf.size();
      L.cdf_count = fm.tri_count;
      float total = 0.0f;
      std::vector<float> areas;
      for (int k = 0; k < fm.tri_count; k++) {
        const int32_t* vi = &b->tri_vidx[3 * (size_t)(fm.tri_begin + k)];
        const V3 v0 = mul_point(in.xf, vert(vi[0])), v1 = mul_point(in.xf, vert(vi[1])), v2 = mul_point(in.xf, vert(vi[2]));
        const float area = length(cross(v1 - v0, v2 - v0)) / 2;
        total += area;
        areas.push_back(area);
      }
      float cum = 0.0f;
      for (int k = 0; k < fm.tri_count; k++) {
        cum += areas[k];
        b->light_cdf.push_back(cum / total);
        b->light_tri.push_back(fm.tri_begin + k);
      }
      L.total_area = total;
      fi.light = (int)b->lights.size();
      b->lights.push_back(L);
    } else if (in.direct) {
      ib = mb;  // HW2: Mesh::get_bounding_box = bvh box (HW2/Mesh.h:17-19)
      M4 id = M4::identity();
      bool ok;
      fill_xf(id, fi.inv, fi.nrm, fi.fwd, ok);
      fi.flags |= B200RT_INST_DIRECT;
    } else {
      bool ok;
      fill_xf(in.xf, fi.inv, fi.nrm, fi.fwd, ok);
      if (!ok) {
        g_host_error = "NOT INVERTIBLE MATRIX (mesh instance)";
        return B200RT_E_SINGULAR;
      }
      ib = apply_transform(mb, in.xf);  // Mesh_instance ctor HW3/include/Mesh.h:101-111
      motion_expand(ib, in.velocity);
      if (in.velocity != V3(0.f)) fi.flags |= B200RT_INST_MOVING;
    }
    put_box(fi.box, fi.box + 3, ib);
    store3(fi.velocity, in.velocity);
    fi.mesh = in.mesh;
    fi.material = in.material;
    fi.texture = b->meshes_in[in.mesh].texture;  // instances take the base mesh's texture (HW7/src/Scene.cpp:1182)
    if (b->semantics >= B200RT_SEM_HW7 && in.material >= 0 && in.material < (int)b->materials.size()) {
      const float* tr = b->materials[in.material].transparency;
      const float* mr = b->materials[in.material].mirror;
      bool refr = (tr[0] != 0.f || tr[1] != 0.f || tr[2] != 0.f);
      // PPM: material_type == mt_refractive, and mirror wins over transparency (PPM/src/Material.cpp:55-61)
      if (b->semantics == B200RT_SEM_PPM && (mr[0] != 0.f || mr[1] != 0.f || mr[2] != 0.f)) refr = false;
      if (refr) fi.flags |= B200RT_INST_REFRACTIVE;
    }
    Obj o;
    o.ref = B200RT_REF(in.is_light ? B200RT_KIND_LIGHTINST : B200RT_KIND_INSTANCE, b->instances.size());
    o.box = ib;
    top.push_back(o);
    b->instances.push_back(fi);
  }
  if (!hw1) {
    for (const TriIn& t : b->tri_in) {
      if (!check_idx(t.idx[0]) || !check_idx(t.idx[1]) || !check_idx(t.idx[2])) {
        g_host_error = "triangle index out of range";
        return B200RT_E_INVALID;
      }
      b200rt_triprim ft;
      std::memset(&ft, 0, sizeof(ft));
      const V3 v0 = vert(t.idx[0]), v1 = vert(t.idx[1]), v2 = vert(t.idx[2]);
      store3(ft.v0, v0); store3(ft.a1, v0 - v1); store3(ft.a2, v0 - v2);
      store3(ft.normal, normalize(cross(v1 - v0, v2 - v0)));
      Box tb = tri_box(v0, v1, v2);
      const bool ident = is_identity(t.xf);  // Triangle ctor HW3/src/Triangle.cpp:20-41
      bool ok;
      fill_xf(t.xf, ft.inv, ft.nrm, nullptr, ok);
      if (!ok) {
        g_host_error = "NOT INVERTIBLE MATRIX (triangle)";
        return B200RT_E_SINGULAR;
      }
      if (ident) ft.flags |= B200RT_XF_IDENTITY; else tb = apply_transform(tb, t.xf);
      ft.material = t.material;
      Obj o;
      o.ref = B200RT_REF(B200RT_KIND_TRIPRIM, b->triprims.size());
      o.box = tb;
      top.push_back(o);
      b->triprims.push_back(ft);
    }
  }
  std::vector<SphIn> all_sph(b->light_sph_in);
  all_sph.insert(all_sph.end(), b->sph_in.begin(), b->sph_in.end());
  for (const SphIn& s : all_sph) {
    if (!check_idx(s.center_vertex)) {
      g_host_error = "sphere centre index out of range";
      return B200RT_E_INVALID;
    }
    b200rt_sphere fs;
    std::memset(&fs, 0, sizeof(fs));
    const V3 c = vert(s.center_vertex);
    store3(fs.center, c);
    fs.radius = s.radius;
    const bool ident = is_identity(s.xf);
    bool ok;
    fill_xf(s.xf, fs.inv, fs.nrm, fs.fwd, ok);
    if (!ok) {
      g_host_error = "NOT INVERTIBLE MATRIX (sphere)";
      return B200RT_E_SINGULAR;
    }
    const V3 delta(s.radius);
    Box sb(c - delta, c + delta);  // Sphere ctor HW3/include/Sphere.h:18-37
    if (ident) fs.flags |= B200RT_XF_IDENTITY; else sb = apply_transform(sb, s.xf);
    motion_expand(sb, s.velocity);
    if (s.velocity != V3(0.f)) fs.flags |= B200RT_XF_MOVING;
    store3(fs.velocity, s.velocity);
    fs.material = s.material;
    fs.texture = s.texture;
    fs.light = -1;
    if (s.is_light) {
      fs.flags |= B200RT_XF_LIGHT;
      b200rt_light L;
      std::memset(&L, 0, sizeof(L));
      L.kind = B200RT_LIGHT_SPHERE;
      store3(L.intensity, s.radiance);
      L.object = (int)b->spheres.size();
      fs.light = (int)b->lights.size();
      b->lights.push_back(L);
    }
    Obj o;
    o.ref = B200RT_REF(B200RT_KIND_SPHERE, b->spheres.size());
    o.box = sb;
    top.push_back(o);
    b->spheres.push_back(fs);
  }

  if (b->semantics == B200RT_SEM_PPM) {
    // PPM loads spheres before meshes (PPM/src/Scene.cpp:487-498): objects = spheres, base instances, instances
    std::vector<Obj> reordered;
    for (const Obj& o : top)
      if (B200RT_REF_KIND(o.ref) == B200RT_KIND_SPHERE) reordered.push_back(o);
    for (const Obj& o : top)
      if (B200RT_REF_KIND(o.ref) != B200RT_KIND_SPHERE) reordered.push_back(o);
    top.swap(reordered);
  }
  // ---- Scene::bvh = BVH::create_bvh(objects) — HW3/src/Scene.cpp:871 --------
  Box root_box;
  std::vector<uint32_t> order;
  if (hw1 || top.empty()) {
    b->desc.root = B200RT_REF_NONE;
    for (const Obj& o : top) order.push_back(o.ref);
  } else if (top.size() == 1) {
    b->desc.root = top[0].ref;
    root_box = top[0].box;
    order.push_back(top[0].ref);
  } else {
    BuildOut bo{&b->nodes, &order, nullptr};
    uint32_t r = build_bvh(top, 0, (int)top.size(), 0, bo, root_box);
    b->desc.root = B200RT_REF(B200RT_KIND_INNER, r);
  }
  put_box(b->desc.root_box, b->desc.root_box + 3, root_box);
  for (size_t k = 0; k < order.size(); k++) {
    const uint32_t r = order[k];
    const uint32_t idx = B200RT_REF_INDEX(r);
    switch (B200RT_REF_KIND(r)) {
      case B200RT_KIND_LIGHTINST:
      case B200RT_KIND_INSTANCE: b->instances[idx].rank = (int)k; break;
      case B200RT_KIND_SPHERE: b->spheres[idx].rank = (int)k; break;
      case B200RT_KIND_TRIPRIM: b->triprims[idx].rank = (int)k; break;
      default: break;
    }
  }

  return B200RT_OK;
}

void fill_desc(b200rt_builder* b) {
  const int nv = (int)(b->verts.size() / 3);
  B200SceneDesc& d = b->desc;
  d.n_nodes = (int)b->nodes.size();           d.nodes = b->nodes.data();
  d.n_tris = (int)(b->tri_info.size() / 2);   d.tri_geom = b->tri_geom.data();
  d.tri_vidx = b->tri_vidx.data();            d.tri_info = b->tri_info.data();
  d.n_fast_nodes = (int)b->fast_nodes.size(); d.fast_nodes = b->fast_nodes.data();
  d.n_fast_tris = (int)b->fast_tri_ref.size();
  d.fast_tri_geom = b->fast_tri_geom.data();  d.fast_tri_ref = b->fast_tri_ref.data();
  d.tri_gate = b->tri_gate.data();
  d.n_verts = nv;                             d.vert_pos = b->verts.data();
  d.vert_normal = b->vert_normal.data();
  d.n_meshes = (int)b->meshes.size();         d.meshes = b->meshes.data();
  d.n_instances = (int)b->instances.size();   d.instances = b->instances.data();
  d.n_spheres = (int)b->spheres.size();       d.spheres = b->spheres.data();
  d.n_triprims = (int)b->triprims.size();     d.triprims = b->triprims.data();
  d.n_materials = (int)b->materials.size();   d.materials = b->materials.data();
  d.n_lights = (int)b->lights.size();         d.lights = b->lights.data();
  d.n_cameras = (int)b->cameras.size();       d.cameras = b->cameras.data();
  d.tri_uv = b->tri_uv.data();                d.tri_tex = b->tri_tex.data();
  d.n_textures = (int)b->textures.size();     d.textures = b->textures.data();
  d.n_texels = (int64_t)b->texels.size();     d.texels = b->texels.data();
  d.n_brdfs = (int)b->brdfs.size();           d.brdfs = b->brdfs.data();
  d.n_light_cdf = (int)b->light_cdf.size();   d.light_cdf = b->light_cdf.data();
  d.light_tri = b->light_tri.data();
}

}  // namespace

extern "C" {

const char* b200rt_last_error(void) { return g_host_error.c_str(); }
// internal (not part of the ABI): lets the CUDA translation units report errors
void b200rt_internal_set_error(const char* msg) { g_host_error = msg ? msg : ""; }

void b200rt_xf_identity(float out[16]) { M4::identity().to(out); }
// Scaling — HW3/src/Transformation.cpp:6-17 (forward matrix only; inverses are
// always recomputed by Arbitrary_transformation on the composed product).
void b200rt_xf_scaling(float x, float y, float z, float out[16]) {
  M4 m;
  m.e[0][0] = x; m.e[1][1] = y; m.e[2][2] = z; m.e[3][3] = 1.0f;
  m.to(out);
}
void b200rt_xf_translation(float x, float y, float z, float out[16]) {
  M4 m = M4::identity();
  m.e[0][3] = x; m.e[1][3] = y; m.e[2][3] = z;
  m.to(out);
}
// Rotation — HW3/src/Transformation.cpp:35-70
void b200rt_xf_rotation(float angle_deg, float x, float y, float z, float out[16]) {
  const float degree_to_pi = M_PI / 180.0f;  // HW3/src/Scene.cpp:322
  const float angle = angle_deg * degree_to_pi;
  const V3 u = normalize(V3(x, y, z));
  const V3 v = normalize((x != 0.0f || y != 0.0f) ? V3(-u.y, u.x, 0.0f) : V3(0.0f, 1.0f, 0.0f));
  const V3 w = cross(u, v);
  M4 m;
  m.e[0][0] = u.x; m.e[0][1] = u.y; m.e[0][2] = u.z;
  m.e[1][0] = v.x; m.e[1][1] = v.y; m.e[1][2] = v.z;
  m.e[2][0] = w.x; m.e[2][1] = w.y; m.e[2][2] = w.z;
  m.e[3][3] = 1.0f;
  M4 rot;
  rot.e[0][0] = 1.0f;
  rot.e[1][1] = std::cos(angle);
  rot.e[1][2] = -std::sin(angle);
  rot.e[2][1] = -rot.e[1][2];
  rot.e[2][2] = rot.e[1][1];
  rot.e[3][3] = 1.0f;
  M4 mt = transpose(m);
  mul(mt, mul(rot, m)).to(out);
}
void b200rt_xf_mul(const float a[16], const float b[16], float out[16]) {
  mul(M4::from(a), M4::from(b)).to(out);
}
int b200rt_xf_invert(const float m[16], float out[16]) try {
  M4 inv;
  if (!invert(M4::from(m), inv)) return B200RT_E_SINGULAR;
  inv.to(out);
  return B200RT_OK;
} B200RT_CATCH

b200rt_builder* b200rt_builder_create(int semantics) {
  if (semantics < B200RT_SEM_HW1 || semantics > B200RT_SEM_PPM) return nullptr;
  b200rt_builder* b = new b200rt_builder();
  b->semantics = semantics;
  std::memset(&b->desc, 0, sizeof(b->desc));
  b->desc.semantics = semantics;
  b->desc.shadow_eps = 0.001f;
  b->desc.isect_eps = (semantics == B200RT_SEM_PPM) ? 0.0001f : 0.000001f;
  b->desc.background_texture = -1;
  b->desc.env_light = -1;
  b->finished = false;
  return b;
}
void b200rt_builder_destroy(b200rt_builder* b) { delete b; }

int b200rt_builder_set_globals(b200rt_builder* b, const float bg[3], const float amb[3],
                               float shadow_eps, int max_depth) try {
  if (!b) return B200RT_E_INVALID;
  std::memcpy(b->desc.background, bg, 12);
  std::memcpy(b->desc.ambient, amb, 12);
  b->desc.shadow_eps = shadow_eps;
  b->desc.max_depth = max_depth;
  if (b->semantics == B200RT_SEM_PPM && b->desc.max_depth > 20) b->desc.max_depth = 20;  // PPM/src/Scene.cpp:426-429
  return B200RT_OK;
} B200RT_CATCH

// Camera::Camera — HW3/include/Camera.h:11-34 (HW1/Camera.h:11-23 has no
// s_u/s_v constants and no DOF; the fields are still filled for uniformity).
int b200rt_builder_add_camera(b200rt_builder* b, const b200rt_camera_in* in) try {
  if (!b || !in) return B200RT_E_INVALID;
  b200rt_camera c;
  std::memset(&c, 0, sizeof(c));
  const V3 e = load3(in->position), up = load3(in->up);
  V3 gaze = load3(in->gaze);
  float left = in->near_plane[0], right = in->near_plane[1], bottom = in->near_plane[2],
        top = in->near_plane[3], distance = in->near_distance;
  if (b->semantics >= B200RT_SEM_HW4 && in->simple) {
    // HW7/src/Scene.cpp:590-607
    const float degrees_to_radians = (float)(M_PI / 180.0f);
    const float half_y_radian = degrees_to_radians * in->fov_y / 2;
    top = tanf(half_y_radian) * distance;
    const float aspect_ratio = 1.0f * in->width / in->height;
    bottom = -1.0f * top;
    right = top * aspect_ratio;
    left = -1.0f * right;
    gaze = normalize(gaze - e);
  }
  const V3 w = -(normalize(gaze));
  V3 u, v;
  if (b->semantics >= B200RT_SEM_HW7 && in->left_handed) {
    u = normalize(cross(w, normalize(up)));
    v = normalize(cross(u, w));
  } else {
    u = normalize(cross(normalize(up), w));
    v = normalize(cross(w, u));
  }
  if (b->semantics >= B200RT_SEM_HW3 && in->aperture_size != 0.0f) {
    const float ratio = in->focus_distance / distance;
    left = ratio * left; right = ratio * right; bottom = ratio * bottom; top = ratio * top;
    distance = in->focus_distance;
  }
  const V3 tl = e - w * distance + left * u + top * v;
  const V3 su = ((right - left) / in->width) * u;
  const V3 sv = ((top - bottom) / in->height) * v;
  store3(c.e, e); store3(c.u, u); store3(c.v, v); store3(c.w, w);
  store3(c.top_left, tl); store3(c.su, su); store3(c.sv, sv);
  c.left = left; c.right = right; c.bottom = bottom; c.top = top; c.distance = distance;
  c.width = in->width; c.height = in->height;
  int s = (int)std::sqrt(in->num_samples);  // HW3/src/Scene.cpp:412 (int -> double sqrt)
  if (s <= 0) s = 1;
  c.samples_sqrt = (b->semantics == B200RT_SEM_HW1) ? 1 : s;
  if (b->semantics == B200RT_SEM_PPM && c.samples_sqrt > 2) c.samples_sqrt = 2;  // PPM/src/Scene.cpp:268
  c.aperture = (b->semantics >= B200RT_SEM_HW3) ? in->aperture_size : 0.0f;
  b->cameras.push_back(c);
  return (int)b->cameras.size() - 1;
} B200RT_CATCH

int b200rt_builder_add_material(b200rt_builder* b, const b200rt_material* m) try {
  if (!b || !m) return B200RT_E_INVALID;
  b200rt_material c = *m;
  if (b->semantics < B200RT_SEM_HW6) c.brdf = -1;
  if (b->semantics >= B200RT_SEM_HW5 && c.degamma) {  // from HW5 on: HW5/src/Scene.cpp:338-348
    // HW7/src/Scene.cpp:951-961: the unqualified pow(float, 2.2f) there is the double pow (probed with the tree's
    // own include list), rounded to float by the assignment
    for (int i = 0; i < 3; i++) {
      c.ambient[i] = (float)std::pow((double)c.ambient[i], (double)2.2f);
      c.diffuse[i] = (float)std::pow((double)c.diffuse[i], (double)2.2f);
      c.specular[i] = (float)std::pow((double)c.specular[i], (double)2.2f);
    }
  }
  b->materials.push_back(c);
  return (int)b->materials.size() - 1;
} B200RT_CATCH
int b200rt_builder_add_light(b200rt_builder* b, const b200rt_light* l) try {
  if (!b || !l) return B200RT_E_INVALID;
  b200rt_light c = *l;
  if (c.kind == B200RT_LIGHT_AREA) {
    // HW3/src/Scene.cpp:484
    store3(c.normal, normalize(cross(load3(c.edge1), load3(c.edge2))));
  }
  if (c.kind == B200RT_LIGHT_SPOT) {
    // Spot_light ctor — HW7/src/Spot_light.cpp:5-12.  In: edge1 = direction, edge2 = {coverage, falloff} in DEGREES
    const float d2r = (float)(M_PI / 180.0f);
    const float cov = c.edge2[0] * d2r, fall = c.edge2[1] * d2r;
    store3(c.edge1, normalize(load3(c.edge1)));
    c.edge2[0] = std::cos(cov / 2);
    c.edge2[1] = std::cos(fall / 2);
    c.edge2[2] = 0.f;
  }
  if (c.kind == B200RT_LIGHT_ENV) b->desc.env_light = (int)b->lights.size();
  b->lights.push_back(c);
  return (int)b->lights.size() - 1;
} B200RT_CATCH
int b200rt_builder_set_vertices(b200rt_builder* b, const float* xyz, int n) try {
  if (!b || (n > 0 && !xyz)) return B200RT_E_INVALID;
  b->verts.assign(xyz, xyz + 3 * (size_t)n);
  return B200RT_OK;
} B200RT_CATCH
int b200rt_builder_add_mesh(b200rt_builder* b, int material, int smooth, const int32_t* faces,
                            int n_faces, const float base_xf[16], const float velocity[3]) try {
  if (!b || n_faces < 0 || (n_faces > 0 && !faces)) return B200RT_E_INVALID;
  MeshIn m;
  m.material = material;
  m.smooth = smooth;
  m.faces.assign(faces, faces + 3 * (size_t)n_faces);
  m.base_xf = base_xf ? M4::from(base_xf) : M4::identity();
  m.velocity = velocity ? load3(velocity) : V3(0.f);
  b->meshes_in.push_back(std::move(m));
  const int id = (int)b->meshes_in.size() - 1;
  if (b->semantics != B200RT_SEM_HW1) {
    InstIn i;
    i.mesh = id;
    i.material = material;
    i.xf = b->meshes_in[id].base_xf;
    i.velocity = b->meshes_in[id].velocity;
    i.direct = (b->semantics == B200RT_SEM_HW2);
    b->inst_in.push_back(i);  // base instances precede explicit ones as long as
                              // all meshes are added first (the loader does so).
  }
  return id;
} B200RT_CATCH
int b200rt_builder_add_instance(b200rt_builder* b, int base_mesh, int material,
                                const float xf[16], const float velocity[3]) try {
  if (!b || base_mesh < 0 || base_mesh >= (int)b->meshes_in.size() || !xf) return B200RT_E_INVALID;
  if (b->semantics < B200RT_SEM_HW3) return B200RT_E_UNSUPPORTED;
  InstIn i;
  i.mesh = base_mesh;
  i.material = material;
  i.xf = M4::from(xf);
  i.velocity = velocity ? load3(velocity) : V3(0.f);
  i.direct = false;
  b->inst_in.push_back(i);
  return (int)b->inst_in.size() - 1;
} B200RT_CATCH
int b200rt_builder_add_triangle(b200rt_builder* b, int material, const int32_t idx[3],
                                const float xf[16]) try {
  if (!b || !idx) return B200RT_E_INVALID;
  TriIn t;
  t.material = material;
  std::memcpy(t.idx, idx, 12);
  t.xf = xf ? M4::from(xf) : M4::identity();
  b->tri_in.push_back(t);
  return (int)b->tri_in.size() - 1;
} B200RT_CATCH
int b200rt_builder_add_sphere(b200rt_builder* b, int material, int center_vertex, float radius,
                              const float xf[16], const float velocity[3]) try {
  if (!b) return B200RT_E_INVALID;
  SphIn s;
  s.material = material;
  s.center_vertex = center_vertex;
  s.radius = radius;
  s.xf = xf ? M4::from(xf) : M4::identity();
  s.velocity = velocity ? load3(velocity) : V3(0.f);
  b->sph_in.push_back(s);
  return (int)b->sph_in.size() - 1;
} B200RT_CATCH

int b200rt_builder_set_texcoords(b200rt_builder* b, const float* uv, int n) try {
  if (!b || (n > 0 && !uv)) return B200RT_E_INVALID;
  b->texcoords.assign(uv, uv + 2 * (size_t)n);
  return B200RT_OK;
} B200RT_CATCH
int b200rt_builder_add_texture(b200rt_builder* b, const b200rt_texture* t, const float* texels) try {
  if (!b || !t) return B200RT_E_INVALID;
  b200rt_texture c = *t;
  c.texel_offset = (int64_t)b->texels.size();
  if (c.kind == B200RT_TEX_IMAGE) {
    if (!texels || c.width <= 0 || c.height <= 0) return B200RT_E_INVALID;
    b->texels.insert(b->texels.end(), texels, texels + (size_t)c.width * c.height * 3);
  }
  b->textures.push_back(c);
  return (int)b->textures.size() - 1;
} B200RT_CATCH
int b200rt_builder_add_brdf(b200rt_builder* b, const b200rt_brdf* r) try {
  if (!b || !r) return B200RT_E_INVALID;
  b->brdfs.push_back(*r);
  return (int)b->brdfs.size() - 1;
} B200RT_CATCH
int b200rt_builder_set_mesh_texture(b200rt_builder* b, int mesh, int texture, int texture_offset) try {
  if (!b || mesh < 0 || mesh >= (int)b->meshes_in.size()) return B200RT_E_INVALID;
  b->meshes_in[mesh].texture = texture;
  b->meshes_in[mesh].texture_offset = texture_offset;
  return B200RT_OK;
} B200RT_CATCH
int b200rt_builder_set_sphere_texture(b200rt_builder* b, int sphere, int texture) try {
  if (!b || sphere < 0 || sphere >= (int)b->sph_in.size()) return B200RT_E_INVALID;
  b->sph_in[sphere].texture = texture;
  return B200RT_OK;
} B200RT_CATCH
int b200rt_builder_add_light_mesh(b200rt_builder* b, int material, const int32_t* faces, int n_faces,
                                  const float xf[16], const float radiance[3]) try {
  if (!b || n_faces <= 0 || !faces || !radiance) return B200RT_E_INVALID;
  if (b->semantics < B200RT_SEM_HW7) return B200RT_E_UNSUPPORTED;
  MeshIn m;
  m.material = material;
  m.smooth = 0;
  m.faces.assign(faces, faces + 3 * (size_t)n_faces);
  m.base_xf = xf ? M4::from(xf) : M4::identity();
  m.velocity = V3(0.f);
  m.is_light = true;
  m.radiance = load3(radiance);
  b->meshes_in.push_back(std::move(m));
  InstIn i;
  i.mesh = (int)b->meshes_in.size() - 1;
  i.material = material;
  i.xf = b->meshes_in.back().base_xf;
  i.velocity = V3(0.f);
  i.direct = false;
  i.is_light = true;
  b->light_inst_in.push_back(i);
  return i.mesh;
} B200RT_CATCH
int b200rt_builder_add_light_sphere(b200rt_builder* b, int material, int center_vertex, float radius,
                                    const float xf[16], const float radiance[3]) try {
  if (!b || !radiance) return B200RT_E_INVALID;
  if (b->semantics < B200RT_SEM_HW7) return B200RT_E_UNSUPPORTED;
  SphIn s;
  s.material = material;
  s.center_vertex = center_vertex;
  s.radius = radius;
  s.xf = xf ? M4::from(xf) : M4::identity();
  s.velocity = V3(0.f);
  s.is_light = true;
  s.radiance = load3(radiance);
  b->light_sph_in.push_back(s);
  return (int)b->light_sph_in.size() - 1;
} B200RT_CATCH
int b200rt_builder_set_integrator(b200rt_builder* b, int path_tracing, int importance_sampling) try {
  if (!b) return B200RT_E_INVALID;
  b->desc.integrator = path_tracing ? 1 : 0;
  b->desc.importance_sampling = importance_sampling ? 1 : 0;
  return B200RT_OK;
} B200RT_CATCH
int b200rt_builder_set_background_texture(b200rt_builder* b, int texture) try {
  if (!b) return B200RT_E_INVALID;
  b->desc.background_texture = texture;
  return B200RT_OK;
} B200RT_CATCH

int b200rt_builder_finish(b200rt_builder* b, const B200SceneDesc** out) try {
  if (!b || !out) return B200RT_E_INVALID;
  if (b->finished) {
    *out = &b->desc;
    return B200RT_OK;
  }
  const int nv = (int)(b->verts.size() / 3);
  auto vert = [&](int i) { return load3(&b->verts[3 * (size_t)i]); };
  auto check_idx = [&](int i) { return i >= 0 && i < nv; };
  const bool hw1 = b->semantics == B200RT_SEM_HW1;

  // ---- mesh triangles: records, vertex-normal accumulation, per-mesh BVH ----
  std::vector<V3> vnorm((size_t)nv, V3(0.f));
  std::vector<char> vhas((size_t)nv, 0);
  b->meshes.clear();
  for (size_t mi = 0; mi < b->meshes_in.size(); mi++) {
    const MeshIn& m = b->meshes_in[mi];
    const int nf = (int)(m.faces.size() / 3);
    const auto T0 = std::chrono::steady_clock::now();
    const auto secs = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) {
      return std::chrono::duration<double>(b - a).count();
    };
    double t_faces = 0.0, t_tree = 0.0;
    std::vector<Obj> objs((size_t)nf);
    std::vector<V3> normals((size_t)nf);
    for (int f = 0; f < nf; f++) {
      const int i0 = m.faces[3 * f], i1 = m.faces[3 * f + 1], i2 = m.faces[3 * f + 2];
      if (!check_idx(i0) || !check_idx(i1) || !check_idx(i2)) {
        g_host_error = "mesh face index out of range";
        return B200RT_E_INVALID;
      }
      const V3 v0 = vert(i0), v1 = vert(i1), v2 = vert(i2);
      const V3 c = cross(v1 - v0, v2 - v0);
      const V3 n = normalize(c);                       // Mesh_triangle.cpp:17
      normals[f] = n;
      objs[f].ref = B200RT_REF(B200RT_KIND_TRI, f);    // local face id for now
      objs[f].box = tri_box(v0, v1, v2);
      if (!hw1) {
        const float area = length(c) / 2;              // get_surface_area, Mesh_triangle.cpp:70-75
        // Vertex::add_vertex_normal — HW3/src/Scene.cpp:694-696
        vnorm[i0] = vnorm[i0] + n * area; vhas[i0] = 1;
        vnorm[i1] = vnorm[i1] + n * area; vhas[i1] = 1;
        vnorm[i2] = vnorm[i2] + n * area; vhas[i2] = 1;
      }
    }
    b200rt_mesh fm;
    std::memset(&fm, 0, sizeof(fm));
    fm.tri_begin = (int)(b->tri_info.size() / 2);
    fm.tri_count = nf;
    fm.node_begin = (int)b->nodes.size();
    std::vector<uint32_t> order;
    std::vector<Box> gate((size_t)nf);
    Box root_box;
    fm.fast_root = B200RT_REF_NONE;
    if (hw1 || nf == 0) {
      // HW1 has no BVH (HW1/Scene.cpp:35-54): keep file order.
      for (int f = 0; f < nf; f++) order.push_back(B200RT_REF(B200RT_KIND_TRI, f));
      fm.root = B200RT_REF_NONE;
    } else if (nf == 1) {
      order.push_back(objs[0].ref);  // BVH::create_bvh size==1 — Bounding_volume_hierarchy.h:13-14
      root_box = objs[0].box;
      fm.root = B200RT_REF(B200RT_KIND_TRI, fm.tri_begin);
    } else {
      const auto T1 = std::chrono::steady_clock::now();
      BuildOut bo{&b->nodes, &order, &gate};
      uint32_t r = build_bvh(objs, 0, nf, 0, bo, root_box);
      fm.root = B200RT_REF(B200RT_KIND_INNER, r);
      t_faces = secs(T0, T1);
      t_tree = secs(T1, std::chrono::steady_clock::now());
    }
    const auto T2 = std::chrono::steady_clock::now();
    fm.node_count = (int)b->nodes.size() - fm.node_begin;
    put_box(fm.root_box, fm.root_box + 3, root_box);
    // local face id -> global triangle index in DFS order
    std::vector<int32_t> remap((size_t)nf);
    for (int k = 0; k < nf; k++) remap[B200RT_REF_INDEX(order[k])] = fm.tri_begin + k;
    for (int n = fm.node_begin; n < fm.node_begin + fm.node_count; n++) {
      b200rt_bvh_node& nd = b->nodes[n];
      if (B200RT_REF_KIND(nd.left) == B200RT_KIND_TRI)
        nd.left = B200RT_REF(B200RT_KIND_TRI, remap[B200RT_REF_INDEX(nd.left)]);
      if (B200RT_REF_KIND(nd.right) == B200RT_KIND_TRI)
        nd.right = B200RT_REF(B200RT_KIND_TRI, remap[B200RT_REF_INDEX(nd.right)]);
    }
    for (int k = 0; k < nf; k++) {
      const int f = (int)B200RT_REF_INDEX(order[k]);
      const int i0 = m.faces[3 * f], i1 = m.faces[3 * f + 1], i2 = m.faces[3 * f + 2];
      const V3 v0 = vert(i0), v1 = vert(i1), v2 = vert(i2);
      float g[12];
      store3(g, v0); store3(g + 3, v0 - v1); store3(g + 6, v0 - v2); store3(g + 9, normals[f]);
      b->tri_geom.insert(b->tri_geom.end(), g, g + 12);
      b->tri_vidx.push_back(i0); b->tri_vidx.push_back(i1); b->tri_vidx.push_back(i2);
      b->tri_info.push_back(m.material);
      b->tri_info.push_back(m.smooth ? 1 : 0);
      {
        // texture coordinates: texture_coord_data[vertex_index + texture_offset], where vertex_index is
        // the face index BEFORE the vertex offset (HW7/src/Mesh_triangle.cpp:119-124); the Python/XML
        // host passes faces with the vertex offset applied and texture_offset relative to that.
        float uv[6] = {0, 0, 0, 0, 0, 0};
        const int ntc = (int)(b->texcoords.size() / 2);
        const int idx3[3] = {i0, i1, i2};
        if (m.texture >= 0)
          for (int c = 0; c < 3; c++) {
            const int ti = idx3[c] + m.texture_offset;
            if (ti >= 0 && ti < ntc) { uv[2 * c] = b->texcoords[2 * (size_t)ti]; uv[2 * c + 1] = b->texcoords[2 * (size_t)ti + 1]; }
          }
        b->tri_uv.insert(b->tri_uv.end(), uv, uv + 6);
        b->tri_tex.push_back(m.texture);
      }
      float gb[6];
      put_box(gb, gb + 3, gate[f]);
      b->tri_gate.insert(b->tri_gate.end(), gb, gb + 6);
    }
    const auto T3 = std::chrono::steady_clock::now();
    if (!hw1 && nf >= 2) {
      // SAH tree for the device over the same triangles (global ids = DFS order)
      std::vector<float> tb((size_t)nf * 6);
      float ext = 0.f;
      for (int k = 0; k < nf; k++) {
        const int f = (int)B200RT_REF_INDEX(order[k]);
        const Box bx = tri_box(vert(m.faces[3 * f]), vert(m.faces[3 * f + 1]), vert(m.faces[3 * f + 2]));
        put_box(&tb[6 * (size_t)k], &tb[6 * (size_t)k + 3], bx);
      }
      for (int a = 0; a < 3; a++) {
        ext = std::fmax(ext, root_box.mx[a] - root_box.mn[a]);
        ext = std::fmax(ext, std::fmax(std::fabs(root_box.mx[a]), std::fabs(root_box.mn[a])));
      }
      // conservative growth: covers the rounding slop of the Cramer test and the
      // HW7/PPM acceptance band (beta, gamma >= -eps) — DESIGN.md §4b
      const float grow = ext * (2e-5f + 4.0f * b->desc.isect_eps) + 1e-30f;
      const size_t o0 = b->fast_tri_ref.size();
      fm.fast_root = b200rt_fast::build(tb.data(), nf, fm.tri_begin, grow, b->fast_nodes, b->fast_tri_ref);
      for (size_t k = o0; k < b->fast_tri_ref.size(); k++) {
        const float* g = &b->tri_geom[12 * (size_t)b->fast_tri_ref[k]];
        b->fast_tri_geom.insert(b->fast_tri_geom.end(), g, g + 12);
      }
    }
    if (std::getenv("B200RT_BUILD_TIMING"))  // where a large mesh's load time goes (profiles/r01m_host_build.txt)
      std::fprintf(stderr, "b200rt build: mesh %zu, %d faces: normals %.3f s, reference tree %.3f s, records %.3f s, SAH tree %.3f s\n",
                   mi, nf, t_faces, t_tree, secs(T2, T3), secs(T3, std::chrono::steady_clock::now()));
    b->meshes.push_back(fm);
  }
  // HW1 stand-alone triangles are plain triangles scanned after the meshes
  // (HW1/Scene.cpp:44-49): append them to the triangle arrays.
  if (hw1) {
    for (const TriIn& t : b->tri_in) {
      if (!check_idx(t.idx[0]) || !check_idx(t.idx[1]) || !check_idx(t.idx[2])) {
        g_host_error = "triangle index out of range";
        return B200RT_E_INVALID;
      }
      const V3 v0 = vert(t.idx[0]), v1 = vert(t.idx[1]), v2 = vert(t.idx[2]);
      float g[12];
      store3(g, v0); store3(g + 3, v0 - v1); store3(g + 6, v0 - v2);
      store3(g + 9, normalize(cross(v1 - v0, v2 - v0)));
      b->tri_geom.insert(b->tri_geom.end(), g, g + 12);
      b->tri_vidx.push_back(t.idx[0]); b->tri_vidx.push_back(t.idx[1]); b->tri_vidx.push_back(t.idx[2]);
      b->tri_info.push_back(t.material);
      b->tri_info.push_back(0);
      b->tri_gate.insert(b->tri_gate.end(), 6, 0.0f);
      b->tri_uv.insert(b->tri_uv.end(), 6, 0.0f);
      b->tri_tex.push_back(-1);
    }
  }
  // Vertex::finalize_normal — HW3/src/Scene.cpp:872-877.  The PPM tree finalises EVERY vertex, with or without a face
  // on it (PPM/src/Scene.cpp:499-501): a vertex no face uses — a sphere centre — ends up 0/0 there.  Nothing reads
  // those normals; they are kept only so that this description equals the one integration/flatten_ppm.cpp copies
  // out of the tree's own objects.
  b->vert_normal.assign((size_t)nv * 3, 0.f);
  for (int i = 0; i < nv; i++) {
    V3 n = vnorm[i];
    if (vhas[i] || b->semantics == B200RT_SEM_PPM) n = normalize(n);
    store3(&b->vert_normal[3 * (size_t)i], n);
  }

  b->n_mesh_nodes = b->nodes.size();
  b->n_user_lights = b->lights.size();
  const int rc_top = finish_top(b);
  if (rc_top) return rc_top;
  fill_desc(b);
  b->finished = true;
  *out = &b->desc;
  return B200RT_OK;
} B200RT_CATCH

/* ---- moving geometry between frames (include/b200rt.h: "refit") ---- */
int b200rt_builder_set_instance_transform(b200rt_builder* b, int instance, const float xf[16], const float velocity[3]) try {
  if (!b || !xf || instance < 0 || instance >= (int)b->inst_in.size()) {
    g_host_error = "b200rt_builder_set_instance_transform: bad instance index";
    return B200RT_E_INVALID;
  }
  InstIn& i = b->inst_in[(size_t)instance];
  if (i.direct) {
    g_host_error = "b200rt_builder_set_instance_transform: HW2 meshes have no transform";
    return B200RT_E_UNSUPPORTED;
  }
  i.xf = M4::from(xf);
  i.velocity = velocity ? load3(velocity) : V3(0.f);
  return B200RT_OK;
} B200RT_CATCH
int b200rt_builder_set_sphere_transform(b200rt_builder* b, int sphere, const float xf[16], const float velocity[3]) try {
  if (!b || !xf || sphere < 0 || sphere >= (int)b->sph_in.size()) {
    g_host_error = "b200rt_builder_set_sphere_transform: bad sphere index";
    return B200RT_E_INVALID;
  }
  b->sph_in[(size_t)sphere].xf = M4::from(xf);
  b->sph_in[(size_t)sphere].velocity = velocity ? load3(velocity) : V3(0.f);
  return B200RT_OK;
} B200RT_CATCH
int b200rt_builder_refit(b200rt_builder* b, const B200SceneDesc** out) try {
  if (!b || !out || !b->finished) {
    g_host_error = "b200rt_builder_refit: b200rt_builder_finish first";
    return B200RT_E_INVALID;
  }
  // everything the top-level phase appended goes; the mesh trees, triangle records and gates stay
  b->nodes.resize(b->n_mesh_nodes);
  b->lights.resize(b->n_user_lights);
  b->light_cdf.clear(); b->light_tri.clear();
  const int rc = finish_top(b);
  if (rc) { b->finished = false; return rc; }
  fill_desc(b);
  *out = &b->desc;
  return B200RT_OK;
} B200RT_CATCH

}  // extern "C"
