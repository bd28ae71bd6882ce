// flatten_hw1.cpp — the reference-side half of the drop-in boundary, for the HW1 tree (cfg1).
//
// HW1 has no BVH, no transforms, no instancing: `Scene` holds meshes (vectors of Triangle), stand-alone triangles and
// spheres, and Scene::trace_ray scans them in exactly that order (HW1/Scene.cpp:33-55) with the order-dependent rule
// `t < best - kEpsilon`.  The description keeps that order: triangles = every mesh's triangles in file order, then
// the stand-alone ones; spheres after them, `rank` = position in the scan.
//
//     g++ -std=c++14 -O2 -fno-access-control -I HW1 -I <repo>/include -c flatten_hw1.cpp
//
// tests/test_flatten_reference.py links this file with the UNTOUCHED objects of /root/reference/HW1
// (oracle/build_ref.sh flat1) and holds its output equal to what b200rt_builder_* makes of the same XML.
#include <cstdint>
#include <cstring>
#include <limits>
#include <vector>

#include "Scene.h"      // the reference's (HW1/)
#include "b200rt.h"     // this repo's include/

namespace {
void put3(float* d, const Vector3& v) { d[0] = v.x; d[1] = v.y; d[2] = v.z; }
}  // namespace

struct b200rt_flat_hw1 {
  B200SceneDesc desc;
  std::vector<float> tri_geom, tri_gate, tri_uv, vert_pos, vert_normal;
  std::vector<int32_t> tri_vidx, tri_info, tri_tex;
  std::vector<b200rt_mesh> meshes;
  std::vector<b200rt_sphere> spheres;
  std::vector<b200rt_material> materials;
  std::vector<b200rt_light> lights;
  std::vector<b200rt_camera> cameras;

  void add_tri(const Scene& scene, const Triangle& t, bool in_mesh) {
    const Vector3& v0 = scene.get_vertex_at(t.index_0);
    const Vector3& v1 = scene.get_vertex_at(t.index_1);
    const Vector3& v2 = scene.get_vertex_at(t.index_2);
    float g[12];
    put3(g, v0); put3(g + 3, v0 - v1); put3(g + 6, v0 - v2); put3(g + 9, t.normal);   // Triangle::intersect's columns (HW1/Triangle.cpp)
    tri_geom.insert(tri_geom.end(), g, g + 12);
    tri_vidx.push_back(t.index_0); tri_vidx.push_back(t.index_1); tri_vidx.push_back(t.index_2);
    tri_info.push_back(t.material_id);
    tri_info.push_back(0);                                    // flat shading only
    const float inf = std::numeric_limits<float>::infinity();
    const float no_gate[6] = {inf, inf, inf, -inf, -inf, -inf}, zero[6] = {0, 0, 0, 0, 0, 0};
    const float* gate = in_mesh ? no_gate : zero;             // there are no boxes in this tree: never read
    tri_gate.insert(tri_gate.end(), gate, gate + 6);
    tri_uv.insert(tri_uv.end(), 6, 0.0f);
    tri_tex.push_back(-1);
  }
};

extern "C" b200rt_flat_hw1* b200rt_flatten_hw1(const Scene& scene) {
  b200rt_flat_hw1* f = new b200rt_flat_hw1();
  B200SceneDesc& d = f->desc;
  std::memset(&d, 0, sizeof(d));
  d.semantics = B200RT_SEM_HW1;
  d.max_depth = scene.max_recursion_depth;
  put3(d.background, scene.background_color);
  put3(d.ambient, scene.ambient_light);
  d.shadow_eps = scene.shadow_ray_epsilon;
  d.isect_eps = kEpsilon;                       // HW1/Vector3.h
  d.background_texture = -1;
  d.env_light = -1;
  d.root = B200RT_REF_NONE;                     // no Scene::bvh
  const float inf = std::numeric_limits<float>::infinity();
  for (int a = 0; a < 3; a++) { d.root_box[a] = inf; d.root_box[3 + a] = -inf; }

  for (const Vector3& v : scene.vertex_data) {
    float p[3];
    put3(p, v);
    f->vert_pos.insert(f->vert_pos.end(), p, p + 3);
    f->vert_normal.insert(f->vert_normal.end(), 3, 0.0f);     // no vertex normals in this tree
  }
  for (const Material& m : scene.materials) {
    b200rt_material fm;
    std::memset(&fm, 0, sizeof(fm));
    put3(fm.ambient, m.ambient); put3(fm.diffuse, m.diffuse); put3(fm.specular, m.specular); put3(fm.mirror, m.mirror);
    fm.refraction_index = 1.0f;
    fm.phong_exponent = m.phong_exponent;
    fm.brdf = -1;
    f->materials.push_back(fm);
  }
  for (const Point_light& l : scene.point_lights) {
    b200rt_light fl;
    std::memset(&fl, 0, sizeof(fl));
    fl.kind = B200RT_LIGHT_POINT;
    put3(fl.position, l.position); put3(fl.intensity, l.intensity);
    fl.object = -1;
    f->lights.push_back(fl);
  }
  for (const Camera& c : scene.cameras) {
    b200rt_camera fc;
    std::memset(&fc, 0, sizeof(fc));
    const Image_plane& ip = c.get_image_plane();
    put3(fc.e, c.e); put3(fc.u, c.u); put3(fc.v, c.v); put3(fc.w, c.w);
    put3(fc.top_left, c.top_left_corner);
    // su / sv stay 0: HW1's Camera has no per-pixel constants, calculate_ray_at works from the plane itself
    // ((right - left) * (x + 0.5f) / width, HW1/Camera.h:24-32) and so does the device for this tree
    fc.left = ip.left; fc.right = ip.right; fc.bottom = ip.bottom; fc.top = ip.top; fc.distance = ip.distance;
    fc.width = ip.width; fc.height = ip.height;
    fc.samples_sqrt = 1;
    f->cameras.push_back(fc);
  }
  for (const Mesh& m : scene.meshes) {          // scan order: meshes, then triangles, then spheres
    b200rt_mesh fm;
    std::memset(&fm, 0, sizeof(fm));
    fm.root = B200RT_REF_NONE;
    fm.fast_root = B200RT_REF_NONE;
    for (int a = 0; a < 3; a++) { fm.root_box[a] = inf; fm.root_box[3 + a] = -inf; }
    fm.tri_begin = (int32_t)f->tri_tex.size();
    for (const Triangle& t : m.triangles) f->add_tri(scene, t, true);
    fm.tri_count = (int32_t)f->tri_tex.size() - fm.tri_begin;
    f->meshes.push_back(fm);
  }
  for (const Triangle& t : scene.triangles) f->add_tri(scene, t, false);
  int rank = 0;
  for (const Sphere& s : scene.spheres) {
    b200rt_sphere fs;
    std::memset(&fs, 0, sizeof(fs));
    put3(fs.center, s.center);
    fs.radius = s.radius;
    for (int i = 0; i < 3; i++) { fs.inv[4 * i + i] = 1.0f; fs.nrm[3 * i + i] = 1.0f; }
    for (int i = 0; i < 4; i++) fs.fwd[4 * i + i] = 1.0f;
    fs.material = s.material_id;
    fs.texture = -1;
    fs.flags = B200RT_XF_IDENTITY;
    fs.rank = rank++;
    fs.light = -1;
    f->spheres.push_back(fs);
  }

  d.n_tris = (int32_t)f->tri_tex.size();         d.tri_geom = f->tri_geom.data();
  d.tri_vidx = f->tri_vidx.data();               d.tri_info = f->tri_info.data();
  d.tri_gate = f->tri_gate.data();               d.tri_uv = f->tri_uv.data();
  d.tri_tex = f->tri_tex.data();
  d.n_verts = (int32_t)scene.vertex_data.size(); d.vert_pos = f->vert_pos.data();
  d.vert_normal = f->vert_normal.data();
  d.n_meshes = (int32_t)f->meshes.size();        d.meshes = f->meshes.data();
  d.n_spheres = (int32_t)f->spheres.size();      d.spheres = f->spheres.data();
  d.n_materials = (int32_t)f->materials.size();  d.materials = f->materials.data();
  d.n_lights = (int32_t)f->lights.size();        d.lights = f->lights.data();
  d.n_cameras = (int32_t)f->cameras.size();      d.cameras = f->cameras.data();
  return f;
}
extern "C" const B200SceneDesc* b200rt_flat_hw1_desc(const b200rt_flat_hw1* f) { return &f->desc; }
extern "C" void b200rt_flat_hw1_free(b200rt_flat_hw1* f) { delete f; }
