// flatten_hw2.cpp — the reference-side half of the drop-in boundary, for the HW2 tree: HW3's walk
// over HW2's classes.  No transforms and no instancing here: a `Mesh` sits in Scene::bvh itself (an instance record
// flagged DIRECT: no gate box of its own, no matrix), `Triangle` is both the mesh triangle and the stand-alone one,
// and `Scene` keeps no mesh list, so meshes are numbered in the order the walk of Scene::bvh meets them.
//
//     g++ -std=c++14 -O2 -fno-access-control -I HW2 -I <repo>/include -c flatten_hw2.cpp
//
// tests/test_flatten_reference.py links this file with the UNTOUCHED objects of /root/reference/HW2
// (oracle/build_ref.sh flat2) and holds its output equal to what b200rt_builder_* makes of the same XML.
#include <cstdint>
#include <cstring>
#include <map>
#include <vector>

#include "Scene.h"      // the reference's (HW2/)
#include "Bounding_volume_hierarchy.h"
#include "Mesh.h"
#include "Sphere.h"
#include "Triangle.h"
#include "b200rt.h"     // this repo's include/

namespace {

void put3(float* d, const Vector3& v) { d[0] = v.x; d[1] = v.y; d[2] = v.z; }
void put_box(float* mn, float* mx, const Bounding_box& b) { put3(mn, b.min_corner); put3(mx, b.max_corner); }
void put_identity(float inv12[12], float nrm9[9], float fwd16[16]) {
  for (int i = 0; i < 3; i++) { inv12[4 * i + i] = 1.0f; nrm9[3 * i + i] = 1.0f; }
  if (fwd16) for (int i = 0; i < 4; i++) fwd16[4 * i + i] = 1.0f;
}

}  // namespace

struct b200rt_flat_hw2 {
  B200SceneDesc desc;
  std::vector<b200rt_bvh_node> nodes;
  std::vector<float> tri_geom, tri_gate, tri_uv, vert_pos, vert_normal;
  std::vector<int32_t> tri_vidx, tri_info, tri_tex;
  std::vector<b200rt_mesh> meshes;
  std::vector<b200rt_instance> instances;
  std::vector<b200rt_sphere> spheres;
  std::vector<b200rt_triprim> triprims;
  std::vector<b200rt_material> materials;
  std::vector<b200rt_light> lights;
  std::vector<b200rt_camera> cameras;
  std::map<const Mesh*, int> mesh_index;
  const Scene* scene;
  int rank;

  // One mesh triangle, in the order the walk meets it: v0, v0-v1, v0-v2, the stored normal (Mesh_triangle::intersect
  // forms the same two differences per ray, HW2/src/Mesh_triangle.cpp:77-86), its vertex ids and its gate box.
  uint32_t add_tri(const Triangle* t, const Bounding_box& gate) {
    const uint32_t id = (uint32_t)tri_tex.size();
    const int i0 = t->index_0, i1 = t->index_1, i2 = t->index_2;
    const Vector3& v0 = scene->get_vertex_at(i0);
    const Vector3& v1 = scene->get_vertex_at(i1);
    const Vector3& v2 = scene->get_vertex_at(i2);
    float g[12];
    put3(g, v0); put3(g + 3, v0 - v1); put3(g + 6, v0 - v2); put3(g + 9, t->normal);
    tri_geom.insert(tri_geom.end(), g, g + 12);
    tri_vidx.push_back(i0); tri_vidx.push_back(i1); tri_vidx.push_back(i2);
    tri_info.push_back(t->material_id);
    tri_info.push_back(0);                                    // flat shading only
    float gb[6];
    put_box(gb, gb + 3, gate);
    tri_gate.insert(tri_gate.end(), gb, gb + 6);
    tri_uv.insert(tri_uv.end(), 6, 0.0f);   // HW3 has no textures
    tri_tex.push_back(-1);
    return B200RT_REF(B200RT_KIND_TRI, id);
  }
  // Mesh::bvh: pre-order node slots, leaves in depth-first order (BVH::intersect visits left then right,
  // HW2/src/Bounding_volume_hierarchy.cpp:31-60).  `gate` = the box of the node that holds the child.
  uint32_t walk_mesh(const Shape* s, const Bounding_box& gate) {
    if (const BVH* n = dynamic_cast<const BVH*>(s)) {
      const uint32_t me = (uint32_t)nodes.size();
      nodes.emplace_back();
      const uint32_t l = walk_mesh(n->left, n->bounding_box);
      const uint32_t r = walk_mesh(n->right, n->bounding_box);
      b200rt_bvh_node d;
      std::memset(&d, 0, sizeof(d));
      put_box(d.lmin, d.lmax, n->left->get_bounding_box());
      put_box(d.rmin, d.rmax, n->right->get_bounding_box());
      d.left = l;
      d.right = r;
      nodes[me] = d;
      return B200RT_REF(B200RT_KIND_INNER, me);
    }
    return add_tri(static_cast<const Triangle*>(s), gate);
  }
  void add_mesh(const Mesh* m) {
    b200rt_mesh fm;
    std::memset(&fm, 0, sizeof(fm));
    fm.tri_begin = (int32_t)tri_tex.size();
    fm.node_begin = (int32_t)nodes.size();
    fm.fast_root = B200RT_REF_NONE;   // the SAH tree is libb200rt.so's own business (b200rt_scene_create)
    if (m->bvh) {
      fm.root = walk_mesh(m->bvh, Bounding_box());
      put_box(fm.root_box, fm.root_box + 3, m->bvh->get_bounding_box());
    } else {
      fm.root = B200RT_REF_NONE;
      put_box(fm.root_box, fm.root_box + 3, Bounding_box());
    }
    fm.tri_count = (int32_t)tri_tex.size() - fm.tri_begin;
    fm.node_count = (int32_t)nodes.size() - fm.node_begin;
    mesh_index[m] = (int)meshes.size();
    meshes.push_back(fm);
  }

  // Mesh in Scene::bvh: Mesh::intersect hands the ray straight to its own bvh (HW2/Mesh.h:13-15)
  uint32_t add_direct_mesh(const Mesh* m) {
    b200rt_instance fi;
    std::memset(&fi, 0, sizeof(fi));
    put_identity(fi.inv, fi.nrm, fi.fwd);
    put_box(fi.box, fi.box + 3, m->get_bounding_box());
    fi.mesh = mesh_index.at(m);
    fi.material = m->material_id;
    fi.texture = m->texture_id;
    fi.flags = B200RT_INST_DIRECT;
    fi.rank = rank++;
    fi.light = -1;
    instances.push_back(fi);
    return B200RT_REF(B200RT_KIND_INSTANCE, instances.size() - 1);
  }
  uint32_t add_sphere(const Sphere* sp) {
    b200rt_sphere fs;
    std::memset(&fs, 0, sizeof(fs));
    put3(fs.center, sp->center);
    fs.radius = sp->radius;
    put_identity(fs.inv, fs.nrm, fs.fwd);
    fs.material = sp->material_id;
    fs.texture = -1;
    fs.flags = B200RT_XF_IDENTITY;
    fs.rank = rank++;
    fs.light = -1;
    spheres.push_back(fs);
    return B200RT_REF(B200RT_KIND_SPHERE, spheres.size() - 1);
  }
  uint32_t add_triprim(const Triangle* t) {
    b200rt_triprim ft;
    std::memset(&ft, 0, sizeof(ft));
    const Vector3& v0 = scene->get_vertex_at(t->index_0);
    const Vector3& v1 = scene->get_vertex_at(t->index_1);
    const Vector3& v2 = scene->get_vertex_at(t->index_2);
    put3(ft.v0, v0); put3(ft.a1, v0 - v1); put3(ft.a2, v0 - v2); put3(ft.normal, t->normal);
    put_identity(ft.inv, ft.nrm, nullptr);
    ft.material = t->material_id;
    ft.flags = B200RT_XF_IDENTITY;
    ft.rank = rank++;
    triprims.push_back(ft);
    return B200RT_REF(B200RT_KIND_TRIPRIM, triprims.size() - 1);
  }
  // Meshes first (their nodes and triangles precede the top-level nodes, as in the other producer), in the order
  // the depth-first walk of Scene::bvh meets them.
  void collect_meshes(const Shape* s) {
    if (const BVH* n = dynamic_cast<const BVH*>(s)) {
      collect_meshes(n->left);
      collect_meshes(n->right);
    } else if (const Mesh* m = dynamic_cast<const Mesh*>(s)) {
      add_mesh(m);
    }
  }
  // Scene::bvh — the objects vector of the constructor is gone by now; the tree is what is left of it.
  uint32_t walk_top(const Shape* s) {
    if (const BVH* n = dynamic_cast<const BVH*>(s)) {
      const uint32_t me = (uint32_t)nodes.size();
      nodes.emplace_back();
      const uint32_t l = walk_top(n->left);
      const uint32_t r = walk_top(n->right);
      b200rt_bvh_node d;
      std::memset(&d, 0, sizeof(d));
      put_box(d.lmin, d.lmax, n->left->get_bounding_box());
      put_box(d.rmin, d.rmax, n->right->get_bounding_box());
      d.left = l;
      d.right = r;
      nodes[me] = d;
      return B200RT_REF(B200RT_KIND_INNER, me);
    }
    if (const Mesh* m = dynamic_cast<const Mesh*>(s)) return add_direct_mesh(m);
    if (const Sphere* sp = dynamic_cast<const Sphere*>(s)) return add_sphere(sp);
    if (const Triangle* t = dynamic_cast<const Triangle*>(s)) return add_triprim(t);
    return B200RT_REF_NONE;
  }
};

extern "C" b200rt_flat_hw2* b200rt_flatten_hw2(const Scene& scene) {
  b200rt_flat_hw2* f = new b200rt_flat_hw2();
  f->scene = &scene;
  f->rank = 0;
  B200SceneDesc& d = f->desc;
  std::memset(&d, 0, sizeof(d));
  d.semantics = B200RT_SEM_HW2;
  d.max_depth = scene.max_recursion_depth;
  put3(d.background, scene.background_color);
  put3(d.ambient, scene.ambient_light);
  d.shadow_eps = scene.shadow_ray_epsilon;
  d.isect_eps = kEpsilon;                       // HW2/Vector3.h:7
  d.background_texture = -1;
  d.env_light = -1;

  for (const Vector3& v : scene.vertex_data) {
    float p[3];
    put3(p, v);
    f->vert_pos.insert(f->vert_pos.end(), p, p + 3);
    f->vert_normal.insert(f->vert_normal.end(), 3, 0.0f);     // no vertex normals in this tree
  }
  for (const Material& m : scene.materials) {
    b200rt_material fm;
    std::memset(&fm, 0, sizeof(fm));
    put3(fm.ambient, m.ambient); put3(fm.diffuse, m.diffuse); put3(fm.specular, m.specular);
    put3(fm.mirror, m.mirror); put3(fm.transparency, m.transparency);
    fm.refraction_index = m.refraction_index;
    fm.phong_exponent = m.phong_exponent;
    fm.brdf = -1;
    f->materials.push_back(fm);
  }
  for (const Point_light& l : scene.point_lights) {   // Scene::trace_ray loops the point lights, then the area lights
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
    put3(fc.top_left, c.top_left_corner); put3(fc.su, c.s_u_constant); put3(fc.sv, c.s_v_constant);
    fc.left = ip.left; fc.right = ip.right; fc.bottom = ip.bottom; fc.top = ip.top; fc.distance = ip.distance;
    fc.width = ip.width; fc.height = ip.height;
    fc.samples_sqrt = c.get_number_of_samples() > 0 ? c.get_number_of_samples() : 1;   // already (int)sqrt(NumSamples): HW2/src/Scene.cpp:411
    f->cameras.push_back(fc);
  }
  if (scene.bvh) {
    f->collect_meshes(scene.bvh);
    d.root = f->walk_top(scene.bvh);
    put_box(d.root_box, d.root_box + 3, scene.bvh->get_bounding_box());
  } else {
    d.root = B200RT_REF_NONE;
    put_box(d.root_box, d.root_box + 3, Bounding_box());
  }

  d.n_nodes = (int32_t)f->nodes.size();          d.nodes = f->nodes.data();
  d.n_tris = (int32_t)f->tri_tex.size();         d.tri_geom = f->tri_geom.data();
  d.tri_vidx = f->tri_vidx.data();               d.tri_info = f->tri_info.data();
  d.tri_gate = f->tri_gate.data();               d.tri_uv = f->tri_uv.data();
  d.tri_tex = f->tri_tex.data();
  d.n_verts = (int32_t)scene.vertex_data.size(); d.vert_pos = f->vert_pos.data();
  d.vert_normal = f->vert_normal.data();
  d.n_meshes = (int32_t)f->meshes.size();        d.meshes = f->meshes.data();
  d.n_instances = (int32_t)f->instances.size();  d.instances = f->instances.data();
  d.n_spheres = (int32_t)f->spheres.size();      d.spheres = f->spheres.data();
  d.n_triprims = (int32_t)f->triprims.size();    d.triprims = f->triprims.data();
  d.n_materials = (int32_t)f->materials.size();  d.materials = f->materials.data();
  d.n_lights = (int32_t)f->lights.size();        d.lights = f->lights.data();
  d.n_cameras = (int32_t)f->cameras.size();      d.cameras = f->cameras.data();
  return f;
}
extern "C" const B200SceneDesc* b200rt_flat_hw2_desc(const b200rt_flat_hw2* f) { return &f->desc; }
extern "C" void b200rt_flat_hw2_free(b200rt_flat_hw2* f) { delete f; }
