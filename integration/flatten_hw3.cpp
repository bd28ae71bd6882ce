// flatten_hw3.cpp — the reference-side half of the drop-in boundary, for the HW3 tree.
//
// Compiled INSIDE the reference tree, next to its own sources, by the reference's compiler:
//     g++ -std=c++14 -O2 -fno-access-control -I HW3/include -I <repo>/include -c flatten_hw3.cpp
// (`-fno-access-control`: Mesh_instance::mesh_ / transformation_, Sphere::transformation_ / is_identity_,
// Triangle::transformation_, Camera's frame and Mesh_triangle::bounding_box_ are private; nothing in the tree has to
// change.)  It reads the objects `Scene::Scene(const std::string&)` built — the real pointer tree, the real matrices,
// the real vertex normals — and lays them out as the B200SceneDesc of include/b200rt.h.  No arithmetic of the
// reference is redone here: every number is copied out of the object that holds it, and the depth-first order the
// device's tie-breaks hang on is the order of the tree that is walked.
//
//   Scene scene(argv[1]);
//   b200rt_flat_hw3* flat = b200rt_flatten_hw3(scene);           // owns the buffers
//   b200rt_scene_create(b200rt_flat_hw3_desc(flat), 0, &gpu);    // libb200rt.so builds its SAH tree from tri_geom
//   ... b200rt_render(gpu, camera_index, &opts, rgbw, &stats) instead of the std::thread fan-out (HW3/src/main.cpp:27-40)
//   b200rt_flat_hw3_free(flat);
//
// tests/test_flatten_reference.py links this file with the UNTOUCHED objects of /root/reference/HW3 (oracle/build_ref.sh
// flat3), writes the description with b200rt_scene_desc_save, and holds it equal — array for array, bit for bit — to
// what b200rt_builder_* (the producer the Python host uses) makes of the same XML.
#include <cstdint>
#include <cstring>
#include <map>
#include <vector>

#include "Scene.h"      // the reference's (HW3/include)
#include "b200rt.h"     // this repo's include/

namespace {

void put3(float* d, const Vector3& v) { d[0] = v.x; d[1] = v.y; d[2] = v.z; }
void put_box(float* mn, float* mx, const Bounding_box& b) { put3(mn, b.min_corner); put3(mx, b.max_corner); }
// rows 0..2 of M^-1, upper 3x3 of the normal matrix, M itself — Transformation's three matrices (HW3/include/Transformation.h:7-20)
void put_xf(const Transformation& t, float inv12[12], float nrm9[9], float fwd16[16]) {
  const Matrix4x4& inv = t.get_inverse_transformation_matrix();
  const Matrix4x4& nrm = t.get_normal_transformation_matrix();
  const Matrix4x4& fwd = t.get_transformation_matrix();
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 4; j++) inv12[4 * i + j] = inv[i][j];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) nrm9[3 * i + j] = nrm[i][j];
  if (fwd16)
    for (int i = 0; i < 4; i++)
      for (int j = 0; j < 4; j++) fwd16[4 * i + j] = fwd[i][j];
}

}  // namespace

struct b200rt_flat_hw3 {
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
  // forms the same two differences per ray, HW3/src/Mesh_triangle.cpp:77-86), its vertex ids and its gate box.
  uint32_t add_tri(const Mesh_triangle* t, const Bounding_box& gate) {
    const uint32_t id = (uint32_t)tri_tex.size();
    const int i0 = t->index_0 + t->offset, i1 = t->index_1 + t->offset, i2 = t->index_2 + t->offset;
    const Vector3& v0 = scene->get_vertex_at(i0).get_vertex_position();
    const Vector3& v1 = scene->get_vertex_at(i1).get_vertex_position();
    const Vector3& v2 = scene->get_vertex_at(i2).get_vertex_position();
    float g[12];
    put3(g, v0); put3(g + 3, v0 - v1); put3(g + 6, v0 - v2); put3(g + 9, t->normal);
    tri_geom.insert(tri_geom.end(), g, g + 12);
    tri_vidx.push_back(i0); tri_vidx.push_back(i1); tri_vidx.push_back(i2);
    tri_info.push_back(t->material_id);
    tri_info.push_back(t->triangle_shading_mode == tsm_smooth ? 1 : 0);
    float gb[6];
    put_box(gb, gb + 3, gate);
    tri_gate.insert(tri_gate.end(), gb, gb + 6);
    tri_uv.insert(tri_uv.end(), 6, 0.0f);   // HW3 has no textures
    tri_tex.push_back(-1);
    return B200RT_REF(B200RT_KIND_TRI, id);
  }
  // Mesh::bvh: pre-order node slots, leaves in depth-first order (BVH::intersect visits left then right,
  // HW3/src/Bounding_volume_hierarchy.cpp:31-60).  `gate` = the box of the node that holds the child.
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
    return add_tri(static_cast<const Mesh_triangle*>(s), gate);
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

  uint32_t add_instance(const Mesh_instance* mi) {
    b200rt_instance fi;
    std::memset(&fi, 0, sizeof(fi));
    put_xf(mi->transformation_, fi.inv, fi.nrm, fi.fwd);
    put_box(fi.box, fi.box + 3, mi->bounding_box_);
    put3(fi.velocity, mi->velocity);
    fi.mesh = mesh_index.at(mi->mesh_);
    fi.material = mi->material_id;
    fi.texture = mi->texture_id;
    if (mi->velocity != Vector3(0.0f)) fi.flags |= B200RT_INST_MOVING;   // HW3/include/Mesh.h:62,104
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
    put_xf(sp->transformation_, fs.inv, fs.nrm, fs.fwd);
    put3(fs.velocity, sp->velocity);
    fs.material = sp->material_id;
    fs.texture = -1;
    if (sp->is_identity_) fs.flags |= B200RT_XF_IDENTITY;               // HW3/include/Sphere.h:20,46
    if (sp->velocity != Vector3(0.0f)) fs.flags |= B200RT_XF_MOVING;    // HW3/include/Sphere.h:30,46,76
    fs.rank = rank++;
    fs.light = -1;
    spheres.push_back(fs);
    return B200RT_REF(B200RT_KIND_SPHERE, spheres.size() - 1);
  }
  uint32_t add_triprim(const Triangle* t) {
    b200rt_triprim ft;
    std::memset(&ft, 0, sizeof(ft));
    const Vector3& v0 = scene->get_vertex_at(t->index_0 + t->offset).get_vertex_position();
    const Vector3& v1 = scene->get_vertex_at(t->index_1 + t->offset).get_vertex_position();
    const Vector3& v2 = scene->get_vertex_at(t->index_2 + t->offset).get_vertex_position();
    put3(ft.v0, v0); put3(ft.a1, v0 - v1); put3(ft.a2, v0 - v2); put3(ft.normal, t->normal);
    put_xf(t->transformation_, ft.inv, ft.nrm, nullptr);
    ft.material = t->material_id;
    if (t->is_identity_) ft.flags |= B200RT_XF_IDENTITY;                // HW3/src/Triangle.cpp:20-41
    ft.rank = rank++;
    triprims.push_back(ft);
    return B200RT_REF(B200RT_KIND_TRIPRIM, triprims.size() - 1);
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
    if (const Mesh_instance* mi = dynamic_cast<const Mesh_instance*>(s)) return add_instance(mi);
    if (const Sphere* sp = dynamic_cast<const Sphere*>(s)) return add_sphere(sp);
    if (const Triangle* t = dynamic_cast<const Triangle*>(s)) return add_triprim(t);
    return B200RT_REF_NONE;
  }
};

extern "C" b200rt_flat_hw3* b200rt_flatten_hw3(const Scene& scene) {
  b200rt_flat_hw3* f = new b200rt_flat_hw3();
  f->scene = &scene;
  f->rank = 0;
  B200SceneDesc& d = f->desc;
  std::memset(&d, 0, sizeof(d));
  d.semantics = B200RT_SEM_HW3;
  d.max_depth = scene.max_recursion_depth;
  put3(d.background, scene.background_color);
  put3(d.ambient, scene.ambient_light);
  d.shadow_eps = scene.shadow_ray_epsilon;
  d.isect_eps = kEpsilon;                       // HW3/include/Vector3.h:6
  d.background_texture = -1;
  d.env_light = -1;

  for (const Vertex& v : scene.vertex_data) {   // normals are final: Scene::Scene ends with finalize_normal (HW3/src/Scene.cpp:872-877)
    float p[3], n[3];
    put3(p, v.get_vertex_position());
    put3(n, v.get_vertex_normal());
    f->vert_pos.insert(f->vert_pos.end(), p, p + 3);
    f->vert_normal.insert(f->vert_normal.end(), n, n + 3);
  }
  for (const Material& m : scene.materials) {
    b200rt_material fm;
    std::memset(&fm, 0, sizeof(fm));
    put3(fm.ambient, m.ambient); put3(fm.diffuse, m.diffuse); put3(fm.specular, m.specular);
    put3(fm.mirror, m.mirror); put3(fm.transparency, m.transparency);
    fm.roughness = m.roughness;
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
  for (const Area_light& l : scene.area_lights) {
    b200rt_light fl;
    std::memset(&fl, 0, sizeof(fl));
    fl.kind = B200RT_LIGHT_AREA;
    put3(fl.position, l.position); put3(fl.intensity, l.intensity);
    put3(fl.edge1, l.edge_vector_1); put3(fl.edge2, l.edge_vector_2); put3(fl.normal, l.normal);
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
    fc.samples_sqrt = c.get_number_of_samples() > 0 ? c.get_number_of_samples() : 1;   // already (int)sqrt(NumSamples): HW3/src/Scene.cpp:411
    fc.aperture = c.get_aperture_size();
    f->cameras.push_back(fc);
  }
  for (const Mesh* m : scene.meshes) f->add_mesh(m);
  if (scene.bvh) {
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
extern "C" const B200SceneDesc* b200rt_flat_hw3_desc(const b200rt_flat_hw3* f) { return &f->desc; }
extern "C" void b200rt_flat_hw3_free(b200rt_flat_hw3* f) { delete f; }
