// flatten_hw7.cpp — the reference-side half of the drop-in boundary, for the HW7 tree: what
// flatten_hw3.cpp does, plus everything HW4-HW7 added to `Scene` — textures (image texels as Texture holds them after
// degamma, the Perlin permutation of the HOST object), BRDF objects, the polymorphic light list (point, area, spot,
// directional, environment, mesh and sphere lights), texture coordinates, the integrator switch.
//
//     g++ -std=c++14 -O2 -fno-access-control -I HW7/include -I <repo>/include -c flatten_hw7.cpp
//
// No arithmetic of the reference is redone: every number is copied out of the object that holds it.
// tests/test_flatten_reference.py links this file with the UNTOUCHED objects of /root/reference/HW7
// (oracle/build_ref.sh flat7) and holds its output equal to what b200rt_builder_* makes of the same XML.
#include <cstdint>
#include <cstring>
#include <map>
#include <vector>

// The HW4, HW5 and HW6 trees are earlier states of the same classes: -DB200RT_TREE=6 leaves out what HW7 added (the
// integrator switch, Spherical_directional_light, Light_mesh, Light_sphere), -DB200RT_TREE=5 also the BRDF objects,
// -DB200RT_TREE=4 also the Light hierarchy (HW4 keeps HW3's point_lights / area_lights vectors).
#ifndef B200RT_TREE
#define B200RT_TREE 7
#endif
#include "Scene.h"      // the reference's (HW7/include)
#if B200RT_TREE >= 7
#include "Light_mesh.h"
#include "Light_sphere.h"
#define TRI_I0(t) (t)->vertex_index_0
#define TRI_I1(t) (t)->vertex_index_1
#define TRI_I2(t) (t)->vertex_index_2
#define TRI_OFFSET(t) (t)->vertex_offset
#else   // HW7 renamed these members
#define TRI_I0(t) (t)->index_0
#define TRI_I1(t) (t)->index_1
#define TRI_I2(t) (t)->index_2
#define TRI_OFFSET(t) (t)->offset
#endif
#include "b200rt.h"     // this repo's include/

namespace {

void put3(float* d, const Vector3& v) { d[0] = v.x; d[1] = v.y; d[2] = v.z; }
void put_box(float* mn, float* mx, const Bounding_box& b) { put3(mn, b.min_corner); put3(mx, b.max_corner); }
// rows 0..2 of M^-1, upper 3x3 of the normal matrix, M itself — Transformation's three matrices (HW7/include/Transformation.h:7-20)
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

struct b200rt_flat_hw7 {
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
  std::vector<b200rt_texture> textures;
  std::vector<float> texels, light_cdf;
  std::vector<b200rt_brdf> brdfs;
  std::vector<int32_t> light_tri;
  std::map<const Mesh*, int> mesh_index;
  std::map<const Shape*, uint32_t> tri_index;     // Mesh_triangle -> its depth-first id (Light_mesh::triangles_)
#if B200RT_TREE >= 5
  std::map<const Light*, int> light_index;
#endif
  const Scene* scene;
  int rank;

  // One mesh triangle, in the order the walk meets it: v0, v0-v1, v0-v2, the stored normal (Mesh_triangle::intersect
  // forms the same two differences per ray, HW7/src/Mesh_triangle.cpp:58-69), its vertex ids and its gate box.
  uint32_t add_tri(const Mesh_triangle* t, const Bounding_box& gate) {
    const uint32_t id = (uint32_t)tri_tex.size();
    const int i0 = TRI_I0(t) + TRI_OFFSET(t), i1 = TRI_I1(t) + TRI_OFFSET(t), i2 = TRI_I2(t) + TRI_OFFSET(t);
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
    // texture_coord_data[vertex_index + texture_offset] — HW7/src/Mesh_triangle.cpp:119-124 (0 when the triangle has
    // no texture or the index falls outside the array: the device never reads them then)
    float uv[6] = {0, 0, 0, 0, 0, 0};
    const int ntc = (int)scene->texture_coord_data.size();
    const int vi[3] = {TRI_I0(t), TRI_I1(t), TRI_I2(t)};
    if (t->texture_id >= 0)
      for (int c = 0; c < 3; c++) {
        const int ti = vi[c] + t->texture_offset;
        if (ti >= 0 && ti < ntc) { uv[2 * c] = scene->texture_coord_data[ti].x; uv[2 * c + 1] = scene->texture_coord_data[ti].y; }
      }
    tri_uv.insert(tri_uv.end(), uv, uv + 6);
    tri_tex.push_back(t->texture_id);
    tri_index[t] = id;
    return B200RT_REF(B200RT_KIND_TRI, id);
  }
  // Mesh::bvh: pre-order node slots, leaves in depth-first order (BVH::intersect visits left then right,
  // HW7/src/Bounding_volume_hierarchy.cpp:31-60).  `gate` = the box of the node that holds the child.
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
    if (mi->velocity != Vector3(0.0f)) fi.flags |= B200RT_INST_MOVING;   // HW7/include/Mesh.h:76,134
#if B200RT_TREE >= 7
    if (mi->is_refractive_) fi.flags |= B200RT_INST_REFRACTIVE;          // culling off inside glass: HW7/include/Mesh.h:73-75
#endif
    fi.rank = rank++;
    fi.light = -1;
    instances.push_back(fi);
    return B200RT_REF(B200RT_KIND_INSTANCE, instances.size() - 1);
  }
#if B200RT_TREE >= 7
  // Light_mesh sits in Scene::bvh itself: entered without a gate of its own, its box is the object-space mesh box,
  // rays go through base_transform (HW7/include/Light_mesh.h:52-68).
  uint32_t add_light_mesh(const Light_mesh* lm) {
    b200rt_instance fi;
    std::memset(&fi, 0, sizeof(fi));
    put_xf(lm->base_transform, fi.inv, fi.nrm, fi.fwd);
    put_box(fi.box, fi.box + 3, lm->get_bounding_box());
    fi.mesh = mesh_index.at(static_cast<const Mesh*>(lm));
    fi.material = lm->material_id;
    fi.texture = lm->texture_id;
    fi.flags |= B200RT_INST_LIGHT;
    fi.rank = rank++;
    fi.light = light_index.at(static_cast<const Light*>(lm));
    lights[fi.light].object = (int32_t)instances.size();
    instances.push_back(fi);
    return B200RT_REF(B200RT_KIND_LIGHTINST, instances.size() - 1);
  }
#endif
  uint32_t add_sphere(const Sphere* sp) {
    b200rt_sphere fs;
    std::memset(&fs, 0, sizeof(fs));
    put3(fs.center, sp->center);
    fs.radius = sp->radius;
    put_xf(sp->transformation_, fs.inv, fs.nrm, fs.fwd);
    put3(fs.velocity, sp->velocity);
    fs.material = sp->material_id;
    fs.texture = sp->texture_id;
    if (sp->is_identity_) fs.flags |= B200RT_XF_IDENTITY;               // HW7/src/Sphere.cpp:16-18
    if (sp->velocity != Vector3(0.0f)) fs.flags |= B200RT_XF_MOVING;    // HW7/src/Sphere.cpp:24,34
    fs.rank = rank++;
    fs.light = -1;
#if B200RT_TREE >= 7
    if (const Light_sphere* ls = dynamic_cast<const Light_sphere*>(sp)) {   // HW7/include/Light_sphere.h:16-24
      fs.flags |= B200RT_XF_LIGHT;
      fs.light = light_index.at(static_cast<const Light*>(ls));
      lights[fs.light].object = (int32_t)spheres.size();
    }
#endif
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
    if (t->is_identity_) ft.flags |= B200RT_XF_IDENTITY;                // HW7/src/Triangle.cpp:20-41
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
#if B200RT_TREE >= 7
    if (const Light_mesh* lm = dynamic_cast<const Light_mesh*>(s)) return add_light_mesh(lm);
#endif
    if (const Sphere* sp = dynamic_cast<const Sphere*>(s)) return add_sphere(sp);
    if (const Triangle* t = dynamic_cast<const Triangle*>(s)) return add_triprim(t);
    return B200RT_REF_NONE;
  }
};

extern "C" b200rt_flat_hw7* b200rt_flatten_hw7(const Scene& scene) {
  b200rt_flat_hw7* f = new b200rt_flat_hw7();
  f->scene = &scene;
  f->rank = 0;
  B200SceneDesc& d = f->desc;
  std::memset(&d, 0, sizeof(d));
  d.semantics = B200RT_TREE;                    // B200RT_SEM_HW5 / HW6 / HW7 are 5 / 6 / 7
  d.max_depth = scene.max_recursion_depth;
  put3(d.background, scene.background_color);
  put3(d.ambient, scene.ambient_light);
  d.shadow_eps = scene.shadow_ray_epsilon;
  d.isect_eps = intersection_test_epsilon;      // HW7/include/Vector3.h:10
  d.background_texture = -1;
  d.env_light = -1;
#if B200RT_TREE >= 7
  d.integrator = scene.integrator_type == it_pathtracing ? 1 : 0;          // HW7/src/Scene.cpp:1573-1582
  d.importance_sampling = (scene.integrator_type == it_pathtracing && !scene.is_uniform_sampling) ? 1 : 0;
#endif

  for (const Vertex& v : scene.vertex_data) {   // normals are final: Scene::Scene ends with finalize_normal (HW7/src/Scene.cpp:872-877)
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
#if B200RT_TREE >= 6
    fm.brdf = m.brdf_id;
#else
    fm.brdf = -1;
#endif
    f->materials.push_back(fm);
  }
  // Texture objects: Scene::textures in order, then the background texture, then the environment map (the order the
  // other producer uses; ids of the first group are what Mesh_triangle::texture_id / Sphere::texture_id hold)
  auto add_texture = [&](const Texture& t) {
    b200rt_texture ft;
    std::memset(&ft, 0, sizeof(ft));
    ft.interp = t.interpolation_type_ == Texture::it_nearest ? B200RT_TEX_NEAREST : B200RT_TEX_BILINEAR;
    ft.decal = t.decal_mode_ == Texture::dm_replace_kd ? B200RT_TEX_REPLACE_KD
             : t.decal_mode_ == Texture::dm_blend_kd ? B200RT_TEX_BLEND_KD : B200RT_TEX_REPLACE_ALL;
    ft.normalizer = t.normalizer_;
    ft.scaling = 1.0f;
    ft.bump_multiplier = t.bumpmap_multiplier_;
    ft.is_bump = t.is_bump_ ? 1 : 0;
    ft.texel_offset = (int64_t)f->texels.size();
    if (t.perlin_noise_) {                                   // HW7/include/Perlin_noise.h:13-17
      ft.kind = B200RT_TEX_PERLIN;
      ft.appearance = t.perlin_noise_->pn_appearance_ == Perlin_noise::pn_vein ? B200RT_TEX_VEIN : B200RT_TEX_PATCH;
      ft.scaling = t.perlin_noise_->scaling_factor_;
      for (int i = 0; i < 16; i++) ft.perm[i] = t.perlin_noise_->p_[i];   // the table THIS host object shuffled
    } else {
      ft.kind = B200RT_TEX_IMAGE;
      ft.appearance = t.appearance_ == Texture::a_repeat ? B200RT_TEX_REPEAT : B200RT_TEX_CLAMP;
      ft.width = t.width_;
      ft.height = t.height_;
#if B200RT_TREE >= 7
      f->texels.insert(f->texels.end(), t.texture_image_.begin(), t.texture_image_.end());   // after degamma
#else   // HW5 / HW6 keep stb_image's bytes and convert per lookup (HW6/src/Texture.cpp:72-74): the same values as floats
      for (size_t k = 0; k < (size_t)t.width_ * t.height_ * 3; k++) f->texels.push_back((float)t.texture_image_[k]);
#endif
    }
    f->textures.push_back(ft);
    return (int32_t)f->textures.size() - 1;
  };
  for (const Texture& t : scene.textures) add_texture(t);
  if (scene.background_texture) d.background_texture = add_texture(*scene.background_texture);
#if B200RT_TREE >= 6
  for (const BRDF* r : scene.brdfs) {
    b200rt_brdf fb;
    std::memset(&fb, 0, sizeof(fb));
    fb.exponent = 1.0f;
    fb.refractive_index = 1.0f;
    if (const Phong_BRDF* x = dynamic_cast<const Phong_BRDF*>(r)) {
      fb.kind = B200RT_BRDF_PHONG; fb.exponent = x->phong_exponent_;
    } else if (const Modified_phong_BRDF* x = dynamic_cast<const Modified_phong_BRDF*>(r)) {
      fb.kind = B200RT_BRDF_MODIFIED_PHONG; fb.exponent = x->phong_exponent_; fb.normalized = x->normalized_;
    } else if (const Blinn_phong_BRDF* x = dynamic_cast<const Blinn_phong_BRDF*>(r)) {
      fb.kind = B200RT_BRDF_BLINN_PHONG; fb.exponent = x->phong_exponent_;
    } else if (const Modified_blinn_phong_BRDF* x = dynamic_cast<const Modified_blinn_phong_BRDF*>(r)) {
      fb.kind = B200RT_BRDF_MODIFIED_BLINN_PHONG; fb.exponent = x->phong_exponent_; fb.normalized = x->normalized_;
    } else if (const Torrance_sparrow_BRDF* x = dynamic_cast<const Torrance_sparrow_BRDF*>(r)) {
      fb.kind = B200RT_BRDF_TORRANCE_SPARROW; fb.exponent = x->exponent_; fb.refractive_index = x->refractive_index_;
    }
    f->brdfs.push_back(fb);
  }
#endif
#if B200RT_TREE >= 5
  // Scene::lights, in the order calculate_diffuse_and_specular_radiance loops them (HW7/src/Scene.cpp:230-262)
  for (const Light* l : scene.lights) {
    b200rt_light fl;
    std::memset(&fl, 0, sizeof(fl));
    fl.object = -1;
    if (const Point_light* x = dynamic_cast<const Point_light*>(l)) {
      fl.kind = B200RT_LIGHT_POINT;
      put3(fl.position, x->position_); put3(fl.intensity, x->intensity_);
    } else if (const Area_light* x = dynamic_cast<const Area_light*>(l)) {
      fl.kind = B200RT_LIGHT_AREA;
      put3(fl.position, x->position_); put3(fl.intensity, x->intensity_);
      put3(fl.edge1, x->edge_vector_1_); put3(fl.edge2, x->edge_vector_2_); put3(fl.normal, x->normal_);
    } else if (const Spot_light* x = dynamic_cast<const Spot_light*>(l)) {
      fl.kind = B200RT_LIGHT_SPOT;
      put3(fl.position, x->position_); put3(fl.intensity, x->intensity_); put3(fl.edge1, x->direction_);
      fl.edge2[0] = x->cos_half_of_coverage_angle_; fl.edge2[1] = x->cos_half_of_falloff_angle_;
    } else if (const Directional_light* x = dynamic_cast<const Directional_light*>(l)) {
      fl.kind = B200RT_LIGHT_DIRECTIONAL;
      put3(fl.edge1, x->direction_); put3(fl.intensity, x->radiance_);
#if B200RT_TREE >= 7
    } else if (const Spherical_directional_light* x = dynamic_cast<const Spherical_directional_light*>(l)) {
      fl.kind = B200RT_LIGHT_ENV;                            // the map rides along as one more image texture
      b200rt_texture ft;
      std::memset(&ft, 0, sizeof(ft));
      ft.kind = B200RT_TEX_IMAGE; ft.interp = B200RT_TEX_NEAREST; ft.decal = B200RT_TEX_REPLACE_ALL;
      ft.appearance = B200RT_TEX_CLAMP; ft.normalizer = 1.0f; ft.scaling = 1.0f; ft.bump_multiplier = 1.0f;
      ft.width = x->width_; ft.height = x->height_;
      ft.texel_offset = (int64_t)f->texels.size();
      for (size_t p = 0; p < (size_t)x->width_ * x->height_; p++)       // tinyexr hands back RGBA
        f->texels.insert(f->texels.end(), x->env_map_ + 4 * p, x->env_map_ + 4 * p + 3);
      f->textures.push_back(ft);
      fl.object = (int32_t)f->textures.size() - 1;
      d.env_light = (int32_t)f->lights.size();
    } else if (const Light_mesh* x = dynamic_cast<const Light_mesh*>(l)) {
      fl.kind = B200RT_LIGHT_MESH;                           // object / cdf range: when the walk meets it
      put3(fl.intensity, x->radiance_);
      fl.total_area = x->total_area_;
    } else if (const Light_sphere* x = dynamic_cast<const Light_sphere*>(l)) {
      fl.kind = B200RT_LIGHT_SPHERE;
      put3(fl.intensity, x->radiance_);
#endif
    }
    f->light_index[l] = (int)f->lights.size();
    f->lights.push_back(fl);
  }
#else   // HW4 still has HW3's two plain vectors; Scene::trace_ray loops the point lights, then the area lights
  for (const Point_light& l : scene.point_lights) {
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
#endif
  for (const Camera& c : scene.cameras) {
    b200rt_camera fc;
    std::memset(&fc, 0, sizeof(fc));
    const Image_plane& ip = c.get_image_plane();
    put3(fc.e, c.e); put3(fc.u, c.u); put3(fc.v, c.v); put3(fc.w, c.w);
    put3(fc.top_left, c.top_left_corner); put3(fc.su, c.s_u_constant); put3(fc.sv, c.s_v_constant);
    fc.left = ip.left; fc.right = ip.right; fc.bottom = ip.bottom; fc.top = ip.top; fc.distance = ip.distance;
    fc.width = ip.width; fc.height = ip.height;
    fc.samples_sqrt = c.get_number_of_samples() > 0 ? c.get_number_of_samples() : 1;   // already (int)sqrt(NumSamples): HW7/src/Scene.cpp:412
    fc.aperture = c.get_aperture_size();
    f->cameras.push_back(fc);
  }
  for (const Mesh* m : scene.meshes) f->add_mesh(m);
#if B200RT_TREE >= 7
  for (const Light* l : scene.lights)                        // a Light_mesh IS a Mesh, kept only in lights / the BVH
    if (const Light_mesh* lm = dynamic_cast<const Light_mesh*>(l)) {
      f->add_mesh(static_cast<const Mesh*>(lm));
      b200rt_light& fl = f->lights[f->light_index.at(l)];
      fl.cdf_begin = (int32_t)f->light_cdf.size();
      fl.cdf_count = (int32_t)lm->cdf_.size();
      for (size_t k = 0; k < lm->cdf_.size(); k++) {         // Light_mesh::cdf_ / triangles_ (HW7/include/Light_mesh.h:19-50)
        f->light_cdf.push_back(lm->cdf_[k]);
        f->light_tri.push_back((int32_t)f->tri_index.at(lm->triangles_[k]));
      }
    }
#endif
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
  d.n_textures = (int32_t)f->textures.size();    d.textures = f->textures.data();
  d.n_texels = (int64_t)f->texels.size();        d.texels = f->texels.data();
  d.n_brdfs = (int32_t)f->brdfs.size();          d.brdfs = f->brdfs.data();
  d.n_light_cdf = (int32_t)f->light_cdf.size();  d.light_cdf = f->light_cdf.data();
  d.light_tri = f->light_tri.data();
  return f;
}
extern "C" const B200SceneDesc* b200rt_flat_hw7_desc(const b200rt_flat_hw7* f) { return &f->desc; }
extern "C" void b200rt_flat_hw7_free(b200rt_flat_hw7* f) { delete f; }
