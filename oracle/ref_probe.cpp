// TEST INFRASTRUCTURE ONLY — known-answer vectors for the leaf functions of the hot path.
// Compiled by oracle/build_ref.sh (-fno-access-control) against the UNMODIFIED objects of the HW7 tree, it calls the
// reference's own functions on seeded inputs and writes inputs and outputs to one file; tests/test_leaf_vectors.py
// feeds the same inputs to the oracle's hooks and demands the same bits.
//   ref_probe <out.bin> <asset_dir>
// File: repeated { char name[24]; int32 rows, n_in, n_out; float32 rows*(n_in+n_out) }.
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "Blinn_phong_BRDF.h"
#include "Bounding_box.h"
#include "Hit_data.h"
#include "Mesh_triangle.h"
#include "Area_light.h"
#include "Scene.h"
#include "Sphere.h"
#include "Spot_light.h"
#include <cmath>
#include "Transformation.h"
#include "Vertex.h"
#include <new>
#include "Modified_blinn_phong_BRDF.h"
#include "Modified_phong_BRDF.h"
#include "Perlin_noise.h"
#include "Phong_BRDF.h"
#include "Ray.h"
#include "Texture.h"
#include "Torrance_sparrow_BRDF.h"

namespace {
struct Lcg {                                   // our own generator: the vectors must not depend on libstdc++'s <random>
  uint64_t s;
  explicit Lcg(uint64_t seed) : s(seed * 2862933555777941757ull + 3037000493ull) {}
  float u01() {
    s = s * 6364136223846793005ull + 1442695040888963407ull;
    return (float)((s >> 40) & 0xFFFFFF) * (1.0f / 16777216.0f);
  }
  float range(float a, float b) { return a + (b - a) * u01(); }
};
std::FILE* g_out;
void section(const char* name, int rows, int n_in, int n_out, const std::vector<float>& data) {
  char nm[24] = {0};
  std::strncpy(nm, name, 23);
  const int32_t hdr[3] = {rows, n_in, n_out};
  std::fwrite(nm, 24, 1, g_out);
  std::fwrite(hdr, sizeof(hdr), 1, g_out);
  std::fwrite(data.data(), sizeof(float), data.size(), g_out);
}
Vector3 unit(Lcg& r) {
  Vector3 v(r.range(-1, 1), r.range(-1, 1), r.range(-1, 1));
  return v.normalize();
}
}  // namespace

int main(int argc, char** argv) {
  if (argc < 3) return 1;
  g_out = std::fopen(argv[1], "wb");
  if (!g_out) return 2;
  const std::string assets = argv[2];
  {  // Bounding_box::intersect — HW7/src/bounding_box.cpp:19-39
    Lcg r(11);
    std::vector<float> d;
    const int rows = 2048;
    for (int k = 0; k < rows; k++) {
      Vector3 a(r.range(-3, 3), r.range(-3, 3), r.range(-3, 3)), b(r.range(-3, 3), r.range(-3, 3), r.range(-3, 3));
      Vector3 mn(std::min(a.x, b.x), std::min(a.y, b.y), std::min(a.z, b.z));
      Vector3 mx(std::max(a.x, b.x), std::max(a.y, b.y), std::max(a.z, b.z));
      Vector3 o(r.range(-5, 5), r.range(-5, 5), r.range(-5, 5));
      Vector3 dir = unit(r);
      if (k % 8 == 1) dir.x = 0.0f;                       // axis-parallel: the |d| < 1e-6 skip
      if (k % 8 == 2) { dir.y = 5e-7f; }
      if (k % 8 == 3) { dir.x = 0.0f; dir.z = -0.0f; }
      if (k % 16 == 4) o = (mn + mx) / 2.0f;             // origin inside the box
      if (k % 16 == 5) dir = dir * 37.5f;                 // un-normalised direction (instances)
      const float t = Bounding_box(mn, mx).intersect(Ray(o, dir, r_primary));
      const float row[13] = {mn.x, mn.y, mn.z, mx.x, mx.y, mx.z, o.x, o.y, o.z, dir.x, dir.y, dir.z, t};
      d.insert(d.end(), row, row + 13);
    }
    section("box", rows, 12, 1, d);
  }
  {  // BRDF::get_reflectance x5 — HW7/src/*_BRDF.cpp
    Phong_BRDF b0(23.0f);
    Modified_phong_BRDF b1(31.5f, true), b1n(12.0f, false);
    Blinn_phong_BRDF b2(40.0f);
    Modified_blinn_phong_BRDF b3(50.0f, true), b3n(7.0f, false);
    Torrance_sparrow_BRDF b4(60.0f, 2.1f);
    const BRDF* all[] = {&b0, &b1, &b1n, &b2, &b3, &b3n, &b4};
    const char* names[] = {"brdf0_e23", "brdf1_e31.5_n1", "brdf1_e12_n0", "brdf2_e40", "brdf3_e50_n1", "brdf3_e7_n0",
                           "brdf4_e60_r2.1"};
    for (int q = 0; q < 7; q++) {
      Lcg r(100 + q);
      std::vector<float> d;
      const int rows = 1024;
      for (int k = 0; k < rows; k++) {
        Hit_data h;
        h.normal = unit(r);
        Vector3 wi = unit(r), wo = unit(r);
        if (k % 4 != 0 && h.normal.dot(wi) < 0) wi = -wi;   // mostly the lit hemisphere, some below it
        if (k % 3 != 0 && h.normal.dot(wo) < 0) wo = -wo;
        if (k % 64 == 7) wi = h.normal;                       // cos = 1 (or a hair above it)
        const Vector3 kd(r.u01(), r.u01(), r.u01()), ks(r.u01(), r.u01(), r.u01());
        const Vector3 f = all[q]->get_reflectance(h, kd, ks, wi, wo);
        const float row[18] = {h.normal.x, h.normal.y, h.normal.z, kd.x, kd.y, kd.z, ks.x, ks.y, ks.z,
                               wi.x, wi.y, wi.z, wo.x, wo.y, wo.z, f.x, f.y, f.z};
        d.insert(d.end(), row, row + 18);
      }
      section(names[q], rows, 15, 3, d);
    }
  }
  {  // Texture::get_color_at, all four interpolation x appearance combinations — HW7/src/Texture.cpp:66-143
    const char* interp[] = {"nearest", "bilinear"};
    const char* app[] = {"repeat", "clamp"};
    for (int a = 0; a < 2; a++)
      for (int b = 0; b < 2; b++) {
        Texture T(assets + "/tex_checker.png", interp[a], "replace_kd", app[b], 255.0f, 1.0f, false, 1.0f, false);
        Lcg r(200 + 2 * a + b);
        std::vector<float> d;
        const int rows = 2048;
        for (int k = 0; k < rows; k++) {
          float u = r.range(-1.5f, 2.5f), v = r.range(-1.5f, 2.5f);
          if (k % 32 == 0) u = 1.0f;
          if (k % 32 == 1) v = 1.0f;
          if (k % 32 == 2) { u = 0.0f; v = 0.0f; }
          if (k % 32 == 3) { u = 0.999999f; v = 0.999999f; }
          const Vector3 c = T.get_color_at(u, v);
          const float row[5] = {u, v, c.x, c.y, c.z};
          d.insert(d.end(), row, row + 5);
        }
        section((std::string("tex_") + interp[a] + "_" + app[b]).c_str(), rows, 2, 3, d);
      }
  }
  {  // Perlin_noise::get_value_at with a FIXED table (the constructor shuffles it from the clock) — Perlin_noise.cpp:23-62
    const int perm[16] = {11, 3, 7, 0, 14, 9, 2, 13, 5, 8, 1, 15, 6, 12, 4, 10};
    const char* app[] = {"vein", "patch"};
    for (int a = 0; a < 2; a++) {
      Perlin_noise P(app[a], 3.5f);
      P.p_.assign(perm, perm + 16);                          // -fno-access-control
      Lcg r(300 + a);
      std::vector<float> d;
      const int rows = 2048;
      for (int k = 0; k < rows; k++) {
        const Vector3 p(r.range(-6, 6), r.range(-6, 6), r.range(-6, 6));
        const float v = P.get_value_at(p);
        const float row[4] = {p.x, p.y, p.z, v};
        d.insert(d.end(), row, row + 4);
      }
      section((std::string("perlin_") + app[a]).c_str(), rows, 3, 1, d);
    }
  }
  {  // Mesh_triangle::intersect — HW7/src/Mesh_triangle.cpp:56-100 (the +-eps band on beta / gamma / t, no culling)
    // The class reads its vertices through Scene::get_vertex_at, i.e. Scene::vertex_data; a Scene cannot be made
    // without an XML file, so only that one member is brought to life inside raw storage (-fno-access-control).
    alignas(Scene) static unsigned char scene_storage[sizeof(Scene)];
    Scene* scene = reinterpret_cast<Scene*>(scene_storage);
    new (&scene->vertex_data) std::vector<Vertex>();
    Lcg r(400);
    std::vector<float> d;
    const int rows = 6144;
    for (int k = 0; k < rows; k++) {
      Vector3 v0(r.range(-2, 2), r.range(-2, 2), r.range(-2, 2)), v1(r.range(-2, 2), r.range(-2, 2), r.range(-2, 2)),
          v2(r.range(-2, 2), r.range(-2, 2), r.range(-2, 2));
      if (k % 32 == 7) v2 = v0 + (v1 - v0) * r.range(0.1f, 0.9f);            // zero-area triangle
      if (k % 32 == 9) { v1 = v0 + Vector3(1e-4f, 0, 0); v2 = v0 + Vector3(0, 1e-4f, 0); }   // tiny
      Vector3 o(r.range(-5, 5), r.range(-5, 5), r.range(-5, 5));
      // aim at a point of the triangle's plane: inside, exactly on an edge, exactly on a vertex, just outside, far outside
      float b = r.u01(), g = r.u01() * (1 - b);
      switch (k % 8) {
        case 1: g = 0.0f; break;                       // on the edge v0-v1
        case 2: b = 0.0f; break;                       // on the edge v0-v2
        case 3: g = 1.0f - b; break;                   // on the edge v1-v2
        case 4: b = (k & 8) ? 1.0f : 0.0f; g = (k & 16) ? 0.0f : 1.0f - b; break;   // a vertex
        case 5: b = -r.u01() * 2e-6f; break;           // inside the acceptance band, or just beyond it
        case 6: b = r.range(-1.5f, 2.5f); g = r.range(-1.5f, 2.5f); break;
        default: break;
      }
      const Vector3 target = v0 + (v1 - v0) * b + (v2 - v0) * g;
      Vector3 dir = target - o;
      if (k % 16 == 11) dir = dir * -1.0f;              // triangle behind the origin: t < 0
      if (k % 4 != 0) dir = dir.normalize();            // un-normalised directions too (instances hand those down)
      if (k % 64 == 13) dir = (v1 - v0) * 0.7f + (v2 - v0) * 0.2f;   // parallel to the plane: det = 0 or nearly
      scene->vertex_data.clear();
      scene->vertex_data.push_back(Vertex(v0));
      scene->vertex_data.push_back(Vertex(v1));
      scene->vertex_data.push_back(Vertex(v2));
      const Mesh_triangle tri(scene, 0, 1, 2, 0, 0, 0, -1, tsm_flat);
      Hit_data hd;
      hd.t = -12345.0f;
      const bool hit = tri.intersect(Ray(o, dir, r_primary), hd, false);
      const float row[17] = {v0.x, v0.y, v0.z, v1.x, v1.y, v1.z, v2.x, v2.y, v2.z, o.x, o.y, o.z, dir.x, dir.y, dir.z,
                             hit ? 1.0f : 0.0f, hit ? hd.t : 0.0f};
      d.insert(d.end(), row, row + 17);
    }
    section("tri7", rows, 15, 2, d);
  }
  {  // Sphere::intersect — HW7/src/Sphere.cpp:32-110: per-ray Translation(time * velocity) * M, inverse, quadratic, the
     // t2 < 0 -> t1 rule, normal through the normal matrix.  M is composed by the tree's own Scaling / Rotation /
     // Translation classes and handed to the test as 16 floats.
    Lcg r(500);
    std::vector<float> d;
    const int rows = 4096;
    for (int k = 0; k < rows; k++) {
      Matrix4x4 M(true);
      if (k % 4 != 0) {
        const Scaling S(r.range(0.3f, 2.0f) * ((k % 8 == 5) ? -1.0f : 1.0f), r.range(0.3f, 2.0f), r.range(0.3f, 2.0f));
        const Rotation R(r.range(0, 360), r.range(-1, 1), r.range(-1, 1), r.range(-1, 1));
        const Translation T(r.range(-2, 2), r.range(-2, 2), r.range(-2, 2));
        M = T.get_transformation_matrix() * (R.get_transformation_matrix() * S.get_transformation_matrix());
      }
      const Vector3 center(r.range(-1, 1), r.range(-1, 1), r.range(-1, 1));
      const float radius = r.range(0.2f, 1.5f);
      const Vector3 velocity = (k % 4 == 3) ? Vector3(r.range(-0.5f, 0.5f), r.range(-0.5f, 0.5f), r.range(-0.5f, 0.5f)) : Vector3(0.0f);
      const float time = (k % 4 == 3) ? r.u01() : 0.0f;
      const Arbitrary_transformation xf(M);
      const Sphere sphere(nullptr, center, radius, 0, -1, xf, velocity);
      Vector3 o(r.range(-6, 6), r.range(-6, 6), r.range(-6, 6));
      const Vector3 world_c = M.multiply(center);
      if (k % 8 == 2) o = world_c + Vector3(r.range(-0.1f, 0.1f), r.range(-0.1f, 0.1f), r.range(-0.1f, 0.1f));   // origin inside: t2 < 0 -> t1
      Vector3 aim = world_c + Vector3(r.range(-1, 1), r.range(-1, 1), r.range(-1, 1)) * (radius * ((k % 16 == 9) ? 2.5f : 0.9f));
      Vector3 dir = aim - o;
      if (k % 16 == 6) dir = dir * -1.0f;                 // sphere behind the origin: negative t is still returned
      if (k % 2) dir = dir.normalize();
      Hit_data hd;
      hd.t = -12345.0f;
      const bool hit = sphere.intersect(Ray(o, dir, r_primary, time), hd, false);
      float row[35];
      for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) row[4 * i + j] = M[i][j];
      const float rest[14] = {center.x, center.y, center.z, radius, velocity.x, velocity.y, velocity.z, o.x, o.y, o.z,
                              dir.x, dir.y, dir.z, time};
      for (int i = 0; i < 14; i++) row[16 + i] = rest[i];
      row[30] = hit ? 1.0f : 0.0f; row[31] = hit ? hd.t : 0.0f;
      row[32] = hit ? hd.normal.x : 0.0f; row[33] = hit ? hd.normal.y : 0.0f; row[34] = hit ? hd.normal.z : 0.0f;
      d.insert(d.end(), row, row + 35);
    }
    section("sphere7", rows, 30, 5, d);
  }
  {  // Light::incoming_radiance — HW7/src/Spot_light.cpp:24-46 (inside the falloff cone, the pow(c, 4) ramp, outside
     // the coverage cone) and HW7/src/Area_light.cpp (cosine at the light over squared distance)
    Lcg r(600);
    std::vector<float> d;
    const int rows = 3072;
    for (int k = 0; k < rows; k++) {
      const Vector3 pos(r.range(-3, 3), r.range(2, 5), r.range(-3, 3)), inten(r.range(100, 3000), r.range(100, 3000), r.range(100, 3000));
      const float cov = r.range(30, 120) * (float)(M_PI / 180.0), fall = cov * r.range(0.2f, 0.9f);
      const Spot_light L(pos, inten, Vector3(r.range(-0.4f, 0.4f), -1.0f, r.range(-0.4f, 0.4f)), cov, fall);
      // a point around the axis: angle from the axis swept over all three zones
      const float ang = r.range(0.0f, 1.3f) * cov * 0.5f, az = r.range(0, 6.2831853f);
      const Vector3 ax = L.direction_;
      const Vector3 t1 = ax.cross(Vector3(0.3f, 0.1f, 0.9f)).normalize(), t2 = ax.cross(t1);
      const Vector3 w = ax * std::cos(ang) + (t1 * std::cos(az) + t2 * std::sin(az)) * std::sin(ang);
      const Vector3 from = pos + w * r.range(0.5f, 8.0f);
      float dist = 0, prob = 0;
      const Vector3 to_light = L.direction_and_distance(from, Vector3(0, 1, 0), dist, prob);
      const Vector3 rad = L.incoming_radiance(to_light, prob);
      const float row[18] = {pos.x, pos.y, pos.z, inten.x, inten.y, inten.z, L.direction_.x, L.direction_.y, L.direction_.z,
                             L.cos_half_of_coverage_angle_, L.cos_half_of_falloff_angle_, to_light.x, to_light.y, to_light.z,
                             dist, rad.x, rad.y, rad.z};
      d.insert(d.end(), row, row + 18);
    }
    section("light_spot", rows, 15, 3, d);
    d.clear();
    for (int k = 0; k < 2048; k++) {
      const Vector3 pos(r.range(-3, 3), r.range(2, 5), r.range(-3, 3)), inten(r.range(100, 3000), r.range(100, 3000), r.range(100, 3000));
      const Area_light L(pos, inten, Vector3(r.range(0.5f, 2), 0, r.range(-0.5f, 0.5f)), Vector3(r.range(-0.5f, 0.5f), 0, r.range(0.5f, 2)));
      const Vector3 to_light(r.range(-4, 4), r.range(-4, 4), r.range(-4, 4));
      const Vector3 rad = L.incoming_radiance(to_light, 1.0f);
      const float row[12] = {inten.x, inten.y, inten.z, L.normal_.x, L.normal_.y, L.normal_.z, to_light.x, to_light.y, to_light.z,
                             rad.x, rad.y, rad.z};
      d.insert(d.end(), row, row + 12);
    }
    section("light_area", 2048, 9, 3, d);
  }
  std::fclose(g_out);
  return 0;
}
