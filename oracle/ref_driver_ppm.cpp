// TEST INFRASTRUCTURE ONLY — pins the deterministic half of the PPM path (rows a16 / a17 of SURVEY.md §8).
// The PPM tree's main.cpp goes straight on to the (randomly seeded) photon pass; this driver, linked by
// oracle/build_ref.sh with the UNMODIFIED objects of the tree, stops after the two deterministic steps:
//   reset_hash_grid(); eye_trace_lines(camera, 0, 1); build_hash_grid(W, H);      (PPM/src/main.cpp:35-61)
// and dumps what they produced:  <out.bin> = int32 n_hp, num_hash; float32 initial_radius^2, hash_scale, bbox_min[3];
//   n_hp x 16 float32 {position, normal, w_o, attenuation, diffuse.x (material tag), pixel, pixel_weight, 0};
//   num_hash x int32 = size of every hash-grid cell.
//   ref_driver_ppm scene.xml out.bin      (NumSamples must be 1: the multi-sample eye pass draws from random_device)
// With a third argument N it goes on, still on this one thread, with the tree's own photon pass and density estimate:
//   trace_n_photons(N, 1); density_estimation(pixels, N);                          (PPM/src/main.cpp:77-98)
// and appends  n_hp x 5 float32 {flux.rgb, radius_squared, n}  and  W*H x 3 float32 Pixel::get_color().
// The tree seeds its three function-local mt19937 engines from std::random_device; this file defines
// std::random_device::_M_getval() itself (the executable's definition wins over libstdc++'s), so every engine starts
// from the constant kSeed and the photon pass becomes one reproducible sequence the oracle can replay draw for draw.
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#include <random>
#include <vector>

#include "Pixel.h"
#include "Scene.h"

static const unsigned kSeed = 795u;
unsigned int std::random_device::_M_getval() { return kSeed; }

int main(int argc, char** argv) {
  if (argc < 3) return 1;
  Scene scene(argv[1]);
  const Image_plane& plane = scene.cameras[0].get_image_plane();
  scene.reset_hash_grid();
  scene.eye_trace_lines(0, 0, 1);
  scene.build_hash_grid(plane.width, plane.height);
  std::FILE* f = std::fopen(argv[2], "wb");
  if (!f) return 2;
  const int32_t n[2] = {(int32_t)scene.hit_points.size(), (int32_t)scene.num_hash};
  std::fwrite(n, sizeof(n), 1, f);
  const float r2 = scene.hit_points.empty() ? 0.0f : scene.hit_points[0]->radius_squared;
  const float hdr[5] = {r2, scene.hash_scale, scene.hit_point_bbox.min_corner.x, scene.hit_point_bbox.min_corner.y,
                        scene.hit_point_bbox.min_corner.z};
  std::fwrite(hdr, sizeof(hdr), 1, f);
  for (const Hit_point* h : scene.hit_points) {
    const float row[16] = {h->position.x, h->position.y, h->position.z, h->normal.x, h->normal.y, h->normal.z,
                           h->w_o.x, h->w_o.y, h->w_o.z, h->attenuation.x, h->attenuation.y, h->attenuation.z,
                           h->material.diffuse.x, (float)h->pixel, h->pixel_weight, 0.0f};
    std::fwrite(row, sizeof(row), 1, f);
  }
  for (const auto& cell : scene.hash_grid) {
    const int32_t k = (int32_t)cell.size();
    std::fwrite(&k, 4, 1, f);
  }
  if (argc > 3) {
    const int n_photons = std::atoi(argv[3]);
    scene.trace_n_photons(n_photons, 1);
    std::vector<Pixel> film((size_t)plane.width * plane.height);
    scene.density_estimation(film.data(), n_photons);
    for (const Hit_point* h : scene.hit_points) {
      const float row[5] = {h->flux.x, h->flux.y, h->flux.z, h->radius_squared, (float)h->n};
      std::fwrite(row, sizeof(row), 1, f);
    }
    for (size_t p = 0; p < film.size(); p++) {
      const Vector3 c = film[p].get_color();
      const float rgb[3] = {c.x, c.y, c.z};
      std::fwrite(rgb, sizeof(rgb), 1, f);
    }
  }
  std::fclose(f);
  return 0;
}
