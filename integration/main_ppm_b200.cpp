// main_ppm_b200.cpp — what PPM/src/main.cpp becomes when its passes go to libb200rt.so.
//
// The reference's own host stays (XML loader, `Scene`, BVHs, lodepng).  Replaced, per camera (PPM/src/main.cpp:33-99):
//   reset_hash_grid + the eye_trace_lines fan-out + build_hash_grid   ->  b200rt_ppm_eye_pass
//   the trace_n_photons fan-out                                        ->  b200rt_ppm_trace_photons
//   density_estimation                                                 ->  b200rt_ppm_density
// with the SAME photon bookkeeping as the tree's main: every one of `thread_count` workers emits
// photon_count_per_iteration / thread_count photons in each of number_of_iterations rounds, and density_estimation is
// handed photon_count_per_iteration * photons_per_thread * thread_count as the photon total (sic — SURVEY Appendix
// B.7); thread_count is hardware_concurrency() here too, so the same XML gives the same normalisation on the same box.
// The 8-bit conversion is main's: int(pow(1 - exp(-x), 1 / 2.2f) * 255 + 0.5f), then write_png's clamp.
// No CPU fallback: without a device the program says why and exits non-zero.
//
// Built by `oracle/build_ref.sh b200mainppm` -> oracle/_ref/ppm_b200_raytracer.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <iostream>
#include <string>
#include <thread>
#include <vector>

#include "Scene.h"
#include "b200_host.h"
#include "b200rt.h"
#include "lodepng.h"

struct b200rt_flat_ppm;
extern "C" b200rt_flat_ppm* b200rt_flatten_ppm(const Scene& scene);
extern "C" const B200SceneDesc* b200rt_flat_ppm_desc(const b200rt_flat_ppm* f);
extern "C" void b200rt_flat_ppm_free(b200rt_flat_ppm* f);

static int fail(const char* what) {
  std::cerr << "b200rt: " << what << ": " << b200rt_last_error() << std::endl;
  return 3;
}

int main(int argc, char* argv[]) {
  if (argc < 2) {
    std::cerr << "Please provide scene file as argument" << std::endl;
    return 1;
  }
  Scene scene(argv[1]);
  std::cout << "Scene is parsed" << std::endl;
  int thread_count = (int)std::thread::hardware_concurrency();
  if (thread_count == 0) thread_count = 1;
  b200rt_flat_ppm* flat = b200rt_flatten_ppm(scene);
  B200Devices gpus;                       // one replica per GPU: main's photon fan-out over threads, over GPUs
  if (gpus.create(b200rt_flat_ppm_desc(flat)) != 0) {
    b200rt_flat_ppm_free(flat);
    return 2;
  }
  b200rt_scene* gpu = gpus.scenes[0];
  const int n_gpus = gpus.size();
  AsyncWriter writer;                     // camera i's encode overlaps camera i+1's passes
  int rc = 0;
  for (int index = 0; index < (int)scene.cameras.size() && rc == 0; index++) {
    const Camera& camera = scene.cameras[index];
    const int width = camera.get_image_plane().width, height = camera.get_image_plane().height;
    const size_t n = (size_t)width * height;
    const uint64_t seed = 795;
    B200PpmStats eye = {}, pass = {};
    const auto t0 = std::chrono::steady_clock::now();
    if ((n_gpus > 1 ? b200rt_ppm_eye_pass_multi(gpus.scenes.data(), n_gpus, index, seed, &eye)
                    : b200rt_ppm_eye_pass(gpu, index, seed, &eye)) != B200RT_OK) { rc = fail("eye pass"); break; }
    // main's bookkeeping, in 64 bits (the tree multiplies ints)
    const long long per_iteration = scene.photon_count_per_iteration;
    const long long photons_per_thread = height < thread_count ? per_iteration : per_iteration / thread_count;
    const long long workers = height < thread_count ? 1 : thread_count;
    B200PpmOpts opts = {};
    opts.photons = photons_per_thread * workers * (long long)scene.number_of_iterations;
    opts.first_photon = 0;
    opts.seed = seed;
    if (n_gpus > 1) {
      // one call per iteration, as main's loop (PPM/src/main.cpp:74-98): the photons of an iteration are split over the
      // GPUs by index and the accepted-photon counts are summed across them before the next iteration starts
      const long long per_call = photons_per_thread * workers;
      B200PpmStats one = {};
      for (long long it = 0; it < (long long)scene.number_of_iterations && rc == 0; it++) {
        B200PpmOpts o = opts;
        o.photons = per_call;
        o.first_photon = (uint64_t)(it * per_call);
        if (b200rt_ppm_trace_photons_multi(gpus.scenes.data(), n_gpus, &o, &one) != B200RT_OK) { rc = fail("photon pass"); break; }
        pass.photons += one.photons; pass.photon_rays += one.photon_rays; pass.launches += one.launches;
      }
      if (rc) break;
    } else if (b200rt_ppm_trace_photons(gpu, &opts, &pass) != B200RT_OK) { rc = fail("photon pass"); break; }
    std::vector<float> rgbw(n * 4);
    const double total = (double)(per_iteration * (per_iteration / thread_count) * thread_count);   // PPM/src/main.cpp:97-98
    if ((n_gpus > 1 ? b200rt_ppm_density_multi(gpus.scenes.data(), n_gpus, total, rgbw.data())
                    : b200rt_ppm_density(gpu, total, rgbw.data())) != B200RT_OK) { rc = fail("density estimation"); break; }
    const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    std::vector<unsigned char> image(n * 4, 255);
    for (size_t p = 0; p < n; p++) {
      const float w = rgbw[4 * p + 3];
      for (int c = 0; c < 3; c++) {
        const float x = (w == 0) ? 0.0f : rgbw[4 * p + c] / w;                                   // Pixel::get_color
        const float v = (float)int(std::pow(1 - std::exp(-x), 1 / 2.2f) * 255 + 0.5f);             // PPM/src/main.cpp:150-157
        image[4 * p + c] = (unsigned char)std::min(255.0f, std::max(0.0f, v));
      }
    }
    const std::string stem = camera.get_filename().substr(0, camera.get_filename().find_last_of("."));
    const std::string png = stem + ".png";
    writer.submit([png, width, height, img = std::move(image)]() -> int {
      const unsigned error = lodepng::encode(png.c_str(), img.data(), width, height);
      if (error) {
        std::cout << "encoder error " << error << ": " << lodepng_error_text(error) << std::endl;
        return 4;
      }
      return 0;
    });
    std::cout << stem << "(" << width << "x" << height << ") is saved in: " << ms << " milliseconds (" << eye.hit_points
              << " hit points, " << pass.photons << " photons, " << pass.photon_rays << " photon rays, "
              << (eye.launches + pass.launches) << " kernel launches)" << std::endl;
  }
  const int wrc = writer.finish();        // every file is on disk before the program reports success
  if (!rc) rc = wrc;
  gpus.destroy();
  b200rt_flat_ppm_free(flat);
  return rc;
}
