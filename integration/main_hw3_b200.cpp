// main_hw3_b200.cpp — what HW3/src/main.cpp becomes when the render fan-out goes to libb200rt.so.
//
// Everything around the hot path is still the reference's own code, linked from its untouched objects: `Scene::Scene`
// parses the XML and builds the BVHs, `Camera` names the output file, lodepng writes it, and the 8-bit conversion is
// `Pixel::get_color`'s rule (colour / weight, clamped to [0, 255], truncated — HW3/include/Pixel.h:18-30,
// HW3/src/main.cpp:44-52).  What is replaced is lines 27-39 of HW3/src/main.cpp: the std::thread fan-out over
// Scene::render_image.  No CPU fallback: without a device the program says why and exits non-zero.
//
//   g++ -std=c++14 -O2 -I HW3/include -I <repo>/include main_hw3_b200.cpp flatten_hw3.o <HW3 objects minus main.o> \
//       -L <repo>/ceng795_b200 -lb200rt -pthread
// Built here by `oracle/build_ref.sh b200main3` -> oracle/_ref/hw3_b200_raytracer; tests/test_z_dropin_binary.py runs it
// on the GPU box and compares its PNG with the one the unmodified HW3 binary wrote for the same XML.
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <iostream>
#include <vector>

#include "Scene.h"
#include "b200_host.h"
#include "b200rt.h"
#include "lodepng.h"

struct b200rt_flat_hw3;
extern "C" b200rt_flat_hw3* b200rt_flatten_hw3(const Scene& scene);
extern "C" const B200SceneDesc* b200rt_flat_hw3_desc(const b200rt_flat_hw3* f);
extern "C" void b200rt_flat_hw3_free(b200rt_flat_hw3* f);

int main(int argc, char* argv[]) {
  if (argc < 2) {
    std::cerr << "Please provide scene file as argument" << std::endl;
    return 1;
  }
  Scene scene(argv[1]);
  std::cout << "Scene is parsed" << std::endl;
  b200rt_flat_hw3* flat = b200rt_flatten_hw3(scene);
  B200Devices gpus;                       // one replica per GPU: the reference's thread_count (b200_host.h)
  if (gpus.create(b200rt_flat_hw3_desc(flat)) != 0) {
    b200rt_flat_hw3_free(flat);
    return 2;
  }
  AsyncWriter writer;                     // camera i's encode overlaps camera i+1's render
  int rc = 0;
  for (int index = 0; index < (int)scene.cameras.size() && rc == 0; index++) {
    const Camera& camera = scene.cameras[index];
    const int width = camera.get_image_plane().width, height = camera.get_image_plane().height;
    std::vector<float> rgbw((size_t)width * height * 4);     // Pixel{color, weight} of every pixel
    B200RenderOpts opts = {};
    opts.row_stride = 1;
    opts.seed = 795;
    B200RenderStats stats = {};
    const auto start = std::chrono::steady_clock::now();
    if (gpus.render(index, &opts, rgbw.data(), &stats) != B200RT_OK) {
      std::cerr << "b200rt: " << b200rt_last_error() << std::endl;
      rc = 3;
      break;
    }
    const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - start).count();
    std::vector<unsigned char> image((size_t)width * height * 4);
    for (size_t p = 0; p < (size_t)width * height; p++) {
      const float w = rgbw[4 * p + 3];
      for (int c = 0; c < 3; c++) {
        float v = (w == 0) ? 0.0f : rgbw[4 * p + c] / w;
        v = std::min(255.0f, std::max(0.0f, v));
        image[4 * p + c] = (unsigned char)(int)v;
      }
      image[4 * p + 3] = 255;
    }
    const std::string filename = camera.get_filename();
    writer.submit([filename, width, height, img = std::move(image)]() -> int {
      const unsigned error = lodepng::encode(filename.c_str(), img.data(), width, height);
      if (error) {
        std::cout << "encoder error " << error << ": " << lodepng_error_text(error) << std::endl;
        return 4;
      }
      return 0;
    });
    std::cout << camera.get_filename() << "(" << width << "x" << height << ") is saved in: " << ms << " milliseconds ("
              << (stats.rays_primary + stats.rays_shadow + stats.rays_secondary) << " rays, " << stats.launches
              << " kernel launches, " << gpus.size() << " GPU" << (gpus.size() > 1 ? "s" : "") << ")" << std::endl;
  }
  const int wrc = writer.finish();        // every file is on disk before the program reports success
  if (!rc) rc = wrc;
  gpus.destroy();
  b200rt_flat_hw3_free(flat);
  return rc;
}
