// main_hw1_b200.cpp — what HW1/main.cpp becomes when the render fan-out goes to libb200rt.so.
// The reference's own host stays (XML loader, `Scene`, lodepng); replaced: the std::thread fan-out over
// Scene::render_image (HW1/main.cpp:24-36).  The 8-bit rule is render_image's own: min(255, max(0, int(c)))
// (HW1/Scene.cpp:20-22).  No CPU fallback: without a device the program says why and exits non-zero.
// Built by `oracle/build_ref.sh b200main1` -> oracle/_ref/hw1_b200_raytracer.
#include <algorithm>
#include <chrono>
#include <iostream>
#include <vector>

#include "Scene.h"
#include "b200_host.h"
#include "b200rt.h"
#include "lodepng/lodepng.h"

struct b200rt_flat_hw1;
extern "C" b200rt_flat_hw1* b200rt_flatten_hw1(const Scene& scene);
extern "C" const B200SceneDesc* b200rt_flat_hw1_desc(const b200rt_flat_hw1* f);
extern "C" void b200rt_flat_hw1_free(b200rt_flat_hw1* f);

int main(int argc, char* argv[]) {
  if (argc < 2) {
    std::cerr << "Please provide scene file as argument" << std::endl;
    return 1;
  }
  Scene scene(argv[1]);
  std::cout << "Scene is parsed" << std::endl;
  b200rt_flat_hw1* flat = b200rt_flatten_hw1(scene);
  B200Devices gpus;                       // one replica per GPU: the reference's thread_count (b200_host.h)
  if (gpus.create(b200rt_flat_hw1_desc(flat)) != 0) {
    b200rt_flat_hw1_free(flat);
    return 2;
  }
  AsyncWriter writer;                     // camera i's encode overlaps camera i+1's render
  int rc = 0;
  for (int index = 0; index < (int)scene.cameras.size() && rc == 0; index++) {
    const Camera& camera = scene.cameras[index];
    const int width = camera.get_image_plane().width, height = camera.get_image_plane().height;
    const size_t n = (size_t)width * height;
    std::vector<float> rgbw(n * 4);
    B200RenderOpts opts = {};
    opts.row_stride = 1;
    B200RenderStats stats = {};
    const auto start = std::chrono::steady_clock::now();
    if (gpus.render(index, &opts, rgbw.data(), &stats) != B200RT_OK) {
      std::cerr << "b200rt: " << b200rt_last_error() << std::endl;
      rc = 3;
      break;
    }
    const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - start).count();
    std::vector<unsigned char> image(n * 4, 255);
    for (size_t p = 0; p < n; p++) {
      const float w = rgbw[4 * p + 3];
      for (int c = 0; c < 3; c++) {
        const float v = (w == 0) ? 0.0f : rgbw[4 * p + c] / w;
        image[4 * p + c] = (unsigned char)std::min(255, std::max(0, int(v)));
      }
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
  b200rt_flat_hw1_free(flat);
  return rc;
}
