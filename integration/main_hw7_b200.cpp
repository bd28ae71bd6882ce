// main_hw7_b200.cpp — what HW7/src/main.cpp becomes when the render fan-out goes to libb200rt.so.
//
// The reference's own host stays: `Scene::Scene` parses the XML, loads textures (stb_image / tinyexr), builds the
// BVHs; `Camera` names the output; lodepng writes it.  Replaced: the std::thread fan-out over Scene::render_image
// (HW7/src/main.cpp:31-47) by b200rt_render, and — for a camera that has a tonemapping operator —
// Photographic_tmo::apply_tmo plus the sRGB loop (HW7/src/main.cpp:59-91) by b200rt_tonemap_photographic on the frame
// that is still in HBM.  (The side .exr of a tonemapped camera is host I/O through tinyexr in the tree's main.cpp and
// is not reproduced here.)  No CPU fallback: without a device the program says why and exits non-zero.
//
// Built by `oracle/build_ref.sh b200main7` -> oracle/_ref/hw7_b200_raytracer (-fno-access-control: the operator's
// three parameters are private); tests/test_z_dropin_binary.py compares its PNG with the unmodified binary's.
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <iostream>
#include <string>
#include <vector>

#include "Photographic_tmo.h"
#include "Scene.h"
#include "b200_host.h"
#include "b200rt.h"
#include "lodepng.h"

struct b200rt_flat_hw7;
extern "C" b200rt_flat_hw7* b200rt_flatten_hw7(const Scene& scene);
extern "C" const B200SceneDesc* b200rt_flat_hw7_desc(const b200rt_flat_hw7* f);
extern "C" void b200rt_flat_hw7_free(b200rt_flat_hw7* f);

int main(int argc, char* argv[]) {
  if (argc < 2) {
    std::cerr << "Please provide scene file as argument" << std::endl;
    return 1;
  }
  Scene scene(argv[1]);
  std::cout << "Scene is parsed" << std::endl;
  b200rt_flat_hw7* flat = b200rt_flatten_hw7(scene);
  B200Devices gpus;                       // one replica per GPU: the reference's thread_count (b200_host.h)
  if (gpus.create(b200rt_flat_hw7_desc(flat)) != 0) {
    b200rt_flat_hw7_free(flat);
    return 2;
  }
  AsyncWriter writer;                     // camera i's encode overlaps camera i+1's render (HW7/src/main.cpp:105-198)
  int rc = 0;
  for (int index = 0; index < (int)scene.cameras.size() && rc == 0; index++) {
    const Camera& camera = scene.cameras[index];
    const int width = camera.get_image_plane().width, height = camera.get_image_plane().height;
    const size_t n = (size_t)width * height;
    std::vector<float> rgbw(n * 4);                          // Pixel{color, weight} of every pixel
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
    std::vector<unsigned char> image(n * 4, 255);
    const Photographic_tmo* tmo = dynamic_cast<const Photographic_tmo*>(camera.get_tmo());
    if (tmo) {
      std::vector<unsigned char> rgb8(n * 3);
      if (b200rt_tonemap_photographic(gpus.scenes[0], index, tmo->image_key_, tmo->saturation_percentage_, tmo->saturation_,
                                      rgb8.data()) != B200RT_OK) {
        std::cerr << "b200rt: " << b200rt_last_error() << std::endl;
        rc = 3;
        break;
      }
      for (size_t p = 0; p < n; p++)
        for (int c = 0; c < 3; c++) image[4 * p + c] = rgb8[3 * p + c];
    } else {
      for (size_t p = 0; p < n; p++) {                        // Pixel::get_color, then write_png's clamp and store
        const float w = rgbw[4 * p + 3];
        for (int c = 0; c < 3; c++) {
          const float v = (w == 0) ? 0.0f : rgbw[4 * p + c] / w;
          image[4 * p + c] = (unsigned char)std::min(255.0f, std::max(0.0f, v));
        }
      }
    }
    const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - start).count();
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
    std::cout << stem << "(" << width << "x" << height << ") is saved in: " << ms << " milliseconds ("
              << (stats.rays_primary + stats.rays_shadow + stats.rays_secondary) << " rays, " << stats.launches
              << " kernel launches, " << gpus.size() << " GPU" << (gpus.size() > 1 ? "s" : "") << ")" << std::endl;
  }
  const int wrc = writer.finish();        // every file is on disk before the program reports success
  if (!rc) rc = wrc;
  gpus.destroy();
  b200rt_flat_hw7_free(flat);
  return rc;
}
