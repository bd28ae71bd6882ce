// TEST INFRASTRUCTURE ONLY — runs integration/flatten_hw3.cpp (the reference-side half of the drop-in boundary) on the
// reference's own `Scene`.  Linked by oracle/build_ref.sh (target flat3; flat7 with -DFLATTEN_HW7) with the UNMODIFIED
// objects of /root/reference/HW3 (HW7) minus main.cpp, plus ceng795_b200/csrc/scene_cache.cpp for the file format:
//   flat3_flatten <scene.xml> <out.b200rt>
// parses the XML with the reference's loader, flattens the objects it built and writes the description with
// b200rt_scene_desc_save.  tests/test_flatten_reference.py compares that file with what b200rt_builder_* makes of the
// same XML.
#include <cstdio>
#include <string>

#include "Scene.h"
#include "b200rt.h"

#ifdef FLATTEN_HW7   // the same driver against integration/flatten_hw7.cpp and the HW7 tree's objects (target flat7)
#define b200rt_flat_hw3 b200rt_flat_hw7
#define b200rt_flatten_hw3 b200rt_flatten_hw7
#define b200rt_flat_hw3_desc b200rt_flat_hw7_desc
#define b200rt_flat_hw3_free b200rt_flat_hw7_free
#endif
#ifdef FLATTEN_HW1   // ... integration/flatten_hw1.cpp and the HW1 tree's objects (target flat1)
#define b200rt_flat_hw3 b200rt_flat_hw1
#define b200rt_flatten_hw3 b200rt_flatten_hw1
#define b200rt_flat_hw3_desc b200rt_flat_hw1_desc
#define b200rt_flat_hw3_free b200rt_flat_hw1_free
#endif
#ifdef FLATTEN_HW2   // ... integration/flatten_hw2.cpp and the HW2 tree's objects (target flat2)
#define b200rt_flat_hw3 b200rt_flat_hw2
#define b200rt_flatten_hw3 b200rt_flatten_hw2
#define b200rt_flat_hw3_desc b200rt_flat_hw2_desc
#define b200rt_flat_hw3_free b200rt_flat_hw2_free
#endif
#ifdef FLATTEN_PPM   // ... and against integration/flatten_ppm.cpp and the PPM tree's objects (target flatppm)
#define b200rt_flat_hw3 b200rt_flat_ppm
#define b200rt_flatten_hw3 b200rt_flatten_ppm
#define b200rt_flat_hw3_desc b200rt_flat_ppm_desc
#define b200rt_flat_hw3_free b200rt_flat_ppm_free
#endif
struct b200rt_flat_hw3;
extern "C" b200rt_flat_hw3* b200rt_flatten_hw3(const Scene& scene);
extern "C" const B200SceneDesc* b200rt_flat_hw3_desc(const b200rt_flat_hw3* f);
extern "C" void b200rt_flat_hw3_free(b200rt_flat_hw3* f);
static std::string g_err;
extern "C" void b200rt_internal_set_error(const char* msg) { g_err = msg ? msg : ""; }   // lives in the CUDA TU of the library

int main(int argc, char** argv) {
  if (argc < 3) {
    std::fprintf(stderr, "usage: %s scene.xml out.b200rt\n", argv[0]);
    return 1;
  }
  Scene scene(argv[1]);
  b200rt_flat_hw3* flat = b200rt_flatten_hw3(scene);
  const int rc = b200rt_scene_desc_save(b200rt_flat_hw3_desc(flat), argv[2]);
  if (rc != B200RT_OK) std::fprintf(stderr, "save failed: %s\n", g_err.c_str());
  b200rt_flat_hw3_free(flat);
  return rc == B200RT_OK ? 0 : 2;
}
