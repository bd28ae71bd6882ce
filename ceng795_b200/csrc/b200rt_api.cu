// temporary stub
#include "../../include/b200rt.h"
extern "C" {
int b200rt_device_count(void) { int n = 0; if (cudaGetDeviceCount(&n) != cudaSuccess) return 0; return n; }
int b200rt_scene_create(const B200SceneDesc*, int, b200rt_scene**) { return B200RT_E_UNSUPPORTED; }
void b200rt_scene_destroy(b200rt_scene*) {}
int b200rt_render(b200rt_scene*, int, const B200RenderOpts*, float*, B200RenderStats*) { return B200RT_E_UNSUPPORTED; }
int b200rt_render_device(b200rt_scene*, int, const B200RenderOpts*, float**, B200RenderStats*) { return B200RT_E_UNSUPPORTED; }
int b200rt_primary_ids(b200rt_scene*, int, const B200RenderOpts*, int32_t*, float*) { return B200RT_E_UNSUPPORTED; }
int b200rt_trace_rays(b200rt_scene*, const float*, int64_t, int32_t, int32_t*, float*, B200RenderStats*) { return B200RT_E_UNSUPPORTED; }
const char* b200rt_version(void) { return "b200rt 0.1"; }
}
