// b200rt_api.cu — device side of the C ABI (include/b200rt.h): scene upload,
// the wavefront driver, and the parity / measurement hooks.
#include <chrono>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include <functional>

#include "fast_bvh.h"
#include "api_guard.h"
#include "wide_bvh.h"
#include "lbvh.cuh"
#include "ppm_kernels.cuh"
#include "rt_hw1_fast.cuh"
#include "tonemap_kernels.cuh"
#include "multi_gpu.cuh"

extern "C" void b200rt_internal_set_error(const char* msg);

namespace {

#define CK(call)                                                                      \
  do {                                                                                \
    cudaError_t e_ = (call);                                                          \
    if (e_ != cudaSuccess) {                                                          \
      char buf_[512];                                                                 \
      snprintf(buf_, sizeof(buf_), "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
      b200rt_internal_set_error(buf_);                                                \
      return B200RT_E_CUDA;                                                           \
    }                                                                                 \
  } while (0)

template <typename T>
int upload(const T* src, size_t n, T** dst) {
  *dst = nullptr;
  if (n == 0) n = 1;  // keep pointers valid
  CK(cudaMalloc((void**)dst, n * sizeof(T)));
  if (src) CK(cudaMemcpy(*dst, src, n * sizeof(T), cudaMemcpyHostToDevice));
  return B200RT_OK;
}

// Scratch device memory of one entry point: released on every return path.
struct DevBuf {
  void* p = nullptr;
  ~DevBuf() { if (p) cudaFree(p); }
  cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, bytes ? bytes : 1); }
  template <typename T> T* as() const { return reinterpret_cast<T*>(p); }
};

constexpr int kMaxWaves = 66;
enum Stage { ST_GENERATE = 0, ST_EXTEND, ST_SHADE, ST_SHADOW, ST_COUNT };

}  // namespace

// One wavefront context: the vertex pool, the shadow queue, the wave table, the counters and the stream they are
// driven on.  A frame is cut into row bands, one per context, and the contexts' launch trains run concurrently: the
// persistent traversal kernels of one band fill the SMs that another band's tail wave (a few long rays on a grid sized
// for the whole machine) leaves idle.  Context 0 also serves the single-stream entry points (PPM, parity hooks).
struct WaveCtx {
  Pool P;
  ShadowQueue Q;
  WaveInfo* d_waves;
  Counters* d_counters;
  cudaStream_t stream;
  cudaEvent_t ev[2];
  std::vector<cudaEvent_t> ev_pool;   // stage marks (Timers)
};
constexpr int kMaxCtx = 8;
constexpr int kDefaultCtx = 1;   // swept on the device (profiles/r02*_tuning.txt)

struct b200rt_scene {
  int device;
  DScene S;                    // device pointers (camera filled per render)
  std::vector<b200rt_camera> cameras;
  bool has_transparency, has_mirror;
  int sm_count;
  int n_meshes;
  int n_nodes, n_triprims, n_light_cdf, n_materials;   // what was uploaded (b200rt_scene_refit checks against it)
  int wide_nodes, wide_depth;   // the 8-wide tree (0 nodes: the scene walks the binary tree)
  int wide_blocks;              // resident CTAs per SM of k_trace_wide (4 / 6 / 8)
  int n_ctx;                    // contexts a frame is cut into (B200RT_CONTEXTS, default kDefaultCtx)
  // device allocations owned by the scene
  std::vector<void*> allocs;
  // wavefront state
  WaveCtx c0;
  std::vector<WaveCtx*> more;   // contexts 1.. (created on first use)
  float4* d_accum;
  size_t accum_pixels;
  // pipelined host copies (b200rt_render_async): two staging frames, a copy stream, events
  float4* d_stage[2];
  size_t stage_pixels;
  cudaStream_t copy_stream;
  cudaEvent_t ev_staged[2], ev_copied[2];
  bool copy_pending[2];
  int stage_next;
  int* d_lit;            // parity hook: lit shadow tests per pixel x light (b200rt_debug_lit_counts)
  size_t lit_entries;
  bool lit_on;
  size_t scene_bytes;
  // PPM state
  HitPoints HP;
  PpmGrid G;
  PpmCounters* d_pc;
  float* d_rtab;
  float4* d_export;
  void* d_scan_tmp;
  uint32_t* d_sort[4];   // PPM: direction keys / photon order (in, out)
  size_t scan_tmp_bytes, items_capacity;
  uint32_t n_hp;
  int ppm_camera;
  // the scan grid of the photon pass, rebuilt as the radii shrink (ppm_kernels.cuh: PpmGrid::adaptive)
  PpmGrid G2;
  size_t entries2_capacity;
  double g2_sumR;      // sum of R(n) over the hit points when G2 was built
  bool g2_valid;
  double* d_sumR;
  // multi-GPU (multi_gpu.cuh).  Transient, set around a render: where wave 0's combine writes, and what it waits for.
  float4* rt_frame;                 // nullptr = d_accum
  const unsigned int* rt_wait_flag; // awaited (>= rt_wait_target) before the first pixel of a frame is written
  unsigned int rt_wait_target;
  unsigned int* d_sync_err;         // raised by a wait kernel that timed out
  // the frame this scene OWNS (exported): counters of b200rt_frame_begin / _collect
  unsigned int own_frame_no, own_parts_seen;
  int own_camera;
  bool own_direct;
  // the frame this scene CONTRIBUTES to (attached)
  float4* tgt_frame;
  FrameTail* tgt_tail;
  size_t tgt_pixels;
  int tgt_w, tgt_h;
  bool tgt_direct, tgt_ipc, tgt_self;
  unsigned int tgt_frame_no;        // render_shared calls since the attach
};

namespace {

template <typename T>
int pool_alloc(T** p, size_t n) {
  CK(cudaMalloc((void**)p, n * sizeof(T)));
  return B200RT_OK;
}

void pool_free(WaveCtx& cx) {
  Pool& P = cx.P;
  void* ptrs[] = {P.ray_o, P.ray_d, P.key, P.pixel, P.sample, P.hit_a, P.hit_b, P.nrm, P.rec0, P.rec1, P.rec2,
                  P.rec3, P.final_c, P.kd, P.rec4, P.omask, cx.Q.o, cx.Q.d, cx.Q.time, cx.Q.vis, cx.Q.lv, cx.Q.omask};
  for (void* p : ptrs)
    if (p) cudaFree(p);
  memset(&cx.P, 0, sizeof(Pool));
  memset(&cx.Q, 0, sizeof(ShadowQueue));
}

int ensure_pool(b200rt_scene* sc, WaveCtx& cx, size_t slots, size_t shadow_slots) {
  if (cx.P.capacity >= slots && cx.Q.capacity >= shadow_slots) return B200RT_OK;
  pool_free(cx);
  Pool& P = cx.P;
  int rc;
#define PA(field, n) if ((rc = pool_alloc(&field, (n))) != B200RT_OK) return rc
  PA(P.ray_o, slots); PA(P.ray_d, slots); PA(P.key, slots); PA(P.pixel, slots); PA(P.sample, slots);
  PA(P.hit_a, slots); PA(P.hit_b, slots); PA(P.nrm, slots); PA(P.rec0, slots); PA(P.rec1, slots);
  PA(P.rec2, slots); PA(P.rec3, slots); PA(P.final_c, slots);
  if (sc->S.semantics >= B200RT_SEM_HW4) { PA(P.kd, slots); PA(P.rec4, slots); }
  if (sc->S.use_wide) { PA(P.omask, slots); PA(cx.Q.omask, shadow_slots); }
  P.capacity = (uint32_t)slots;
  ShadowQueue& Q = cx.Q;
  PA(Q.o, shadow_slots); PA(Q.d, shadow_slots); PA(Q.time, shadow_slots);
  PA(Q.vis, slots * (size_t)(sc->S.n_lights > 0 ? sc->S.n_lights : 1));
  if (sc->S.semantics >= B200RT_SEM_HW5 && sc->S.semantics <= B200RT_SEM_HW7) PA(Q.lv, slots * (size_t)(sc->S.n_lights > 0 ? sc->S.n_lights : 1));
#undef PA
  Q.capacity = (uint32_t)shadow_slots;
  return B200RT_OK;
}

int ctx_init(WaveCtx& cx) {
  memset(&cx.P, 0, sizeof(Pool));
  memset(&cx.Q, 0, sizeof(ShadowQueue));
  cx.d_waves = nullptr; cx.d_counters = nullptr; cx.stream = nullptr; cx.ev[0] = nullptr; cx.ev[1] = nullptr;
  CK(cudaMalloc((void**)&cx.d_waves, sizeof(WaveInfo) * (kMaxWaves + 2)));
  CK(cudaMalloc((void**)&cx.d_counters, sizeof(Counters)));
  CK(cudaMemset(cx.d_waves, 0, sizeof(WaveInfo) * (kMaxWaves + 2)));
  CK(cudaStreamCreateWithFlags(&cx.stream, cudaStreamNonBlocking));
  CK(cudaEventCreate(&cx.ev[0]));
  CK(cudaEventCreate(&cx.ev[1]));
  return B200RT_OK;
}

void ctx_destroy(WaveCtx& cx) {
  pool_free(cx);
  if (cx.d_waves) cudaFree(cx.d_waves);
  if (cx.d_counters) cudaFree(cx.d_counters);
  if (cx.stream) cudaStreamDestroy(cx.stream);
  if (cx.ev[0]) cudaEventDestroy(cx.ev[0]);
  if (cx.ev[1]) cudaEventDestroy(cx.ev[1]);
  for (cudaEvent_t e : cx.ev_pool) cudaEventDestroy(e);
  cx.ev_pool.clear();
  cx.d_waves = nullptr; cx.d_counters = nullptr; cx.stream = nullptr; cx.ev[0] = nullptr; cx.ev[1] = nullptr;
}

// Context k of the scene (0 = the scene's own), created on first use.
int get_ctx(b200rt_scene* sc, int k, WaveCtx** out) {
  if (k == 0) { *out = &sc->c0; return B200RT_OK; }
  while ((int)sc->more.size() < k) {
    WaveCtx* cx = new WaveCtx();
    const int rc = ctx_init(*cx);
    if (rc) { ctx_destroy(*cx); delete cx; return rc; }
    sc->more.push_back(cx);
  }
  *out = sc->more[(size_t)k - 1];
  return B200RT_OK;
}

int ensure_accum(b200rt_scene* sc, size_t pixels) {
  if (sc->accum_pixels >= pixels) return B200RT_OK;
  if (sc->d_accum) cudaFree(sc->d_accum);
  sc->d_accum = nullptr;
  CK(cudaMalloc((void**)&sc->d_accum, pixels * sizeof(float4) + sizeof(FrameTail)));   // FrameTail: multi_gpu.cuh
  CK(cudaMemset(sc->d_accum + pixels, 0, sizeof(FrameTail)));
  sc->accum_pixels = pixels;
  return B200RT_OK;
}

// Stage marks of a frame.  The events come from a pool the scene owns and grows on first use, so a steady-state frame
// creates and destroys nothing inside its timed region.
struct Timers {
  std::vector<cudaEvent_t>* pool;
  std::vector<int> stage;
  size_t used;
  cudaStream_t s;
  bool on;
  void mark(int st) {
    if (!on) return;
    if (used == pool->size()) {
      cudaEvent_t e;
      if (cudaEventCreate(&e) != cudaSuccess) { on = false; return; }
      pool->push_back(e);
    }
    cudaEventRecord((*pool)[used++], s);
    stage.push_back(st);
  }
  void reset() { used = 0; stage.clear(); }
  void collect(float* ms) {
    for (int i = 0; i < ST_COUNT; i++) ms[i] = 0.f;
    for (size_t i = 1; i < used; i++) {
      float t = 0.f;
      cudaEventElapsedTime(&t, (*pool)[i - 1], (*pool)[i]);
      ms[stage[i]] += t;
    }
    reset();
  }
};

// k_trace_wide<ANY, TRI_MODE, MINB> on the context's stream, one persistent CTA set per SM
void launch_wide(b200rt_scene* sc, WaveCtx& cx, bool any, bool mode7, int d) {
  const int mb = sc->wide_blocks;
  const int grid = sc->sm_count * mb;
  cudaStream_t st = cx.stream;
  if (any) k_top_mask<true><<<sc->sm_count * 8, 256, 0, st>>>(sc->S, cx.P, cx.Q, cx.d_waves, d, cx.d_counters);
  else k_top_mask<false><<<sc->sm_count * 8, 256, 0, st>>>(sc->S, cx.P, cx.Q, cx.d_waves, d, cx.d_counters);
#define LW(A, M, B) k_trace_wide<A, M, B><<<grid, RT_BLOCK, 0, st>>>(sc->S, cx.P, cx.Q, cx.d_waves, d, cx.d_counters)
#define LWB(A, M) do { if (mb == 4) LW(A, M, 4); else if (mb == 8) LW(A, M, 8); else LW(A, M, 6); } while (0)
  if (any) { if (mode7) LWB(true, 7); else LWB(true, 2); }
  else { if (mode7) LWB(false, 7); else LWB(false, 2); }
#undef LWB
#undef LW
}

// Which tree a wave loop walks.  The 8-wide kernel wins where the rays of a warp part company (1-sample frames with
// their shadow and mirror waves, photon waves, caller-supplied rays: cfg2 3.54 against 3.99 ms); on a multi-sample
// camera the strata of a pixel travel together through whole-warp batches and the binary walk's cheaper steps win
// (cfg3 189 against 211 ms); path tracing is a draw (cfg4 858 / 860 ms).  B200RT_WIDE_MULTISAMPLE=1 forces the wide tree.
bool wide_for(const b200rt_scene* sc, bool exact, bool frame) {
  if (!sc->S.use_wide || exact || sc->S.semantics == B200RT_SEM_HW1) return false;
  if (frame && sc->S.cam.samples_sqrt > 1 && !sc->S.integrator) {
    const char* e = getenv("B200RT_WIDE_MULTISAMPLE");
    return e && atoi(e) != 0;
  }
  return true;
}

Timers make_timers(WaveCtx& cx, bool on) {
  Timers tm;
  tm.pool = &cx.ev_pool; tm.used = 0; tm.s = cx.stream; tm.on = on;
  return tm;
}

// Runs the wave loop of context cx for paths already placed in its wave 0 by k_generate / k_import_rays.
// primary_only: stop after the first extend (parity hooks).
int run_waves(b200rt_scene* sc, WaveCtx& cx, bool exact, bool primary_only, bool with_shading, Timers& tm, int* launches,
              int* waves_run) {
  const DScene& S = sc->S;
  cudaStream_t st = cx.stream;
  const int grid_p = sc->sm_count * (RT_FAST_MIN_BLOCKS > 8 ? RT_FAST_MIN_BLOCKS : 8);  // persistent traversal grid: resident CTAs of 128 threads per SM
  const int grid_s = sc->sm_count * 8;
  const bool hw1 = S.semantics == B200RT_SEM_HW1;
  const bool wide = wide_for(sc, exact, with_shading);
  const bool mode7 = S.semantics >= B200RT_SEM_HW4;
  const int max_d = primary_only ? 0 : S.max_depth;
  int d_last = 0;
  // wave bookkeeping: one <<<1,1>>> launch per depth (k_finish_wave also opens the next wave); the shadow-queue clamp
  // is done by the traversal kernel itself on the wide path
  k_next_wave<<<1, 1, 0, st>>>(cx.d_waves, 0, cx.d_counters);
  (*launches)++;
  for (int d = 0; d <= max_d; d++) {
    const bool hw1_fast = hw1 && S.hw1_fast_root != B200RT_REF_NONE && !exact;
    if (hw1_fast) {
      k_trace_hw1_fast<false><<<grid_s, RT_BLOCK, 0, st>>>(S, cx.P, cx.Q, cx.d_waves, d, cx.d_counters);
    } else if (hw1) {
      k_extend_hw1<<<grid_p, RT_BLOCK, 0, st>>>(S, cx.P, cx.d_waves, d, cx.d_counters);
    } else if (wide) {
      launch_wide(sc, cx, false, mode7, d);
      (*launches)++;   // k_top_mask (the traversal kernel itself is counted below)
    } else if (mode7) {
      if (exact) k_extend<true, 7><<<grid_p, RT_BLOCK, 0, st>>>(S, cx.P, cx.d_waves, d, cx.d_counters);
      else k_trace_fast<false, 7><<<grid_p, RT_BLOCK, 0, st>>>(S, cx.P, cx.Q, cx.d_waves, d, cx.d_counters);
    } else {
      if (exact) k_extend<true, 2><<<grid_p, RT_BLOCK, 0, st>>>(S, cx.P, cx.d_waves, d, cx.d_counters);
      else k_trace_fast<false, 2><<<grid_p, RT_BLOCK, 0, st>>>(S, cx.P, cx.Q, cx.d_waves, d, cx.d_counters);
    }
    (*launches)++;
    tm.mark(ST_EXTEND);
    d_last = d;
    if (!with_shading) break;
    const bool hw7 = S.semantics >= B200RT_SEM_HW5;  // HW5/HW6 shade like HW7's ray-tracing mode (flags inside)
    const bool hw4 = S.semantics == B200RT_SEM_HW4;
    if (hw7) k_shade7_a<<<grid_s * RT_BLOCK / RT_SHADE7_BLOCK, RT_SHADE7_BLOCK, 0, st>>>(S, cx.P, cx.Q, cx.d_waves, d, cx.d_counters);
    else if (hw4) k_shade4_a<<<grid_s, RT_BLOCK, 0, st>>>(S, cx.P, cx.Q, cx.d_waves, d, cx.d_counters);
    else k_shade_a<<<grid_s, RT_BLOCK, 0, st>>>(S, cx.P, cx.Q, cx.d_waves, d, cx.d_counters);
    (*launches)++;
    if (!wide) { k_clamp_shadow_count<<<1, 1, 0, st>>>(cx.Q, cx.d_counters); (*launches)++; }
    tm.mark(ST_SHADE);
    if (hw1_fast) {
      k_trace_hw1_fast<true><<<grid_s, RT_BLOCK, 0, st>>>(S, cx.P, cx.Q, cx.d_waves, d, cx.d_counters);
    } else if (hw1) {
      k_shadow_hw1<<<grid_p, RT_BLOCK, 0, st>>>(S, cx.Q, cx.d_counters);
    } else if (wide) {
      launch_wide(sc, cx, true, mode7, d);
      (*launches)++;   // k_top_mask
    } else if (mode7) {
      if (exact) k_shadow<true, 7><<<grid_p, RT_BLOCK, 0, st>>>(S, cx.Q, cx.d_counters);
      else k_trace_fast<true, 7><<<grid_p, RT_BLOCK, 0, st>>>(S, cx.P, cx.Q, cx.d_waves, d, cx.d_counters);
    } else {
      if (exact) k_shadow<true, 2><<<grid_p, RT_BLOCK, 0, st>>>(S, cx.Q, cx.d_counters);
      else k_trace_fast<true, 2><<<grid_p, RT_BLOCK, 0, st>>>(S, cx.P, cx.Q, cx.d_waves, d, cx.d_counters);
    }
    (*launches)++;
    tm.mark(ST_SHADOW);
    if (hw7) k_shade7_b<<<grid_s * RT_BLOCK / RT_SHADE7_BLOCK, RT_SHADE7_BLOCK, 0, st>>>(S, cx.P, cx.Q, cx.d_waves, d, cx.d_counters);
    else if (hw4) k_shade4_b<<<grid_s, RT_BLOCK, 0, st>>>(S, cx.P, cx.Q, cx.d_waves, d, cx.d_counters);
    else k_shade_b<<<grid_s, RT_BLOCK, 0, st>>>(S, cx.P, cx.Q, cx.d_waves, d, cx.d_counters);
    k_finish_wave<<<1, 1, 0, st>>>(cx.d_waves, d, cx.P, cx.d_counters);
    (*launches) += 2;
    tm.mark(ST_SHADE);
  }
  if (with_shading) {
    float4* frame = sc->rt_frame ? sc->rt_frame : sc->d_accum;
    for (int d = d_last; d >= 0; d--) {
      if (d == 0 && sc->rt_wait_flag) {   // a shared frame: its owner must have opened it (multi_gpu.cuh)
        k_frame_wait<<<1, 1, 0, st>>>(sc->rt_wait_flag, sc->rt_wait_target, sc->d_sync_err, 20000000000ull);
        (*launches)++;
      }
      if (S.semantics >= B200RT_SEM_HW5) k_combine7<<<grid_s, 256, 0, st>>>(S, cx.P, cx.d_waves, d, frame);
      else k_combine<<<grid_s, 256, 0, st>>>(S, cx.P, cx.d_waves, d, frame);
      (*launches)++;
    }
    tm.mark(ST_SHADE);
  }
  *waves_run = d_last + 1;
  CK(cudaGetLastError());
  return B200RT_OK;
}

int resolve_rows(const b200rt_camera& c, const B200RenderOpts* o, RenderParams* R) {
  int rb = 0, re = c.height, rs = 1, rg = 1;
  if (o) {
    if (o->row_end > 0) { rb = o->row_begin; re = o->row_end; }
    if (o->row_group > 1) rg = o->row_group;
    if (o->row_stride > 1) rs = o->row_stride;
  }
  if (rs < rg) rs = rg;   // groups do not overlap
  if (rb < 0 || re > c.height || rb > re) return B200RT_E_INVALID;
  R->row_begin = rb; R->row_end = re; R->row_stride = rs; R->row_group = rg; R->row_index0 = 0;
  int rows = 0;
  for (int start = rb; start < re; start += rs) rows += (re - start < rg) ? (re - start) : rg;
  R->n_rows = rows;
  const int total = c.samples_sqrt * c.samples_sqrt;
  int sb = 0, se = total;
  if (o && o->sample_end > 0) { sb = o->sample_begin; se = o->sample_end; }
  if (sb < 0 || se > total || sb > se) return B200RT_E_INVALID;
  R->sample_begin = sb; R->sample_count = se - sb;
  const uint64_t seed = o ? o->seed : 0;
  R->seed_key = rng_mix((uint32_t)seed, (uint32_t)(seed >> 32));
  return B200RT_OK;
}

int render_impl(b200rt_scene* sc, int camera_index, const B200RenderOpts* opts, B200RenderStats* stats) {
  if (!sc || camera_index < 0 || camera_index >= (int)sc->cameras.size()) {
    b200rt_internal_set_error("b200rt_render: bad scene or camera index");
    return B200RT_E_INVALID;
  }
  CK(cudaSetDevice(sc->device));
  sc->S.cam = sc->cameras[camera_index];
  const b200rt_camera& c = sc->S.cam;
  RenderParams R;
  if (resolve_rows(c, opts, &R) != B200RT_OK) {
    b200rt_internal_set_error("b200rt_render: bad row / sample range");
    return B200RT_E_INVALID;
  }
  if (sc->S.max_depth + 2 > kMaxWaves) {
    b200rt_internal_set_error("MaxRecursionDepth too large");
    return B200RT_E_UNSUPPORTED;
  }
  const bool exact = opts && opts->exact_traversal;
  const size_t pixels = (size_t)c.width * c.height;
  int rc = ensure_accum(sc, pixels);
  if (rc) return rc;
  sc->S.lit_count = nullptr;
  if (sc->lit_on) {
    const size_t need = pixels * (size_t)(sc->S.n_lights > 0 ? sc->S.n_lights : 1);
    if (sc->lit_entries < need) {
      if (sc->d_lit) cudaFree(sc->d_lit);
      sc->d_lit = nullptr; sc->lit_entries = 0;
      CK(cudaMalloc((void**)&sc->d_lit, need * sizeof(int)));
      sc->lit_entries = need;
    }
    sc->S.lit_count = sc->d_lit;
  }

  // The rows of the frame are cut into K contiguous bands (multiples of 4 rows: k_generate's 8x4 tiles), one per
  // wavefront context; each band runs its own generate -> (extend -> shade_a -> shadow -> shade_b)* -> combine train on
  // its own stream.  Results do not depend on K: every sample's random stream is keyed by its pixel.
  int K = (opts && opts->contexts > 0) ? opts->contexts : sc->n_ctx;
  if (K > kMaxCtx) K = kMaxCtx;
  while (K > 1 && R.n_rows < 16 * K) K--;
  WaveCtx* cxs[kMaxCtx];
  for (int k = 0; k < K; k++) {
    rc = get_ctx(sc, k, &cxs[k]);
    if (rc) return rc;
  }
  int band_begin[kMaxCtx + 1];
  for (int k = 0; k <= K; k++) {
    int b = (int)((long long)R.n_rows * k / K);
    if (k > 0 && k < K) b &= ~3;
    band_begin[k] = b;
  }

  // batches of pixel rows so that the vertex pool stays bounded.  The recursion is a TREE
  // (mirror + two dielectric children + the path tracer's hemisphere sample), so the pool is
  // sized from an estimate and, if a batch overflows it, the frame is redone with half the rows
  // per batch (same pool: twice the slots per camera path).
  const size_t paths_per_row = (size_t)c.width * R.sample_count;
  const size_t fan = (sc->has_mirror ? 1 : 0) + (sc->has_transparency ? 2 : 0) + (sc->S.integrator ? 1 : 0);
  const size_t growth = 1 + (size_t)sc->S.max_depth * (fan > 1 ? fan : 1);
  const size_t max_paths = ((size_t)16 << 20) / (size_t)K;  // 16 Mi camera paths per batch over all contexts
  size_t rows_per_batch = max_paths / (paths_per_row ? paths_per_row : 1);
  if (rows_per_batch < 1) rows_per_batch = 1;
  if (rows_per_batch >= 4) rows_per_batch &= ~(size_t)3;  // keeps k_generate's 8x4 tiling available
  int max_band = 0;
  for (int k = 0; k < K; k++) max_band = std::max(max_band, band_begin[k + 1] - band_begin[k]);
  if (rows_per_batch > (size_t)max_band) rows_per_batch = (size_t)max_band;
  const size_t nl = (size_t)(sc->S.n_lights > 0 ? sc->S.n_lights : 1);
  {
    const size_t batch_paths = rows_per_batch * paths_per_row;
    size_t slots = batch_paths * growth + 1024;
    const size_t cap = 0xF0000000ull / nl;
    if (slots > cap) slots = cap;
    if (slots < batch_paths + 1024) {
      b200rt_internal_set_error("vertex pool too large for 32-bit slots");
      return B200RT_E_UNSUPPORTED;
    }
    const size_t per_wave = batch_paths * (fan > 1 ? 2 * fan : 1);
    for (int k = 0; k < K; k++) {
      rc = ensure_pool(sc, *cxs[k], slots, per_wave * nl + 1024);
      if (rc) return rc;
    }
  }

  WaveCtx& c0 = sc->c0;
  Timers tms[kMaxCtx];
  for (int k = 0; k < K; k++) tms[k] = make_timers(*cxs[k], stats != nullptr);
  int launches = 0, waves_total = 0;
  Counters hc;
  for (int attempt = 0;; attempt++) {
    if (!sc->rt_frame) CK(cudaMemsetAsync(sc->d_accum, 0, pixels * sizeof(float4), c0.stream));   // a shared frame is its owner's to clear
    if (sc->S.lit_count) CK(cudaMemsetAsync(sc->d_lit, 0, sc->lit_entries * sizeof(int), c0.stream));
    launches = 0; waves_total = 0;
    CK(cudaEventRecord(c0.ev[0], c0.stream));   // the fork: every context starts after the frame buffer is cleared
    for (int k = 0; k < K; k++) {
      WaveCtx& cx = *cxs[k];
      if (k > 0) CK(cudaStreamWaitEvent(cx.stream, c0.ev[0], 0));
      CK(cudaMemsetAsync(cx.d_counters, 0, sizeof(Counters), cx.stream));
      tms[k].reset();
      tms[k].mark(ST_GENERATE);
    }
    // batch b of every context is issued before batch b+1 of any: the streams fill evenly from one host thread
    int r0s[kMaxCtx];
    for (int k = 0; k < K; k++) r0s[k] = band_begin[k];
    for (bool any = true; any;) {
      any = false;
      for (int k = 0; k < K; k++) {
        if (r0s[k] >= band_begin[k + 1]) continue;
        any = true;
        WaveCtx& cx = *cxs[k];
        RenderParams B = R;
        const int left = band_begin[k + 1] - r0s[k];
        const int nr = (size_t)left < rows_per_batch ? left : (int)rows_per_batch;
        B.row_index0 = r0s[k];
        B.n_rows = nr;
        B.n_paths = (uint32_t)((size_t)nr * paths_per_row);
        k_generate<<<sc->sm_count * 8, 256, 0, cx.stream>>>(sc->S, cx.P, B, cx.d_waves, cx.d_counters);
        launches++;
        tms[k].mark(ST_GENERATE);
        int wr = 0;
        rc = run_waves(sc, cx, exact, false, true, tms[k], &launches, &wr);
        if (rc) return rc;
        waves_total += wr;
        r0s[k] += nr;
      }
    }
    for (int k = 1; k < K; k++) {   // the join: context 0 waits for the others
      CK(cudaEventRecord(cxs[k]->ev[1], cxs[k]->stream));
      CK(cudaStreamWaitEvent(c0.stream, cxs[k]->ev[1], 0));
    }
    CK(cudaEventRecord(c0.ev[1], c0.stream));
    CK(cudaStreamSynchronize(c0.stream));
    memset(&hc, 0, sizeof(hc));
    for (int k = 0; k < K; k++) {
      Counters one;
      CK(cudaMemcpy(&one, cxs[k]->d_counters, sizeof(one), cudaMemcpyDeviceToHost));
      hc.rays_primary += one.rays_primary; hc.rays_shadow += one.rays_shadow; hc.rays_secondary += one.rays_secondary;
      hc.nodes_visited += one.nodes_visited; hc.tris_tested += one.tris_tested; hc.spheres_tested += one.spheres_tested;
      hc.instances_entered += one.instances_entered;
      hc.sh_nodes_visited += one.sh_nodes_visited; hc.sh_tris_tested += one.sh_tris_tested;
      hc.sh_spheres_tested += one.sh_spheres_tested; hc.sh_instances_entered += one.sh_instances_entered;
      if (one.overflow > hc.overflow) hc.overflow = one.overflow;
    }
    if (!hc.overflow) break;
    if (hc.overflow == 2) {
      b200rt_internal_set_error("traversal stack overflow (BVH deeper than the traversal stack)");
      return B200RT_E_UNSUPPORTED;
    }
    if (rows_per_batch <= 1 || attempt >= 8) {
      b200rt_internal_set_error("vertex pool / shadow queue overflow (recursion tree larger than the pool)");
      return B200RT_E_UNSUPPORTED;
    }
    rows_per_batch = (rows_per_batch + 1) / 2;
    if (rows_per_batch >= 4) rows_per_batch &= ~(size_t)3;
  }
  if (stats) {
    memset(stats, 0, sizeof(*stats));
    stats->rays_primary = hc.rays_primary; stats->rays_shadow = hc.rays_shadow;
    stats->rays_secondary = hc.rays_secondary; stats->nodes_visited = hc.nodes_visited;
    stats->tris_tested = hc.tris_tested; stats->spheres_tested = hc.spheres_tested;
    stats->instances_entered = hc.instances_entered;
    stats->sh_nodes_visited = hc.sh_nodes_visited; stats->sh_tris_tested = hc.sh_tris_tested;
    stats->sh_spheres_tested = hc.sh_spheres_tested; stats->sh_instances_entered = hc.sh_instances_entered;
    // stage times: event intervals on each context's own stream, summed over the contexts.  With more than one
    // context the intervals of different streams overlap in time (that is the point), so their sum exceeds ms_total;
    // a one-context render (opts->contexts = 1) gives the serial per-kernel times the roofline is quoted on.
    for (int k = 0; k < K; k++) {
      float ms[ST_COUNT];
      tms[k].collect(ms);
      stats->ms_generate += ms[ST_GENERATE]; stats->ms_extend += ms[ST_EXTEND];
      stats->ms_shade += ms[ST_SHADE]; stats->ms_shadow += ms[ST_SHADOW];
    }
    cudaEventElapsedTime(&stats->ms_total, c0.ev[0], c0.ev[1]);
    stats->waves = waves_total; stats->launches = launches;
    stats->node_bytes = wide_for(sc, exact, true) ? (int)sizeof(b200rt_wide::Node) : 64;
    stats->top_objects = exact ? 0 : sc->S.n_top;
    stats->contexts = K;
  }
  return B200RT_OK;
}

void ppm_free(b200rt_scene* sc) {
  HitPoints& H = sc->HP;
  void* ptrs[] = {H.count, H.pos, H.nrm, H.wo, H.att, H.n, H.n_base, H.delta, H.state, H.dep, sc->G.cell_count, sc->G.cell_start, sc->G.items,
                  sc->G.cell_pos,
                  sc->d_pc, sc->d_rtab, sc->d_export, sc->d_scan_tmp, sc->d_sort[0], sc->d_sort[1], sc->d_sort[2],
                  sc->d_sort[3]};
  for (void* p : ptrs)
    if (p) cudaFree(p);
  void* g2[] = {sc->G2.cell_count, sc->G2.cell_start, sc->G2.cell_pos, sc->d_sumR};
  for (void* p : g2)
    if (p) cudaFree(p);
  memset(&sc->G2, 0, sizeof(PpmGrid));
  sc->entries2_capacity = 0; sc->g2_valid = false; sc->d_sumR = nullptr;
  {
    const uint32_t part = sc->HP.part, parts = sc->HP.parts;   // the share of a multi-GPU run outlives the buffers
    memset(&sc->HP, 0, sizeof(HitPoints));
    sc->HP.part = part; sc->HP.parts = parts;
  }
  memset(&sc->G, 0, sizeof(PpmGrid));
  sc->d_pc = nullptr; sc->d_rtab = nullptr; sc->d_export = nullptr; sc->d_scan_tmp = nullptr;
  for (int q = 0; q < 4; q++) sc->d_sort[q] = nullptr;
  sc->scan_tmp_bytes = 0; sc->items_capacity = 0; sc->n_hp = 0;
}

int ppm_ensure_hit_points(b200rt_scene* sc, size_t cap) {
  HitPoints& H = sc->HP;
  if (H.capacity >= cap) return B200RT_OK;
  void* ptrs[] = {H.pos, H.nrm, H.wo, H.att, H.n, H.n_base, H.delta, H.state, H.dep, sc->d_export};
  for (void* p : ptrs)
    if (p) cudaFree(p);
  H.pos = H.nrm = H.wo = H.att = H.state = H.dep = nullptr; H.n = H.n_base = H.delta = nullptr;
  sc->d_export = nullptr; H.capacity = 0;
  if (!H.count) CK(cudaMalloc((void**)&H.count, sizeof(unsigned int)));
  CK(cudaMalloc((void**)&H.pos, cap * sizeof(float4)));
  CK(cudaMalloc((void**)&H.nrm, cap * sizeof(float4)));
  CK(cudaMalloc((void**)&H.wo, cap * sizeof(float4)));
  CK(cudaMalloc((void**)&H.att, cap * sizeof(float4)));
  CK(cudaMalloc((void**)&H.n, cap * sizeof(unsigned int)));
  CK(cudaMalloc((void**)&H.n_base, cap * sizeof(unsigned int)));
  CK(cudaMalloc((void**)&H.delta, cap * sizeof(unsigned int)));
  CK(cudaMalloc((void**)&H.state, cap * sizeof(float4)));
  CK(cudaMalloc((void**)&H.dep, cap * 4 * sizeof(float4)));
  CK(cudaMalloc((void**)&sc->d_export, cap * sizeof(float4)));
  H.capacity = (uint32_t)cap;
  return B200RT_OK;
}

// R(n) table in the reference's own arithmetic (PPM/src/Scene.cpp:138-141): float product of
// g = (float)((n*a + a) / (n*a + 1.0))
int ppm_ensure_rtab(b200rt_scene* sc) {
  if (sc->d_rtab) return B200RT_OK;
  std::vector<float> r((size_t)PPM_RTAB);
  float acc = 1.0f;
  for (unsigned int n = 0; n < PPM_RTAB; n++) {
    r[n] = acc;
    const float g = (float)(((double)((float)n * PPM_ALPHA + PPM_ALPHA)) / ((double)((float)n * PPM_ALPHA) + 1.0));
    acc = acc * g;
  }
  CK(cudaMalloc((void**)&sc->d_rtab, sizeof(float) * PPM_RTAB));
  CK(cudaMemcpy(sc->d_rtab, r.data(), sizeof(float) * PPM_RTAB, cudaMemcpyHostToDevice));
  return B200RT_OK;
}

// Queues "frame in d_accum -> host memory" without waiting for it: a device-to-device copy into one of two staging
// frames on the scene's stream (tens of microseconds), then the host copy on a stream of its own, so that the next
// frame can be rendered into d_accum while this one crosses PCIe.  The staging frame is reused two frames later, after
// its host copy has landed.
int queue_host_copy(b200rt_scene* sc, float* out, size_t pixels) {
  if (!sc->copy_stream) {
    CK(cudaStreamCreateWithFlags(&sc->copy_stream, cudaStreamNonBlocking));
    for (int k = 0; k < 2; k++) {
      CK(cudaEventCreateWithFlags(&sc->ev_staged[k], cudaEventDisableTiming));
      CK(cudaEventCreateWithFlags(&sc->ev_copied[k], cudaEventDisableTiming));
    }
  }
  if (sc->stage_pixels < pixels) {
    CK(cudaStreamSynchronize(sc->copy_stream));
    for (int k = 0; k < 2; k++) {
      if (sc->d_stage[k]) cudaFree(sc->d_stage[k]);
      sc->d_stage[k] = nullptr; sc->copy_pending[k] = false;
    }
    sc->stage_pixels = 0;
    for (int k = 0; k < 2; k++) CK(cudaMalloc((void**)&sc->d_stage[k], pixels * sizeof(float4)));
    sc->stage_pixels = pixels;
  }
  const int k = sc->stage_next;
  sc->stage_next ^= 1;
  if (sc->copy_pending[k]) CK(cudaEventSynchronize(sc->ev_copied[k]));
  CK(cudaMemcpyAsync(sc->d_stage[k], sc->d_accum, pixels * sizeof(float4), cudaMemcpyDeviceToDevice, sc->c0.stream));
  CK(cudaEventRecord(sc->ev_staged[k], sc->c0.stream));
  CK(cudaStreamWaitEvent(sc->copy_stream, sc->ev_staged[k], 0));
  CK(cudaMemcpyAsync(out, sc->d_stage[k], pixels * sizeof(float4), cudaMemcpyDeviceToHost, sc->copy_stream));
  CK(cudaEventRecord(sc->ev_copied[k], sc->copy_stream));
  sc->copy_pending[k] = true;
  return B200RT_OK;
}

// A traversal-stack or pool overflow inside a wave loop must not come back as B200RT_OK with wrong results.
int check_overflow(b200rt_scene* sc, const char* who) {
  unsigned int ov = 0;
  CK(cudaMemcpy(&ov, &sc->c0.d_counters->overflow, sizeof(ov), cudaMemcpyDeviceToHost));
  if (!ov) return B200RT_OK;
  std::string msg = std::string(who) + (ov == 2 ? ": traversal stack overflow (BVH deeper than the traversal stack)"
                                                : ": vertex pool / queue overflow");
  b200rt_internal_set_error(msg.c_str());
  return B200RT_E_UNSUPPORTED;
}

// Rebuilds the scan grid of the photon pass when the radii have shrunk enough since the last build (or when there is
// none yet and they have left r0 behind).  Cells become twice the typical current radius wide; a hit point is entered
// into every cell its current radius reaches.  cub's scan: a library call on a step that runs a few dozen times a frame.
int ppm_rebuild_scan_grid(b200rt_scene* sc, int* launches, int* rebuilds) {
  const uint32_t n_hp = sc->n_hp;
  if (n_hp == 0 || getenv("B200RT_PPM_STATIC_GRID")) return B200RT_OK;
  cudaStream_t st = sc->c0.stream;
  if (!sc->d_sumR) CK(cudaMalloc((void**)&sc->d_sumR, sizeof(double)));
  CK(cudaMemsetAsync(sc->d_sumR, 0, sizeof(double), st));
  k_ppm_sum_R<<<sc->sm_count * 4, 256, 0, st>>>(sc->HP, sc->G, n_hp, sc->d_sumR);
  double sumR = 0.0;
  CK(cudaMemcpyAsync(&sumR, sc->d_sumR, sizeof(double), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  (*launches)++;
  const double meanR = sumR / (double)n_hp;          // mean of r2 / r0^2
  if (sc->g2_valid ? sumR > 0.8 * sc->g2_sumR : meanR > 0.7) return B200RT_OK;
  const bool timing = getenv("B200RT_BUILD_TIMING") != nullptr;
  const auto T0 = std::chrono::steady_clock::now();
  auto since = [&](std::chrono::steady_clock::time_point a) {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - a).count();
  };
  PpmGrid& G2 = sc->G2;
  unsigned int* cc = G2.cell_count;
  unsigned int* cs = G2.cell_start;
  float4* cp = G2.cell_pos;
  G2 = sc->G;                                         // r0, r0sq, bbox_min, num_hash, rtab: the reference's
  G2.cell_count = cc; G2.cell_start = cs; G2.cell_pos = cp; G2.items = nullptr;
  G2.adaptive = 1;
  if (!G2.cell_count) {
    CK(cudaMalloc((void**)&G2.cell_count, sizeof(unsigned int) * ((size_t)n_hp + 1)));
    CK(cudaMalloc((void**)&G2.cell_start, sizeof(unsigned int) * ((size_t)n_hp + 1)));
  }
  float cell_r = sc->G.r0 * (float)sqrt(meanR);       // the typical current radius
  const size_t max_entries = (size_t)24 * n_hp + 1024;   // the reference's grid holds ~8 per hit point
  size_t need = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, need, G2.cell_count, G2.cell_start, (int)n_hp + 1, st);
  if (need > sc->scan_tmp_bytes) {
    if (sc->d_scan_tmp) cudaFree(sc->d_scan_tmp);
    sc->d_scan_tmp = nullptr; sc->scan_tmp_bytes = 0;
    CK(cudaMalloc(&sc->d_scan_tmp, need));
    sc->scan_tmp_bytes = need;
  }
  unsigned int total = 0;
  for (int attempt = 0;; attempt++) {
    if (cell_r > sc->G.r0) cell_r = sc->G.r0;
    G2.hash_scale = (float)(1.0 / ((double)cell_r * 2.0));
    CK(cudaMemsetAsync(G2.cell_count, 0, sizeof(unsigned int) * ((size_t)n_hp + 1), st));
    k_ppm_grid<<<sc->sm_count * 4, 256, 0, st>>>(sc->HP, G2, n_hp, 2, 0, sc->d_pc, sc->S.mats);
    size_t nb = sc->scan_tmp_bytes;
    cub::DeviceScan::ExclusiveSum(sc->d_scan_tmp, nb, G2.cell_count, G2.cell_start, (int)n_hp + 1, st);
    CK(cudaMemcpyAsync(&total, G2.cell_start + n_hp, sizeof(total), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    *launches += 2;
    if (timing) fprintf(stderr, "b200rt ppm grid: attempt %d cell_r %.5f entries %u (max %zu) at %.2f ms\n", attempt, cell_r, total, max_entries, since(T0));
    if (total <= max_entries || cell_r >= sc->G.r0 || attempt >= 6) break;
    cell_r *= 1.3f;   // hit points that still have a large radius reach too many of the small cells: larger cells
  }
  if ((size_t)total + 1 > sc->entries2_capacity) {
    // with half as much again: the entry count wanders by a per cent from rebuild to rebuild, and a cudaFree +
    // cudaMalloc of 300 MB inside a pass was seen to cost anything between 1 and 270 ms
    const size_t cap = (size_t)total + (size_t)total / 2 + 1024;
    if (G2.cell_pos) cudaFree(G2.cell_pos);
    G2.cell_pos = nullptr; sc->entries2_capacity = 0;
    CK(cudaMalloc((void**)&G2.cell_pos, sizeof(float4) * cap));
    sc->entries2_capacity = cap;
  }
  CK(cudaMemcpyAsync(G2.cell_count, G2.cell_start, sizeof(unsigned int) * ((size_t)n_hp + 1), cudaMemcpyDeviceToDevice, st));
  k_ppm_grid<<<sc->sm_count * 4, 256, 0, st>>>(sc->HP, G2, n_hp, 1, (uint32_t)sc->entries2_capacity, sc->d_pc, sc->S.mats);
  CK(cudaGetLastError());
  if (timing) {
    CK(cudaStreamSynchronize(st));
    fprintf(stderr, "b200rt ppm grid: rebuilt, mean R %.4f, %u entries, capacity %zu, %.2f ms\n", meanR, total, sc->entries2_capacity, since(T0));
  }
  (*launches)++;
  (*rebuilds)++;
  sc->g2_valid = true;
  sc->g2_sumR = sumR;
  return B200RT_OK;
}

// Canonical numbering of the hit points (multi_gpu.cuh): sort by (pixel, stratum, tree code), then — stable — by home
// cell when `by_cell`.  cub radix sorts: library calls on a once-per-frame step.
int ppm_order_hit_points(b200rt_scene* sc, uint32_t n_hp, bool by_cell) {
  if (n_hp < 2) return B200RT_OK;
  cudaStream_t st = sc->c0.stream;
  HitPoints& H = sc->HP;
  unsigned long long *k64a = nullptr, *k64b = nullptr;
  uint32_t *ia = nullptr, *ib = nullptr, *k32a = nullptr, *k32b = nullptr;
  void* tmp = nullptr;
  int rc = B200RT_OK;
  const int grid = sc->sm_count * 4;
  do {
    if (cudaMalloc((void**)&k64a, 8 * (size_t)n_hp) != cudaSuccess || cudaMalloc((void**)&k64b, 8 * (size_t)n_hp) != cudaSuccess ||
        cudaMalloc((void**)&ia, 4 * (size_t)n_hp) != cudaSuccess || cudaMalloc((void**)&ib, 4 * (size_t)n_hp) != cudaSuccess) {
      rc = B200RT_E_CUDA; break;
    }
    size_t need64 = 0, need32 = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, need64, k64a, k64b, ia, ib, (int)n_hp, 0, 64, st);
    k32a = reinterpret_cast<uint32_t*>(k64a); k32b = reinterpret_cast<uint32_t*>(k64b);
    cub::DeviceRadixSort::SortPairs(nullptr, need32, k32a, k32b, ia, ib, (int)n_hp, 0, 32, st);
    const size_t need = std::max(need64, need32);
    if (cudaMalloc(&tmp, need) != cudaSuccess) { rc = B200RT_E_CUDA; break; }
    k_ppm_order_keys<<<grid, 256, 0, st>>>(H.wo, H.pos, n_hp, k64a, ia);
    size_t n1 = need;
    cub::DeviceRadixSort::SortPairs(tmp, n1, k64a, k64b, ia, ib, (int)n_hp, 0, 64, st);
    k_ppm_gather4<<<grid, 256, 0, st>>>(H.pos, H.nrm, H.wo, H.att, ib, n_hp, H.dep);
    k_ppm_scatter4<<<grid, 256, 0, st>>>(H.dep, n_hp, H.pos, H.nrm, H.wo, H.att);
    if (by_cell) {
      k_ppm_home_cell_keys<<<grid, 256, 0, st>>>(H.pos, sc->G, n_hp, k32a, ia);
      size_t n2 = need;
      cub::DeviceRadixSort::SortPairs(tmp, n2, k32a, k32b, ia, ib, (int)n_hp, 0, 32, st);
      k_ppm_gather4<<<grid, 256, 0, st>>>(H.pos, H.nrm, H.wo, H.att, ib, n_hp, H.dep);
      k_ppm_scatter4<<<grid, 256, 0, st>>>(H.dep, n_hp, H.pos, H.nrm, H.wo, H.att);
    }
    if (cudaStreamSynchronize(st) != cudaSuccess || cudaGetLastError() != cudaSuccess) rc = B200RT_E_CUDA;
  } while (0);
  if (rc) b200rt_internal_set_error("PPM: ordering the hit points failed (CUDA error / out of memory)");
  cudaFree(k64a); cudaFree(k64b); cudaFree(ia); cudaFree(ib); cudaFree(tmp);
  return rc;
}

}  // namespace

extern "C" {

int b200rt_ppm_eye_pass(b200rt_scene* sc, int camera_index, uint64_t seed, B200PpmStats* stats) try {
  if (!sc || camera_index < 0 || camera_index >= (int)sc->cameras.size() || sc->S.semantics != B200RT_SEM_PPM) {
    b200rt_internal_set_error("b200rt_ppm_eye_pass: needs a PPM scene and a valid camera index");
    return B200RT_E_INVALID;
  }
  CK(cudaSetDevice(sc->device));
  sc->S.cam = sc->cameras[camera_index];
  sc->ppm_camera = camera_index;
  sc->g2_valid = false;
  const b200rt_camera& c = sc->S.cam;
  if (sc->S.max_depth + 2 > kMaxWaves) return B200RT_E_UNSUPPORTED;
  cudaStream_t st = sc->c0.stream;
  RenderParams R;
  B200RenderOpts o;
  memset(&o, 0, sizeof(o));
  o.row_stride = 1; o.seed = seed;
  if (resolve_rows(c, &o, &R) != B200RT_OK) return B200RT_E_INVALID;
  const size_t paths = (size_t)c.width * c.height * R.sample_count;
  R.n_paths = (uint32_t)paths;
  // every glass vertex can fork: size for one fork per path and level, detect overflow otherwise
  const size_t slots = paths * (size_t)(1 + (sc->has_transparency ? 2 * sc->S.max_depth : sc->S.max_depth)) + 1024;
  if (slots > 0xF0000000ull) return B200RT_E_UNSUPPORTED;
  int rc = ensure_pool(sc, sc->c0, slots, 1024);
  if (rc) return rc;
  rc = ppm_ensure_hit_points(sc, slots);
  if (rc) return rc;
  rc = ppm_ensure_rtab(sc);
  if (rc) return rc;
  if (!sc->d_pc) CK(cudaMalloc((void**)&sc->d_pc, sizeof(PpmCounters)));
  PpmCounters hpc;
  memset(&hpc, 0, sizeof(hpc));
  for (int i = 0; i < 3; i++) { hpc.bbox[i] = 0xffffffffu; hpc.bbox[3 + i] = 0u; }
  CK(cudaMemcpyAsync(sc->d_pc, &hpc, sizeof(hpc), cudaMemcpyHostToDevice, st));
  CK(cudaMemsetAsync(sc->HP.count, 0, sizeof(unsigned int), st));
  CK(cudaMemsetAsync(sc->c0.d_counters, 0, sizeof(Counters), st));
  const int grid_p = sc->sm_count * (RT_FAST_MIN_BLOCKS > 8 ? RT_FAST_MIN_BLOCKS : 8);
  int launches = 0;
  cudaEvent_t e0, e1, e2;
  cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventCreate(&e2);
  cudaEventRecord(e0, st);
  k_generate<<<sc->sm_count * 8, 256, 0, st>>>(sc->S, sc->c0.P, R, sc->c0.d_waves, sc->c0.d_counters);
  launches++;
  k_next_wave<<<1, 1, 0, st>>>(sc->c0.d_waves, 0, sc->c0.d_counters);
  launches++;
  for (int d = 0; d <= sc->S.max_depth; d++) {
    if (sc->S.use_wide) { launch_wide(sc, sc->c0, false, true, d); launches++; }
    else k_trace_fast<false, 7><<<grid_p, RT_BLOCK, 0, st>>>(sc->S, sc->c0.P, sc->c0.Q, sc->c0.d_waves, d, sc->c0.d_counters);
    k_ppm_eye_shade<<<sc->sm_count * 8, RT_BLOCK, 0, st>>>(sc->S, sc->c0.P, sc->c0.d_waves, d, sc->HP, sc->c0.d_counters, sc->d_pc);
    k_finish_wave<<<1, 1, 0, st>>>(sc->c0.d_waves, d, sc->c0.P, sc->c0.d_counters);
    launches += 3;
  }
  cudaEventRecord(e1, st);
  CK(cudaStreamSynchronize(st));
  Counters hc;
  CK(cudaMemcpy(&hc, sc->c0.d_counters, sizeof(hc), cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(&hpc, sc->d_pc, sizeof(hpc), cudaMemcpyDeviceToHost));
  unsigned int n_hp = 0;
  CK(cudaMemcpy(&n_hp, sc->HP.count, sizeof(n_hp), cudaMemcpyDeviceToHost));
  if (hc.overflow || n_hp > sc->HP.capacity) {
    b200rt_internal_set_error("PPM eye pass: vertex pool / hit point overflow");
    return B200RT_E_UNSUPPORTED;
  }
  sc->n_hp = n_hp;
  // Scene::build_hash_grid — PPM/src/Scene.cpp:53-93 (host: the scalars; device: counting sort into CSR)
  PpmGrid& G = sc->G;
  const float* rtab = sc->d_rtab;
  if (G.cell_count) cudaFree(G.cell_count);
  if (G.cell_start) cudaFree(G.cell_start);
  G.cell_count = nullptr; G.cell_start = nullptr;
  G.rtab = rtab;
  G.num_hash = n_hp;
  float r0 = 0.f;
  size_t entries = 0;
  if (n_hp > 0) {
    float mn[3], mx[3];
    for (int i = 0; i < 3; i++) { mn[i] = ord2f_bits(hpc.bbox[i]); mx[i] = ord2f_bits(hpc.bbox[3 + i]); }
    const float sx = mx[0] - mn[0], sy = mx[1] - mn[1], sz = mx[2] - mn[2];
    r0 = ((sx + sy + sz) / 3.0f) / ((c.width + c.height) / 2.0f) * 2.0f * 4.0f;
    G.r0 = r0; G.r0sq = r0 * r0;
    G.bbox_min = mk3(mn[0] - r0, mn[1] - r0, mn[2] - r0);  // min_i fl(p_i - r0) == fl(min_i p_i - r0)
    G.hash_scale = (float)(1.0 / ((double)r0 * 2.0));
    G.adaptive = 0;
    {
      const char* e = getenv("B200RT_PPM_ORDER");   // "none": append order (one GPU only), "pixel", "cell" (default)
      if (!(e && !strcmp(e, "none"))) {
        rc = ppm_order_hit_points(sc, n_hp, !(e && !strcmp(e, "pixel")));
        if (rc) return rc;
        launches += 8;
      }
    }
    CK(cudaMalloc((void**)&G.cell_count, sizeof(unsigned int) * ((size_t)n_hp + 1)));
    CK(cudaMalloc((void**)&G.cell_start, sizeof(unsigned int) * ((size_t)n_hp + 1)));
    CK(cudaMemsetAsync(G.cell_count, 0, sizeof(unsigned int) * ((size_t)n_hp + 1), st));
    k_ppm_grid<<<sc->sm_count * 4, 256, 0, st>>>(sc->HP, G, n_hp, 0, 0, sc->d_pc, sc->S.mats);
    size_t need = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, need, G.cell_count, G.cell_start, (int)n_hp + 1, st);
    if (need > sc->scan_tmp_bytes) {
      if (sc->d_scan_tmp) cudaFree(sc->d_scan_tmp);
      CK(cudaMalloc(&sc->d_scan_tmp, need));
      sc->scan_tmp_bytes = need;
    }
    cub::DeviceScan::ExclusiveSum(sc->d_scan_tmp, need, G.cell_count, G.cell_start, (int)n_hp + 1, st);
    unsigned int total = 0;
    CK(cudaMemcpyAsync(&total, G.cell_start + n_hp, sizeof(total), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    entries = total;
    if (entries > sc->items_capacity) {
      if (G.items) cudaFree(G.items);
      if (G.cell_pos) cudaFree(G.cell_pos);
      G.items = nullptr; G.cell_pos = nullptr;
      CK(cudaMalloc((void**)&G.items, sizeof(unsigned int) * (entries + 1)));
      CK(cudaMalloc((void**)&G.cell_pos, sizeof(float4) * (entries + 1)));
      sc->items_capacity = entries + 1;
    }
    CK(cudaMemcpyAsync(G.cell_count, G.cell_start, sizeof(unsigned int) * ((size_t)n_hp + 1), cudaMemcpyDeviceToDevice, st));
    k_ppm_grid<<<sc->sm_count * 4, 256, 0, st>>>(sc->HP, G, n_hp, 1, (uint32_t)sc->items_capacity, sc->d_pc, sc->S.mats);
    launches += 4;
  }
  cudaEventRecord(e2, st);
  CK(cudaStreamSynchronize(st));
  CK(cudaGetLastError());
  if (stats) {
    memset(stats, 0, sizeof(*stats));
    stats->hit_points = n_hp; stats->grid_entries = entries; stats->initial_radius = r0;
    cudaEventElapsedTime(&stats->ms_eye, e0, e1);
    cudaEventElapsedTime(&stats->ms_grid, e1, e2);
    stats->launches = launches;
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1); cudaEventDestroy(e2);
  return B200RT_OK;
} B200RT_CATCH

int b200rt_ppm_trace_photons(b200rt_scene* sc, const B200PpmOpts* opts, B200PpmStats* stats) try {
  if (!sc || !opts || opts->photons < 0 || sc->S.semantics != B200RT_SEM_PPM || sc->ppm_camera < 0) {
    b200rt_internal_set_error("b200rt_ppm_trace_photons: call b200rt_ppm_eye_pass on a PPM scene first");
    return B200RT_E_INVALID;
  }
  if (sc->S.n_lights < 1) {
    b200rt_internal_set_error("b200rt_ppm_trace_photons: the reference emits from lights[0]; the scene has no light");
    return B200RT_E_INVALID;
  }
  CK(cudaSetDevice(sc->device));
  cudaStream_t st = sc->c0.stream;
  const size_t batch = (size_t)8 << 20;  // photons per wavefront batch
  int rc = ensure_pool(sc, sc->c0, 2 * batch + 1024, 1024);
  if (rc) return rc;
  const uint32_t half = (uint32_t)(sc->c0.P.capacity / 2);
  const uint32_t s0 = rng_mix((uint32_t)opts->seed, (uint32_t)(opts->seed >> 32));
  const int grid_p = sc->sm_count * (RT_FAST_MIN_BLOCKS > 8 ? RT_FAST_MIN_BLOCKS : 8);
  PpmCounters hpc0, hpc1;
  CK(cudaMemcpy(&hpc0, sc->d_pc, sizeof(hpc0), cudaMemcpyDeviceToHost));
  CK(cudaMemsetAsync(sc->c0.d_counters, 0, sizeof(Counters), st));
  Timers tm = make_timers(sc->c0, stats != nullptr);
  int launches = 0;
  cudaEventRecord(sc->c0.ev[0], st);
  tm.mark(ST_GENERATE);
  int rebuilds = 0;
  for (int64_t done = 0; done < opts->photons && sc->S.max_depth > 1; done += (int64_t)batch) {
    const uint32_t n = (uint32_t)((opts->photons - done) < (int64_t)batch ? (opts->photons - done) : (int64_t)batch);
    rc = ppm_rebuild_scan_grid(sc, &launches, &rebuilds);   // before every batch: cheap when nothing is to be done
    if (rc) return rc;
    const PpmGrid& Gs = sc->g2_valid ? sc->G2 : sc->G;
    const uint32_t* order = nullptr;
    if (!getenv("B200RT_PPM_NO_SORT")) {
      // trace the batch in direction order (ppm_kernels.cuh: k_ppm_emit_keys); cub radix sort on 16-bit keys is a
      // library call on a once-per-batch step
      if (!sc->d_sort[0]) {
        for (int q = 0; q < 4; q++) CK(cudaMalloc((void**)&sc->d_sort[q], batch * sizeof(uint32_t)));
      }
      int bits = PPM_SORT_BITS;
      if (const char* e = getenv("B200RT_PPM_SORT_BITS")) bits = std::max(1, std::min(16, atoi(e)));
      k_ppm_emit_keys<<<sc->sm_count * 8, 256, 0, st>>>(s0, opts->first_photon + (uint64_t)done, n, bits, sc->d_sort[0],
                                                        sc->d_sort[1]);
      size_t need = 0;
      cub::DeviceRadixSort::SortPairs(nullptr, need, sc->d_sort[0], sc->d_sort[2], sc->d_sort[1], sc->d_sort[3], (int)n,
                                      0, 16, st);
      if (need > sc->scan_tmp_bytes) {
        CK(cudaStreamSynchronize(st));
        if (sc->d_scan_tmp) cudaFree(sc->d_scan_tmp);
        CK(cudaMalloc(&sc->d_scan_tmp, need));
        sc->scan_tmp_bytes = need;
      }
      cub::DeviceRadixSort::SortPairs(sc->d_scan_tmp, need, sc->d_sort[0], sc->d_sort[2], sc->d_sort[1], sc->d_sort[3],
                                      (int)n, 0, bits, st);
      order = sc->d_sort[3];
      launches += 3;
    }
    k_ppm_emit<<<sc->sm_count * 8, 256, 0, st>>>(sc->S, sc->c0.P, sc->c0.d_waves, s0, opts->first_photon + (uint64_t)done, n,
                                                 order);
    launches++;
    tm.mark(ST_GENERATE);
    for (int d = 0; d + 1 < sc->S.max_depth; d++) {
      k_ppm_next_wave<<<1, 1, 0, st>>>(sc->c0.d_waves, d, half, sc->c0.d_counters);
      k_ppm_count_rays<<<1, 1, 0, st>>>(sc->c0.d_waves, d, sc->d_pc);
      if (sc->S.use_wide) { launch_wide(sc, sc->c0, false, true, d); launches++; }
      else k_trace_fast<false, 7><<<grid_p, RT_BLOCK, 0, st>>>(sc->S, sc->c0.P, sc->c0.Q, sc->c0.d_waves, d, sc->c0.d_counters);
      tm.mark(ST_EXTEND);
      k_ppm_photon_shade<<<sc->sm_count * 8, RT_BLOCK, 0, st>>>(sc->S, sc->c0.P, sc->c0.d_waves, d, sc->HP, Gs,
                                                                sc->c0.d_counters, sc->d_pc, sc->G.cell_start, sc->G.hash_scale);
      tm.mark(ST_SHADE);
      launches += 4;
    }
  }
  cudaEventRecord(sc->c0.ev[1], st);
  CK(cudaStreamSynchronize(st));
  CK(cudaGetLastError());
  rc = check_overflow(sc, "b200rt_ppm_trace_photons");
  if (rc) return rc;
  if (stats) {
    CK(cudaMemcpy(&hpc1, sc->d_pc, sizeof(hpc1), cudaMemcpyDeviceToHost));
    Counters hc;
    CK(cudaMemcpy(&hc, sc->c0.d_counters, sizeof(hc), cudaMemcpyDeviceToHost));
    memset(stats, 0, sizeof(*stats));
    stats->hit_points = sc->n_hp; stats->photons = (uint64_t)opts->photons;
    stats->photon_rays = hpc1.photon_rays - hpc0.photon_rays;
    stats->cell_scans = hpc1.cell_scans - hpc0.cell_scans;
    stats->deposits = hpc1.deposits - hpc0.deposits;
    stats->entries_scanned = hpc1.entries_scanned - hpc0.entries_scanned;
    stats->scan_cell = 1.0f / (sc->g2_valid ? sc->G2.hash_scale : sc->G.hash_scale);
    stats->grid_rebuilds = rebuilds;
    stats->nodes_visited = hc.nodes_visited; stats->tris_tested = hc.tris_tested;
    float ms[ST_COUNT];
    tm.collect(ms);
    stats->ms_extend = ms[ST_EXTEND]; stats->ms_deposit = ms[ST_SHADE];
    cudaEventElapsedTime(&stats->ms_photon, sc->c0.ev[0], sc->c0.ev[1]);
    stats->initial_radius = sc->G.r0;
    stats->launches = launches;
    stats->node_bytes = sc->S.use_wide ? (int)sizeof(b200rt_wide::Node) : 64;
    stats->top_objects = sc->S.n_top;
  }
  return B200RT_OK;
} B200RT_CATCH

int b200rt_ppm_export(b200rt_scene* sc, float** d_S_n, int64_t* n_hit_points) try {
  if (!sc || !d_S_n || sc->ppm_camera < 0) return B200RT_E_INVALID;
  CK(cudaSetDevice(sc->device));
  if (sc->n_hp) k_ppm_export<<<sc->sm_count * 4, 256, 0, sc->c0.stream>>>(sc->HP, sc->n_hp, sc->d_export);
  CK(cudaStreamSynchronize(sc->c0.stream));
  *d_S_n = reinterpret_cast<float*>(sc->d_export);
  if (n_hit_points) *n_hit_points = sc->n_hp;
  return B200RT_OK;
} B200RT_CATCH

int b200rt_ppm_density_from(b200rt_scene* sc, const float* d_S_n, double total, float* rgbw_out) try {
  if (!sc || !rgbw_out || sc->ppm_camera < 0 || (!d_S_n && sc->n_hp)) return B200RT_E_INVALID;
  CK(cudaSetDevice(sc->device));
  const b200rt_camera& c = sc->cameras[sc->ppm_camera];
  const size_t pixels = (size_t)c.width * c.height;
  int rc = ensure_accum(sc, pixels);
  if (rc) return rc;
  cudaStream_t st = sc->c0.stream;
  CK(cudaMemsetAsync(sc->d_accum, 0, pixels * sizeof(float4), st));
  if (sc->n_hp)
    k_ppm_density<<<sc->sm_count * 4, 256, 0, st>>>(sc->HP, sc->n_hp, reinterpret_cast<const float4*>(d_S_n), sc->G.r0sq,
                                                    total, sc->d_accum);
  CK(cudaMemcpyAsync(rgbw_out, sc->d_accum, pixels * sizeof(float4), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  CK(cudaGetLastError());
  return B200RT_OK;
} B200RT_CATCH

int b200rt_ppm_density(b200rt_scene* sc, double total, float* rgbw_out) try {
  float* p = nullptr;
  int rc = b200rt_ppm_export(sc, &p, nullptr);
  if (rc) return rc;
  return b200rt_ppm_density_from(sc, p, total, rgbw_out);
} B200RT_CATCH

int b200rt_ppm_get_hit_points(b200rt_scene* sc, float* out16, int64_t capacity, int64_t* n_out) try {
  if (!sc || sc->ppm_camera < 0) return B200RT_E_INVALID;
  CK(cudaSetDevice(sc->device));
  if (n_out) *n_out = sc->n_hp;
  if (!out16 || capacity <= 0 || sc->n_hp == 0) return B200RT_OK;
  const uint32_t n = (uint32_t)((int64_t)sc->n_hp < capacity ? (int64_t)sc->n_hp : capacity);
  float* d = nullptr;
  CK(cudaMalloc((void**)&d, sizeof(float) * 16 * (size_t)n));
  k_ppm_pack_hit_points<<<sc->sm_count * 4, 256, 0, sc->c0.stream>>>(sc->HP, n, d);
  CK(cudaMemcpyAsync(out16, d, sizeof(float) * 16 * (size_t)n, cudaMemcpyDeviceToHost, sc->c0.stream));
  CK(cudaStreamSynchronize(sc->c0.stream));
  cudaFree(d);
  return B200RT_OK;
} B200RT_CATCH

int b200rt_device_count(void) try {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
} B200RT_CATCH

const char* b200rt_version(void) { return "b200rt 0.1 (sm_100a)"; }

int b200rt_scene_create(const B200SceneDesc* d, int device, b200rt_scene** out) try {
  if (!d || !out) {
    b200rt_internal_set_error("b200rt_scene_create: null argument");
    return B200RT_E_INVALID;
  }
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    b200rt_internal_set_error("no CUDA device: the B200 core has no CPU fallback");
    return B200RT_E_NOGPU;
  }
  if (device < 0 || device >= ndev) {
    b200rt_internal_set_error("b200rt_scene_create: bad device index");
    return B200RT_E_INVALID;
  }
  if (d->semantics != B200RT_SEM_HW1 && d->semantics != B200RT_SEM_HW2 && d->semantics != B200RT_SEM_HW3 &&
      d->semantics != B200RT_SEM_HW4 && d->semantics != B200RT_SEM_HW5 && d->semantics != B200RT_SEM_HW6 &&
      d->semantics != B200RT_SEM_HW7 && d->semantics != B200RT_SEM_PPM) {
    b200rt_internal_set_error("b200rt_scene_create: semantics not implemented on the device yet");
    return B200RT_E_UNSUPPORTED;
  }
  CK(cudaSetDevice(device));
  b200rt_scene* sc = new b200rt_scene();
  struct Guard {   // every early return below releases what has been allocated so far
    b200rt_scene* p;
    ~Guard() { if (p) b200rt_scene_destroy(p); }
  } guard{sc};
  sc->c0.d_waves = nullptr; sc->c0.d_counters = nullptr; sc->c0.stream = nullptr; sc->c0.ev[0] = nullptr; sc->c0.ev[1] = nullptr;
  memset(&sc->S, 0, sizeof(DScene));
  memset(&sc->c0.P, 0, sizeof(Pool));
  memset(&sc->c0.Q, 0, sizeof(ShadowQueue));
  sc->device = device;
  sc->d_accum = nullptr; sc->accum_pixels = 0;
  sc->d_lit = nullptr; sc->lit_entries = 0; sc->lit_on = false;
  sc->d_stage[0] = sc->d_stage[1] = nullptr; sc->stage_pixels = 0; sc->copy_stream = nullptr;
  sc->ev_staged[0] = sc->ev_staged[1] = sc->ev_copied[0] = sc->ev_copied[1] = nullptr;
  sc->copy_pending[0] = sc->copy_pending[1] = false; sc->stage_next = 0;
  memset(&sc->HP, 0, sizeof(HitPoints));
  memset(&sc->G, 0, sizeof(PpmGrid));
  sc->d_pc = nullptr; sc->d_rtab = nullptr; sc->d_export = nullptr; sc->d_scan_tmp = nullptr;
  for (int q = 0; q < 4; q++) sc->d_sort[q] = nullptr;
  sc->scan_tmp_bytes = 0; sc->items_capacity = 0; sc->n_hp = 0; sc->ppm_camera = -1;
  memset(&sc->G2, 0, sizeof(PpmGrid));
  sc->entries2_capacity = 0; sc->g2_sumR = 0.0; sc->g2_valid = false; sc->d_sumR = nullptr;
  sc->rt_frame = nullptr; sc->rt_wait_flag = nullptr; sc->rt_wait_target = 0; sc->d_sync_err = nullptr;
  sc->own_frame_no = 0; sc->own_parts_seen = 0; sc->own_camera = -1; sc->own_direct = false;
  sc->tgt_frame = nullptr; sc->tgt_tail = nullptr; sc->tgt_pixels = 0; sc->tgt_w = 0; sc->tgt_h = 0;
  sc->tgt_direct = false; sc->tgt_ipc = false; sc->tgt_self = false; sc->tgt_frame_no = 0;
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, device));
  sc->sm_count = prop.multiProcessorCount;
  DScene& S = sc->S;
  S.semantics = d->semantics; S.max_depth = d->max_depth;
  S.refill_closest = RT_REFILL_MIN_IDLE; S.refill_any = RT_REFILL_MIN_IDLE;
  S.node_vote = RT_NODE_VOTE;
  if (const char* e = getenv("B200RT_NODE_VOTE")) S.node_vote = std::max(1, std::min(32, atoi(e)));
  if (const char* e = getenv("B200RT_REFILL_CLOSEST")) S.refill_closest = std::max(1, std::min(32, atoi(e)));  // tuning knobs
  if (const char* e = getenv("B200RT_REFILL_ANY")) S.refill_any = std::max(1, std::min(32, atoi(e)));
  S.background = ld3(d->background); S.ambient = ld3(d->ambient);
  S.shadow_eps = d->shadow_eps; S.isect_eps = d->isect_eps;
  S.root = d->root;
  memcpy(S.root_box, d->root_box, sizeof(S.root_box));
  S.n_lights = d->n_lights; S.n_tris = d->n_tris; S.n_spheres = d->n_spheres; S.n_inst = d->n_instances;
  int rc = B200RT_OK;
  size_t bytes = 0;
#define UP(dst, src, n, T)                                                 \
  do {                                                                     \
    T* p_ = nullptr;                                                       \
    rc = upload<T>((const T*)(src), (size_t)(n), &p_);                     \
    if (rc) return rc;                                                     \
    sc->allocs.push_back(p_);                                              \
    dst = reinterpret_cast<decltype(dst)>(p_);                             \
    bytes += (size_t)(n) * sizeof(T);                                      \
  } while (0)
  UP(S.nodes, d->nodes, d->n_nodes, b200rt_bvh_node);
  {
    // Flat list of the top-level objects with the innermost box of the reference's gate chain above each
    // (b200rt_wide::flatten_top, csrc/wide_bvh.cpp): a short list tested by all lanes in lock step replaces the
    // divergent walk of the top-level tree (25 % of the traversal kernels' instructions on cfg4, profiles/r01g).
    std::vector<float> flat;
    const bool ok = b200rt_wide::flatten_top(d, flat);
    const size_t n_top = flat.size() / 8;
    if (!ok || n_top == 0 || n_top > RT_TOP_LINEAR_MAX || getenv("B200RT_TOP_TREE")) {
      S.n_top = 0; S.top_objs = nullptr;
    } else {
      S.n_top = (int)n_top;
      UP(S.top_objs, flat.data(), flat.size(), float);
    }
  }
  {
    // 12 floats per triangle -> 3 float4 (same bytes)
    float* p = nullptr;
    rc = upload<float>(d->tri_geom, (size_t)d->n_tris * 12, &p);
    if (rc) return rc;
    sc->allocs.push_back(p);
    S.tris = reinterpret_cast<const float4*>(p);
    bytes += (size_t)d->n_tris * 48;
  }
  // A producer that walks the reference's own objects (flatten.cpp) supplies the reference tree
  // and tri_gate but need not know how to build the SAH tree the device walks: build it here from tri_geom
  // (v1 = v0 - a1, v2 = v0 - a2; the boxes are grown anyway, only tri_gate carries the reference's decision).
  std::vector<b200rt_bvh_node> own_nodes;
  std::vector<int32_t> own_ref;
  std::vector<float> own_geom;
  std::vector<b200rt_mesh> own_meshes;
  const b200rt_bvh_node* fast_nodes = d->fast_nodes;
  const float* fast_geom = d->fast_tri_geom;
  const int32_t* fast_ref = d->fast_tri_ref;
  const b200rt_mesh* meshes = d->meshes;
  int n_fast_nodes = d->n_fast_nodes, n_fast_tris = d->n_fast_tris;
  if (d->semantics != B200RT_SEM_HW1 && d->n_fast_tris == 0 && d->n_tris > 0 && d->tri_gate) {
    own_meshes.assign(d->meshes, d->meshes + d->n_meshes);
    for (b200rt_mesh& m : own_meshes) {
      m.fast_root = B200RT_REF_NONE;
      if (m.tri_count < 2) continue;
      std::vector<float> tb((size_t)m.tri_count * 6);
      float ext = 0.f;
      for (int k = 0; k < m.tri_count; k++) {
        const float* g = d->tri_geom + 12 * (size_t)(m.tri_begin + k);
        for (int a = 0; a < 3; a++) {
          const float v0 = g[a], v1 = g[a] - g[3 + a], v2 = g[a] - g[6 + a];
          tb[6 * (size_t)k + a] = fminf(v0, fminf(v1, v2));
          tb[6 * (size_t)k + 3 + a] = fmaxf(v0, fmaxf(v1, v2));
        }
      }
      for (int a = 0; a < 3; a++) {
        ext = fmaxf(ext, m.root_box[3 + a] - m.root_box[a]);
        ext = fmaxf(ext, fmaxf(fabsf(m.root_box[a]), fabsf(m.root_box[3 + a])));
      }
      const float grow = ext * (3e-5f + 4.0f * d->isect_eps) + 1e-30f;
      m.fast_root = b200rt_fast::build(tb.data(), m.tri_count, m.tri_begin, grow, own_nodes, own_ref);
    }
    own_geom.resize(own_ref.size() * 12);
    for (size_t k = 0; k < own_ref.size(); k++)
      memcpy(&own_geom[12 * k], d->tri_geom + 12 * (size_t)own_ref[k], 48);
    fast_nodes = own_nodes.data(); fast_geom = own_geom.data(); fast_ref = own_ref.data();
    meshes = own_meshes.data();
    n_fast_nodes = (int)own_nodes.size(); n_fast_tris = (int)own_ref.size();
  }
  S.hw1_fast_root = B200RT_REF_NONE;
  if (d->semantics == B200RT_SEM_HW1 && d->n_tris >= 2 && !getenv("B200RT_HW1_BRUTE")) {
    // HW1 has no BVH of its own (the reference tests every primitive): one SAH tree over all triangles, walked by
    // k_trace_hw1_fast, which replays HW1's order-dependent rule on the near-tie set (rt_hw1_fast.cuh)
    std::vector<float> tb((size_t)d->n_tris * 6);
    float lo[3] = {3e38f, 3e38f, 3e38f}, hi[3] = {-3e38f, -3e38f, -3e38f};
    for (int k = 0; k < d->n_tris; k++) {
      const float* g = d->tri_geom + 12 * (size_t)k;
      for (int a = 0; a < 3; a++) {
        const float v0 = g[a], v1 = g[a] - g[3 + a], v2 = g[a] - g[6 + a];
        tb[6 * (size_t)k + a] = fminf(v0, fminf(v1, v2));
        tb[6 * (size_t)k + 3 + a] = fmaxf(v0, fmaxf(v1, v2));
        lo[a] = fminf(lo[a], tb[6 * (size_t)k + a]); hi[a] = fmaxf(hi[a], tb[6 * (size_t)k + 3 + a]);
      }
    }
    float ext = 0.f;
    for (int a = 0; a < 3; a++) ext = fmaxf(ext, fmaxf(hi[a] - lo[a], fmaxf(fabsf(lo[a]), fabsf(hi[a]))));
    const float grow = ext * (3e-5f + 4.0f * d->isect_eps) + 1e-30f;
    own_nodes.clear(); own_ref.clear();
    S.hw1_fast_root = b200rt_fast::build(tb.data(), d->n_tris, 0, grow, own_nodes, own_ref);
    own_geom.resize(own_ref.size() * 12);
    for (size_t k = 0; k < own_ref.size(); k++) memcpy(&own_geom[12 * k], d->tri_geom + 12 * (size_t)own_ref[k], 48);
    fast_nodes = own_nodes.data(); fast_geom = own_geom.data(); fast_ref = own_ref.data();
    n_fast_nodes = (int)own_nodes.size(); n_fast_tris = (int)own_ref.size();
  }
  UP(S.fast_nodes, fast_nodes, n_fast_nodes, b200rt_bvh_node);
  UP(S.fast_tris, fast_geom, (size_t)n_fast_tris * 12, float);
  UP(S.fast_tri_ref, fast_ref, (size_t)n_fast_tris, int32_t);
  {
    // The tree the default kernel walks inside a mesh: the SAH tree collapsed 8 wide with quantised child boxes
    // (csrc/wide_bvh.h).  Triangles are stored once more, in the order the wide leaves want them.
    std::vector<b200rt_wide::Node> wnodes;
    std::vector<int32_t> worder;
    std::vector<int32_t> wroot((size_t)(d->n_meshes > 0 ? d->n_meshes : 1), -1);
    b200rt_wide::Stats ws = {0, 0, 0, 0};
    const bool want = d->semantics != B200RT_SEM_HW1 && S.n_top > 0 && n_fast_tris > 0 && !getenv("B200RT_NO_WIDE");
    if (want) {
      for (int mi = 0; mi < d->n_meshes; mi++) {
        const b200rt_mesh& m = meshes[mi];
        if (m.tri_count < 2 || m.fast_root == B200RT_REF_NONE) continue;
        float ext = 0.f;
        for (int a = 0; a < 3; a++) {
          ext = fmaxf(ext, m.root_box[3 + a] - m.root_box[a]);
          ext = fmaxf(ext, fmaxf(fabsf(m.root_box[a]), fabsf(m.root_box[3 + a])));
        }
        const float grow = ext * (3e-5f + 4.0f * d->isect_eps) + 1e-30f;
        wroot[mi] = b200rt_wide::build(fast_nodes, m.fast_root, fast_geom, grow, wnodes, worder, &ws);
      }
    }
    std::vector<float> wgeom(worder.size() * 12);
    std::vector<int32_t> wref(worder.size());
    for (size_t k = 0; k < worder.size(); k++) {
      memcpy(&wgeom[12 * k], fast_geom + 12 * (size_t)worder[k], 48);
      wref[k] = fast_ref[worder[k]];
    }
    UP(S.wide_nodes, wnodes.data(), wnodes.size() * (sizeof(b200rt_wide::Node) / 4), float);   // 5 or 8 float4 per node
    UP(S.wide_tris, wgeom.data(), wgeom.size(), float);
    UP(S.wide_tri_ref, wref.data(), wref.size(), int32_t);
    UP(S.wide_root, wroot.data(), wroot.size(), int32_t);
    // the shared-memory stack of k_trace_wide holds one node group per level
    S.use_wide = (want && ws.depth <= RT_WSTACK) ? 1 : 0;
    S.wide_node_vote = RT_WIDE_NODE_VOTE;
    if (const char* e = getenv("B200RT_WIDE_NODE_VOTE")) S.wide_node_vote = std::max(1, std::min(32, atoi(e)));
    sc->wide_nodes = (int)wnodes.size(); sc->wide_depth = ws.depth;
    sc->wide_blocks = RT_WIDE_MIN_BLOCKS;
    if (const char* e = getenv("B200RT_WIDE_BLOCKS")) { const int v = atoi(e); if (v == 4 || v == 6 || v == 8) sc->wide_blocks = v; }
  }
  UP(S.tri_gate, d->tri_gate, (size_t)d->n_tris * 6, float);
  UP(S.tri_info, d->tri_info, (size_t)d->n_tris * 2, int32_t);
  UP(S.tri_vidx, d->tri_vidx, (size_t)d->n_tris * 3, int32_t);
  {
    std::vector<float4> vn((size_t)(d->n_verts > 0 ? d->n_verts : 1));
    for (int i = 0; i < d->n_verts; i++)
      vn[i] = make_float4(d->vert_normal[3 * (size_t)i], d->vert_normal[3 * (size_t)i + 1],
                          d->vert_normal[3 * (size_t)i + 2], 0.f);
    UP(S.vnormal, vn.data(), vn.size(), float4);
  }
  UP(S.meshes, meshes, d->n_meshes, b200rt_mesh);
  sc->n_meshes = d->n_meshes;
  sc->n_nodes = d->n_nodes; sc->n_triprims = d->n_triprims; sc->n_light_cdf = d->n_light_cdf; sc->n_materials = d->n_materials;
  UP(S.inst, d->instances, d->n_instances, b200rt_instance);
  UP(S.spheres, d->spheres, d->n_spheres, b200rt_sphere);
  UP(S.triprims, d->triprims, d->n_triprims, b200rt_triprim);
  UP(S.mats, d->materials, d->n_materials, b200rt_material);
  UP(S.lights, d->lights, d->n_lights, b200rt_light);
  UP(S.vert_pos, d->vert_pos, (size_t)d->n_verts * 3, float);
  UP(S.tri_uv, d->tri_uv, (size_t)d->n_tris * 6, float);
  UP(S.tri_tex, d->tri_tex, (size_t)d->n_tris, int32_t);
  UP(S.textures, d->textures, d->n_textures, b200rt_texture);
  UP(S.texels, d->texels, (size_t)d->n_texels, float);
  UP(S.brdfs, d->brdfs, d->n_brdfs, b200rt_brdf);
  UP(S.light_cdf, d->light_cdf, d->n_light_cdf, float);
  UP(S.light_tri, d->light_tri, d->n_light_cdf, int32_t);
  S.background_texture = d->background_texture; S.env_light = d->env_light;
  S.integrator = d->integrator; S.importance_sampling = d->importance_sampling;
#undef UP
  sc->scene_bytes = bytes;
  sc->cameras.assign(d->cameras, d->cameras + d->n_cameras);
  sc->has_transparency = false;
  sc->has_mirror = false;
  for (int i = 0; i < d->n_materials; i++) {
    const float* t = d->materials[i].transparency;
    if (t[0] != 0.f || t[1] != 0.f || t[2] != 0.f) sc->has_transparency = true;
    const float* mr = d->materials[i].mirror;
    if (mr[0] != 0.f || mr[1] != 0.f || mr[2] != 0.f) sc->has_mirror = true;
  }
  rc = ctx_init(sc->c0);
  if (rc) return rc;
  CK(cudaMalloc((void**)&sc->d_sync_err, sizeof(unsigned int)));
  CK(cudaMemset(sc->d_sync_err, 0, sizeof(unsigned int)));
  sc->n_ctx = kDefaultCtx;
  if (const char* e = getenv("B200RT_CONTEXTS")) sc->n_ctx = std::max(1, std::min(kMaxCtx, atoi(e)));
  guard.p = nullptr;
  *out = sc;
  return B200RT_OK;
} B200RT_CATCH

void b200rt_scene_destroy(b200rt_scene* sc) {
  if (!sc) return;
  cudaSetDevice(sc->device);
  ctx_destroy(sc->c0);
  for (WaveCtx* cx : sc->more) { ctx_destroy(*cx); delete cx; }
  for (void* p : sc->allocs)
    if (p) cudaFree(p);
  ppm_free(sc);
  if (sc->copy_stream) {
    cudaStreamSynchronize(sc->copy_stream);
    cudaStreamDestroy(sc->copy_stream);
  }
  for (int k = 0; k < 2; k++) {
    if (sc->d_stage[k]) cudaFree(sc->d_stage[k]);
    if (sc->ev_staged[k]) cudaEventDestroy(sc->ev_staged[k]);
    if (sc->ev_copied[k]) cudaEventDestroy(sc->ev_copied[k]);
  }
  if (sc->tgt_frame && sc->tgt_ipc) cudaIpcCloseMemHandle(sc->tgt_frame);
  if (sc->d_sync_err) cudaFree(sc->d_sync_err);
  if (sc->d_accum) cudaFree(sc->d_accum);
  if (sc->d_lit) cudaFree(sc->d_lit);
  delete sc;
}

int b200rt_tonemap_photographic(b200rt_scene* sc, int camera_index, float key, float burn_percent, float saturation,
                                uint8_t* rgb8_out) try {
  if (!sc || !rgb8_out || camera_index < 0 || camera_index >= (int)sc->cameras.size() || !sc->d_accum) {
    b200rt_internal_set_error("b200rt_tonemap_photographic: render the camera first");
    return B200RT_E_INVALID;
  }
  CK(cudaSetDevice(sc->device));
  const b200rt_camera& c = sc->cameras[camera_index];
  const uint32_t n = (uint32_t)c.width * (uint32_t)c.height;
  if (n == 0 || n > sc->accum_pixels) return B200RT_E_INVALID;
  cudaStream_t st = sc->c0.stream;
  float* d_f = nullptr;        // lum | scaled | keys | sorted
  double* d_sum = nullptr;
  unsigned char* d_rgb = nullptr;
  void* d_tmp = nullptr;
  int rc = B200RT_OK;
  do {
    if (cudaMalloc((void**)&d_f, sizeof(float) * 4 * (size_t)n) != cudaSuccess ||
        cudaMalloc((void**)&d_sum, sizeof(double)) != cudaSuccess ||
        cudaMalloc((void**)&d_rgb, 3 * (size_t)n) != cudaSuccess) { rc = B200RT_E_CUDA; break; }
    float *lum = d_f, *scaled = d_f + n, *keys = d_f + 2 * (size_t)n, *sorted = d_f + 3 * (size_t)n;
    cudaMemsetAsync(d_sum, 0, sizeof(double), st);
    k_tm_luminance<<<sc->sm_count * 4, 256, 0, st>>>(sc->d_accum, n, lum, d_sum);
    k_tm_scale<<<sc->sm_count * 4, 256, 0, st>>>(n, key, d_sum, lum, scaled, keys);
    size_t need = 0;
    cub::DeviceRadixSort::SortKeys(nullptr, need, keys, sorted, (int)n, 0, 32, st);
    if (cudaMalloc(&d_tmp, need) != cudaSuccess) { rc = B200RT_E_CUDA; break; }
    cub::DeviceRadixSort::SortKeys(d_tmp, need, keys, sorted, (int)n, 0, 32, st);
    // white point: sorted[size - white_count], the last element when white_count == 0 (Photographic_tmo.cpp:30-37)
    float wc = (float)n * burn_percent / 100.0f;
    int white_count = (int)wc;
    if (white_count < 0) white_count = 0;
    if (white_count > (int)n) white_count = (int)n;
    const uint32_t white_index = white_count == 0 ? n - 1 : n - (uint32_t)white_count;
    k_tm_apply<<<sc->sm_count * 4, 256, 0, st>>>(sc->d_accum, n, lum, scaled, sorted, white_index, saturation, d_rgb);
    if (cudaMemcpyAsync(rgb8_out, d_rgb, 3 * (size_t)n, cudaMemcpyDeviceToHost, st) != cudaSuccess ||
        cudaStreamSynchronize(st) != cudaSuccess || cudaGetLastError() != cudaSuccess) rc = B200RT_E_CUDA;
  } while (0);
  if (rc == B200RT_E_CUDA) b200rt_internal_set_error("b200rt_tonemap_photographic: CUDA error");
  if (d_f) cudaFree(d_f);
  if (d_sum) cudaFree(d_sum);
  if (d_rgb) cudaFree(d_rgb);
  if (d_tmp) cudaFree(d_tmp);
  return rc;
} B200RT_CATCH

int b200rt_debug_lit_counts(b200rt_scene* sc, int enable) try {
  if (!sc) return B200RT_E_INVALID;
  sc->lit_on = enable != 0;
  return B200RT_OK;
} B200RT_CATCH

int b200rt_debug_get_lit_counts(b200rt_scene* sc, int camera_index, int32_t* out) try {
  if (!sc || !out || camera_index < 0 || camera_index >= (int)sc->cameras.size() || !sc->d_lit) return B200RT_E_INVALID;
  const b200rt_camera& c = sc->cameras[camera_index];
  const size_t n = (size_t)c.width * c.height * (size_t)(sc->S.n_lights > 0 ? sc->S.n_lights : 1);
  if (n > sc->lit_entries) return B200RT_E_INVALID;
  CK(cudaSetDevice(sc->device));
  CK(cudaMemcpy(out, sc->d_lit, n * sizeof(int), cudaMemcpyDeviceToHost));
  return B200RT_OK;
} B200RT_CATCH

int b200rt_debug_pool(b200rt_scene* sc, int field, void* out, int64_t count) try {
  if (!sc || !out || count < 0 || (uint64_t)count > sc->c0.P.capacity) return B200RT_E_INVALID;
  const Pool& P = sc->c0.P;
  const void* src[] = {P.ray_o, P.ray_d, P.hit_a, P.nrm, P.rec0, P.rec1, P.rec2, P.rec4, P.final_c, P.kd,
                       P.pixel, P.key, P.hit_b, P.rec3};
  const size_t esz[] = {16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 4, 4, 8, 4};
  if (field < 0 || field >= 14 || !src[field]) return B200RT_E_INVALID;
  CK(cudaSetDevice(sc->device));
  CK(cudaMemcpy(out, src[field], esz[field] * (size_t)count, cudaMemcpyDeviceToHost));
  return B200RT_OK;
} B200RT_CATCH

void* b200rt_scene_stream(b200rt_scene* sc) { return sc ? (void*)sc->c0.stream : nullptr; }
// Read-bandwidth probe ("L2 bandwidth is not in MEASURED_PEAKS.json, the builder must measure it").
// Every thread streams 16-byte loads over `bytes` of device memory `iters` times; a working set below the
// 126 MB L2 measures L2, one far above it HBM.  The sum is written so the loads cannot be dropped.
__global__ void __launch_bounds__(256) k_read_probe(const uint4* __restrict__ src, size_t n16, int iters,
                                                    unsigned* sink) {
  unsigned acc = 0;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (int it = 0; it < iters; it++)
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride) {
      uint4 v;
      asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                   : "l"(src + i));
      acc += v.x ^ v.y ^ v.z ^ v.w;
    }
  if (acc == 0x9e3779b9u) *sink = acc;
}

int b200rt_measure_read_bandwidth(int device, uint64_t bytes, int iters, double* gbs_out) try {
  if (!gbs_out || bytes < 4096 || iters < 1) {
    b200rt_internal_set_error("measure_read_bandwidth: bad argument");
    return B200RT_E_INVALID;
  }
  CK(cudaSetDevice(device));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, device));
  uint4* buf = nullptr;
  unsigned* sink = nullptr;
  const size_t n16 = bytes / 16;
  CK(cudaMalloc(&buf, n16 * 16));
  CK(cudaMalloc(&sink, 4));
  CK(cudaMemset(buf, 1, n16 * 16));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  const int grid = prop.multiProcessorCount * 8;
  k_read_probe<<<grid, 256>>>(buf, n16, 2, sink);  // warm: brings a sub-L2 working set into L2
  CK(cudaEventRecord(e0));
  k_read_probe<<<grid, 256>>>(buf, n16, iters, sink);
  CK(cudaEventRecord(e1));
  CK(cudaEventSynchronize(e1));
  float ms = 0.f;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  cudaFree(buf); cudaFree(sink);
  *gbs_out = (double)n16 * 16.0 * iters / (ms * 1e-3) / 1e9;
  return B200RT_OK;
} B200RT_CATCH

// Leaf functions of the device path on caller-supplied inputs (tests/test_leaf_vectors.py feeds the known-answer
// vectors produced by the reference's own functions).
__global__ void k_debug_leaf(int which, DScene S, b200rt_texture T, b200rt_brdf B, const float* in, int rows, int n_in,
                             float* out, int n_out) {
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < rows; k += gridDim.x * blockDim.x) {
    const float* r = in + (size_t)k * n_in;
    float* o = out + (size_t)k * n_out;
    if (which == 0) {          // Bounding_box::intersect: the exact gate and the filtered gate must take the same decision
      const v3 org = ld3(r + 6), dir = ld3(r + 9);
      float e0 = 0.f, e1 = 0.f;
      o[0] = box_gate_exact(org, dir, r[0], r[1], r[2], r[3], r[4], r[5], S.isect_eps, &e0) ? 1.0f : 0.0f;
      const RayCtx rc = make_ray(org, dir, S.isect_eps);
      o[1] = box_gate(rc, r[0], r[1], r[2], r[3], r[4], r[5], S.isect_eps, &e1) ? 1.0f : 0.0f;
    } else if (which == 1) {   // BRDF::get_reflectance
      const v3 f = brdf_eval7(B, ld3(r), ld3(r + 3), ld3(r + 6), ld3(r + 9), ld3(r + 12));
      o[0] = f.x; o[1] = f.y; o[2] = f.z;
    } else if (which == 2) {   // Texture::get_color_at
      const v3 c = tex_color7(S, T, r[0], r[1]);
      o[0] = c.x; o[1] = c.y; o[2] = c.z;
    } else {                   // Perlin_noise::get_value_at
      o[0] = perlin_value7(T, ld3(r));
    }
  }
}

int b200rt_debug_leaf(int device, int which, const b200rt_texture* tex, const float* texels, int64_t n_texels,
                      const b200rt_brdf* brdf, const float* in, int rows, int n_in, float* out, int n_out) try {
  if (which < 0 || which > 3 || !in || !out || rows <= 0 || n_in <= 0 || n_out <= 0) return B200RT_E_INVALID;
  CK(cudaSetDevice(device));
  DScene S;
  memset(&S, 0, sizeof(S));
  S.isect_eps = 0.000001f;
  b200rt_texture T;
  memset(&T, 0, sizeof(T));
  if (tex) T = *tex;
  b200rt_brdf B;
  memset(&B, 0, sizeof(B));
  if (brdf) B = *brdf;
  DevBuf b_tex, b_in, b_out;
  if (texels && n_texels > 0) {
    CK(b_tex.alloc((size_t)n_texels * 4));
    CK(cudaMemcpy(b_tex.p, texels, (size_t)n_texels * 4, cudaMemcpyHostToDevice));
    S.texels = b_tex.as<float>();
  }
  CK(b_in.alloc((size_t)rows * n_in * 4));
  CK(b_out.alloc((size_t)rows * n_out * 4));
  float *d_in = b_in.as<float>(), *d_out = b_out.as<float>();
  CK(cudaMemcpy(d_in, in, (size_t)rows * n_in * 4, cudaMemcpyHostToDevice));
  k_debug_leaf<<<(rows + 127) / 128, 128>>>(which, S, T, B, d_in, rows, n_in, d_out, n_out);
  CK(cudaGetLastError());
  CK(cudaMemcpy(out, d_out, (size_t)rows * n_out * 4, cudaMemcpyDeviceToHost));
  return B200RT_OK;
} B200RT_CATCH

int b200rt_scene_build_bvh_device(b200rt_scene* sc, float* ms_out) try {
  if (!sc) return B200RT_E_INVALID;
  if (sc->S.semantics == B200RT_SEM_HW1) {
    b200rt_internal_set_error("build_bvh_device: HW1 has no per-mesh trees");
    return B200RT_E_UNSUPPORTED;
  }
  CK(cudaSetDevice(sc->device));
  cudaStream_t st = sc->c0.stream;
  const int n_meshes = sc->n_meshes;
  std::vector<b200rt_mesh> meshes((size_t)n_meshes);
  CK(cudaMemcpy(meshes.data(), sc->S.meshes, sizeof(b200rt_mesh) * n_meshes, cudaMemcpyDeviceToHost));
  size_t total_tris = 0, max_tris = 0;
  for (const b200rt_mesh& m : meshes)
    if (m.tri_count >= 2) { total_tris += (size_t)m.tri_count; max_tris = std::max(max_tris, (size_t)m.tri_count); }
  if (total_tris == 0) { if (ms_out) *ms_out = 0.f; return B200RT_OK; }
  // outputs (replace the host-built SAH arrays; the old ones stay allocated until the scene is destroyed)
  b200rt_bvh_node* d_nodes = nullptr;
  float4* d_ftris = nullptr;
  int* d_fref = nullptr;
  CK(cudaMalloc((void**)&d_nodes, sizeof(b200rt_bvh_node) * total_tris));
  CK(cudaMalloc((void**)&d_ftris, sizeof(float4) * 3 * total_tris));
  CK(cudaMalloc((void**)&d_fref, sizeof(int) * total_tris));
  sc->allocs.push_back(d_nodes); sc->allocs.push_back(d_ftris); sc->allocs.push_back(d_fref);
  // scratch for the largest mesh
  uint64_t *d_keys = nullptr, *d_keys2 = nullptr;
  uint32_t *d_idx = nullptr, *d_idx2 = nullptr;
  int* d_parent = nullptr;
  float* d_leaf_box = nullptr;
  unsigned int* d_visits = nullptr;
  void* d_tmp = nullptr;
  size_t tmp_bytes = 0;
  CK(cudaMalloc((void**)&d_keys, 8 * max_tris));
  CK(cudaMalloc((void**)&d_keys2, 8 * max_tris));
  CK(cudaMalloc((void**)&d_idx, 4 * max_tris));
  CK(cudaMalloc((void**)&d_idx2, 4 * max_tris));
  CK(cudaMalloc((void**)&d_parent, 4 * 2 * max_tris));
  CK(cudaMalloc((void**)&d_leaf_box, 4 * 6 * max_tris));
  CK(cudaMalloc((void**)&d_visits, 4 * max_tris));
  cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, d_keys, d_keys2, d_idx, d_idx2, (int)max_tris, 0, 63, st);
  CK(cudaMalloc(&d_tmp, tmp_bytes));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  CK(cudaEventRecord(e0, st));
  const int grid = sc->sm_count * 8;
  size_t base = 0;   // nodes and leaves of one mesh share the offset: n-1 nodes, n leaves
  for (b200rt_mesh& m : meshes) {
    if (m.tri_count < 2) continue;
    const int n = m.tri_count;
    float ext = 0.f;
    for (int q = 0; q < 3; q++) {
      ext = fmaxf(ext, m.root_box[3 + q] - m.root_box[q]);
      ext = fmaxf(ext, fmaxf(fabsf(m.root_box[q]), fabsf(m.root_box[3 + q])));
    }
    const float grow = ext * (3e-5f + 4.0f * sc->S.isect_eps) + 1e-30f;   // as the host builder's SAH boxes
    const float3 lo = make_float3(m.root_box[0], m.root_box[1], m.root_box[2]);
    const float3 ie = make_float3(1.0f / fmaxf(m.root_box[3] - m.root_box[0], 1e-30f),
                                  1.0f / fmaxf(m.root_box[4] - m.root_box[1], 1e-30f),
                                  1.0f / fmaxf(m.root_box[5] - m.root_box[2], 1e-30f));
    k_lbvh_keys<<<grid, 256, 0, st>>>(sc->S.tris, m.tri_begin, n, lo, ie, d_keys, d_idx);
    size_t need = tmp_bytes;
    cub::DeviceRadixSort::SortPairs(d_tmp, need, d_keys, d_keys2, d_idx, d_idx2, n, 0, 63, st);
    k_lbvh_leaves<<<grid, 256, 0, st>>>(sc->S.tris, m.tri_begin, n, d_idx2, (int)base, d_ftris, d_fref, d_leaf_box);
    CK(cudaMemsetAsync(d_visits, 0, 4 * (size_t)n, st));
    k_lbvh_hierarchy<<<grid, 256, 0, st>>>(d_keys2, n, (int)base, (int)base, d_nodes, d_parent);
    k_lbvh_refit<<<grid, 256, 0, st>>>(n, (int)base, d_parent, d_leaf_box, d_nodes, d_visits, grow);
    m.fast_root = B200RT_REF(B200RT_KIND_FNODE, (uint32_t)base);
    base += (size_t)n;
  }
  CK(cudaEventRecord(e1, st));
  CK(cudaMemcpyAsync(const_cast<b200rt_mesh*>(sc->S.meshes), meshes.data(), sizeof(b200rt_mesh) * n_meshes,
                     cudaMemcpyHostToDevice, st));
  CK(cudaStreamSynchronize(st));
  CK(cudaGetLastError());
  float ms = 0.f;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  if (ms_out) *ms_out = ms;
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  cudaFree(d_keys); cudaFree(d_keys2); cudaFree(d_idx); cudaFree(d_idx2); cudaFree(d_parent); cudaFree(d_leaf_box); cudaFree(d_visits); cudaFree(d_tmp);
  sc->S.fast_nodes = reinterpret_cast<const float4*>(d_nodes);
  sc->S.fast_tris = d_ftris;
  sc->S.fast_tri_ref = d_fref;
  sc->S.use_wide = 0;   // the device-built tree is binary: k_trace_fast walks it
  return B200RT_OK;
} B200RT_CATCH

// Moving geometry between frames: the description's top-level part (b200rt_builder_refit, or a producer that walked the
// reference's objects again) replaces the resident one; meshes, their trees and every triangle record stay in HBM.
int b200rt_scene_refit(b200rt_scene* sc, const B200SceneDesc* d, float* ms_out) try {
  if (!sc || !d) {
    b200rt_internal_set_error("b200rt_scene_refit: null argument");
    return B200RT_E_INVALID;
  }
  const DScene& S = sc->S;
  if (d->semantics != S.semantics || d->n_tris != S.n_tris || d->n_meshes != sc->n_meshes || d->n_nodes != sc->n_nodes ||
      d->n_instances != S.n_inst || d->n_spheres != S.n_spheres || d->n_triprims != sc->n_triprims ||
      d->n_lights != S.n_lights || d->n_light_cdf != sc->n_light_cdf || d->n_materials != sc->n_materials) {
    b200rt_internal_set_error("b200rt_scene_refit: the description has other objects than the resident scene (refit moves objects, it does not add or remove them)");
    return B200RT_E_INVALID;
  }
  std::vector<float> flat;
  const bool ok = b200rt_wide::flatten_top(d, flat);
  const size_t n_top = ok ? flat.size() / 8 : 0;
  if ((S.n_top > 0) != (n_top > 0 && n_top <= RT_TOP_LINEAR_MAX) || (S.n_top > 0 && (int)n_top != S.n_top)) {
    b200rt_internal_set_error("b200rt_scene_refit: the flat top-level list changed size");
    return B200RT_E_UNSUPPORTED;
  }
  CK(cudaSetDevice(sc->device));
  cudaStream_t st = sc->c0.stream;
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  CK(cudaEventRecord(e0, st));
#define PUT(dst, src, n, T) \
  if ((n) > 0) CK(cudaMemcpyAsync(const_cast<void*>(reinterpret_cast<const void*>(dst)), (src), sizeof(T) * (size_t)(n), cudaMemcpyHostToDevice, st))
  PUT(S.nodes, d->nodes, d->n_nodes, b200rt_bvh_node);   // Scene::bvh's nodes follow the mesh trees' in one array
  PUT(S.inst, d->instances, d->n_instances, b200rt_instance);
  PUT(S.spheres, d->spheres, d->n_spheres, b200rt_sphere);
  PUT(S.triprims, d->triprims, d->n_triprims, b200rt_triprim);
  PUT(S.lights, d->lights, d->n_lights, b200rt_light);
  PUT(S.light_cdf, d->light_cdf, d->n_light_cdf, float);
  PUT(S.light_tri, d->light_tri, d->n_light_cdf, int32_t);
  if (S.n_top > 0) PUT(S.top_objs, flat.data(), flat.size(), float);
#undef PUT
  CK(cudaEventRecord(e1, st));
  CK(cudaStreamSynchronize(st));   // the host arrays may go away after the call
  sc->S.root = d->root;
  memcpy(sc->S.root_box, d->root_box, sizeof(sc->S.root_box));
  if (ms_out) cudaEventElapsedTime(ms_out, e0, e1);
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  return B200RT_OK;
} B200RT_CATCH

uint64_t b200rt_scene_bytes(b200rt_scene* sc) { return sc ? (uint64_t)sc->scene_bytes : 0; }

int b200rt_render_device(b200rt_scene* sc, int camera_index, const B200RenderOpts* opts, float** d_rgbw_out,
                         B200RenderStats* stats) try {
  int rc = render_impl(sc, camera_index, opts, stats);
  if (rc) return rc;
  if (d_rgbw_out) *d_rgbw_out = reinterpret_cast<float*>(sc->d_accum);
  return B200RT_OK;
} B200RT_CATCH

int b200rt_render(b200rt_scene* sc, int camera_index, const B200RenderOpts* opts, float* rgbw_out,
                  B200RenderStats* stats) try {
  if (!rgbw_out) {
    b200rt_internal_set_error("b200rt_render: null output");
    return B200RT_E_INVALID;
  }
  int rc = render_impl(sc, camera_index, opts, stats);
  if (rc) return rc;
  const b200rt_camera& c = sc->cameras[camera_index];
  CK(cudaMemcpy(rgbw_out, sc->d_accum, (size_t)c.width * c.height * sizeof(float4), cudaMemcpyDeviceToHost));
  return B200RT_OK;
} B200RT_CATCH

int b200rt_render_async(b200rt_scene* sc, int camera_index, const B200RenderOpts* opts, float* rgbw_out,
                        B200RenderStats* stats) try {
  if (!rgbw_out) {
    b200rt_internal_set_error("b200rt_render_async: null output");
    return B200RT_E_INVALID;
  }
  int rc = render_impl(sc, camera_index, opts, stats);
  if (rc) return rc;
  const b200rt_camera& c = sc->cameras[camera_index];
  return queue_host_copy(sc, rgbw_out, (size_t)c.width * c.height);
} B200RT_CATCH

int b200rt_render_wait(b200rt_scene* sc) try {
  if (!sc) return B200RT_E_INVALID;
  CK(cudaSetDevice(sc->device));
  for (int k = 0; k < 2; k++)
    if (sc->copy_pending[k]) {
      CK(cudaEventSynchronize(sc->ev_copied[k]));
      sc->copy_pending[k] = false;
    }
  return B200RT_OK;
} B200RT_CATCH

int b200rt_primary_ids(b200rt_scene* sc, int camera_index, const B200RenderOpts* opts, int32_t* ids_out,
                       float* t_out) try {
  if (!sc || !ids_out || !t_out || camera_index < 0 || camera_index >= (int)sc->cameras.size()) {
    b200rt_internal_set_error("b200rt_primary_ids: bad argument");
    return B200RT_E_INVALID;
  }
  CK(cudaSetDevice(sc->device));
  sc->S.cam = sc->cameras[camera_index];
  const b200rt_camera& c = sc->S.cam;
  if (c.samples_sqrt != 1) {
    b200rt_internal_set_error("b200rt_primary_ids: defined for 1-spp cameras only");
    return B200RT_E_UNSUPPORTED;
  }
  RenderParams R;
  if (resolve_rows(c, opts, &R) != B200RT_OK) return B200RT_E_INVALID;
  R.n_paths = (uint32_t)((size_t)R.n_rows * c.width);
  const size_t pixels = (size_t)c.width * c.height;
  int rc = ensure_pool(sc, sc->c0, R.n_paths + 1024, 1024);
  if (rc) return rc;
  cudaStream_t st = sc->c0.stream;
  DevBuf b_ids, b_t;
  CK(b_ids.alloc(pixels * 2 * sizeof(int32_t)));
  CK(b_t.alloc(pixels * sizeof(float)));
  int32_t* d_ids = b_ids.as<int32_t>();
  float* d_t = b_t.as<float>();
  CK(cudaMemsetAsync(d_ids, 0xff, pixels * 2 * sizeof(int32_t), st));
  {
    std::vector<float> inf(pixels, __builtin_inff());
    CK(cudaMemcpyAsync(d_t, inf.data(), pixels * sizeof(float), cudaMemcpyHostToDevice, st));
    CK(cudaStreamSynchronize(st));
  }
  CK(cudaMemsetAsync(sc->c0.d_counters, 0, sizeof(Counters), st));
  k_generate<<<sc->sm_count * 8, 256, 0, st>>>(sc->S, sc->c0.P, R, sc->c0.d_waves, sc->c0.d_counters);
  Timers tm = make_timers(sc->c0, false);
  int launches = 0, wr = 0;
  rc = run_waves(sc, sc->c0, opts && opts->exact_traversal, true, false, tm, &launches, &wr);
  if (rc) return rc;
  k_export_ids<<<sc->sm_count * 4, 256, 0, st>>>(sc->c0.P, R.n_paths, d_ids, d_t);
  CK(cudaMemcpyAsync(ids_out, d_ids, pixels * 2 * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(t_out, d_t, pixels * sizeof(float), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  return check_overflow(sc, "b200rt_primary_ids");
} B200RT_CATCH

int b200rt_trace_rays(b200rt_scene* sc, const float* rays, int64_t n, int32_t exact, int32_t* ids_out, float* t_out,
                      B200RenderStats* stats) try {
  if (!sc || !rays || n < 0 || !ids_out || !t_out || n > 0x7fffffffll) {
    b200rt_internal_set_error("b200rt_trace_rays: bad argument");
    return B200RT_E_INVALID;
  }
  if (n == 0) return B200RT_OK;
  CK(cudaSetDevice(sc->device));
  int rc = ensure_pool(sc, sc->c0, (size_t)n + 1024, 1024);
  if (rc) return rc;
  cudaStream_t st = sc->c0.stream;
  DevBuf b_rays, b_ids, b_t;
  CK(b_rays.alloc((size_t)n * 8 * sizeof(float)));
  CK(b_ids.alloc((size_t)n * 2 * sizeof(int32_t)));
  CK(b_t.alloc((size_t)n * sizeof(float)));
  float* d_rays = b_rays.as<float>();
  int32_t* d_ids = b_ids.as<int32_t>();
  float* d_t = b_t.as<float>();
  CK(cudaMemcpyAsync(d_rays, rays, (size_t)n * 8 * sizeof(float), cudaMemcpyHostToDevice, st));
  CK(cudaMemsetAsync(sc->c0.d_counters, 0, sizeof(Counters), st));
  k_import_rays<<<sc->sm_count * 4, 256, 0, st>>>(sc->c0.P, d_rays, (uint32_t)n, sc->c0.d_waves);
  Timers tm = make_timers(sc->c0, stats != nullptr);
  tm.mark(ST_GENERATE);
  int launches = 0, wr = 0;
  rc = run_waves(sc, sc->c0, exact != 0, true, false, tm, &launches, &wr);
  if (rc) return rc;
  k_export_ids<<<sc->sm_count * 4, 256, 0, st>>>(sc->c0.P, (uint32_t)n, d_ids, d_t);
  CK(cudaMemcpyAsync(ids_out, d_ids, (size_t)n * 2 * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(t_out, d_t, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  rc = check_overflow(sc, "b200rt_trace_rays");
  if (rc) return rc;
  if (stats) {
    Counters hc;
    CK(cudaMemcpy(&hc, sc->c0.d_counters, sizeof(hc), cudaMemcpyDeviceToHost));
    memset(stats, 0, sizeof(*stats));
    stats->nodes_visited = hc.nodes_visited; stats->tris_tested = hc.tris_tested;
    stats->spheres_tested = hc.spheres_tested; stats->instances_entered = hc.instances_entered;
    stats->rays_primary = (uint64_t)n;
    float ms[ST_COUNT];
    tm.collect(ms);
    stats->ms_extend = ms[ST_EXTEND];
    stats->launches = launches; stats->waves = 1;
    stats->node_bytes = (sc->S.use_wide && !exact && sc->S.semantics != B200RT_SEM_HW1) ? (int)sizeof(b200rt_wide::Node) : 64;
    stats->top_objects = exact ? 0 : sc->S.n_top;
  }
  return B200RT_OK;
} B200RT_CATCH

}  // extern "C"
#include "api_multi.inc"
