// multi_gpu.cuh — device side of the multi-GPU frame and PPM-count exchange (include/b200rt.h "multi-GPU").
//
// The reference's threads share ONE Pixel array (HW7/src/main.cpp:35-46: `j = i; j += thread_count`, every thread
// calls Pixel::add_color on the same array) and ONE Hit_point array (PPM/src/Scene.cpp:135-169).  Across GPUs the
// Pixel array lives on the owner GPU and is mapped into every other GPU's address space (peer access inside one
// process, cudaIpc* between processes); contributors write into it over NVLink from their own kernels:
//   1-sample cameras   k_combine / k_combine7 of wave 0 store Pixel{color, 1} in place (the kernel's `accum` argument
//                      simply points at the peer mapping) — the transfer IS the last kernel of the frame;
//   other cameras      k_frame_add: one 16-byte vector atomic per touched pixel of the locally accumulated frame.
// Hand-shake, all on the device: FrameTail sits right after the pixels.  `begun` = number of the frame the owner has
// opened (cleared, previous one collected); `arrived` = shares delivered since the frame was exported.  A share is
// delivered by a fence + system-scope atomic after the stores; the owner's stream waits for k * parts arrivals.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

struct FrameTail {
  unsigned int begun;     // frame number opened by the owner (release store, stream-ordered after the previous collect)
  unsigned int arrived;   // shares delivered, all frames (atomicAdd_system by contributors)
  unsigned int pad[62];
};

__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

// One thread: spins until *flag has reached `target` (wrap-safe); gives up after timeout_ns and raises *err so that a
// dead peer turns into an error code instead of a hung GPU.
__global__ void k_frame_wait(const unsigned int* flag, unsigned int target, unsigned int* err,
                             unsigned long long timeout_ns) {
  const unsigned long long t0 = global_ns();
  while ((int)(ld_acquire_sys(flag) - target) < 0) {
    if (global_ns() - t0 > timeout_ns) { *err = 1u; return; }
    __nanosleep(256);
  }
}
// Delivery of one share: everything this stream wrote before (previous kernels) is ordered before the bump.
__global__ void k_frame_signal(unsigned int* arrived) {
  __threadfence_system();
  atomicAdd_system(arrived, 1u);
}
__global__ void k_frame_open(unsigned int* begun, unsigned int frame_no) {
  __threadfence_system();
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(begun), "r"(frame_no) : "memory");
}
// Pixel::add_color of a locally accumulated frame into the shared one: pixels this GPU did not touch are skipped.
__global__ void __launch_bounds__(256) k_frame_add(const float4* __restrict__ local, float4* target, size_t pixels) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < pixels; i += (size_t)gridDim.x * blockDim.x) {
    const float4 v = local[i];
    if (v.x != 0.0f || v.y != 0.0f || v.z != 0.0f || v.w != 0.0f) atomicAdd(target + i, v);
  }
}

// ---- PPM: accepted-photon counts across GPUs -------------------------------------------------------------------
// A GPU that is share `part` of `parts` claims the indices n_base + part + j * parts (ppm_kernels.cuh: HitPoints), so its
// n holds n_base + part + parts * (photons it accepted since the last sync).
__global__ void k_ppm_counts_delta(const unsigned int* n, const unsigned int* n_base, unsigned int* delta, uint32_t n_hp,
                                   uint32_t part, uint32_t parts) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_hp; i += gridDim.x * blockDim.x)
    delta[i] = (n[i] - n_base[i] - part) / parts;
}
__global__ void k_ppm_counts_apply(unsigned int* n, unsigned int* n_base, const unsigned int* delta_sum, uint32_t n_hp,
                                   uint32_t part) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_hp; i += gridDim.x * blockDim.x) {
    const unsigned int v = n_base[i] + delta_sum[i];
    n_base[i] = v; n[i] = v + part;
  }
}
struct PeerPtrs {
  const void* p[16];
  int n;
};
// Runs on the owner: delta[i] += sum over the peers' delta[i] (P2P loads over NVLink), fixed device order.
__global__ void k_ppm_sum_peer_counts(unsigned int* mine, PeerPtrs peers, uint32_t n_hp) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_hp; i += gridDim.x * blockDim.x) {
    unsigned int s = mine[i];
    for (int k = 0; k < peers.n; k++) s += reinterpret_cast<const unsigned int*>(peers.p[k])[i];
    mine[i] = s;
  }
}
// Runs on the owner: out[i] = {sum of S.rgb over all devices, n} — the additive PPM state, fixed device order.
__global__ void k_ppm_sum_peer_state(const float4* mine, PeerPtrs peers, const unsigned int* n_base, float4* out, uint32_t n_hp) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_hp; i += gridDim.x * blockDim.x) {
    float4 s = mine[i];
    for (int k = 0; k < peers.n; k++) {
      const float4 v = reinterpret_cast<const float4*>(peers.p[k])[i];
      s.x += v.x; s.y += v.y; s.z += v.z;
    }
    s.w = (float)n_base[i];   // the counts as of the last synchronisation (= all of them after a pass)
    out[i] = s;
  }
}

// ---- PPM: canonical hit-point numbering ------------------------------------------------------------------------
// The eye pass appends hit points with warp-aggregated atomics, so their order differs from run to run and from GPU to
// GPU.  They are renumbered by the key (pixel, stratum, position of the vertex in the pixel's reflect / transmit tree)
// so that index i names the same hit point on every GPU (sums of S and of the counts are per index), and optionally
// grouped by home cell of the hash grid so that the counts a photon gathers while scanning a cell share cache lines.
// pos.w carries (stratum << 20 | tree code), written by k_ppm_eye_shade; wo.w the pixel.
__global__ void k_ppm_order_keys(const float4* wo, const float4* pos, uint32_t n_hp, unsigned long long* keys,
                                 uint32_t* idx) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_hp; i += gridDim.x * blockDim.x) {
    keys[i] = ((unsigned long long)__float_as_uint(wo[i].w) << 32) | (unsigned long long)__float_as_uint(pos[i].w);
    idx[i] = i;
  }
}
__global__ void k_ppm_gather4(const float4* a0, const float4* a1, const float4* a2, const float4* a3, const uint32_t* idx,
                              uint32_t n_hp, float4* out4) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_hp; i += gridDim.x * blockDim.x) {
    const uint32_t s = idx[i];
    out4[4 * (size_t)i + 0] = a0[s]; out4[4 * (size_t)i + 1] = a1[s];
    out4[4 * (size_t)i + 2] = a2[s]; out4[4 * (size_t)i + 3] = a3[s];
  }
}
__global__ void k_ppm_scatter4(const float4* in4, uint32_t n_hp, float4* a0, float4* a1, float4* a2, float4* a3) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_hp; i += gridDim.x * blockDim.x) {
    a0[i] = in4[4 * (size_t)i + 0]; a1[i] = in4[4 * (size_t)i + 1];
    a2[i] = in4[4 * (size_t)i + 2]; a3[i] = in4[4 * (size_t)i + 3];
  }
}
