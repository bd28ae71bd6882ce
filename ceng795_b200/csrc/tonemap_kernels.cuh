// tonemap_kernels.cuh — SURVEY.md §8(f) rank 2: the step right after the hot path, on the device.
//   Photographic_tmo::apply_tmo      HW7/src/Photographic_tmo.cpp:10-49
//   sRGB curve + write_png's clamp   HW7/src/main.cpp:59-91, 105-115
// log-average luminance (block reduction in double; the reference sums floats sequentially), percentile
// white point (cub radix sort of the scaled luminances, a library call on a once-per-frame step), per-pixel curve.
#pragma once
#include "rt_common.cuh"

__global__ void k_tm_luminance(const float4* accum, uint32_t n, float* lum, double* log_sum) {
  __shared__ double part[8];
  double s = 0.0;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float4 a = accum[i];
    v3 c = mk3(0, 0, 0);
    if (a.w != 0.0f) c = mk3(a.x / a.w, a.y / a.w, a.z / a.w);   // Pixel::get_color
    const float l = (0.21f * c.x + 0.72f * c.y + 0.07f * c.z);
    lum[i] = l;
    s += (double)logf(l + 0.001f);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); w++) t += part[w];
    atomicAdd(log_sum, t);
  }
}
__global__ void k_tm_scale(uint32_t n, float key, const double* log_sum, const float* lum, float* scaled, float* sort_keys) {
  const float log_avg = expf((float)(*log_sum) / (float)n);
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float v = key * lum[i] / log_avg;
    scaled[i] = v;
    sort_keys[i] = v;
  }
}
__global__ void k_tm_apply(const float4* accum, uint32_t n, const float* lum, const float* scaled, const float* sorted,
                           uint32_t white_index, float saturation, unsigned char* rgb8) {
  const float white = sorted[white_index];
  const float inverse = 1.0f / 2.4f;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float4 a = accum[i];
    v3 c = mk3(0, 0, 0);
    if (a.w != 0.0f) c = mk3(a.x / a.w, a.y / a.w, a.z / a.w);
    const float l = scaled[i];
    const float ld = (l * (1 + l / (white * white))) / (l + 1);
    const float li = lum[i];
    const float ch[3] = {c.x, c.y, c.z};
#pragma unroll
    for (int k = 0; k < 3; k++) {
      const float t = powf(ch[k] / li, saturation) * ld;
      // clamp(0.0f, 1.0f, x) with the reference's argument order = min(1, max(0, x)); NaN -> 0
      const float g = fminf(1.0f, fmaxf(0.0f, 1.055f * powf(fminf(1.0f, fmaxf(0.0f, t)), inverse) - 0.055f)) * 255.0f;
      rgb8[3 * (size_t)i + k] = (unsigned char)fminf(255.0f, fmaxf(0.0f, g));
    }
  }
}
