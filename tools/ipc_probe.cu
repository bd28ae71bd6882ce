// ipc_probe — can two PROCESSES on one box share device memory (cudaIpc*) and what do peer stores / atomics cost?
//   ipc_probe owner <dir>   : device 0 allocates, exports a handle to <dir>/handle, waits for <dir>/done, checks
//   ipc_probe peer  <dir>   : device 1 opens the handle, stores + CAS/add atomics into it over NVLink, times them
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -o build/ipc_probe tools/ipc_probe.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <unistd.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

__global__ void k_store(float4* dst, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    dst[i] = make_float4((float)i, 1.f, 2.f, 3.f);
}
__global__ void k_cas(unsigned* cnt, size_t n_slots, size_t n_ops, unsigned seed) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n_ops; i += (size_t)gridDim.x * blockDim.x) {
    unsigned h = (unsigned)i * 2654435761u + seed; h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
    unsigned* p = cnt + (h % n_slots);
    unsigned old = *(volatile unsigned*)p;
    for (;;) { const unsigned got = atomicCAS(p, old, old + 1u); if (got == old) break; old = got; }
  }
}
__global__ void k_add4(float4* acc, size_t n_slots, size_t n_ops, unsigned seed) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n_ops; i += (size_t)gridDim.x * blockDim.x) {
    unsigned h = (unsigned)i * 2654435761u + seed; h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
    atomicAdd(acc + (h % n_slots), make_float4(1.f, 1.f, 1.f, 1.f));
  }
}

static float timed(cudaStream_t s, void (*f)(void*), void* a) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(a); cudaStreamSynchronize(s);
  cudaEventRecord(e0, s); f(a); cudaEventRecord(e1, s); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}

int main(int argc, char** argv) {
  if (argc < 3) return 1;
  const std::string mode = argv[1], dir = argv[2];
  const size_t n_slots = 2u << 20, n_px = 1920 * 1080;
  if (mode == "owner") {
    CK(cudaSetDevice(0));
    char* buf; CK(cudaMalloc(&buf, n_slots * 4 + n_slots * 16 + n_px * 16));
    CK(cudaMemset(buf, 0, n_slots * 4 + n_slots * 16 + n_px * 16));
    cudaIpcMemHandle_t h; CK(cudaIpcGetMemHandle(&h, buf));
    FILE* f = fopen((dir + "/handle.tmp").c_str(), "wb"); fwrite(&h, sizeof h, 1, f); fclose(f);
    rename((dir + "/handle.tmp").c_str(), (dir + "/handle").c_str());
    for (int i = 0; i < 600 && access((dir + "/done").c_str(), F_OK) != 0; i++) usleep(100000);
    unsigned* host = (unsigned*)malloc(n_slots * 4);
    CK(cudaMemcpy(host, buf, n_slots * 4, cudaMemcpyDeviceToHost));
    unsigned long long sum = 0; for (size_t i = 0; i < n_slots; i++) sum += host[i];
    float4 px; CK(cudaMemcpy(&px, buf + n_slots * 20 + 16 * 12345, 16, cudaMemcpyDeviceToHost));
    printf("owner: sum of counters %llu, pixel 12345 = %g %g %g %g\n", sum, px.x, px.y, px.z, px.w);
    return 0;
  }
  int ndev = 0; CK(cudaGetDeviceCount(&ndev));
  const int dev = ndev > 1 ? 1 : 0;
  CK(cudaSetDevice(dev));
  for (int i = 0; i < 600 && access((dir + "/handle").c_str(), F_OK) != 0; i++) usleep(100000);
  cudaIpcMemHandle_t h; FILE* f = fopen((dir + "/handle").c_str(), "rb"); if (!f) { printf("no handle\n"); return 3; }
  fread(&h, sizeof h, 1, f); fclose(f);
  char* peer = nullptr;
  cudaError_t e = cudaIpcOpenMemHandle((void**)&peer, h, cudaIpcMemLazyEnablePeerAccess);
  printf("peer (device %d of %d): cudaIpcOpenMemHandle -> %s\n", dev, ndev, cudaGetErrorString(e));
  if (e != cudaSuccess) return 4;
  char* local; CK(cudaMalloc(&local, n_slots * 4 + n_slots * 16 + n_px * 16)); CK(cudaMemset(local, 0, n_slots * 20 + n_px * 16));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int where = 0; where < 2; where++) {
    char* b = where ? peer : local;
    const char* name = where ? "peer " : "local";
    for (size_t n_ops : {size_t(1) << 20, size_t(8) << 20}) {
      float ms;
      k_cas<<<148 * 8, 256>>>((unsigned*)b, n_slots, n_ops, 1u); CK(cudaDeviceSynchronize());
      cudaEventRecord(e0); k_cas<<<148 * 8, 256>>>((unsigned*)b, n_slots, n_ops, 2u); cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
      cudaEventElapsedTime(&ms, e0, e1);
      printf("%s read+CAS     %8zu ops: %.3f ms  (%.0f M/s)\n", name, n_ops, ms, n_ops / ms / 1e3);
      k_add4<<<148 * 8, 256>>>((float4*)(b + n_slots * 4), n_slots, n_ops, 1u); CK(cudaDeviceSynchronize());
      cudaEventRecord(e0); k_add4<<<148 * 8, 256>>>((float4*)(b + n_slots * 4), n_slots, n_ops, 2u); cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
      cudaEventElapsedTime(&ms, e0, e1);
      printf("%s float4 add   %8zu ops: %.3f ms  (%.0f M/s)\n", name, n_ops, ms, n_ops / ms / 1e3);
    }
    float ms;
    k_store<<<148 * 8, 256>>>((float4*)(b + n_slots * 20), n_px); CK(cudaDeviceSynchronize());
    cudaEventRecord(e0); k_store<<<148 * 8, 256>>>((float4*)(b + n_slots * 20), n_px); cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
    cudaEventElapsedTime(&ms, e0, e1);
    printf("%s float4 store %8zu px: %.3f ms  (%.1f GB/s)\n", name, n_px, ms, n_px * 16 / ms / 1e6);
    cudaEventRecord(e0); CK(cudaMemcpyAsync(b + n_slots * 20, local + n_slots * 20, n_px * 16 / 8, cudaMemcpyDefault)); cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
    cudaEventElapsedTime(&ms, e0, e1);
    printf("%s memcpy 1/8 frame (%zu B): %.3f ms\n", name, n_px * 2, ms);
  }
  CK(cudaIpcCloseMemHandle(peer));
  f = fopen((dir + "/done").c_str(), "wb"); fclose(f);
  return 0;
}
