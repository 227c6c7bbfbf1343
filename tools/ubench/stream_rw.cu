// HBM stream micro-benchmark: read-only, write-only and copy bandwidth of plain 128-bit grid-stride kernels on a
// buffer far larger than L2 (the write-dominated kernels of the generator -- up-sampling conv, noise conv -- all sit
// near 2 TB/s; this is what a pure store stream gets on the part).
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_write(float4* p, size_t n, float v) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = make_float4(v, v, v, v);
}
__global__ void k_read(const float4* p, size_t n, float* out) {
  float a = 0.f;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) { const float4 t = p[i]; a += t.x + t.y + t.z + t.w; }
  if (a == 12345.678f) *out = a;
}
__global__ void k_copy(const float4* p, float4* q, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) q[i] = p[i];
}
// block-contiguous variant: every CTA owns a contiguous 64 KB span per iteration (like a conv epilogue writing its tile)
__global__ void k_write_tile(float4* p, size_t n, float v) {
  const size_t tile = 4096;  // float4 per tile = 64 KB
  for (size_t t = blockIdx.x; t * tile < n; t += gridDim.x)
    for (size_t i = threadIdx.x; i < tile && t * tile + i < n; i += blockDim.x) p[t * tile + i] = make_float4(v, v, v, v);
}
int main() {
  const size_t bytes = 3ull << 30, n = bytes / 16;
  float4 *a, *b; float* o;
  cudaMalloc(&a, bytes); cudaMalloc(&b, bytes); cudaMalloc(&o, 4);
  cudaMemset(a, 0, bytes); cudaMemset(b, 0, bytes);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int blocks : {148 * 4, 148 * 8, 148 * 16}) {
    for (int mode = 0; mode < 4; mode++) {
      float best = 1e9f;
      for (int rep = 0; rep < 4; rep++) {
        cudaEventRecord(e0);
        if (mode == 0) k_write<<<blocks, 256>>>(a, n, 1.f + rep);
        if (mode == 1) k_read<<<blocks, 256>>>(a, n, o);
        if (mode == 2) k_copy<<<blocks, 256>>>(a, b, n);
        if (mode == 3) k_write_tile<<<blocks, 256>>>(a, n, 2.f + rep);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
      }
      const double gb = (mode == 2 ? 2.0 : 1.0) * bytes / 1e9;
      printf("blocks %5d  %-10s %7.3f ms  %7.1f GB/s\n", blocks, mode == 0 ? "write" : mode == 1 ? "read" : mode == 2 ? "copy" : "write-tile", best, gb / (best * 1e-3));
    }
  }
  return cudaGetLastError() != cudaSuccess;
}
