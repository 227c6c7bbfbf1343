// Micro-benchmark: clock64 rate vs globaltimer, FFMA vs FFMA2 throughput per SM (8 warps/SM like the LSTM kernel,
// 16 independent accumulator chains per thread), LDS.128 broadcast throughput.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fma_rate fma_rate.cu && ./fma_rate
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }

template <int MODE>
__global__ void __launch_bounds__(256, 1) k(float* out, long long* res, int iters) {
  __shared__ float4 sh[1024];
  for (int i = threadIdx.x; i < 1024; i += 256) sh[i] = make_float4(1e-3f * i, 1.f, 2.f, 3.f);
  __syncthreads();
  float2 acc[16];
  for (int i = 0; i < 16; i++) acc[i] = make_float2(threadIdx.x * 1e-3f, i);
  float2 w = make_float2(1.0001f, 0.9999f), h = make_float2(1e-4f, 2e-4f);
  const int q = (threadIdx.x & 31) >> 3;
  const unsigned long long g0 = gtime();
  const long long c0 = clock64();
  for (int it = 0; it < iters; it++) {
    if (MODE == 0) {          // 32 FFMA2 per iteration
#pragma unroll
      for (int r = 0; r < 2; r++)
#pragma unroll
        for (int i = 0; i < 16; i++) acc[i] = __ffma2_rn(w, h, acc[i]);
    } else if (MODE == 1) {   // 64 FFMA per iteration (same flops)
#pragma unroll
      for (int r = 0; r < 2; r++)
#pragma unroll
        for (int i = 0; i < 16; i++) { acc[i].x = fmaf(w.x, h.x, acc[i].x); acc[i].y = fmaf(w.y, h.y, acc[i].y); }
    } else {                  // LSTM-like: 8 broadcast LDS.128 + 32 FFMA2
#pragma unroll
      for (int g = 0; g < 8; g++) {
        const float4 v = sh[((it & 15) * 8 + g) * 4 + q + (it >> 10)];
        acc[2 * g] = __ffma2_rn(w, make_float2(v.x, v.y), acc[2 * g]);
        acc[2 * g + 1] = __ffma2_rn(w, make_float2(v.x, v.y), acc[2 * g + 1]);
        acc[2 * g] = __ffma2_rn(w, make_float2(v.z, v.w), acc[2 * g]);
        acc[2 * g + 1] = __ffma2_rn(w, make_float2(v.z, v.w), acc[2 * g + 1]);
      }
    }
  }
  const long long c1 = clock64();
  const unsigned long long g1 = gtime();
  float s = 0;
  for (int i = 0; i < 16; i++) s += acc[i].x + acc[i].y;
  out[blockIdx.x * 256 + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) { res[0] = c1 - c0; res[1] = (long long)(g1 - g0); }
}

int main() {
  float* out; long long* res;
  cudaMalloc(&out, 148 * 256 * 4); cudaMallocManaged(&res, 16);
  const int iters = 200000;
  const char* names[3] = {"FFMA2 x32/iter", "FFMA x64/iter", "8 LDS.128 + 32 FFMA2 / iter"};
  for (int m = 0; m < 3; m++) {
    for (int rep = 0; rep < 2; rep++) {
      if (m == 0) k<0><<<148, 256>>>(out, res, iters);
      if (m == 1) k<1><<<148, 256>>>(out, res, iters);
      if (m == 2) k<2><<<148, 256>>>(out, res, iters);
      cudaDeviceSynchronize();
    }
    const double clk = (double)res[0], ns = (double)res[1];
    printf("%-30s clock64 %.0f  globaltimer %.0f ns  -> clock64 rate %.3f GHz; %.2f clock64/iter, %.2f ns/iter; "
           "FMA/clk64/SM = %.1f, FMA/ns/SM = %.1f\n", names[m], clk, ns, clk / ns, clk / iters, ns / iters,
           256.0 * 64 / (clk / iters), 256.0 * 64 / (ns / iters));
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
