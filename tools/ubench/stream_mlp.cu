// stream_mlp.cu -- what bounds the operand producers of the fused res-block conv (kernels_arb.cu)?
// A persistent kernel (one CTA per SM, ~200 KB of dynamic smem so that L1 is at its minimum like in the real kernel)
// whose loader threads read the fp32 activations with the producers' exact access pattern -- 64-channel half rows
// (256 B of every 512 B row), 8 lanes x 32 B per row, GP row passes per group -- keeping DEPTH groups in flight in
// registers, while (optionally) 8 writer warps store the bf16 output.  No math, no smem traffic: the GB/s it reaches is
// what the memory system gives this pattern at this memory-level parallelism.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/ubench/bin/stream_mlp tools/ubench/stream_mlp.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <cstdint>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

constexpr int C = 128;            // channels (fp32 row = 512 B)
constexpr int MT = 256;           // rows per tile

// FULLROW = 0: producers' pattern (chunk-major: 256-byte half rows, the other half one chunk later)
// FULLROW = 1: 16 lanes cover a whole 512-byte row (row-major)
template <int NW, int DEPTH, int GP, int FULLROW, int WRITE>
__global__ void __launch_bounds__(512, 1) k(const float4* __restrict__ x, uint32_t* __restrict__ out, long long rows, float* sink) {
  extern __shared__ uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long tiles = rows / MT;
  const long long t0 = tiles * blockIdx.x / gridDim.x, t1 = tiles * (blockIdx.x + 1) / gridDim.x;
  if (warp < NW) {
    const int pt = threadIdx.x;
    constexpr int NT = NW * 32;
    constexpr int LPR = FULLROW ? 16 : 8;          // lanes per row
    constexpr int RP = NT / LPR;                   // rows per pass
    constexpr int GR = GP * RP;                    // rows per group
    constexpr int NCH = FULLROW ? 1 : 2;
    const int ngc = (MT + GR - 1) / GR;
    const int cg = pt % LPR, rl = pt / LPR;
    const long long F = (t1 - t0) * NCH * ngc;
    float4 buf[DEPTH][GP][2];
    float acc = 0.f;
    long long li = 0;   // next group to issue
    auto issue = [&](float4 (&b)[GP][2]) {
      const long long tile = t0 + li / (NCH * ngc);
      const int rem = (int)(li % (NCH * ngc));
      const int c = rem / ngc, g = rem % ngc;
      const long long row = tile * MT + g * GR + rl;
      const float4* p = x + (row * C + c * 64 + cg * 8) / 4;
#pragma unroll
      for (int q = 0; q < GP; q++) {
        if (g * GR + rl + q * RP < MT) {
          b[q][0] = p[(size_t)q * RP * C / 4];
          b[q][1] = p[(size_t)q * RP * C / 4 + 1];
        }
      }
      li++;
    };
    auto consume = [&](float4 (&b)[GP][2]) {
#pragma unroll
      for (int q = 0; q < GP; q++) acc += (b[q][0].x + b[q][0].y) + (b[q][0].z + b[q][0].w) + (b[q][1].x + b[q][1].y) + (b[q][1].z + b[q][1].w);
    };
#pragma unroll
    for (int d = 0; d < DEPTH; d++)
#pragma unroll
      for (int q = 0; q < GP; q++) { buf[d][q][0] = make_float4(0, 0, 0, 0); buf[d][q][1] = buf[d][q][0]; }
#pragma unroll
    for (int d = 0; d < DEPTH - 1; d++) if (li < F) issue(buf[d]);
#pragma unroll 1
    for (long long f = 0; f < F; f += DEPTH) {
#pragma unroll
      for (int d = 0; d < DEPTH; d++) {
        if (f + d < F) {
          if (li < F) issue(buf[(d + DEPTH - 1) % DEPTH]);
          asm volatile("" ::: "memory");      // keep the loads ahead of the consumption of the oldest buffer
          consume(buf[d]);
          asm volatile("" ::: "memory");
        }
      }
    }
    if (acc == 1.2345f) sink[0] = acc;
  } else if (WRITE && warp >= 8) {
    // 8 writer warps: bf16 rows (256 B); a warp-wide store covers 128 contiguous bytes; warp w takes 32 rows of each tile
    const int w = warp - 8;
    for (long long t = t0; t < t1; t++) {
      uint32_t* o = out + (t * MT + w * 32) * (C / 2) + lane;
#pragma unroll 8
      for (int r = 0; r < 32; r++) { o[r * (C / 2)] = (uint32_t)r; o[r * (C / 2) + 32] = (uint32_t)lane; }
    }
  }
  (void)smem;
}

template <int NW, int DEPTH, int GP, int FULLROW, int WRITE>
void run(const float4* x, uint32_t* out, long long rows, float* sink, int nsm) {
  auto kern = k<NW, DEPTH, GP, FULLROW, WRITE>;
  const int smem = 200 * 1024;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  for (int i = 0; i < 2; i++) kern<<<nsm, 512, smem>>>(x, out, rows, sink);
  CK(cudaDeviceSynchronize());
  cudaEventRecord(a);
  const int reps = 5;
  for (int i = 0; i < reps; i++) kern<<<nsm, 512, smem>>>(x, out, rows, sink);
  cudaEventRecord(b);
  CK(cudaEventSynchronize(b));
  float ms = 0;
  cudaEventElapsedTime(&ms, a, b);
  ms /= reps;
  const double bytes = (double)rows * C * 4 + (WRITE ? (double)rows * C * 2 : 0.0);
  printf("loader warps %2d  depth %d  GP %d  bytes in flight/SM %6.1f KB  %s  %s : %.3f ms  %.0f GB/s\n", NW, DEPTH, GP,
         NW * 32 * GP * 32 * (DEPTH - 1 > 0 ? DEPTH - 1 : 1) / 1024.0, FULLROW ? "full rows" : "half rows", WRITE ? "read+write" : "read only ",
         ms, bytes / ms / 1e6);
}

int main() {
  const long long rows = 5746720 / MT * MT;
  float4* x; uint32_t* out; float* sink;
  CK(cudaMalloc(&x, (size_t)rows * C * 4));
  CK(cudaMalloc(&out, (size_t)rows * C * 2));
  CK(cudaMalloc(&sink, 4));
  CK(cudaMemset(x, 0, (size_t)rows * C * 4));
  int nsm = 0;
  CK(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0));
  //   NW DEPTH GP FULL WRITE
  run<6, 2, 4, 0, 1>(x, out, rows, sink, nsm);    // today's producers (one group in flight while the other is consumed)
  run<6, 3, 4, 0, 1>(x, out, rows, sink, nsm);
  run<6, 4, 4, 0, 1>(x, out, rows, sink, nsm);
  run<8, 2, 4, 0, 1>(x, out, rows, sink, nsm);
  run<8, 3, 4, 0, 1>(x, out, rows, sink, nsm);
  run<8, 4, 4, 0, 1>(x, out, rows, sink, nsm);
  run<8, 3, 2, 0, 1>(x, out, rows, sink, nsm);
  run<8, 5, 2, 0, 1>(x, out, rows, sink, nsm);
  run<6, 2, 4, 1, 1>(x, out, rows, sink, nsm);
  run<6, 3, 4, 1, 1>(x, out, rows, sink, nsm);
  run<8, 3, 4, 1, 1>(x, out, rows, sink, nsm);
  run<6, 2, 4, 0, 0>(x, out, rows, sink, nsm);
  run<6, 3, 4, 0, 0>(x, out, rows, sink, nsm);
  run<6, 4, 4, 0, 0>(x, out, rows, sink, nsm);
  run<8, 4, 4, 0, 0>(x, out, rows, sink, nsm);
  run<8, 4, 4, 1, 0>(x, out, rows, sink, nsm);
  return 0;
}
