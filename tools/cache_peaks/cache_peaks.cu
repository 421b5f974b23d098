// Development aid, not part of librt_b200.so: measures the on-chip bandwidths the traversal kernel lives on, as
// denominators for bench.py's roofline (SURVEY.md §8d asks for an L2 figure measured on the box):
//   L1  every block streams its own 16 KB window with 16-byte loads, over and over (L1-resident after one pass)
//   L2  all blocks stream one 32 MB buffer with 16-byte loads, each from its own starting point (bigger than all
//       L1s together, inside the 126 MB L2)
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -shared -Xcompiler -fPIC -o libcache_peaks.so cache_peaks.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__global__ void k_read(const float4 *__restrict__ buf, size_t window_f4, size_t windows, int repeats, float4 *sink) {
  // block b reads window (b % windows); thread-strided 16-byte loads, 8 independent loads in flight per thread
  const float4 *w = buf + (size_t)(blockIdx.x % windows) * window_f4;
  // one shared window (the L2 case): every block starts somewhere else, so that the blocks of one SM are
  // megabytes apart and nothing is served by L1
  const size_t sweep = 8 * (size_t)blockDim.x;
  const size_t start = windows == 1 ? ((size_t)blockIdx.x * (window_f4 / gridDim.x) / sweep) * sweep : 0;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int r = 0; r < repeats; r++) {
    for (size_t i0 = threadIdx.x; i0 + 7 * (size_t)blockDim.x < window_f4; i0 += sweep) {
      size_t i = i0 + start;
      if (i + 7 * (size_t)blockDim.x >= window_f4)
        i -= (window_f4 / sweep) * sweep;
#pragma unroll
      for (int k = 0; k < 8; k++) {
        float4 v = __ldg(w + i + (size_t)k * blockDim.x);
        acc.x += v.x;
        acc.y += v.y;
        acc.z += v.z;
        acc.w += v.w;
      }
    }
  }
  if (acc.x == 123.456f) // keeps the loads alive
    sink[0] = acc;
}

extern "C" int cache_peaks(double *l1_gbs, double *l2_gbs, int *sm_count) {
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, 0) != cudaSuccess)
    return 1;
  *sm_count = prop.multiProcessorCount;
  const size_t bytes = 32u << 20;
  float4 *buf = nullptr, *sink = nullptr;
  if (cudaMalloc(&buf, bytes) != cudaSuccess || cudaMalloc(&sink, 64) != cudaSuccess)
    return 2;
  cudaMemset(buf, 0, bytes);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int threads = 128, blocks = prop.multiProcessorCount * 8; // 8 x 16 KB windows per SM stay in L1
  double best[2] = {0.0, 0.0};
  for (int mode = 0; mode < 2; mode++) {
    // L1: 16 KB per block, one window per block.  L2: the whole buffer is one window shared by all blocks.
    size_t window_f4 = mode == 0 ? (16u << 10) / 16 : bytes / 16;
    size_t windows = mode == 0 ? bytes / (16u << 10) : 1;
    int repeats = mode == 0 ? 2000 : 4;
    for (int trial = 0; trial < 5; trial++) {
      k_read<<<blocks, threads>>>(buf, window_f4, windows, repeats, sink); // warm
      cudaEventRecord(e0);
      k_read<<<blocks, threads>>>(buf, window_f4, windows, repeats, sink);
      cudaEventRecord(e1);
      if (cudaEventSynchronize(e1) != cudaSuccess)
        return 3;
      float ms = 0.f;
      cudaEventElapsedTime(&ms, e0, e1);
      size_t per_pass = (window_f4 / (8 * (size_t)threads)) * 8 * threads * 16; // bytes one block reads per repeat
      double gbs = (double)per_pass * repeats * blocks / (ms * 1e-3) / 1e9;
      if (gbs > best[mode])
        best[mode] = gbs;
    }
  }
  *l1_gbs = best[0];
  *l2_gbs = best[1];
  cudaFree(buf);
  cudaFree(sink);
  return 0;
}
