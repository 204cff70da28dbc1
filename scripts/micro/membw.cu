// Micro-benchmark (not product code): what launch shape streams HBM fastest on B200?
// Reads (and optionally writes) a bf16 tensor with 16-byte accesses; prints GB/s per configuration.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

template <int U, bool WRITE>
__global__ void stream_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, size_t n, float* sink) {
  float acc = 0.f;
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  for (; i + (U - 1) * stride < n; i += U * stride) {
    uint4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) v[u] = __ldg(in + i + u * stride);
#pragma unroll
    for (int u = 0; u < U; ++u) {
      acc += __uint_as_float(v[u].x) + __uint_as_float(v[u].w);
      if (WRITE) out[i + u * stride] = v[u];
    }
  }
  for (; i < n; i += stride) {
    uint4 v = __ldg(in + i);
    acc += __uint_as_float(v.x);
    if (WRITE) out[i] = v;
  }
  if (acc == 123.456f) *sink = acc;
}

template <int U, bool WRITE>
float run(const uint4* in, uint4* out, size_t n, float* sink, int blocks, int threads) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  stream_kernel<U, WRITE><<<blocks, threads>>>(in, out, n, sink);
  cudaEventRecord(e0);
  for (int r = 0; r < 5; ++r) stream_kernel<U, WRITE><<<blocks, threads>>>(in, out, n, sink);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  return ms / 5;
}

int main() {
  const size_t bytes = 1ull << 30;   // 1 GiB per tensor (>> L2)
  const size_t n = bytes / 16;
  uint4 *in, *out; float* sink;
  cudaMalloc(&in, bytes); cudaMalloc(&out, bytes); cudaMalloc(&sink, 4);
  cudaMemset(in, 1, bytes);
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  printf("SMs %d\n", sms);
  for (int threads : {256, 512, 1024}) {
    for (int bps : {1, 2, 4, 8, 16, 32}) {
      if (threads * bps > 2048 && bps > 2048 / threads * 8) continue;
      const int blocks = sms * bps;
      float r2 = run<2, false>(in, out, n, sink, blocks, threads);
      float r4 = run<4, false>(in, out, n, sink, blocks, threads);
      float r8 = run<8, false>(in, out, n, sink, blocks, threads);
      float w4 = run<4, true>(in, out, n, sink, blocks, threads);
      printf("threads %4d blocks/SM %2d : read U2 %6.0f U4 %6.0f U8 %6.0f GB/s | copy U4 %6.0f GB/s (r+w)\n", threads, bps,
             bytes / r2 / 1e6, bytes / r4 / 1e6, bytes / r8 / 1e6, 2.0 * bytes / w4 / 1e6);
    }
  }
  return 0;
}
