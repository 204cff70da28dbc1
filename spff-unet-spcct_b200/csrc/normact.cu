// Bandwidth-bound kernels around the convolutions: InstanceNorm3d statistics and apply, LeakyReLU,
// the collapsed SPFF tail  out = lrelu(IN(x)) * P[n,d,c] + Q[n,d,c]  (SURVEY.md §7.3), the fused
// (1,2,2) max-pool, and the matching backward passes. They replace nn.InstanceNorm3d / LeakyReLU
// (reference innovative3D/models.py:168-181), the elementwise halves of EnergyFiLM3D / FourierGate3D /
// _SpectralSE / _SEChannelLite (models.py:1505-1512, 1527-1544, 600-614) and nn.MaxPool3d((1,2,2))
// (models.py:658-665).
//
// Access pattern: activations are position-major bf16 [positions][ld]; a thread owns one 16-byte
// vector (8 channels) of a position, so a warp reads/writes 512 contiguous bytes per instruction and
// its per-channel coefficients (A, B, P, Q ...) stay in registers for a whole (sample, plane) chunk.
// Reductions over positions are per-thread fp32 partials -> shared-memory tree across the threads
// that own the same channels -> one atomic per channel per block.
#include "common.h"

#include <cuda_bf16.h>

namespace spff {
namespace {

constexpr int kBlock = 256;

struct PlaneGrid {
  int c8;     // 16-byte vectors per position (C / 8)
  int rpi;    // positions ("rows") per block iteration = kBlock / c8
  int hw;     // positions per plane
  int chunk;  // positions per block
};

__device__ __forceinline__ void unpack8(const uint4& v, float (&f)[8]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 t = __bfloat1622float2(h[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint4 v;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&v);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  return v;
}
__device__ __forceinline__ uint4 ldg16(const __nv_bfloat16* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }
__device__ __forceinline__ void stg16(__nv_bfloat16* p, const uint4& v) { *reinterpret_cast<uint4*>(p) = v; }

// Sum `NV` per-thread values over the threads of the block that own the same channel vector
// (same tid % c8); the result is valid in the threads with tid < c8. `red` holds NV*kBlock floats.
template <int NV>
__device__ __forceinline__ void reduce_rows(float (&v)[NV], float* red, int c8, int rpi) {
  const int tid = threadIdx.x;
#pragma unroll
  for (int i = 0; i < NV; ++i) red[i * kBlock + tid] = v[i];
  __syncthreads();
  if (tid < c8) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      float s = 0.f;
      for (int r = 0; r < rpi; ++r) s += red[i * kBlock + r * c8 + tid];
      v[i] = s;
    }
  }
  __syncthreads();
}


// Fast path of the same reduction for c8 in {1,2,4,...,32} (every in-scope channel count): rows that
// share a warp are folded with shuffles, the 8 warps meet in shared memory, and thread t of the
// first L*NV threads ends up owning output element t (vector t / NV... laid out [vec][NV]) so that
// the global atomics that follow are contiguous. `red` holds 8 * 32 * NV floats.
// Returns true when this thread holds a valid result in `out` for output index `idx`
// (idx = vec * NV + i); blocks with more than 256 outputs loop through `round`.
template <int NV>
__device__ __forceinline__ void reduce_rows_fast(float (&v)[NV], float* red, int c8) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int L = c8 < 32 ? c8 : 32;
  for (int off = 16; off >= c8; off >>= 1) {
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] += __shfl_xor_sync(0xffffffffu, v[i], off);
  }
  if (lane < L) {
#pragma unroll
    for (int i = 0; i < NV; ++i) red[(warp * L + lane) * NV + i] = v[i];
  }
  __syncthreads();
}
// sum over the 8 warps of output element idx (vec = idx / NV, i = idx % NV)
template <int NV>
__device__ __forceinline__ float reduce_rows_fetch(const float* red, int c8, int idx) {
  const int L = c8 < 32 ? c8 : 32;
  float t = 0.f;
#pragma unroll
  for (int w = 0; w < kBlock / 32; ++w) t += red[w * L * NV + idx];
  return t;
}
__device__ __forceinline__ bool fast_reduce_ok(int c8) { return c8 <= 32 && (c8 & (c8 - 1)) == 0; }
inline bool fast_reduce_ok_host(int c8) { return c8 <= 32 && (c8 & (c8 - 1)) == 0; }

// ---------------------------------------------------------------------------------------------
// InstanceNorm statistics: stats[n][c] += {sum x, sum x^2} over the block's positions (double).
// grid = (chunks, d, n)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kBlock, 5) in_stats_kernel(const __nv_bfloat16* __restrict__ x, long long ld,
                                                          PlaneGrid g, int d, double* __restrict__ stats, int c) {
  extern __shared__ float red[];
  const int n = blockIdx.z, dd = blockIdx.y;
  const int v = threadIdx.x % g.c8, r = threadIdx.x / g.c8;
  const long long base = (static_cast<long long>(n) * d + dd) * g.hw;
  const int p0 = blockIdx.x * g.chunk;
  const int p1 = min(g.hw, p0 + g.chunk);
  float s[8], q[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) s[i] = q[i] = 0.f;
  if (r < g.rpi) {
    auto body = [&](const uint4& raw) {
      float f[8];
      unpack8(raw, f);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        s[i] += f[i];
        q[i] = fmaf(f[i], f[i], q[i]);
      }
    };
    constexpr int U = 4;   // independent 16-byte loads in flight per thread
    int cnt = (p0 + r < p1) ? (p1 - p0 - r + g.rpi - 1) / g.rpi : 0;
    const long long sx = static_cast<long long>(g.rpi) * ld;
    const __nv_bfloat16* px = x + (base + p0 + r) * ld + v * 8;
    for (; cnt >= U; cnt -= U) {
      uint4 raw[U];
#pragma unroll
      for (int u = 0; u < U; ++u) raw[u] = ldg16(px + u * sx);
      px += U * sx;
#pragma unroll
      for (int u = 0; u < U; ++u) body(raw[u]);
    }
    for (; cnt > 0; --cnt) {
      body(ldg16(px));
      px += sx;
    }
  }
  float sq[16];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    sq[2 * i] = s[i];          // interleaved {sum, sumsq} = the layout of stats[n][c][2]
    sq[2 * i + 1] = q[i];
  }
  if (fast_reduce_ok(g.c8)) {
    reduce_rows_fast<16>(sq, red, g.c8);
    const int L = g.c8 < 32 ? g.c8 : 32;
    for (int idx = threadIdx.x; idx < L * 16; idx += kBlock)
      atomicAdd(stats + static_cast<long long>(n) * c * 2 + idx, static_cast<double>(reduce_rows_fetch<16>(red, g.c8, idx)));
    return;
  }
  reduce_rows<16>(sq, red, g.c8, g.rpi);
  if (threadIdx.x < g.c8) {
    double* dst = stats + (static_cast<long long>(n) * c + v * 8) * 2;
#pragma unroll
    for (int i = 0; i < 16; ++i) atomicAdd(dst + i, static_cast<double>(sq[i]));
  }
}

// coef[n][c] = {A, B, mean, rstd}: IN(x) = x*A + B with A = rstd*gamma, B = beta - mean*A.
// batch_stats != 0 (BatchNorm3d training mode): statistics are summed over the samples first.
__global__ void in_coeffs_kernel(const double* __restrict__ stats, const float* __restrict__ gamma,
                                 const float* __restrict__ beta, float eps, int n, int c, double count,
                                 int batch_stats, float* __restrict__ coef) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * c) return;
  const int ch = i % c;
  double s = 0, q = 0, cnt = count;
  if (batch_stats) {
    for (int k = 0; k < n; ++k) {
      s += stats[(static_cast<long long>(k) * c + ch) * 2];
      q += stats[(static_cast<long long>(k) * c + ch) * 2 + 1];
    }
    cnt = count * n;
  } else {
    s = stats[static_cast<long long>(i) * 2];
    q = stats[static_cast<long long>(i) * 2 + 1];
  }
  const double mean = s / cnt;
  double var = q / cnt - mean * mean;
  if (var < 0) var = 0;
  const float rstd = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
  const float ga = gamma ? gamma[ch] : 1.f;
  const float be = beta ? beta[ch] : 0.f;
  const float A = rstd * ga;
  float4 o;
  o.x = A;
  o.y = be - static_cast<float>(mean) * A;
  o.z = static_cast<float>(mean);
  o.w = rstd;
  reinterpret_cast<float4*>(coef)[i] = o;
}

// coef from the per-item partial statistics the conv epilogue wrote: partial[n][slots][2][c].
// Fixed summation order (double) -> bit-reproducible coefficients.
__global__ void in_coeffs_partial_kernel(const float* __restrict__ partial, int slots, const float* __restrict__ gamma,
                                         const float* __restrict__ beta, float eps, int n, int c, double count,
                                         float* __restrict__ coef) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * c) return;
  const int s = i / c, ch = i % c;
  const float* base = partial + static_cast<size_t>(s) * slots * 2 * c + ch;
  double sum = 0, sq = 0;
  for (int k = 0; k < slots; ++k) {
    sum += base[static_cast<size_t>(k) * 2 * c];
    sq += base[static_cast<size_t>(k) * 2 * c + c];
  }
  const double mean = sum / count;
  double var = sq / count - mean * mean;
  if (var < 0) var = 0;
  const float rstd = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
  const float A = rstd * (gamma ? gamma[ch] : 1.f);
  float4 o;
  o.x = A;
  o.y = (beta ? beta[ch] : 0.f) - static_cast<float>(mean) * A;
  o.z = static_cast<float>(mean);
  o.w = rstd;
  reinterpret_cast<float4*>(coef)[i] = o;
}

__device__ __forceinline__ void load_ab(const float* __restrict__ coef, int n, int c, int v, float (&A)[8],
                                        float (&B)[8]) {
  const float4* cf = reinterpret_cast<const float4*>(coef) + static_cast<long long>(n) * c + v * 8;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float4 t = __ldg(cf + i);
    A[i] = t.x;
    B[i] = t.y;
  }
}
__device__ __forceinline__ void load8(const float* __restrict__ p, float (&o)[8]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p));
  const float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.w;
  o[4] = b.x; o[5] = b.y; o[6] = b.z; o[7] = b.w;
}

// ---------------------------------------------------------------------------------------------
// y = lrelu(x*A+B) [* P + Q]   and/or   S[n][d][c] += sum_hw lrelu(x*A+B).   grid = (chunks, d, n)
// ---------------------------------------------------------------------------------------------
template <bool WRITE, bool REDUCE, bool AFFINE>
__global__ void __launch_bounds__(kBlock)
norm_act_kernel(const __nv_bfloat16* __restrict__ x, long long ldx, const float* __restrict__ coef,
                const float* __restrict__ P, const float* __restrict__ Q, __nv_bfloat16* __restrict__ y,
                long long ldy, float* __restrict__ S, PlaneGrid g, int d, int c, float slope,
                float* __restrict__ part = nullptr /* REDUCE: [plane][chunk][c] block partials instead of atomics */) {
  extern __shared__ float red[];
  const int n = blockIdx.z, dd = blockIdx.y;
  const int v = threadIdx.x % g.c8, r = threadIdx.x / g.c8;
  const long long base = (static_cast<long long>(n) * d + dd) * g.hw;
  const int p0 = blockIdx.x * g.chunk;
  const int p1 = min(g.hw, p0 + g.chunk);
  float A[8], B[8], Pv[8], Qv[8], acc[8];
  load_ab(coef, n, c, v, A, B);
  if (AFFINE) {
    const long long pq = (static_cast<long long>(n) * d + dd) * c + v * 8;
    load8(P + pq, Pv);
    load8(Q + pq, Qv);
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  if (r < g.rpi) {
    constexpr int U = 4;   // independent 16-byte loads in flight per thread
    for (int p = p0 + r; p < p1; p += U * g.rpi) {
      uint4 raw[U];
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (p + u * g.rpi < p1) raw[u] = ldg16(x + (base + p + u * g.rpi) * ldx + v * 8);
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (p + u * g.rpi < p1) {
          float f[8];
          unpack8(raw[u], f);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            float z = fmaf(f[i], A[i], B[i]);
            float a = z > 0.f ? z : z * slope;
            if (REDUCE) acc[i] += a;
            f[i] = AFFINE ? fmaf(a, Pv[i], Qv[i]) : a;
          }
          if (WRITE) stg16(y + (base + p + u * g.rpi) * ldy + v * 8, pack8(f));
        }
      }
    }
  }
  if (REDUCE) {
    if (fast_reduce_ok(g.c8)) {
      reduce_rows_fast<8>(acc, red, g.c8);
      const int L = g.c8 < 32 ? g.c8 : 32;
      if (threadIdx.x < L * 8) {
        const float t = reduce_rows_fetch<8>(red, g.c8, threadIdx.x);
        const long long plane = static_cast<long long>(n) * d + dd;
        if (part)
          part[(plane * gridDim.x + blockIdx.x) * c + threadIdx.x] = t;
        else
          atomicAdd(S + plane * c + threadIdx.x, t);
      }
      return;
    }
    reduce_rows<8>(acc, red, g.c8, g.rpi);
    if (threadIdx.x < g.c8) {
      float* dst = S + (static_cast<long long>(n) * d + dd) * c + v * 8;
#pragma unroll
      for (int i = 0; i < 8; ++i) atomicAdd(dst + i, acc[i]);
    }
  }
}

// S[n][d][c] = sum_hw lrelu(x*A+B) alone (the gate statistic pass): the loop shape of in_stats_kernel — strided
// pointers, U independent loads issued before the first use, a tail of single rows — which the generic kernel above
// does not get from the compiler for its read-only variant (5.1 TB/s; this one is measured in DESIGN.md).
__global__ void __launch_bounds__(kBlock, 4)
norm_act_sum_kernel(const __nv_bfloat16* __restrict__ x, long long ldx, const float* __restrict__ coef, float* __restrict__ S,
                    PlaneGrid g, int d, int c, float slope, float* __restrict__ part) {
  extern __shared__ float red[];
  const int n = blockIdx.z, dd = blockIdx.y;
  const int v = threadIdx.x % g.c8, r = threadIdx.x / g.c8;
  const long long base = (static_cast<long long>(n) * d + dd) * g.hw;
  const int p0 = blockIdx.x * g.chunk;
  const int p1 = min(g.hw, p0 + g.chunk);
  float A[8], B[8], acc[8];
  load_ab(coef, n, c, v, A, B);
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  if (r < g.rpi) {
    auto body = [&](const uint4& raw) {
      float f[8];
      unpack8(raw, f);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float z = fmaf(f[i], A[i], B[i]);
        acc[i] += z > 0.f ? z : z * slope;
      }
    };
    constexpr int U = 4;
    int cnt = (p0 + r < p1) ? (p1 - p0 - r + g.rpi - 1) / g.rpi : 0;
    const long long sx = static_cast<long long>(g.rpi) * ldx;
    const __nv_bfloat16* px = x + (base + p0 + r) * ldx + v * 8;
    for (; cnt >= U; cnt -= U) {
      uint4 raw[U];
#pragma unroll
      for (int u = 0; u < U; ++u) raw[u] = ldg16(px + u * sx);
      px += U * sx;
#pragma unroll
      for (int u = 0; u < U; ++u) body(raw[u]);
    }
    for (; cnt > 0; --cnt) {
      body(ldg16(px));
      px += sx;
    }
  }
  reduce_rows_fast<8>(acc, red, g.c8);     // c8 is a power of two <= 32 here (host dispatch)
  const int L = g.c8 < 32 ? g.c8 : 32;
  if (threadIdx.x < L * 8) {
    const float t = reduce_rows_fetch<8>(red, g.c8, threadIdx.x);
    const long long plane = static_cast<long long>(n) * d + dd;
    if (part)
      part[(plane * gridDim.x + blockIdx.x) * c + threadIdx.x] = t;
    else
      atomicAdd(S + plane * c + threadIdx.x, t);
  }
}

// out = lrelu(x*A+B)*P+Q written at full resolution AND its (1,2,2) max-pool. A thread owns the
// 2x2 window of one pooled position. grid = (chunks over pooled positions, d, n). h, w even.
template <bool AFFINE>
__global__ void __launch_bounds__(kBlock)
norm_act_pool_kernel(const __nv_bfloat16* __restrict__ x, long long ldx, const float* __restrict__ coef,
                     const float* __restrict__ P, const float* __restrict__ Q, __nv_bfloat16* __restrict__ y,
                     long long ldy, __nv_bfloat16* __restrict__ yp, long long ldp, uint8_t* __restrict__ argmax, PlaneGrid g,
                     int d, int c, int h, int w, float slope) {
  const int n = blockIdx.z, dd = blockIdx.y;
  const int v = threadIdx.x % g.c8, r = threadIdx.x / g.c8;
  const int w2 = w / 2;
  const long long base = (static_cast<long long>(n) * d + dd) * (static_cast<long long>(h) * w);
  const long long basep = (static_cast<long long>(n) * d + dd) * g.hw;  // g.hw = pooled positions per plane
  const int p0 = blockIdx.x * g.chunk;
  const int p1 = min(g.hw, p0 + g.chunk);
  float A[8], B[8], Pv[8], Qv[8];
  load_ab(coef, n, c, v, A, B);
  if (AFFINE) {
    const long long pq = (static_cast<long long>(n) * d + dd) * c + v * 8;
    load8(P + pq, Pv);
    load8(Q + pq, Qv);
  }
  if (r >= g.rpi) return;
  for (int p = p0 + r; p < p1; p += g.rpi) {
    const int hp = p / w2, wp = p % w2;
    float mx[8];
    uint32_t code[2] = {0u, 0u};    // per channel: which corner of the window holds the (first) maximum, one byte each
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const long long pos = base + static_cast<long long>(2 * hp + (k >> 1)) * w + 2 * wp + (k & 1);
      float f[8];
      unpack8(ldg16(x + pos * ldx + v * 8), f);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float z = fmaf(f[i], A[i], B[i]);
        float a = z > 0.f ? z : z * slope;
        f[i] = AFFINE ? fmaf(a, Pv[i], Qv[i]) : a;
      }
      const uint4 o = pack8(f);
      stg16(y + pos * ldy + v * 8, o);
      unpack8(o, f);  // pool the values as stored (bf16), so that backward finds the same arg-max
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (k == 0) {
          mx[i] = f[i];
        } else if (f[i] > mx[i]) {   // strict: the first maximum in (kh, kw) order wins, as in max_pool backward
          mx[i] = f[i];
          code[i >> 2] = (code[i >> 2] & ~(0xffu << (8 * (i & 3)))) | (static_cast<uint32_t>(k) << (8 * (i & 3)));
        }
      }
    }
    stg16(yp + (basep + p) * ldp + v * 8, pack8(mx));
    if (argmax) *reinterpret_cast<uint2*>(argmax + (basep + p) * c + v * 8) = make_uint2(code[0], code[1]);
  }
}

// ---------------------------------------------------------------------------------------------
// backward pass 1: R[n][d][c][6] += sums over (h,w) of
//   {dout*a, dout, dout*m, m, dout*m*xhat, m*xhat},  z = x*A+B, m = lrelu'(z), a = lrelu(z),
//   xhat = (x-mean)*rstd.     grid = (chunks, d, n)
// ---------------------------------------------------------------------------------------------
// PLAIN = true: only {dout*m, dout*m*xhat} (slots 2 and 4) are produced - all the plain
// InstanceNorm + LeakyReLU backward (no gate tail) needs.
template <bool PLAIN>
__global__ void __launch_bounds__(kBlock)
norm_act_bwd_reduce_kernel(const __nv_bfloat16* __restrict__ dout, long long lddo, const __nv_bfloat16* __restrict__ x,
                           long long ldx, const float* __restrict__ coef, float* __restrict__ R, PlaneGrid g, int d,
                           int c, float slope) {
  extern __shared__ float red[];
  constexpr int NK = PLAIN ? 2 : 6;
  const int n = blockIdx.z, dd = blockIdx.y;
  const int v = threadIdx.x % g.c8, r = threadIdx.x / g.c8;
  const long long base = (static_cast<long long>(n) * d + dd) * g.hw;
  const int p0 = blockIdx.x * g.chunk;
  const int p1 = min(g.hw, p0 + g.chunk);
  float A[8], B[8], Mn[8], Rs[8];
  {
    const float4* cf = reinterpret_cast<const float4*>(coef) + static_cast<long long>(n) * c + v * 8;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float4 t = __ldg(cf + i);
      A[i] = t.x; B[i] = t.y; Mn[i] = t.z; Rs[i] = t.w;
    }
  }
  float acc[NK][8];
#pragma unroll
  for (int k = 0; k < NK; ++k)
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[k][i] = 0.f;
  if (r < g.rpi) {
    constexpr int U = 2;   // 2 x (x, dout) 16-byte loads in flight per thread
    for (int p = p0 + r; p < p1; p += U * g.rpi) {
      uint4 rx[U], rg[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const bool ok = p + u * g.rpi < p1;
        rx[u] = ok ? ldg16(x + (base + p + u * g.rpi) * ldx + v * 8) : make_uint4(0, 0, 0, 0);
        rg[u] = ok ? ldg16(dout + (base + p + u * g.rpi) * lddo + v * 8) : make_uint4(0, 0, 0, 0);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        float f[8], go[8];
        unpack8(rx[u], f);
        unpack8(rg[u], go);   // out-of-range rows carry dout = 0 and contribute nothing to slots 0,1,2,4
        const float live = (p + u * g.rpi < p1) ? 1.f : 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float z = fmaf(f[i], A[i], B[i]);
          const float m = z > 0.f ? 1.f : slope;
          const float xh = (f[i] - Mn[i]) * Rs[i];
          const float gm = go[i] * m;
          if (PLAIN) {
            acc[0][i] += gm;
            acc[1][i] = fmaf(gm, xh, acc[1][i]);
          } else {
            const float ml = m * live;
            acc[0][i] = fmaf(go[i], z * m, acc[0][i]);
            acc[1][i] += go[i];
            acc[2][i] += gm;
            acc[3][i] += ml;
            acc[4][i] = fmaf(gm, xh, acc[4][i]);
            acc[5][i] = fmaf(ml, xh, acc[5][i]);
          }
        }
      }
    }
  }
  float* Rp = R + (static_cast<long long>(n) * d + dd) * c * 6;
  const bool fast = fast_reduce_ok(g.c8);
#pragma unroll
  for (int k = 0; k < NK; ++k) {
    const int slot = PLAIN ? (k == 0 ? 2 : 4) : k;
    if (fast) {
      reduce_rows_fast<8>(acc[k], red, g.c8);
      const int L = g.c8 < 32 ? g.c8 : 32;
      if (threadIdx.x < L * 8) atomicAdd(Rp + threadIdx.x * 6 + slot, reduce_rows_fetch<8>(red, g.c8, threadIdx.x));
      __syncthreads();
    } else {
      reduce_rows<8>(acc[k], red, g.c8, g.rpi);
      if (threadIdx.x < g.c8) {
        float* dst = Rp + v * 8 * 6 + slot;
#pragma unroll
        for (int i = 0; i < 8; ++i) atomicAdd(dst + 6 * i, acc[k][i]);
      }
    }
  }
}

// backward pass 2: dx = c1*(dz - c2 - xhat*c3), dz = (dout*P + dSa)*m.  bcoef[n][c] = {c1,c2,c3,-}
// Folded per channel: dx = m * (K1*dout + K1d) + K2*x + K3 with
//   K1 = c1*P, K1d = c1*dSa, K2 = -c1*c3*rstd, K3 = c1*(c3*rstd*mean - c2),  m = lrelu'(x*A + B).
template <bool AFFINE>
__global__ void __launch_bounds__(kBlock)
norm_act_bwd_apply_kernel(const __nv_bfloat16* __restrict__ dout, long long lddo, const __nv_bfloat16* __restrict__ x,
                          long long ldx, const float* __restrict__ coef, const float* __restrict__ bcoef,
                          const float* __restrict__ P, const float* __restrict__ dSa, __nv_bfloat16* __restrict__ dx,
                          long long lddx, PlaneGrid g, int d, int c, float slope) {
  const int n = blockIdx.z, dd = blockIdx.y;
  const int v = threadIdx.x % g.c8, r = threadIdx.x / g.c8;
  const long long base = (static_cast<long long>(n) * d + dd) * g.hw;
  const int p0 = blockIdx.x * g.chunk;
  const int p1 = min(g.hw, p0 + g.chunk);
  float A[8], B[8], K1[8], K1d[8], K2[8], K3[8];
  {
    const float4* cf = reinterpret_cast<const float4*>(coef) + static_cast<long long>(n) * c + v * 8;
    const float4* bf = reinterpret_cast<const float4*>(bcoef) + static_cast<long long>(n) * c + v * 8;
    float Pv[8], Dv[8];
    if (AFFINE) {
      const long long pq = (static_cast<long long>(n) * d + dd) * c + v * 8;
      load8(P + pq, Pv);
      load8(dSa + pq, Dv);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float4 t = __ldg(cf + i);   // {A, B, mean, rstd}
      const float4 u = __ldg(bf + i);   // {c1, c2, c3, -}
      A[i] = t.x;
      B[i] = t.y;
      K1[i] = AFFINE ? u.x * Pv[i] : u.x;
      K1d[i] = AFFINE ? u.x * Dv[i] : 0.f;
      K2[i] = -u.x * u.z * t.w;
      K3[i] = u.x * (u.z * t.w * t.z - u.y);
    }
  }
  if (r >= g.rpi) return;
  constexpr int U = 2;
  for (int p = p0 + r; p < p1; p += U * g.rpi) {
    uint4 rx[U], rg[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (p + u * g.rpi < p1) {
        rx[u] = ldg16(x + (base + p + u * g.rpi) * ldx + v * 8);
        rg[u] = ldg16(dout + (base + p + u * g.rpi) * lddo + v * 8);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (p + u * g.rpi < p1) {
        float f[8], go[8];
        unpack8(rx[u], f);
        unpack8(rg[u], go);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float z = fmaf(f[i], A[i], B[i]);
          const float m = z > 0.f ? 1.f : slope;
          f[i] = fmaf(m, fmaf(K1[i], go[i], K1d[i]), fmaf(K2[i], f[i], K3[i]));
        }
        stg16(dx + (base + p + u * g.rpi) * lddx + v * 8, pack8(f));
      }
    }
  }
}

// The same from the arg-max codes the forward wrote (one byte per pooled element: corner 0..3 of its window): the
// full-resolution activation is not read again (a third of this kernel's traffic; half-line reads of the 2C-pitch concat
// buffer are fetched as whole lines, so more than a third at level 1).
__global__ void __launch_bounds__(kBlock)
maxpool_bwd_codes_kernel(const __nv_bfloat16* __restrict__ dpool, long long ldp, const uint8_t* __restrict__ argmax,
                         __nv_bfloat16* __restrict__ dskip, long long ldd, PlaneGrid g, int d, int c, int h, int w,
                         int accumulate) {
  const int n = blockIdx.z, dd = blockIdx.y;
  const int v = threadIdx.x % g.c8, r = threadIdx.x / g.c8;
  const int w2 = w / 2;
  const long long base = (static_cast<long long>(n) * d + dd) * (static_cast<long long>(h) * w);
  const long long basep = (static_cast<long long>(n) * d + dd) * g.hw;
  const int p0 = blockIdx.x * g.chunk;
  const int p1 = min(g.hw, p0 + g.chunk);
  if (r >= g.rpi) return;
  const long long od[4] = {0, ldd, static_cast<long long>(w) * ldd, static_cast<long long>(w + 1) * ldd};
  __nv_bfloat16* db = dskip + base * ldd + v * 8;
  for (int p = p0 + r; p < p1; p += g.rpi) {
    const int hp = p / w2, wp = p % w2;
    __nv_bfloat16* dc = db + (static_cast<long long>(2 * hp) * w + 2 * wp) * ldd;
    const uint4 rawg = ldg16(dpool + (basep + p) * ldp + v * 8);
    const uint2 cd = __ldg(reinterpret_cast<const uint2*>(argmax + (basep + p) * c + v * 8));
    uint4 rawd[4];
    if (accumulate) {
#pragma unroll
      for (int k = 0; k < 4; ++k) rawd[k] = *reinterpret_cast<const uint4*>(dc + od[k]);
    }
    float gp[8];
    unpack8(rawg, gp);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float o[8];
      if (accumulate) {
        unpack8(rawd[k], o);
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = 0.f;
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const uint32_t a = ((i < 4 ? cd.x : cd.y) >> (8 * (i & 3))) & 0xffu;
        o[i] += (a == static_cast<uint32_t>(k)) ? gp[i] : 0.f;
      }
      stg16(dc + od[k], pack8(o));
    }
  }
}

// (1,2,2) max-pool backward fused with the skip add: dskip[arg-max of y in the window] += dpool.
// grid = (chunks over pooled positions, d, n)
__global__ void __launch_bounds__(kBlock)
maxpool_bwd_add_kernel(const __nv_bfloat16* __restrict__ dpool, long long ldp, const __nv_bfloat16* __restrict__ y,
                       long long ldy, __nv_bfloat16* __restrict__ dskip, long long ldd, PlaneGrid g, int d, int h, int w,
                       int accumulate) {
  const int n = blockIdx.z, dd = blockIdx.y;
  const int v = threadIdx.x % g.c8, r = threadIdx.x / g.c8;
  const int w2 = w / 2;
  const long long base = (static_cast<long long>(n) * d + dd) * (static_cast<long long>(h) * w);
  const long long basep = (static_cast<long long>(n) * d + dd) * g.hw;
  const int p0 = blockIdx.x * g.chunk;
  const int p1 = min(g.hw, p0 + g.chunk);
  if (r >= g.rpi) return;
  // (hp, wp) of the pooled position advance incrementally; the four window corners sit at fixed element offsets
  int hp = (p0 + r) / w2, wp = (p0 + r) % w2;
  const int dh = g.rpi / w2, dw = g.rpi % w2;
  const long long oy[4] = {0, ldy, static_cast<long long>(w) * ldy, static_cast<long long>(w + 1) * ldy};
  const long long od[4] = {0, ldd, static_cast<long long>(w) * ldd, static_cast<long long>(w + 1) * ldd};
  const __nv_bfloat16* yb = y + base * ldy + v * 8;
  __nv_bfloat16* db = dskip + base * ldd + v * 8;
  const __nv_bfloat16* pp = dpool + (basep + p0 + r) * ldp + v * 8;
  const long long sp = static_cast<long long>(g.rpi) * ldp;
  for (int p = p0 + r; p < p1; p += g.rpi) {
    const long long corner = static_cast<long long>(2 * hp) * w + 2 * wp;
    const __nv_bfloat16* yc = yb + corner * ldy;
    __nv_bfloat16* dc = db + corner * ldd;
    float gp[8], yv[4][8];
    uint4 raw[4], rawd[4];
    const uint4 rawg = ldg16(pp);
#pragma unroll
    for (int k = 0; k < 4; ++k) raw[k] = ldg16(yc + oy[k]);
    if (accumulate) {
#pragma unroll
      for (int k = 0; k < 4; ++k) rawd[k] = *reinterpret_cast<const uint4*>(dc + od[k]);
    }
    unpack8(rawg, gp);
#pragma unroll
    for (int k = 0; k < 4; ++k) unpack8(raw[k], yv[k]);
    int arg[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      int a = 0;
      float best = yv[0][i];
#pragma unroll
      for (int k = 1; k < 4; ++k)
        if (yv[k][i] > best) {
          best = yv[k][i];
          a = k;
        }
      arg[i] = a;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float o[8];
      if (accumulate) {
        unpack8(rawd[k], o);
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = 0.f;
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] += (arg[i] == k) ? gp[i] : 0.f;
      stg16(dc + od[k], pack8(o));
    }
    pp += sp;
    hp += dh;
    wp += dw;
    if (wp >= w2) {
      wp -= w2;
      ++hp;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Backward passes, 4 channels per thread (8-byte accesses). Half the per-thread state of the
// 8-channel kernels above -> 3-4 resident blocks per SM instead of 2, which is what these
// instruction-heavy streaming kernels need to cover HBM latency. cv = c / 4 vectors per position,
// a power of two <= 256; thread -> (row r = tid / cv, vector v = tid % cv).
// ---------------------------------------------------------------------------------------------
struct PlaneGrid4 {
  int cv, rpi, hw, chunk;
};

__device__ __forceinline__ void unpack4(const uint2& v, float (&f)[4]) {
  const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&v.x));
  const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&v.y));
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y;
}
__device__ __forceinline__ uint2 pack4(const float (&f)[4]) {
  uint2 v;
  *reinterpret_cast<__nv_bfloat162*>(&v.x) = __floats2bfloat162_rn(f[0], f[1]);
  *reinterpret_cast<__nv_bfloat162*>(&v.y) = __floats2bfloat162_rn(f[2], f[3]);
  return v;
}
__device__ __forceinline__ uint2 ldg8(const __nv_bfloat16* p) { return __ldg(reinterpret_cast<const uint2*>(p)); }

// rows sharing a warp fold with shuffles, row-warps meet in shared memory ([row-warp][vec][NV]);
// afterwards reduce_cv_fetch(idx) returns output element idx = vec * NV + i. red: 8*32*NV floats.
template <int NV>
__device__ __forceinline__ void reduce_cv(float (&v)[NV], float* red, int cv) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int L = cv < 32 ? cv : 32, wpr = cv > 32 ? cv / 32 : 1;
  for (int off = 16; off >= cv; off >>= 1) {
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] += __shfl_xor_sync(0xffffffffu, v[i], off);
  }
  if (lane < L) {
#pragma unroll
    for (int i = 0; i < NV; ++i) red[((warp / wpr) * (wpr * L) + (warp % wpr) * L + lane) * NV + i] = v[i];
  }
  __syncthreads();
}
template <int NV>
__device__ __forceinline__ float reduce_cv_fetch(const float* red, int cv, int idx) {
  const int L = cv < 32 ? cv : 32, wpr = cv > 32 ? cv / 32 : 1;
  const int nrw = (kBlock / 32) / wpr, span = wpr * L * NV;
  float t = 0.f;
  for (int w = 0; w < nrw; ++w) t += red[w * span + idx];
  return t;
}

template <bool PLAIN>
__global__ void __launch_bounds__(kBlock, PLAIN ? 4 : 3)
norm_act_bwd_reduce4_kernel(const __nv_bfloat16* __restrict__ dout, long long lddo, const __nv_bfloat16* __restrict__ x,
                            long long ldx, const float* __restrict__ coef, float* __restrict__ R, PlaneGrid4 g, int d,
                            int c, float slope, float* __restrict__ part /* [plane][chunk][c][NK] or null (atomics) */) {
  extern __shared__ float red[];
  constexpr int NK = PLAIN ? 2 : 6;
  const int n = blockIdx.z, dd = blockIdx.y;
  const int v = threadIdx.x % g.cv, r = threadIdx.x / g.cv;
  const long long base = (static_cast<long long>(n) * d + dd) * g.hw;
  const int p0 = blockIdx.x * g.chunk;
  const int p1 = min(g.hw, p0 + g.chunk);
  float A[4], B[4], Mn[4], Rs[4];
  {
    const float4* cf = reinterpret_cast<const float4*>(coef) + static_cast<long long>(n) * c + v * 4;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float4 t = __ldg(cf + i);
      A[i] = t.x; B[i] = t.y; Mn[i] = t.z; Rs[i] = t.w;
    }
  }
  float acc[NK][4];
#pragma unroll
  for (int k = 0; k < NK; ++k)
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[k][i] = 0.f;
  constexpr int U = 4;   // 4 x (x, dout) 8-byte loads in flight per thread
  for (int p = p0 + r; p < p1; p += U * g.rpi) {
    uint2 rx[U], rg[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const bool ok = p + u * g.rpi < p1;
      rx[u] = ok ? ldg8(x + (base + p + u * g.rpi) * ldx + v * 4) : make_uint2(0, 0);
      rg[u] = ok ? ldg8(dout + (base + p + u * g.rpi) * lddo + v * 4) : make_uint2(0, 0);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      float f[4], go[4];
      unpack4(rx[u], f);
      unpack4(rg[u], go);   // rows past the chunk carry dout = 0: nothing reaches slots 0, 1, 2, 4
      const float live = (p + u * g.rpi < p1) ? 1.f : 0.f;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float z = fmaf(f[i], A[i], B[i]);
        const float m = z > 0.f ? 1.f : slope;
        const float xh = (f[i] - Mn[i]) * Rs[i];
        const float gm = go[i] * m;
        if (PLAIN) {
          acc[0][i] += gm;
          acc[1][i] = fmaf(gm, xh, acc[1][i]);
        } else {
          const float ml = m * live;
          acc[0][i] = fmaf(gm, z, acc[0][i]);
          acc[1][i] += go[i];
          acc[2][i] += gm;
          acc[3][i] += ml;
          acc[4][i] = fmaf(gm, xh, acc[4][i]);
          acc[5][i] = fmaf(ml, xh, acc[5][i]);
        }
      }
    }
  }
  float* Rp = R + (static_cast<long long>(n) * d + dd) * c * 6;
  const int nout = (g.cv < kBlock ? g.cv : kBlock) * 4;   // = c
#pragma unroll
  for (int k = 0; k < NK; ++k) {
    const int slot = PLAIN ? (k == 0 ? 2 : 4) : k;
    reduce_cv<4>(acc[k], red, g.cv);
    for (int idx = threadIdx.x; idx < nout; idx += kBlock) {
      const float t = reduce_cv_fetch<4>(red, g.cv, idx);
      if (part)
        part[(((static_cast<long long>(n) * d + dd) * gridDim.x + blockIdx.x) * c + idx) * NK + k] = t;
      else
        atomicAdd(Rp + idx * 6 + slot, t);
    }
    __syncthreads();
  }
}

// Leaner first stage for the fixed-order path. With a = lrelu(z) = m*z and xhat = (z - beta)/gamma the two xhat sums
// of R follow from sums the kernel needs anyway:
//   sum dout*m*xhat = (sum dout*m*z - beta * sum dout*m) / gamma        (slot 4 from slots 0 and 2)
//   sum m*xhat      = (S - beta * sum m) / gamma,  S = sum a = sum m*z   (slot 5 from the forward statistic S and slot 3)
// so only {dout*m*z, dout, dout*m, m} (PLAIN: {dout*m, dout*m*z}) are accumulated, with the per-channel state
// reduced to (A, B): about half the instructions per element of norm_act_bwd_reduce4_kernel, which was
// issue-bound (ncu: sm throughput 71-74 %, DRAM 44-58 %). sum_chunks_bwd_kernel folds the chunks in a fixed order
// and applies the two identities. Raw partial layout: part[plane][chunk][c][NKR], NKR = 2 (PLAIN) or 4.
template <bool PLAIN>
__global__ void __launch_bounds__(kBlock, PLAIN ? 5 : 4)
norm_act_bwd_reduce4v2_kernel(const __nv_bfloat16* __restrict__ dout, long long lddo, const __nv_bfloat16* __restrict__ x,
                              long long ldx, const float* __restrict__ coef, PlaneGrid4 g, int d, int c, float slope,
                              float* __restrict__ part) {
  extern __shared__ float red[];
  constexpr int NKR = PLAIN ? 2 : 4;
  const int n = blockIdx.z, dd = blockIdx.y;
  const int v = threadIdx.x % g.cv, r = threadIdx.x / g.cv;
  const long long base = (static_cast<long long>(n) * d + dd) * g.hw;
  const int p0 = blockIdx.x * g.chunk;
  const int p1 = min(g.hw, p0 + g.chunk);
  float A[4], B[4];
  {
    const float4* cf = reinterpret_cast<const float4*>(coef) + static_cast<long long>(n) * c + v * 4;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float4 t = __ldg(cf + i);
      A[i] = t.x; B[i] = t.y;
    }
  }
  float acc[NKR][4];
#pragma unroll
  for (int k = 0; k < NKR; ++k)
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[k][i] = 0.f;
  // rows p0 + r, p0 + r + rpi, ... < p1 of this thread: pointers advance by a fixed stride (no per-load index
  // arithmetic or predication: the first version of this loop spent 22 instructions per element, most of them
  // 64-bit address math and branches around masked loads); full groups of U rows, then a tail of single rows
  auto body = [&](const uint2& rxv, const uint2& rgv) {
    float f[4], go[4];
    unpack4(rxv, f);
    unpack4(rgv, go);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float z = fmaf(f[i], A[i], B[i]);
      const float m = z > 0.f ? 1.f : slope;
      const float gm = go[i] * m;
      if (PLAIN) {
        acc[0][i] += gm;
        acc[1][i] = fmaf(gm, z, acc[1][i]);
      } else {
        acc[0][i] = fmaf(gm, z, acc[0][i]);
        acc[1][i] += go[i];
        acc[2][i] += gm;
        acc[3][i] += m;
      }
    }
  };
  constexpr int U = 4;
  int cnt = (p0 + r < p1) ? (p1 - p0 - r + g.rpi - 1) / g.rpi : 0;
  const long long sx = static_cast<long long>(g.rpi) * ldx, sg = static_cast<long long>(g.rpi) * lddo;
  const __nv_bfloat16* px = x + (base + p0 + r) * ldx + v * 4;
  const __nv_bfloat16* pg = dout + (base + p0 + r) * lddo + v * 4;
  for (; cnt >= U; cnt -= U) {
    uint2 rx[U], rg[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      rx[u] = ldg8(px + u * sx);
      rg[u] = ldg8(pg + u * sg);
    }
    px += U * sx;
    pg += U * sg;
#pragma unroll
    for (int u = 0; u < U; ++u) body(rx[u], rg[u]);
  }
  for (; cnt > 0; --cnt) {
    const uint2 rx = ldg8(px), rg = ldg8(pg);
    px += sx;
    pg += sg;
    body(rx, rg);
  }
  const int nout = (g.cv < kBlock ? g.cv : kBlock) * 4;   // = c
#pragma unroll
  for (int k = 0; k < NKR; ++k) {
    reduce_cv<4>(acc[k], red, g.cv);
    for (int idx = threadIdx.x; idx < nout; idx += kBlock)
      part[(((static_cast<long long>(n) * d + dd) * gridDim.x + blockIdx.x) * c + idx) * NKR + k] =
          reduce_cv_fetch<4>(red, g.cv, idx);
    __syncthreads();
  }
}

// Second stage: fold the chunks in order, apply the xhat identities, write the R slots (PLAIN: 2 and 4; else all six).
template <bool PLAIN>
__global__ void sum_chunks_bwd_kernel(const float* __restrict__ part, int chunks, int c, int d, const float* __restrict__ coef,
                                      const float* __restrict__ S, float* __restrict__ R, long long total) {
  constexpr int NKR = PLAIN ? 2 : 4;
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;   // (plane, channel)
  if (i >= total) return;
  const int ch = static_cast<int>(i % c);
  const long long plane = i / c;
  const long long n = plane / d;
  float t[NKR];
#pragma unroll
  for (int k = 0; k < NKR; ++k) t[k] = 0.f;
  for (int q = 0; q < chunks; ++q) {
    const float* src = part + ((plane * chunks + q) * c + ch) * NKR;
#pragma unroll
    for (int k = 0; k < NKR; ++k) t[k] += src[k];
  }
  const float4 cf = __ldg(reinterpret_cast<const float4*>(coef) + n * c + ch);   // {A, B, mean, rstd}
  const float gamma = cf.x / cf.w;
  const float beta = fmaf(cf.z, cf.x, cf.y);
  const float inv_g = gamma != 0.f ? 1.f / gamma : 0.f;   // gamma == 0: xhat is not recoverable from z; those sums read 0
  float* out = R + i * 6;
  if (PLAIN) {
    out[2] = t[0];
    out[4] = (t[1] - beta * t[0]) * inv_g;
  } else {
    out[0] = t[0];
    out[1] = t[1];
    out[2] = t[2];
    out[3] = t[3];
    out[4] = (t[0] - beta * t[2]) * inv_g;
    out[5] = (S[i] - beta * t[3]) * inv_g;
  }
}

template <bool AFFINE>
__global__ void __launch_bounds__(kBlock, 4)
norm_act_bwd_apply4_kernel(const __nv_bfloat16* __restrict__ dout, long long lddo, const __nv_bfloat16* __restrict__ x,
                           long long ldx, const float* __restrict__ coef, const float* __restrict__ bcoef,
                           const float* __restrict__ P, const float* __restrict__ dSa, __nv_bfloat16* __restrict__ dx,
                           long long lddx, PlaneGrid4 g, int d, int c, float slope) {
  const int n = blockIdx.z, dd = blockIdx.y;
  const int v = threadIdx.x % g.cv, r = threadIdx.x / g.cv;
  const long long base = (static_cast<long long>(n) * d + dd) * g.hw;
  const int p0 = blockIdx.x * g.chunk;
  const int p1 = min(g.hw, p0 + g.chunk);
  float A[4], B[4], K1[4], K1d[4], K2[4], K3[4];
  {
    const float4* cf = reinterpret_cast<const float4*>(coef) + static_cast<long long>(n) * c + v * 4;
    const float4* bf = reinterpret_cast<const float4*>(bcoef) + static_cast<long long>(n) * c + v * 4;
    float4 pv = make_float4(1.f, 1.f, 1.f, 1.f), dv = make_float4(0.f, 0.f, 0.f, 0.f);
    if (AFFINE) {
      const long long pq = (static_cast<long long>(n) * d + dd) * c + v * 4;
      pv = __ldg(reinterpret_cast<const float4*>(P + pq));
      dv = __ldg(reinterpret_cast<const float4*>(dSa + pq));
    }
    const float Pv[4] = {pv.x, pv.y, pv.z, pv.w}, Dv[4] = {dv.x, dv.y, dv.z, dv.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float4 t = __ldg(cf + i);   // {A, B, mean, rstd}
      const float4 u = __ldg(bf + i);   // {c1, c2, c3, -}
      A[i] = t.x;
      B[i] = t.y;
      K1[i] = u.x * Pv[i];
      K1d[i] = u.x * Dv[i];
      K2[i] = -u.x * u.z * t.w;
      K3[i] = u.x * (u.z * t.w * t.z - u.y);
    }
  }
  constexpr int U = 4;
  for (int p = p0 + r; p < p1; p += U * g.rpi) {
    uint2 rx[U], rg[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (p + u * g.rpi < p1) {
        rx[u] = ldg8(x + (base + p + u * g.rpi) * ldx + v * 4);
        rg[u] = ldg8(dout + (base + p + u * g.rpi) * lddo + v * 4);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (p + u * g.rpi < p1) {
        float f[4], go[4];
        unpack4(rx[u], f);
        unpack4(rg[u], go);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float z = fmaf(f[i], A[i], B[i]);
          const float m = z > 0.f ? 1.f : slope;
          f[i] = fmaf(m, fmaf(K1[i], go[i], K1d[i]), fmaf(K2[i], f[i], K3[i]));
        }
        *reinterpret_cast<uint2*>(dx + (base + p + u * g.rpi) * lddx + v * 4) = pack4(f);
      }
    }
  }
}

// Fixed-order second stage of the plane reductions: out[(plane*c + ch)*out_stride + slot0 + k*slot_step] =
// sum over the plane's chunks of part[((plane*chunks + chunk)*c + ch)*nk + k]  (overwrites: no zeroing,
// no atomics -> bit-reproducible S and R).
__global__ void sum_chunks_kernel(const float* __restrict__ part, int chunks, int c, int nk, float* __restrict__ out,
                                  int out_stride, int slot0, int slot_step, long long total) {
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i >= total) return;
  const int k = static_cast<int>(i % nk);
  const int ch = static_cast<int>((i / nk) % c);
  const long long plane = i / (static_cast<long long>(nk) * c);
  float t = 0.f;
  for (int q = 0; q < chunks; ++q) t += part[((plane * chunks + q) * c + ch) * nk + k];
  out[(plane * c + ch) * out_stride + slot0 + k * slot_step] = t;
}

// grid for the 4-channel kernels; returns false when the channel count does not fit them
bool make_grid4(int c, long long hw, spff_shape s, PlaneGrid4* g, dim3* grid) {
  if (c % 4 != 0 || c <= 0) return false;
  const int cv = c / 4;
  if (cv > kBlock || (cv & (cv - 1)) != 0 || s.d > 65535 || s.n > 65535) return false;
  g->cv = cv;
  g->rpi = kBlock / cv;
  g->hw = static_cast<int>(hw);
  // ~6 waves of blocks over the device (3-4 resident per SM), at least 8 iterations per thread
  const long long planes = static_cast<long long>(s.n) * s.d;
  const long long per_sm = debug_flag(3) > 0 ? debug_flag(3) : 24;   // test hook (spff_debug_set key 3)
  long long want_chunks = (per_sm * num_sms() + planes - 1) / planes;
  if (want_chunks < 1) want_chunks = 1;
  long long chunk = (hw + want_chunks - 1) / want_chunks;
  const long long min_chunk = 8LL * g->rpi;
  if (chunk < min_chunk) chunk = min_chunk;
  chunk = ((chunk + g->rpi - 1) / g->rpi) * g->rpi;
  g->chunk = static_cast<int>(chunk);
  *grid = dim3(static_cast<unsigned>((hw + chunk - 1) / chunk), s.d, s.n);
  return true;
}

// blocks_per_sm: 8 for the kernels that end in a block reduction (their epilogue favours long blocks), 32 for the pure
// streaming ones (measured: norm_act_apply 5.9 -> 6.5 TB/s, maxpool_bwd_add 4.6 -> 5.0 TB/s at level 1)
int make_grid(int c, long long hw, spff_shape s, PlaneGrid* g, dim3* grid, int blocks_per_sm = 8) {
  if (c % 8 != 0 || c <= 0 || c > 8 * kBlock) {
    set_error("channel count %d must be a multiple of 8 and <= %d", c, 8 * kBlock);
    return SPFF_ERR_BAD_ARGUMENT;
  }
  g->c8 = c / 8;
  g->rpi = kBlock / g->c8;
  g->hw = static_cast<int>(hw);
  // aim at >= ~8 blocks per SM overall while keeping >= 16 iterations per block where the plane allows
  const long long planes = static_cast<long long>(s.n) * s.d;
  const long long per_sm = debug_flag(4) > 0 ? debug_flag(4) : blocks_per_sm;    // test hook (spff_debug_set key 4)
  long long want_chunks = (per_sm * num_sms() + planes - 1) / planes;
  if (want_chunks < 1) want_chunks = 1;
  long long chunk = (hw + want_chunks - 1) / want_chunks;
  const long long min_chunk = 16LL * g->rpi;
  if (chunk < min_chunk) chunk = min_chunk;
  chunk = ((chunk + g->rpi - 1) / g->rpi) * g->rpi;
  g->chunk = static_cast<int>(chunk);
  const int chunks = static_cast<int>((hw + chunk - 1) / chunk);
  *grid = dim3(chunks, s.d, s.n);
  if (s.d > 65535 || s.n > 65535) {
    set_error("too many planes/samples for one launch");
    return SPFF_ERR_BAD_ARGUMENT;
  }
  return 0;
}

}  // namespace
}  // namespace spff

using spff::kBlock;
using spff::PlaneGrid;
typedef __nv_bfloat16 bf16;

#define SPFF_ENTRY_CHECK()          \
  do {                              \
    int _e = spff_device_check();   \
    if (_e) return _e;              \
  } while (0)

extern "C" {

int spff_in_stats(const void* x, long long ldx, int c, spff_shape s, double* stats, void* stream) {
  SPFF_ENTRY_CHECK();
  PlaneGrid g;
  dim3 grid;
  int e = spff::make_grid(c, static_cast<long long>(s.h) * s.w, s, &g, &grid);
  if (e) return e;
  spff::in_stats_kernel<<<grid, kBlock, 16 * kBlock * sizeof(float), static_cast<cudaStream_t>(stream)>>>(
      static_cast<const bf16*>(x), ldx, g, s.d, stats, c);
  SPFF_CUDA(cudaGetLastError());
  return 0;
}

int spff_in_coeffs(const double* stats, const float* gamma, const float* beta, float eps, int n, int c,
                   long long count, int batch_stats, float* coef, void* stream) {
  SPFF_ENTRY_CHECK();
  const int total = n * c;
  spff::in_coeffs_kernel<<<(total + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(
      stats, gamma, beta, eps, n, c, static_cast<double>(count), batch_stats, coef);
  SPFF_CUDA(cudaGetLastError());
  return 0;
}

int spff_in_coeffs_from_partials(const float* partial, int slots, const float* gamma, const float* beta, float eps, int n,
                                 int c, long long count, float* coef, void* stream) {
  SPFF_ENTRY_CHECK();
  SPFF_REQUIRE(partial && coef && slots > 0 && n > 0 && c > 0, "in_coeffs_from_partials: bad arguments");
  const int total = n * c;
  spff::in_coeffs_partial_kernel<<<(total + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(
      partial, slots, gamma, beta, eps, n, c, static_cast<double>(count), coef);
  SPFF_CUDA(cudaGetLastError());
  return 0;
}

int spff_norm_act_apply(const void* x, long long ldx, const float* coef, void* y, long long ldy, int c, spff_shape s,
                        float slope, void* stream) {
  SPFF_ENTRY_CHECK();
  PlaneGrid g;
  dim3 grid;
  int e = spff::make_grid(c, static_cast<long long>(s.h) * s.w, s, &g, &grid, 32);
  if (e) return e;
  spff::norm_act_kernel<true, false, false><<<grid, kBlock, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const bf16*>(x), ldx, coef, nullptr, nullptr, static_cast<bf16*>(y), ldy, nullptr, g, s.d, c, slope);
  SPFF_CUDA(cudaGetLastError());
  return 0;
}

size_t spff_norm_act_reduce_workspace(int c, spff_shape s) {
  PlaneGrid g;
  dim3 grid;
  if (c <= 0 || s.n <= 0 || s.d <= 0 || s.h <= 0 || s.w <= 0) return 0;
  if (spff::make_grid(c, static_cast<long long>(s.h) * s.w, s, &g, &grid) || !spff::fast_reduce_ok_host(g.c8)) return 0;
  return static_cast<size_t>(s.n) * s.d * grid.x * c * sizeof(float);
}

int spff_norm_act_reduce(const void* x, long long ldx, const float* coef, float* S, int c, spff_shape s, float slope,
                         void* workspace, size_t workspace_bytes, void* stream) {
  SPFF_ENTRY_CHECK();
  PlaneGrid g;
  dim3 grid;
  int e = spff::make_grid(c, static_cast<long long>(s.h) * s.w, s, &g, &grid);
  if (e) return e;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t need = spff_norm_act_reduce_workspace(c, s);
  float* part = (workspace && need > 0 && workspace_bytes >= need) ? static_cast<float*>(workspace) : nullptr;
  SPFF_REQUIRE(!workspace || part, "norm_act_reduce: workspace too small or channel count without a fixed-order path");
  if (spff::fast_reduce_ok_host(g.c8))
    spff::norm_act_sum_kernel<<<grid, kBlock, 8 * kBlock * sizeof(float), st>>>(static_cast<const bf16*>(x), ldx, coef, S, g,
                                                                              s.d, c, slope, part);
  else
    spff::norm_act_kernel<false, true, false><<<grid, kBlock, 8 * kBlock * sizeof(float), st>>>(
        static_cast<const bf16*>(x), ldx, coef, nullptr, nullptr, nullptr, 0, S, g, s.d, c, slope, part);
  if (part) {
    const long long total = static_cast<long long>(s.n) * s.d * c;
    spff::sum_chunks_kernel<<<static_cast<int>((total + 255) / 256), 256, 0, st>>>(part, grid.x, c, 1, S, 1, 0, 1, total);
  }
  SPFF_CUDA(cudaGetLastError());
  return 0;
}

int spff_norm_act_affine_apply(const void* x, long long ldx, const float* coef, const float* P, const float* Q,
                               void* y, long long ldy, void* ypool, long long ldp, uint8_t* pool_argmax, int c, spff_shape s,
                               float slope, void* stream) {
  SPFF_ENTRY_CHECK();
  SPFF_REQUIRE((P == nullptr) == (Q == nullptr), "norm_act_affine_apply: P and Q must both be given or both be NULL");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  PlaneGrid g;
  dim3 grid;
  if (ypool) {
    SPFF_REQUIRE(s.h % 2 == 0 && s.w % 2 == 0, "norm_act_affine_apply: pooling needs even H, W (got %d x %d)", s.h, s.w);
    int e = spff::make_grid(c, static_cast<long long>(s.h / 2) * (s.w / 2), s, &g, &grid, 32);
    if (e) return e;
    if (P)
      spff::norm_act_pool_kernel<true><<<grid, kBlock, 0, st>>>(static_cast<const bf16*>(x), ldx, coef, P, Q,
                                                               static_cast<bf16*>(y), ldy, static_cast<bf16*>(ypool),
                                                               ldp, pool_argmax, g, s.d, c, s.h, s.w, slope);
    else
      spff::norm_act_pool_kernel<false><<<grid, kBlock, 0, st>>>(static_cast<const bf16*>(x), ldx, coef, P, Q,
                                                                static_cast<bf16*>(y), ldy, static_cast<bf16*>(ypool),
                                                                ldp, pool_argmax, g, s.d, c, s.h, s.w, slope);
  } else {
    int e = spff::make_grid(c, static_cast<long long>(s.h) * s.w, s, &g, &grid, 32);
    if (e) return e;
    if (P)
      spff::norm_act_kernel<true, false, true><<<grid, kBlock, 0, st>>>(
          static_cast<const bf16*>(x), ldx, coef, P, Q, static_cast<bf16*>(y), ldy, nullptr, g, s.d, c, slope);
    else
      spff::norm_act_kernel<true, false, false><<<grid, kBlock, 0, st>>>(
          static_cast<const bf16*>(x), ldx, coef, P, Q, static_cast<bf16*>(y), ldy, nullptr, g, s.d, c, slope);
  }
  SPFF_CUDA(cudaGetLastError());
  return 0;
}

size_t spff_norm_act_bwd_reduce_workspace(int c, spff_shape s, int plain) {
  spff::PlaneGrid4 g4;
  dim3 grid;
  if (c <= 0 || s.n <= 0 || s.d <= 0 || s.h <= 0 || s.w <= 0) return 0;
  if (!spff::make_grid4(c, static_cast<long long>(s.h) * s.w, s, &g4, &grid)) return 0;
  return static_cast<size_t>(s.n) * s.d * grid.x * c * (plain ? 2 : 6) * sizeof(float);
}

int spff_norm_act_bwd_reduce(const void* dout, long long lddo, const void* x, long long ldx, const float* coef,
                             float* R, const float* S, int c, spff_shape s, float slope, int plain, void* workspace,
                             size_t workspace_bytes, void* stream) {
  SPFF_ENTRY_CHECK();
  PlaneGrid g;
  dim3 grid;
  {
    spff::PlaneGrid4 g4;
    if (spff::make_grid4(c, static_cast<long long>(s.h) * s.w, s, &g4, &grid)) {
      cudaStream_t st4 = static_cast<cudaStream_t>(stream);
      const size_t smem = 8 * 32 * 4 * sizeof(float);
      const size_t need = spff_norm_act_bwd_reduce_workspace(c, s, plain);
      float* part = (workspace && workspace_bytes >= need) ? static_cast<float*>(workspace) : nullptr;
      SPFF_REQUIRE(!workspace || part, "norm_act_bwd_reduce: workspace too small");
      if (part && (plain || S)) {   // fixed-order path, lean first stage (the xhat sums follow from the others)
        const long long total = static_cast<long long>(s.n) * s.d * c;
        const int blocks = static_cast<int>((total + 255) / 256);
        if (plain) {
          spff::norm_act_bwd_reduce4v2_kernel<true><<<grid, kBlock, smem, st4>>>(
              static_cast<const bf16*>(dout), lddo, static_cast<const bf16*>(x), ldx, coef, g4, s.d, c, slope, part);
          spff::sum_chunks_bwd_kernel<true><<<blocks, 256, 0, st4>>>(part, grid.x, c, s.d, coef, nullptr, R, total);
        } else {
          spff::norm_act_bwd_reduce4v2_kernel<false><<<grid, kBlock, smem, st4>>>(
              static_cast<const bf16*>(dout), lddo, static_cast<const bf16*>(x), ldx, coef, g4, s.d, c, slope, part);
          spff::sum_chunks_bwd_kernel<false><<<blocks, 256, 0, st4>>>(part, grid.x, c, s.d, coef, S, R, total);
        }
        SPFF_CUDA(cudaGetLastError());
        return 0;
      }
      if (plain)
        spff::norm_act_bwd_reduce4_kernel<true><<<grid, kBlock, smem, st4>>>(
            static_cast<const bf16*>(dout), lddo, static_cast<const bf16*>(x), ldx, coef, R, g4, s.d, c, slope, part);
      else
        spff::norm_act_bwd_reduce4_kernel<false><<<grid, kBlock, smem, st4>>>(
            static_cast<const bf16*>(dout), lddo, static_cast<const bf16*>(x), ldx, coef, R, g4, s.d, c, slope, part);
      if (part) {
        const int nk = plain ? 2 : 6;
        const long long total = static_cast<long long>(s.n) * s.d * c * nk;
        spff::sum_chunks_kernel<<<static_cast<int>((total + 255) / 256), 256, 0, st4>>>(part, grid.x, c, nk, R, 6, plain ? 2 : 0,
                                                                                       plain ? 2 : 1, total);
      }
      SPFF_CUDA(cudaGetLastError());
      return 0;
    }
  }
  SPFF_REQUIRE(!workspace, "norm_act_bwd_reduce: this channel count has no fixed-order path (pass no workspace)");
  int e = spff::make_grid(c, static_cast<long long>(s.h) * s.w, s, &g, &grid);
  if (e) return e;
  if (plain)
    spff::norm_act_bwd_reduce_kernel<true><<<grid, kBlock, 8 * kBlock * sizeof(float), static_cast<cudaStream_t>(stream)>>>(
        static_cast<const bf16*>(dout), lddo, static_cast<const bf16*>(x), ldx, coef, R, g, s.d, c, slope);
  else
    spff::norm_act_bwd_reduce_kernel<false><<<grid, kBlock, 8 * kBlock * sizeof(float), static_cast<cudaStream_t>(stream)>>>(
        static_cast<const bf16*>(dout), lddo, static_cast<const bf16*>(x), ldx, coef, R, g, s.d, c, slope);
  SPFF_CUDA(cudaGetLastError());
  return 0;
}

int spff_norm_act_bwd_apply(const void* dout, long long lddo, const void* x, long long ldx, const float* coef,
                            const float* bcoef, const float* P, const float* dSa, void* dx, long long lddx, int c,
                            spff_shape s, float slope, void* stream) {
  SPFF_ENTRY_CHECK();
  SPFF_REQUIRE((P == nullptr) == (dSa == nullptr), "norm_act_bwd_apply: P and dSa must both be given or both be NULL");
  PlaneGrid g;
  dim3 grid;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  {
    spff::PlaneGrid4 g4;
    if (spff::make_grid4(c, static_cast<long long>(s.h) * s.w, s, &g4, &grid)) {
      if (P)
        spff::norm_act_bwd_apply4_kernel<true><<<grid, kBlock, 0, st>>>(
            static_cast<const bf16*>(dout), lddo, static_cast<const bf16*>(x), ldx, coef, bcoef, P, dSa,
            static_cast<bf16*>(dx), lddx, g4, s.d, c, slope);
      else
        spff::norm_act_bwd_apply4_kernel<false><<<grid, kBlock, 0, st>>>(
            static_cast<const bf16*>(dout), lddo, static_cast<const bf16*>(x), ldx, coef, bcoef, P, dSa,
            static_cast<bf16*>(dx), lddx, g4, s.d, c, slope);
      SPFF_CUDA(cudaGetLastError());
      return 0;
    }
  }
  int e = spff::make_grid(c, static_cast<long long>(s.h) * s.w, s, &g, &grid);
  if (e) return e;
  if (P)
    spff::norm_act_bwd_apply_kernel<true><<<grid, kBlock, 0, st>>>(static_cast<const bf16*>(dout), lddo,
                                                                  static_cast<const bf16*>(x), ldx, coef, bcoef, P, dSa,
                                                                  static_cast<bf16*>(dx), lddx, g, s.d, c, slope);
  else
    spff::norm_act_bwd_apply_kernel<false><<<grid, kBlock, 0, st>>>(static_cast<const bf16*>(dout), lddo,
                                                                   static_cast<const bf16*>(x), ldx, coef, bcoef, P, dSa,
                                                                   static_cast<bf16*>(dx), lddx, g, s.d, c, slope);
  SPFF_CUDA(cudaGetLastError());
  return 0;
}

int spff_maxpool_bwd_add(const void* dpool, long long ldp, const void* y, long long ldy, void* dskip, long long ldd,
                         int c, spff_shape s, int accumulate, void* stream) {
  SPFF_ENTRY_CHECK();
  SPFF_REQUIRE(s.h % 2 == 0 && s.w % 2 == 0, "maxpool_bwd_add: needs even H, W");
  PlaneGrid g;
  dim3 grid;
  int e = spff::make_grid(c, static_cast<long long>(s.h / 2) * (s.w / 2), s, &g, &grid, 32);
  if (e) return e;
  spff::maxpool_bwd_add_kernel<<<grid, kBlock, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const bf16*>(dpool), ldp, static_cast<const bf16*>(y), ldy, static_cast<bf16*>(dskip), ldd, g, s.d,
      s.h, s.w, accumulate);
  SPFF_CUDA(cudaGetLastError());
  return 0;
}

int spff_maxpool_bwd_add_argmax(const void* dpool, long long ldp, const uint8_t* pool_argmax, void* dskip, long long ldd,
                                int c, spff_shape s, int accumulate, void* stream) {
  SPFF_ENTRY_CHECK();
  SPFF_REQUIRE(dpool && pool_argmax && dskip, "maxpool_bwd_add_argmax: null pointer");
  SPFF_REQUIRE(s.h % 2 == 0 && s.w % 2 == 0, "maxpool_bwd_add_argmax: needs even H, W");
  SPFF_REQUIRE(c % 8 == 0 && (reinterpret_cast<uintptr_t>(pool_argmax) & 7) == 0, "maxpool_bwd_add_argmax: c %% 8 and 8-byte aligned codes");
  PlaneGrid g;
  dim3 grid;
  int e = spff::make_grid(c, static_cast<long long>(s.h / 2) * (s.w / 2), s, &g, &grid, 32);
  if (e) return e;
  spff::maxpool_bwd_codes_kernel<<<grid, kBlock, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const bf16*>(dpool), ldp, pool_argmax, static_cast<bf16*>(dskip), ldd, g, s.d, c, s.h, s.w, accumulate);
  SPFF_CUDA(cudaGetLastError());
  return 0;
}

}  // extern "C"
