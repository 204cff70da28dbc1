// Kernels that only the "3DUNet" control needs (Cicek3DUNet + depth adapter, reference
// innovative3D/models.py:718-777, 844-846): depth-only resampling (trilinear with H, W unchanged is a
// [Dout x Din] matrix over the planes, models.py:153-163), BatchNorm3d training / eval coefficients
// and backward coefficients (models.py:721 `nn.BatchNorm3d(c)`), MaxPool3d(2) forward / backward
// (models.py:728-731) and the SGD-with-momentum update (models.py:844-846). The 3x3x3 convolutions,
// the normalise + ReLU passes, the head and the loss run through the same kernels as SPFF-UNet
// (BatchNorm coefficients are broadcast to the per-sample layout coef[n][c][4] those kernels read).
// All of these are bandwidth kernels: 16-byte accesses, channels innermost, one pass over the data.
#include "common.h"

#include <cuda_bf16.h>

namespace spff {
namespace {

constexpr int kThreadsEw = 256;
constexpr int kMaxPlanesRs = 32;   // planes on either side of a depth resample

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

// ---------------------------------------------------------------------------------------------
// y[n][do][e] = sum_di M[do][di] * x[n][di][e],  e over the `inner` elements of a plane.
// A thread owns one 16-byte vector of e for one sample; every input vector it needs is read once per
// output plane that uses it (2 taps for linear interpolation; re-reads hit L1/L2), every output
// written once. BF16 = 8 elements per vector, fp32 = 4.
// ---------------------------------------------------------------------------------------------
template <bool BF16>
__global__ void __launch_bounds__(kThreadsEw) depth_resample_kernel(const void* __restrict__ xv, void* __restrict__ yv,
                                                                    const float* __restrict__ M, int din, int dout,
                                                                    long long inner_vec) {
  __shared__ float m[kMaxPlanesRs * kMaxPlanesRs];
  for (int i = threadIdx.x; i < din * dout; i += blockDim.x) m[i] = M[i];
  __syncthreads();
  const long long n = blockIdx.y;
  const uint4* x = static_cast<const uint4*>(xv) + n * din * inner_vec;
  uint4* y = static_cast<uint4*>(yv) + n * dout * inner_vec;
  for (long long e = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; e < inner_vec;
       e += static_cast<long long>(gridDim.x) * blockDim.x) {
    for (int o = 0; o < dout; ++o) {
      float acc[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] = 0.f;
      for (int k = 0; k < din; ++k) {
        const float wgt = m[o * din + k];
        if (wgt == 0.f) continue;
        const uint4 raw = __ldg(x + k * inner_vec + e);
        if (BF16) {
          float f[8];
          unpack8(raw, f);
#pragma unroll
          for (int i = 0; i < 8; ++i) acc[i] = fmaf(wgt, f[i], acc[i]);
        } else {
          acc[0] = fmaf(wgt, __uint_as_float(raw.x), acc[0]);
          acc[1] = fmaf(wgt, __uint_as_float(raw.y), acc[1]);
          acc[2] = fmaf(wgt, __uint_as_float(raw.z), acc[2]);
          acc[3] = fmaf(wgt, __uint_as_float(raw.w), acc[3]);
        }
      }
      if (BF16) {
        y[o * inner_vec + e] = pack8(acc);
      } else {
        y[o * inner_vec + e] = make_uint4(__float_as_uint(acc[0]), __float_as_uint(acc[1]), __float_as_uint(acc[2]),
                                          __float_as_uint(acc[3]));
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// BatchNorm3d statistics over the whole batch, in a fixed order.
// Stage 1: block (32 channels x 8 row-warps, blockIdx.y = split) sums its rows of either
//   partial[row][2][c] fp32 (rows = n * slots, the conv epilogue's per-item {sum, sumsq}) or
//   stats[row][c][2] fp64 (rows = n, spff_in_stats) into ws[split][c][2] (double).
// Stage 2: one thread per channel folds the splits, updates the running statistics and broadcasts
//   coef[k][c] = {A, B, mean, rstd} to every sample k.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) bn_sum_kernel(const float* __restrict__ partial, const double* __restrict__ stats,
                                                     long long rows, int c, double* __restrict__ ws) {
  __shared__ double red[8][32][2];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int ch = blockIdx.x * 32 + lane;
  const long long r0 = rows * blockIdx.y / gridDim.y, r1 = rows * (blockIdx.y + 1) / gridDim.y;
  double s = 0, q = 0;
  if (ch < c) {
    for (long long r = r0 + warp; r < r1; r += 8) {
      if (partial) {
        s += static_cast<double>(partial[(r * 2) * c + ch]);
        q += static_cast<double>(partial[(r * 2 + 1) * c + ch]);
      } else {
        s += stats[(r * c + ch) * 2];
        q += stats[(r * c + ch) * 2 + 1];
      }
    }
  }
  red[warp][lane][0] = s;
  red[warp][lane][1] = q;
  __syncthreads();
  if (warp == 0 && ch < c) {
    double ts = 0, tq = 0;
    for (int w = 0; w < 8; ++w) {
      ts += red[w][lane][0];
      tq += red[w][lane][1];
    }
    ws[(static_cast<long long>(blockIdx.y) * c + ch) * 2] = ts;
    ws[(static_cast<long long>(blockIdx.y) * c + ch) * 2 + 1] = tq;
  }
}

__global__ void bn_finish_kernel(const double* __restrict__ ws, int splits, const float* __restrict__ gamma,
                                 const float* __restrict__ beta, float eps, int n, int c, double count, float momentum,
                                 float* __restrict__ running_mean, float* __restrict__ running_var, int eval,
                                 float* __restrict__ coef) {
  const int ch = blockIdx.x * blockDim.x + threadIdx.x;
  if (ch >= c) return;
  double mean, var;
  if (eval) {   // inference: the running statistics are the statistics
    mean = running_mean[ch];
    var = running_var[ch];
  } else {
    double s = 0, q = 0;
    for (int k = 0; k < splits; ++k) {
      s += ws[(static_cast<long long>(k) * c + ch) * 2];
      q += ws[(static_cast<long long>(k) * c + ch) * 2 + 1];
    }
    mean = s / count;
    var = q / count - mean * mean;
    if (var < 0) var = 0;
    if (running_mean) {   // F.batch_norm(training=True): biased variance normalises, unbiased one is tracked
      const double unbiased = count > 1 ? var * count / (count - 1) : var;
      running_mean[ch] = static_cast<float>((1.0 - momentum) * running_mean[ch] + momentum * mean);
      running_var[ch] = static_cast<float>((1.0 - momentum) * running_var[ch] + momentum * unbiased);
    }
  }
  const float rstd = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
  const float A = rstd * (gamma ? gamma[ch] : 1.f);
  const float4 o = make_float4(A, (beta ? beta[ch] : 0.f) - static_cast<float>(mean) * A, static_cast<float>(mean), rstd);
  for (int k = 0; k < n; ++k) reinterpret_cast<float4*>(coef)[static_cast<long long>(k) * c + ch] = o;
}

// BatchNorm backward coefficients from the plane sums R[n][d][c][6] (slots 2 = sum dz, 4 = sum dz*xhat):
//   bcoef[k][c] = {gamma*rstd, mean(dz), mean(dz*xhat), 0} for every sample k (means over the BATCH),
//   dgamma[c] += sum dz*xhat, dbeta[c] += sum dz. One block per 32 channels, fixed summation order.
__global__ void __launch_bounds__(256) bn_bwd_coeffs_kernel(const float* __restrict__ R, const float* __restrict__ coef,
                                                            const float* __restrict__ gamma, int n, int d, int c,
                                                            double inv_count, float* __restrict__ bcoef,
                                                            float* __restrict__ dgamma, float* __restrict__ dbeta) {
  __shared__ double red[8][32][2];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int ch = blockIdx.x * 32 + lane;
  const long long rows = static_cast<long long>(n) * d;
  double s = 0, q = 0;
  if (ch < c) {
    for (long long r = warp; r < rows; r += 8) {
      const float* p = R + (r * c + ch) * 6;
      s += static_cast<double>(p[2]);
      q += static_cast<double>(p[4]);
    }
  }
  red[warp][lane][0] = s;
  red[warp][lane][1] = q;
  __syncthreads();
  if (ch >= c) return;
  double ts = 0, tq = 0;
  for (int w = 0; w < 8; ++w) {
    ts += red[w][lane][0];
    tq += red[w][lane][1];
  }
  const float4 cf = reinterpret_cast<const float4*>(coef)[ch];   // sample 0: all samples hold the same coefficients
  const float4 bc = make_float4((gamma ? gamma[ch] : 1.f) * cf.w, static_cast<float>(ts * inv_count),
                                static_cast<float>(tq * inv_count), 0.f);
  for (int k = warp; k < n; k += 8) reinterpret_cast<float4*>(bcoef)[static_cast<long long>(k) * c + ch] = bc;
  if (warp == 0) {
    if (dgamma) dgamma[ch] += static_cast<float>(tq);
    if (dbeta) dbeta[ch] += static_cast<float>(ts);
  }
}

// ---------------------------------------------------------------------------------------------
// MaxPool3d(2): a thread owns one pooled position x 8 channels.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreadsEw) maxpool222_fwd_kernel(const __nv_bfloat16* __restrict__ y, long long ldy,
                                                                    __nv_bfloat16* __restrict__ yp, long long ldp, int c8,
                                                                    int n, int d, int h, int w) {
  const int dp = d / 2, hp = h / 2, wp = w / 2;
  const long long total = static_cast<long long>(n) * dp * hp * wp * c8;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    long long r = i;
    const int v = static_cast<int>(r % c8); r /= c8;
    const int ww = static_cast<int>(r % wp); r /= wp;
    const int hh = static_cast<int>(r % hp); r /= hp;
    const int dd = static_cast<int>(r % dp);
    const long long nn = r / dp;
    uint4 raw[8];
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      const long long pos = ((nn * d + 2 * dd + (t >> 2)) * h + 2 * hh + ((t >> 1) & 1)) * w + 2 * ww + (t & 1);
      raw[t] = __ldg(reinterpret_cast<const uint4*>(y + pos * ldy + v * 8));
    }
    float best[8];
    unpack8(raw[0], best);
#pragma unroll
    for (int t = 1; t < 8; ++t) {
      float f[8];
      unpack8(raw[t], f);
#pragma unroll
      for (int k = 0; k < 8; ++k) best[k] = (f[k] > best[k] || f[k] != f[k]) ? f[k] : best[k];   // ATen: NaN propagates
    }
    const long long pp = ((nn * dp + dd) * hp + hh) * wp + ww;
    *reinterpret_cast<uint4*>(yp + pp * ldp + v * 8) = pack8(best);
  }
}

// dskip (full res) = (accumulate ? dskip : 0) + dpool scattered to the arg-max of y in each 2x2x2 window
// (first maximum in (d, h, w) scan order wins, as ATen's max_pool3d_with_indices).
__global__ void __launch_bounds__(kThreadsEw) maxpool222_bwd_kernel(const __nv_bfloat16* __restrict__ dpool, long long ldp,
                                                                    const __nv_bfloat16* __restrict__ y, long long ldy,
                                                                    __nv_bfloat16* __restrict__ dskip, long long ldd, int c8,
                                                                    int n, int d, int h, int w, int accumulate) {
  const int dp = d / 2, hp = h / 2, wp = w / 2;
  const long long total = static_cast<long long>(n) * dp * hp * wp * c8;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    long long r = i;
    const int v = static_cast<int>(r % c8); r /= c8;
    const int ww = static_cast<int>(r % wp); r /= wp;
    const int hh = static_cast<int>(r % hp); r /= hp;
    const int dd = static_cast<int>(r % dp);
    const long long nn = r / dp;
    const long long pp = ((nn * dp + dd) * hp + hh) * wp + ww;
    float g[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(dpool + pp * ldp + v * 8)), g);
    long long pos[8];
    uint4 raw[8];
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      pos[t] = ((nn * d + 2 * dd + (t >> 2)) * h + 2 * hh + ((t >> 1) & 1)) * w + 2 * ww + (t & 1);
      raw[t] = __ldg(reinterpret_cast<const uint4*>(y + pos[t] * ldy + v * 8));
    }
    float best[8];
    int arg[8];
    unpack8(raw[0], best);
#pragma unroll
    for (int k = 0; k < 8; ++k) arg[k] = 0;
#pragma unroll
    for (int t = 1; t < 8; ++t) {
      float f[8];
      unpack8(raw[t], f);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        if (f[k] > best[k] || f[k] != f[k]) {
          best[k] = f[k];
          arg[k] = t;
        }
      }
    }
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      float o[8];
      uint4* dst = reinterpret_cast<uint4*>(dskip + pos[t] * ldd + v * 8);
      if (accumulate) {
        unpack8(*dst, o);
      } else {
#pragma unroll
        for (int k = 0; k < 8; ++k) o[k] = 0.f;
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) o[k] += (arg[k] == t) ? g[k] : 0.f;
      *dst = pack8(o);
    }
  }
}

// torch.optim.SGD(momentum, dampening 0, nesterov, weight_decay): g = grad*scale + wd*p;
// buf = first ? g : momentum*buf + g; p -= lr * (nesterov ? g + momentum*buf : buf).
__global__ void sgd_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ buf, long long n,
                           float lr, float momentum, float weight_decay, int nesterov, int first, float grad_scale) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    float gi = g[i] * grad_scale;
    const float pi = p[i];
    if (weight_decay != 0.f) gi = fmaf(weight_decay, pi, gi);
    float step = gi;
    if (momentum != 0.f) {
      const float b = first ? gi : fmaf(momentum, buf[i], gi);
      buf[i] = b;
      step = nesterov ? fmaf(momentum, b, gi) : b;
    }
    p[i] = pi - lr * step;
  }
}

int ew_blocks(long long total) {
  long long b = (total + kThreadsEw - 1) / kThreadsEw;
  const long long cap = 16LL * num_sms();
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return static_cast<int>(b);
}

}  // namespace
}  // namespace spff

extern "C" {

int spff_depth_resample(const void* x, void* y, int elem_bytes, int n, int din, int dout, long long inner,
                        const float* matrix, void* stream) {
  int e = spff_device_check();
  if (e) return e;
  SPFF_REQUIRE(x && y && matrix, "depth_resample: null pointer");
  SPFF_REQUIRE(elem_bytes == 2 || elem_bytes == 4, "depth_resample: elements must be bf16 (2) or fp32 (4)");
  SPFF_REQUIRE(n > 0 && n <= 65535 && din > 0 && dout > 0 && din <= spff::kMaxPlanesRs && dout <= spff::kMaxPlanesRs,
               "depth_resample: bad shape n %d, %d -> %d planes (<= %d)", n, din, dout, spff::kMaxPlanesRs);
  const int per = 16 / elem_bytes;
  SPFF_REQUIRE(inner > 0 && inner % per == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(y) & 15) == 0,
               "depth_resample: planes must be 16-byte aligned multiples of 16 bytes (inner %lld)", inner);
  const long long iv = inner / per;
  long long bx = (iv + spff::kThreadsEw - 1) / spff::kThreadsEw;
  const long long cap = (16LL * spff::num_sms() + n - 1) / n;
  if (bx > cap) bx = cap;
  if (bx < 1) bx = 1;
  dim3 grid(static_cast<unsigned>(bx), n);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (elem_bytes == 2)
    spff::depth_resample_kernel<true><<<grid, spff::kThreadsEw, 0, st>>>(x, y, matrix, din, dout, iv);
  else
    spff::depth_resample_kernel<false><<<grid, spff::kThreadsEw, 0, st>>>(x, y, matrix, din, dout, iv);
  SPFF_CUDA(cudaGetLastError());
  return 0;
}

size_t spff_bn_coeffs_workspace(int c) { return c > 0 ? static_cast<size_t>(64) * c * 2 * sizeof(double) : 0; }

int spff_bn_coeffs(const float* partial, int slots, const double* stats, const float* gamma, const float* beta, float eps,
                   int n, int c, long long count, float momentum, float* running_mean, float* running_var, int eval,
                   float* coef, void* workspace, size_t workspace_bytes, void* stream) {
  int e = spff_device_check();
  if (e) return e;
  SPFF_REQUIRE(coef && n > 0 && c > 0 && count > 0, "bn_coeffs: bad arguments");
  SPFF_REQUIRE(eval ? (running_mean && running_var) : ((partial != nullptr) != (stats != nullptr)),
               "bn_coeffs: training needs exactly one of partial / stats, eval needs the running statistics");
  SPFF_REQUIRE((running_mean == nullptr) == (running_var == nullptr), "bn_coeffs: running_mean and running_var go together");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int splits = 0;
  if (!eval) {
    SPFF_REQUIRE(!partial || slots > 0, "bn_coeffs: slots must be positive");
    if (!workspace || workspace_bytes < spff_bn_coeffs_workspace(c)) {
      spff::set_error("bn_coeffs: workspace %zu < %zu bytes", workspace_bytes, spff_bn_coeffs_workspace(c));
      return SPFF_ERR_WORKSPACE;
    }
    const long long rows = partial ? static_cast<long long>(n) * slots : n;
    splits = static_cast<int>(rows / 64 < 1 ? 1 : (rows / 64 > 64 ? 64 : rows / 64));
    dim3 grid((c + 31) / 32, splits);
    spff::bn_sum_kernel<<<grid, 256, 0, st>>>(partial, stats, rows, c, static_cast<double*>(workspace));
  }
  spff::bn_finish_kernel<<<(c + 127) / 128, 128, 0, st>>>(static_cast<const double*>(workspace), splits, gamma, beta, eps, n,
                                                         c, static_cast<double>(count) * n, momentum, running_mean,
                                                         running_var, eval, coef);
  SPFF_CUDA(cudaGetLastError());
  return 0;
}

int spff_bn_bwd_coeffs(const float* R, const float* coef, const float* gamma, int c, spff_shape s, float* bcoef,
                       float* dgamma, float* dbeta, void* stream) {
  int e = spff_device_check();
  if (e) return e;
  SPFF_REQUIRE(R && coef && bcoef && c > 0 && s.n > 0 && s.d > 0 && s.h > 0 && s.w > 0, "bn_bwd_coeffs: bad arguments");
  const double inv = 1.0 / (static_cast<double>(s.n) * s.d * s.h * s.w);
  spff::bn_bwd_coeffs_kernel<<<(c + 31) / 32, 256, 0, static_cast<cudaStream_t>(stream)>>>(R, coef, gamma, s.n, s.d, c, inv,
                                                                                           bcoef, dgamma, dbeta);
  SPFF_CUDA(cudaGetLastError());
  return 0;
}

int spff_maxpool222_fwd(const void* y, long long ldy, void* ypool, long long ldp, int c, spff_shape s, void* stream) {
  int e = spff_device_check();
  if (e) return e;
  SPFF_REQUIRE(y && ypool && c > 0 && c % 8 == 0 && ldy % 8 == 0 && ldp % 8 == 0, "maxpool222_fwd: bad arguments");
  SPFF_REQUIRE(s.d % 2 == 0 && s.h % 2 == 0 && s.w % 2 == 0 && s.n > 0, "maxpool222_fwd: needs even D, H, W (got %d x %d x %d)",
               s.d, s.h, s.w);
  const long long total = static_cast<long long>(s.n) * (s.d / 2) * (s.h / 2) * (s.w / 2) * (c / 8);
  spff::maxpool222_fwd_kernel<<<spff::ew_blocks(total), spff::kThreadsEw, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(y), ldy, static_cast<__nv_bfloat16*>(ypool), ldp, c / 8, s.n, s.d, s.h, s.w);
  SPFF_CUDA(cudaGetLastError());
  return 0;
}

int spff_maxpool222_bwd_add(const void* dpool, long long ldp, const void* y, long long ldy, void* dskip, long long ldd, int c,
                            spff_shape s, int accumulate, void* stream) {
  int e = spff_device_check();
  if (e) return e;
  SPFF_REQUIRE(dpool && y && dskip && c > 0 && c % 8 == 0 && ldy % 8 == 0 && ldp % 8 == 0 && ldd % 8 == 0,
               "maxpool222_bwd_add: bad arguments");
  SPFF_REQUIRE(s.d % 2 == 0 && s.h % 2 == 0 && s.w % 2 == 0 && s.n > 0, "maxpool222_bwd_add: needs even D, H, W");
  const long long total = static_cast<long long>(s.n) * (s.d / 2) * (s.h / 2) * (s.w / 2) * (c / 8);
  spff::maxpool222_bwd_kernel<<<spff::ew_blocks(total), spff::kThreadsEw, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(dpool), ldp, static_cast<const __nv_bfloat16*>(y), ldy,
      static_cast<__nv_bfloat16*>(dskip), ldd, c / 8, s.n, s.d, s.h, s.w, accumulate);
  SPFF_CUDA(cudaGetLastError());
  return 0;
}

int spff_sgd_step(float* param, const float* grad, float* momentum_buf, long long n, float lr, float momentum,
                  float weight_decay, int nesterov, int first_step, float grad_scale, void* stream) {
  int e = spff_device_check();
  if (e) return e;
  SPFF_REQUIRE(param && grad && n >= 0 && (momentum == 0.f || momentum_buf), "sgd_step: bad arguments");
  if (n == 0) return 0;
  spff::sgd_kernel<<<spff::ew_blocks(n), spff::kThreadsEw, 0, static_cast<cudaStream_t>(stream)>>>(
      param, grad, momentum_buf, n, lr, momentum, weight_decay, nesterov, first_step, grad_scale);
  SPFF_CUDA(cudaGetLastError());
  return 0;
}

}  // extern "C"
