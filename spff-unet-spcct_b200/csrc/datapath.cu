// Device side of the data path that feeds the training step:
//   * phantom label rasterisation — the per-pixel Python loop over elliptical ROIs of
//     create_image_and_labels_for_dataset (reference innovative3D/helpers.py:125-129, 197-206);
//   * TrainGridAug (reference innovative3D/datasets.py:134-206) with the separable stripe shuffle
//     (_shuffle_stripes, datasets.py:56-121): flips, rot90 and the stripe permutations compose into ONE gather per
//     sample, out[f][h][w] = in[f][A[u]][B[v]] with (u, v) = (h, w) or (w, h) — the host draws the random
//     decisions in the reference's order and passes the two index tables — fused with the intensity jitter, the
//     Gaussian noise and the visibility stamp, so images and labels are read once and written once.
// Integer / index work is bit-exact with the reference; the jitter uses separately rounded multiply and add like
// the two ATen ops it replaces; only the noise values differ (Philox here, the CPU generator there).
// Bandwidth kernels: coalesced along w on the output side, gathers along rows of the input.
#include "common.h"

#include <math.h>

namespace spff {
namespace {

constexpr int kMaxRois = 64;
struct RoiList {
  int n;
  int v[kMaxRois][5];   // x0, y0, w0, h0, label
};

// labels[f][py][px] (int64) = label of the LAST roi whose ellipse contains the pixel, 0 if none. The reference
// writes lb_arr[f, py, px] for px in range(x0, x0+w0), py in range(y0, y0+h0) (numpy indexing: negative indices
// wrap once), testing ((px-cx)**2)/(a*a) + ((py-cy)**2)/(b*b) <= 1 in double with cx = x0 + w0/2, a = w0/2.
__global__ void roi_labels_kernel(RoiList rois, int frames, int height, int width, long long* __restrict__ labels) {
  const int px_out = blockIdx.x * blockDim.x + threadIdx.x;
  const int py_out = blockIdx.y;
  if (px_out >= width) return;
  long long lab = 0;
  for (int r = 0; r < rois.n; ++r) {
    const int x0 = rois.v[r][0], y0 = rois.v[r][1], w0 = rois.v[r][2], h0 = rois.v[r][3];
    // the loop index that lands on this pixel: itself, or itself - extent when the loop index is negative
#pragma unroll
    for (int wrapx = 0; wrapx < 2; ++wrapx) {
      const int px = px_out - wrapx * width;
      if (px < x0 || px >= x0 + w0 || (wrapx && px >= 0)) continue;
#pragma unroll
      for (int wrapy = 0; wrapy < 2; ++wrapy) {
        const int py = py_out - wrapy * height;
        if (py < y0 || py >= y0 + h0 || (wrapy && py >= 0)) continue;
        const double cx = __dadd_rn(static_cast<double>(x0), static_cast<double>(w0) / 2.0);
        const double cy = __dadd_rn(static_cast<double>(y0), static_cast<double>(h0) / 2.0);
        const double a = static_cast<double>(w0) / 2.0, b = static_cast<double>(h0) / 2.0;
        const double dx = __dsub_rn(static_cast<double>(px), cx), dy = __dsub_rn(static_cast<double>(py), cy);
        const double t = __dadd_rn(__ddiv_rn(__dmul_rn(dx, dx), __dmul_rn(a, a)), __ddiv_rn(__dmul_rn(dy, dy), __dmul_rn(b, b)));
        if (t <= 1.0) lab = rois.v[r][4];
      }
    }
  }
  for (int f = 0; f < frames; ++f) labels[(static_cast<long long>(f) * height + py_out) * width + px_out] = lab;
}

// ---- Philox4x32-10 + Box-Muller: one normal per (seed, sample, element) -----------------------------------------
__device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t (&k)[2]) {
  const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
  const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
  const uint32_t n0 = hi1 ^ c[1] ^ k[0], n1 = lo1, n2 = hi0 ^ c[3] ^ k[1], n3 = lo0;
  c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
  k[0] += 0x9E3779B9u; k[1] += 0xBB67AE85u;
}
__device__ __forceinline__ float philox_normal(unsigned long long seed, unsigned long long idx) {
  uint32_t c[4] = {static_cast<uint32_t>(idx), static_cast<uint32_t>(idx >> 32), 0u, 0u};
  uint32_t k[2] = {static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32)};
#pragma unroll
  for (int r = 0; r < 10; ++r) philox_round(c, k);
  const float u1 = (static_cast<float>(c[0]) + 0.5f) * 2.3283064365386963e-10f;   // (0, 1)
  const float u2 = (static_cast<float>(c[1]) + 0.5f) * 2.3283064365386963e-10f;
  return sqrtf(-2.0f * __logf(u1)) * __cosf(6.283185307179586f * u2);
}

__device__ __forceinline__ int float_order(float f) {   // monotone float -> int for atomicMax
  const int i = __float_as_int(f);
  return i >= 0 ? i : i ^ 0x7fffffff;
}
__device__ __forceinline__ float order_float(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

// per-sample {sum, sum of squares} of the input (double): the noise amplitude is min(noise_std, 0.25 * std(x))
__global__ void __launch_bounds__(256) sample_stats_kernel(const float* __restrict__ x, long long per_sample,
                                                           double* __restrict__ stats) {
  __shared__ double red[2][8];
  const long long n = blockIdx.y;
  const float* xs = x + n * per_sample;
  double s = 0, q = 0;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < per_sample;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const double v = xs[i];
    s += v;
    q += v * v;
  }
  for (int off = 16; off > 0; off >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, off);
    q += __shfl_xor_sync(0xffffffffu, q, off);
  }
  if ((threadIdx.x & 31) == 0) {
    red[0][threadIdx.x >> 5] = s;
    red[1][threadIdx.x >> 5] = q;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double ts = 0, tq = 0;
    for (int w = 0; w < 8; ++w) {
      ts += red[0][w];
      tq += red[1][w];
    }
    atomicAdd(stats + 2 * n, ts);
    atomicAdd(stats + 2 * n + 1, tq);
  }
}

struct AugParams {
  int n, frames, h, w;
  const int* amap;          // [n][h] source row (or column, transposed samples) per output index u
  const int* bmap;          // [n][w]
  const int* transposed;    // [n]
  const float* scale;       // [n] jitter (1 = none)
  const float* shift;       // [n]
  const float* noise_cap;   // [n] noise_std of the sample, 0 = no noise
  const unsigned long long* seed;   // [n]
  const double* stats;      // [n][2] or null
  int* maxima;              // [n][2] ordered-int {max over the stamp region of frame 0, max |x|}, or null
};

template <typename LabelT>
__global__ void __launch_bounds__(256) grid_aug_kernel(const float* __restrict__ x, const LabelT* __restrict__ y,
                                                       float* __restrict__ xo, LabelT* __restrict__ yo, AugParams p) {
  const int n = blockIdx.y;
  const int pos = blockIdx.x * blockDim.x + threadIdx.x;     // flattened (h, w): full blocks whatever the row length
  const bool ok = pos < p.h * p.w;
  const int hh = ok ? pos / p.w : 0, ww = ok ? pos % p.w : 0;
  const int tr = p.transposed[n];
  int sh = 0, sw = 0;
  if (ok) {
    const int u = tr ? ww : hh, v = tr ? hh : ww;
    sh = p.amap[static_cast<long long>(n) * p.h + u];
    sw = p.bmap[static_cast<long long>(n) * p.w + v];
  }
  const float scale = p.scale[n], shift = p.shift[n];
  const bool jitter = !(scale == 1.0f && shift == 0.0f);
  float nstd = 0.f;
  if (p.noise_cap[n] > 0.f && p.stats) {
    const double cnt = static_cast<double>(p.frames) * p.h * p.w;
    const double mean = p.stats[2 * n] / cnt;
    double var = (p.stats[2 * n + 1] - cnt * mean * mean) / (cnt - 1.0);   // unbiased, as Tensor.std()
    if (var < 0) var = 0;
    const float v = fabsf(scale) * static_cast<float>(sqrt(var));            // std after the jitter
    nstd = v > 0.f ? fminf(p.noise_cap[n], 0.25f * v) : 0.f;
  }
  float region_max = -INFINITY, abs_max = 0.f;
  const long long plane = static_cast<long long>(p.h) * p.w;
  for (int f = 0; f < p.frames; ++f) {
    if (!ok) break;
    const long long src = (static_cast<long long>(n) * p.frames + f) * plane + static_cast<long long>(sh) * p.w + sw;
    const long long dst = (static_cast<long long>(n) * p.frames + f) * plane + static_cast<long long>(hh) * p.w + ww;
    float v = __ldg(x + src);
    if (jitter) v = __fadd_rn(__fmul_rn(v, scale), shift);                   // two ATen ops: no contraction
    if (nstd > 0.f) v = __fadd_rn(v, __fmul_rn(philox_normal(p.seed[n], static_cast<unsigned long long>(dst)), nstd));
    xo[dst] = v;
    if (y) yo[dst] = __ldg(y + src);
    abs_max = fmaxf(abs_max, fabsf(v));
    if (f == 0 && hh < 32 && ww < 32) region_max = fmaxf(region_max, v);
  }
  if (p.maxima) {
    for (int off = 16; off > 0; off >>= 1) {
      abs_max = fmaxf(abs_max, __shfl_xor_sync(0xffffffffu, abs_max, off));
      region_max = fmaxf(region_max, __shfl_xor_sync(0xffffffffu, region_max, off));
    }
    if ((threadIdx.x & 31) == 0) {
      atomicMax(p.maxima + 2 * n + 1, float_order(abs_max));
      if (region_max > -INFINITY) atomicMax(p.maxima + 2 * n, float_order(region_max));
    }
  }
}

// x[n][0][:32][:32] = max(region) + max(max|x|, 1) * 0.25 for the stamped samples (datasets.py:196-201)
__global__ void stamp_kernel(float* __restrict__ xo, const int* __restrict__ stamp, const int* __restrict__ maxima, int frames,
                             int h, int w) {
  const int n = blockIdx.x;
  if (!stamp[n]) return;
  const float val = __fadd_rn(order_float(maxima[2 * n]), __fmul_rn(fmaxf(order_float(maxima[2 * n + 1]), 1.0f), 0.25f));
  const int rh = h < 32 ? h : 32, rw = w < 32 ? w : 32;
  float* base = xo + static_cast<long long>(n) * frames * h * w;
  for (int i = threadIdx.x; i < rh * rw; i += blockDim.x) base[(i / rw) * w + (i % rw)] = val;
}

__global__ void init_maxima_kernel(int* maxima, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    maxima[2 * i] = float_order(-INFINITY);
    maxima[2 * i + 1] = float_order(0.f);
  }
}

}  // namespace
}  // namespace spff

extern "C" {

int spff_roi_labels(const int* rois_host, int nroi, int frames, int height, int width, long long* labels, void* stream) {
  int e = spff_device_check();
  if (e) return e;
  SPFF_REQUIRE(labels && frames > 0 && height > 0 && width > 0 && nroi >= 0 && (nroi == 0 || rois_host),
               "roi_labels: bad arguments");
  SPFF_REQUIRE(nroi <= spff::kMaxRois, "roi_labels: at most %d rois (got %d)", spff::kMaxRois, nroi);
  spff::RoiList L;
  L.n = nroi;
  for (int r = 0; r < nroi; ++r) {
    for (int k = 0; k < 5; ++k) L.v[r][k] = rois_host[5 * r + k];
    const int x0 = L.v[r][0], y0 = L.v[r][1], w0 = L.v[r][2], h0 = L.v[r][3];
    // numpy would raise IndexError outside [-extent, extent)
    SPFF_REQUIRE(w0 <= 0 || h0 <= 0 || (x0 >= -width && x0 + w0 <= width && y0 >= -height && y0 + h0 <= height),
                 "roi_labels: roi %d (%d,%d,%d,%d) leaves the %dx%d image (IndexError in the reference)", r, x0, y0, w0, h0,
                 width, height);
  }
  dim3 grid((width + 127) / 128, height);
  spff::roi_labels_kernel<<<grid, 128, 0, static_cast<cudaStream_t>(stream)>>>(L, frames, height, width, labels);
  SPFF_CUDA(cudaGetLastError());
  return 0;
}

size_t spff_grid_aug_workspace(int n) { return n > 0 ? static_cast<size_t>(n) * (2 * sizeof(double) + 2 * sizeof(int)) : 0; }

int spff_grid_aug(const float* x, const void* y, int label_bytes, float* xo, void* yo, int n, int frames, int h, int w,
                  const int* amap, const int* bmap, const int* transposed, const float* scale, const float* shift,
                  const float* noise_cap, const unsigned long long* seed, const int* stamp, int any_noise, int any_stamp,
                  void* workspace, size_t workspace_bytes, void* stream) {
  int e = spff_device_check();
  if (e) return e;
  SPFF_REQUIRE(x && xo && amap && bmap && transposed && scale && shift && noise_cap && seed && stamp, "grid_aug: null pointer");
  SPFF_REQUIRE(x != xo && (!y || y != yo), "grid_aug: the gather cannot run in place");
  SPFF_REQUIRE((y == nullptr) == (yo == nullptr) && (!y || label_bytes == 1 || label_bytes == 8),
               "grid_aug: labels must be uint8 or int64, given with their output");
  SPFF_REQUIRE(n > 0 && n <= 65535 && frames > 0 && h > 0 && w > 0 && static_cast<long long>(h) * w < (1LL << 30),
               "grid_aug: bad shape");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  double* stats = nullptr;
  int* maxima = nullptr;
  if (any_noise || any_stamp) {
    if (!workspace || workspace_bytes < spff_grid_aug_workspace(n)) {
      spff::set_error("grid_aug: workspace %zu < %zu bytes", workspace_bytes, spff_grid_aug_workspace(n));
      return SPFF_ERR_WORKSPACE;
    }
    stats = static_cast<double*>(workspace);
    maxima = reinterpret_cast<int*>(stats + 2 * static_cast<size_t>(n));
  }
  const long long per = static_cast<long long>(frames) * h * w;
  if (any_noise) {
    SPFF_CUDA(cudaMemsetAsync(stats, 0, 2 * sizeof(double) * n, st));
    long long bx = (per + 256 * 16 - 1) / (256 * 16);
    const long long cap = (8LL * spff::num_sms() + n - 1) / n;
    if (bx > cap) bx = cap;
    if (bx < 1) bx = 1;
    spff::sample_stats_kernel<<<dim3(static_cast<unsigned>(bx), n), 256, 0, st>>>(x, per, stats);
  }
  if (any_stamp) spff::init_maxima_kernel<<<(n + 127) / 128, 128, 0, st>>>(maxima, n);
  spff::AugParams p{n, frames, h, w, amap, bmap, transposed, scale, shift, noise_cap, seed, any_noise ? stats : nullptr,
                    any_stamp ? maxima : nullptr};
  dim3 grid((h * w + 255) / 256, n);
  if (!y)
    spff::grid_aug_kernel<unsigned char><<<grid, 256, 0, st>>>(x, nullptr, xo, nullptr, p);
  else if (label_bytes == 1)
    spff::grid_aug_kernel<unsigned char><<<grid, 256, 0, st>>>(x, static_cast<const unsigned char*>(y), xo,
                                                               static_cast<unsigned char*>(yo), p);
  else
    spff::grid_aug_kernel<long long><<<grid, 256, 0, st>>>(x, static_cast<const long long*>(y), xo, static_cast<long long*>(yo), p);
  if (any_stamp) spff::stamp_kernel<<<n, 256, 0, st>>>(xo, stamp, maxima, frames, h, w);
  SPFF_CUDA(cudaGetLastError());
  return 0;
}

}  // extern "C"
