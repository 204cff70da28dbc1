// Small per-step kernels that keep the fused training step free of framework (ATen) launches: the number of valid voxels
// of a label batch (the CE normaliser), the loss scalar from the device-side tally, and the column sums of the conv
// epilogue's per-item statistics partials (bias gradient of the transposed convolutions).
#include "common.h"

namespace spff {
namespace {

template <typename LabelT>
__global__ void count_valid_kernel(const LabelT* __restrict__ labels, long long total, int ignore_index,
                                   unsigned long long* __restrict__ out) {
  unsigned int c = 0;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x)
    c += static_cast<int>(labels[i]) != ignore_index;
  c = __reduce_add_sync(0xffffffffu, c);
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(out, static_cast<unsigned long long>(c));   // integers: order-independent
}

// uint8 labels, 16 per thread and load
__global__ void count_valid_u8x16_kernel(const uint4* __restrict__ labels, long long nvec, int ignore_index,
                                         unsigned long long* __restrict__ out) {
  unsigned int c = 0;
  const unsigned int ign = static_cast<unsigned int>(ignore_index) & 0xffu;
  const bool representable = ignore_index >= 0 && ignore_index <= 255;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < nvec;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const uint4 v = labels[i];
    const unsigned int w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
      for (int b = 0; b < 4; ++b) c += !representable || ((w[k] >> (8 * b)) & 0xffu) != ign;
  }
  c = __reduce_add_sync(0xffffffffu, c);
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(out, static_cast<unsigned long long>(c));
}

// ce_plus_macro_dice_loss from the tally (helpers.py:782-803): CE mean + 0.5 * (1 - mean_{c=1..K-1} dice_c),
// dice_c = (2 tp + s) / (2 tp + fp + fn + s); confusion is [label][argmax].
__global__ void loss_from_tally_kernel(const double* __restrict__ nll, const unsigned long long* __restrict__ count,
                                       const unsigned long long* __restrict__ conf, int K, double smooth,
                                       float* __restrict__ out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double dice = 1.0;
  if (K > 1) {
    double acc = 0.0;
    for (int c = 1; c < K; ++c) {
      double tp = static_cast<double>(conf[c * K + c]), col = 0.0, row = 0.0;
      for (int j = 0; j < K; ++j) {
        col += static_cast<double>(conf[j * K + c]);
        row += static_cast<double>(conf[c * K + j]);
      }
      acc += (2.0 * tp + smooth) / (2.0 * tp + (col - tp) + (row - tp) + smooth);
    }
    dice = acc / (K - 1);
  }
  const unsigned long long n = count[0] > 0 ? count[0] : 1;
  out[0] = static_cast<float>(nll[0] / static_cast<double>(n) + 0.5 * (1.0 - dice));
}

// out[c] += sum over `rows` rows of m[r * row_stride + c], c < cols, in two fixed-order stages: kColChunks row chunks per
// 32 columns (double partials in the workspace), then one block folds the chunks in order.
constexpr int kColChunks = 128;

__global__ void partial_colsum_stage1_kernel(const float* __restrict__ m, long long rows, long long row_stride, int cols,
                                             double* __restrict__ part /* [kColChunks][cols] */) {
  __shared__ double red[8][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + lane;
  const long long r0 = rows * blockIdx.y / kColChunks, r1 = rows * (blockIdx.y + 1) / kColChunks;
  double a = 0.0;
  if (c < cols)
    for (long long r = r0 + warp; r < r1; r += 8) a += static_cast<double>(m[r * row_stride + c]);
  red[warp][lane] = a;
  __syncthreads();
  if (warp == 0 && c < cols) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += red[w][lane];
    part[static_cast<size_t>(blockIdx.y) * cols + c] = t;
  }
}

__global__ void partial_colsum_stage2_kernel(const double* __restrict__ part, int cols, float* __restrict__ out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cols) return;
  double t = 0.0;
  for (int k = 0; k < kColChunks; ++k) t += part[static_cast<size_t>(k) * cols + c];
  out[c] += static_cast<float>(t);
}

// g[i] *= factor / max(count, 1)   (0 when count == 0): the CE normaliser applied to finished gradients
__global__ void scale_by_count_kernel(float* __restrict__ g, long long n, const unsigned long long* __restrict__ count,
                                      float factor) {
  const unsigned long long c = count[0];
  const float sc = c > 0 ? factor / static_cast<float>(c) : 0.f;
  const long long n4 = n / 4;
  float4* g4 = reinterpret_cast<float4*>(g);
  const bool vec = (reinterpret_cast<uintptr_t>(g) & 15) == 0;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  const long long i0 = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (vec) {
    for (long long i = i0; i < n4; i += stride) {
      float4 v = g4[i];
      v.x *= sc; v.y *= sc; v.z *= sc; v.w *= sc;
      g4[i] = v;
    }
    for (long long i = n4 * 4 + i0; i < n; i += stride) g[i] *= sc;
  } else {
    for (long long i = i0; i < n; i += stride) g[i] *= sc;
  }
}

}  // namespace
}  // namespace spff

extern "C" {

int spff_count_valid(const void* labels, int label_bytes, long long total, int ignore_index, unsigned long long* out,
                     void* stream) {
  int e = spff_device_check();
  if (e) return e;
  SPFF_REQUIRE(labels && out && total >= 0 && (label_bytes == 1 || label_bytes == 8), "count_valid: bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  SPFF_CUDA(cudaMemsetAsync(out, 0, sizeof(unsigned long long), st));
  if (total == 0) return 0;
  const int cap = spff::num_sms() * 8;
  if (label_bytes == 1 && (reinterpret_cast<uintptr_t>(labels) & 15) == 0) {
    const long long nvec = total / 16;
    if (nvec > 0) {
      const long long b = (nvec + 255) / 256;
      spff::count_valid_u8x16_kernel<<<static_cast<int>(b < cap ? b : cap), 256, 0, st>>>(static_cast<const uint4*>(labels), nvec,
                                                                                           ignore_index, out);
    }
    const long long rest = total - nvec * 16;
    if (rest > 0)
      spff::count_valid_kernel<uint8_t><<<1, 32, 0, st>>>(static_cast<const uint8_t*>(labels) + nvec * 16, rest, ignore_index, out);
  } else {
    const long long b = (total + 255) / 256;
    const int blocks = static_cast<int>(b < cap ? b : cap);
    if (label_bytes == 1)
      spff::count_valid_kernel<uint8_t><<<blocks, 256, 0, st>>>(static_cast<const uint8_t*>(labels), total, ignore_index, out);
    else
      spff::count_valid_kernel<long long><<<blocks, 256, 0, st>>>(static_cast<const long long*>(labels), total, ignore_index, out);
  }
  SPFF_CUDA(cudaGetLastError());
  return 0;
}

int spff_scale_by_count(float* g, long long n, const unsigned long long* count, float factor, void* stream) {
  int e = spff_device_check();
  if (e) return e;
  SPFF_REQUIRE(g && count && n >= 0, "scale_by_count: bad arguments");
  if (n == 0) return 0;
  const long long b = (n / 4 + 255) / 256 + 1;
  const int cap = spff::num_sms() * 8;
  spff::scale_by_count_kernel<<<static_cast<int>(b < cap ? b : cap), 256, 0, static_cast<cudaStream_t>(stream)>>>(g, n, count,
                                                                                                                  factor);
  SPFF_CUDA(cudaGetLastError());
  return 0;
}

int spff_loss_from_tally(const double* nll, const unsigned long long* count, const unsigned long long* confusion, int k,
                         double smooth, float* out, void* stream) {
  int e = spff_device_check();
  if (e) return e;
  SPFF_REQUIRE(nll && count && confusion && out && k > 0, "loss_from_tally: bad arguments");
  spff::loss_from_tally_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(nll, count, confusion, k, smooth, out);
  SPFF_CUDA(cudaGetLastError());
  return 0;
}

size_t spff_partial_colsum_workspace(int cols) { return static_cast<size_t>(spff::kColChunks) * (cols > 0 ? cols : 0) * sizeof(double); }

int spff_partial_colsum(const float* m, long long rows, long long row_stride, int cols, float* out, void* workspace,
                        size_t workspace_bytes, void* stream) {
  int e = spff_device_check();
  if (e) return e;
  SPFF_REQUIRE(m && out && workspace && rows >= 0 && cols > 0 && row_stride >= cols, "partial_colsum: bad arguments");
  if (workspace_bytes < spff_partial_colsum_workspace(cols)) {
    spff::set_error("partial_colsum: workspace %zu < %zu bytes", workspace_bytes, spff_partial_colsum_workspace(cols));
    return SPFF_ERR_WORKSPACE;
  }
  if (rows == 0) return 0;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  double* part = static_cast<double*>(workspace);
  spff::partial_colsum_stage1_kernel<<<dim3((cols + 31) / 32, spff::kColChunks), 256, 0, st>>>(m, rows, row_stride, cols, part);
  spff::partial_colsum_stage2_kernel<<<(cols + 127) / 128, 128, 0, st>>>(part, cols, out);
  SPFF_CUDA(cudaGetLastError());
  return 0;
}

}  // extern "C"
