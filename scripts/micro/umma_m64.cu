// (see umma_m64.cu header below)
// Micro-experiment: can a K-major, TMA-swizzled A tile be read by tcgen05.mma from a start address that is shifted by
// a number of rows that is NOT a multiple of the 8-row swizzle atom, and with a group stride (SBO) that is not the
// atom size? This decides whether one (h,w)-halo tile in shared memory can serve all nine spatial taps of the 3x3x3
// convolution through descriptor offsets alone (conv3 "halo" design) instead of one TMA load per kh and a lane-shift
// epilogue for kw.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I../../spff-unet-spcct_b200/csrc umma_rowshift.cu -o umma_rowshift
//
// D = A_shifted * I (identity weights), so D[m][n] must equal X[row(m) + s][n].
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#include "ptx.cuh"

using namespace spff;

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  return reinterpret_cast<EncodeTiledFn>(p);
}

static CUtensorMap make_map(void* base, int cols, int rows, int box_rows, int swizzle_bytes) {
  CUtensorMap m;
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t str[1] = {static_cast<cuuint64_t>(cols) * 2};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(cols), static_cast<cuuint32_t>(box_rows)};
  cuuint32_t es[2] = {1, 1};
  CUresult r = encode_fn()(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    printf("encode failed %d\n", (int)r);
    exit(1);
  }
  return m;
}


// ---------------------------------------------------------------------------------------------------------------
// umma_m64: where do the 64 rows of an M = 64 (cta_group::1) accumulator live in TMEM? Pass 1 fills all 128 lanes with
// an M = 128 MMA over rows 64..191 of X (lane l holds row 64 + l), pass 2 overwrites with an M = 64 MMA over rows
// 0..63; reading all 128 lanes back shows which lane received which row.
__global__ void __launch_bounds__(128, 1)
m64_kernel(const __grid_constant__ CUtensorMap tx, const __grid_constant__ CUtensorMap tw, float* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
  uint8_t* sA = smem;
  uint8_t* sW = smem + 32768;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 49152);
  uint64_t* mbar = bar + 1;
  uint32_t* slot = reinterpret_cast<uint32_t*>(smem + 49152 + 64);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    mbar_init(mbar, 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    tmem_alloc(slot, 32);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot;
  if (threadIdx.x == 0) {
    mbar_expect_tx(bar, 192 * 64 + 32 * 64);
    tma_load_2d(sA, &tx, bar, 0, 0);
    tma_load_2d(sW, &tw, bar, 0, 0);
  }
  mbar_wait(bar, 0);
  tc_fence_after();
  if (threadIdx.x == 0) {
    const uint64_t hi = make_smem_desc_hi(16, 512, kSwizzle64);
    const uint64_t wdesc = smem_desc(hi, smem_u32(sW));
    const uint64_t a128 = smem_desc(hi, smem_u32(sA) + 64 * 64);
    const uint64_t a64 = smem_desc(hi, smem_u32(sA));
    const uint32_t id128 = make_idesc_bf16(128, 32, 0, 0), id64 = make_idesc_bf16(64, 32, 0, 0);
    for (int k = 0; k < 2; ++k) umma_bf16(tmem, a128 + k * 2, wdesc + k * 2, id128, k > 0 ? 1u : 0u);
    for (int k = 0; k < 2; ++k) umma_bf16(tmem, a64 + k * 2, wdesc + k * 2, id64, k > 0 ? 1u : 0u);
    umma_commit(mbar);
  }
  mbar_wait(mbar, 0);
  tc_fence_after();
  uint32_t v[32];
  tmem_ld_32x32(tmem + (static_cast<uint32_t>(warp * 32) << 16), v);
  tmem_ld_wait();
  float* o = out + (warp * 32 + lane) * 32;
  for (int j = 0; j < 32; ++j) o[j] = __uint_as_float(v[j]);
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 32);
}

int main() {
  const int R = 200, kc = 32;
  std::vector<__nv_bfloat16> hx(R * kc), hw(32 * kc);
  for (int r = 0; r < R; ++r)
    for (int c = 0; c < kc; ++c) hx[r * kc + c] = __float2bfloat16(static_cast<float>(r));   // every column of row r holds r
  for (int n = 0; n < 32; ++n)
    for (int k = 0; k < kc; ++k) hw[n * kc + k] = __float2bfloat16(n == k ? 1.f : 0.f);
  __nv_bfloat16 *dx, *dw;
  float* dout;
  cudaMalloc(&dx, hx.size() * 2);
  cudaMalloc(&dw, hw.size() * 2);
  cudaMalloc(&dout, 128 * 32 * 4);
  cudaMemcpy(dx, hx.data(), hx.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dw, hw.data(), hw.size() * 2, cudaMemcpyHostToDevice);
  CUtensorMap tx = make_map(dx, kc, R, 192, 64);
  CUtensorMap tw = make_map(dw, kc, 32, 32, 64);
  cudaFuncSetAttribute(m64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  m64_kernel<<<1, 128, 64 * 1024, 0>>>(tx, tw, dout);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("kernel failed: %s\n", cudaGetErrorString(e));
    return 1;
  }
  std::vector<float> ho(128 * 32);
  cudaMemcpy(ho.data(), dout, ho.size() * 4, cudaMemcpyDeviceToHost);
  printf("lane -> source row of X after [M=128 over rows 64..191] then [M=64 over rows 0..63]:\n");
  for (int l = 0; l < 128; ++l) printf("%s%3d:%3d", (l % 16) ? " " : "\n  ", l, static_cast<int>(ho[l * 32]));
  printf("\n");
  return 0;
}
