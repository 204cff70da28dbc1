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

struct Case {
  int shift_rows;   // start address = base + shift_rows * row_bytes
  int sbo_bytes;    // stride between 8-row groups
  int base_mode;    // 0: base_offset 0; 1: (addr >> 7) & 7
};

constexpr int kMaxCases = 64;
struct Params {
  int ncases;
  Case c[kMaxCases];
  int kc;          // channels per row: 32 (64B swizzle) or 64 (128B swizzle)
  int tile_rows;   // rows loaded
};

__global__ void __launch_bounds__(128, 1)
rowshift_kernel(const __grid_constant__ CUtensorMap tx, const __grid_constant__ CUtensorMap tw, const Params p, float* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
  uint8_t* sA = smem;                      // tile_rows * kc * 2 bytes
  uint8_t* sW = smem + 32768;              // 32 x kc identity
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
  const int row_bytes = p.kc * 2;
  if (threadIdx.x == 0) {
    mbar_expect_tx(bar, p.tile_rows * row_bytes + 32 * row_bytes);
    tma_load_2d(sA, &tx, bar, 0, 0);
    tma_load_2d(sW, &tw, bar, 0, 0);
  }
  mbar_wait(bar, 0);
  tc_fence_after();
  const uint32_t swz = (p.kc == 64) ? kSwizzle128 : kSwizzle64;
  const uint32_t idesc = make_idesc_bf16(128, 32, 0, 0);
  uint32_t par = 0;
  for (int ci = 0; ci < p.ncases; ++ci) {
    const Case c = p.c[ci];
    if (threadIdx.x == 0) {
      const uint32_t a_addr = smem_u32(sA) + c.shift_rows * row_bytes;
      uint64_t ahi = make_smem_desc_hi(16, c.sbo_bytes, swz);
      if (c.base_mode == 1) ahi |= static_cast<uint64_t>((a_addr >> 7) & 7) << 49;
      const uint64_t adesc = smem_desc(ahi, a_addr);
      const uint64_t wdesc = smem_desc(make_smem_desc_hi(16, (p.kc == 64) ? 1024 : 512, swz), smem_u32(sW));
      for (int k = 0; k < p.kc / 16; ++k) umma_bf16(tmem, adesc + k * 2, wdesc + k * 2, idesc, k > 0 ? 1u : 0u);
      umma_commit(mbar);
    }
    mbar_wait(mbar, par);
    par ^= 1;
    tc_fence_after();
    uint32_t v[32];
    tmem_ld_32x32(tmem + (static_cast<uint32_t>(warp * 32) << 16), v);
    tmem_ld_wait();
    float* o = out + (static_cast<size_t>(ci) * 128 + warp * 32 + lane) * 32;
    for (int j = 0; j < 32; ++j) o[j] = __uint_as_float(v[j]);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
  }
  if (warp == 0) tmem_dealloc(tmem, 32);
}

static float xval(int r, int c) { return static_cast<float>((r % 13) - 6) + 0.25f * static_cast<float>(c % 7) + 8.f * ((r / 13) % 5); }

int run(int kc) {
  const int R = 216, tile_rows = 208;
  std::vector<__nv_bfloat16> hx(R * kc), hw(32 * kc);
  for (int r = 0; r < R; ++r)
    for (int c = 0; c < kc; ++c) hx[r * kc + c] = __float2bfloat16(xval(r, c));
  for (int n = 0; n < 32; ++n)
    for (int k = 0; k < kc; ++k) hw[n * kc + k] = __float2bfloat16(n == k ? 1.f : 0.f);   // D[m][n] = A[m][n], n < 32
  __nv_bfloat16 *dx, *dw;
  float* dout;
  cudaMalloc(&dx, hx.size() * 2);
  cudaMalloc(&dw, hw.size() * 2);
  cudaMemcpy(dx, hx.data(), hx.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dw, hw.data(), hw.size() * 2, cudaMemcpyHostToDevice);
  Params p;
  p.kc = kc;
  p.tile_rows = tile_rows;
  p.ncases = 0;
  const int row_bytes = kc * 2;
  const int atom = 8 * row_bytes;
  const int shifts[] = {0, 1, 2, 3, 7, 8, 9, 10, 11, 17};
  const int group_rows[] = {8, 10, 12};      // SBO = group_rows * row_bytes: 8 = dense tile, 10 / 12 = 8-wide box rows of a halo tile
  for (int g : group_rows)
    for (int s : shifts)
      for (int bm = 0; bm < 2; ++bm)
        if (p.ncases < kMaxCases && (s <= 11 || g == 8)) p.c[p.ncases++] = Case{s, g * row_bytes, bm};
  cudaMalloc(&dout, sizeof(float) * p.ncases * 128 * 32);
  cudaMemset(dout, 0xff, sizeof(float) * p.ncases * 128 * 32);
  CUtensorMap tx = make_map(dx, kc, R, tile_rows, row_bytes);
  CUtensorMap tw = make_map(dw, kc, 32, 32, row_bytes);
  cudaFuncSetAttribute(rowshift_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  rowshift_kernel<<<1, 128, 64 * 1024, 0>>>(tx, tw, p, dout);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("kernel failed: %s\n", cudaGetErrorString(e));
    return 1;
  }
  std::vector<float> ho(static_cast<size_t>(p.ncases) * 128 * 32);
  cudaMemcpy(ho.data(), dout, ho.size() * 4, cudaMemcpyDeviceToHost);
  printf("kc=%d (row %d B, swizzle atom %d B)\n", kc, row_bytes, atom);
  for (int ci = 0; ci < p.ncases; ++ci) {
    const Case c = p.c[ci];
    const int g = c.sbo_bytes / row_bytes;
    int bad = 0, first_bad = -1;
    for (int m = 0; m < 128; ++m) {
      const int src = (m / 8) * g + (m % 8) + c.shift_rows;
      for (int n = 0; n < 32; ++n) {
        const float want = xval(src, n), got = ho[(static_cast<size_t>(ci) * 128 + m) * 32 + n];
        if (want != got) {
          ++bad;
          if (first_bad < 0) first_bad = m;
        }
      }
    }
    printf("  group %2d rows  shift %2d  base_offset %s : %s", g, c.shift_rows, c.base_mode ? "(addr>>7)&7" : "0          ",
           bad ? "MISMATCH" : "ok");
    if (bad) printf(" (%d wrong, first row %d)", bad, first_bad);
    printf("\n");
  }
  return 0;
}

int main() {
  if (run(32)) return 1;
  if (run(64)) return 1;
  return 0;
}
