// Weight gradient of the 3x3x3 convolution on tcgen05:
//   dw[co][ci][kd][kh][kw] = sum_{n,d,h,w} dy[n,d,h,w][co] * x[n,d+kd-1,h+kh-1,w+kw-1][ci]
// (the grad_weight half of ATen convolution_backward for `_conv3x3xk`, reference
// innovative3D/models.py:616-618).
//
// GEMM view: the contraction runs over POSITIONS, so both operands are "MN-major" for the tensor
// core: a TMA box [positions][channels] of the position-major activations is used as it lands in
// shared memory (128B / 64B swizzle), A = dy, B = x shifted by the tap.
//   * kw is folded into N: the three kw-shifted x tiles of one (plane, kh) are loaded back to
//     back, so one MMA has N = 3*CIB (96 or 192) — wide enough to feed the tensor pipe from smem.
//   * kd is folded into M: the dy tiles of consecutive planes sit in consecutive 8 KB slots, so an
//     A descriptor with M = 128 spans 4 (COB = 32) or 2 (COB = 64) planes; against the x tile of
//     plane d' the row blocks are the taps kd = 2,1,0,(unused). One x tile thus serves all nine
//     (kd,kw) taps of its kh.
//   * work item = (kh, block of COB output channels, block of CIB input channels, split of the
//     position range); accumulators stay in TMEM for the whole range, partial sums go to a
//     workspace and a second kernel reduces them in a fixed order (deterministic, no atomics).
#include "common.h"
#include "ptx.cuh"

namespace spff {

namespace {

constexpr int kThreads = 192;
constexpr int kSlotBytes = 8192;   // one dy plane tile: KT rows x COB channels x 2 B
constexpr int kMaxG = 5;           // x planes per group
constexpr int kSlots = kMaxG + 3;  // dy planes d0-1 .. d0+G+1

struct WgradParams {
  int n, d, h, w;
  int bw, bh;        // position tile (box) extents; rows = bw*bh (multiple of 16)
  int tiles_w, tiles_h;
  long long ntiles;  // n * tiles_h * tiles_w
  int G, ngroups;
  int ncob, ncib;    // channel blocks
  int ksplit;
  int halo_w;        // all-kh variant: 1 = ONE x tile with a w halo serves the three kw shifts (bw == 8)
  float* partial;    // [item][split][mma][128][3*CIB]
};

template <int COB, int CIB>
struct WgCfg {
  static constexpr int KT = kSlotBytes / (COB * 2);   // max rows per tile: 128 (COB 32) / 64 (COB 64)
  static constexpr int SP = 128 / COB;                // planes stacked in M
  static constexpr int NMMA = (3 + SP - 1) / SP;      // MMAs per x tile to cover kd = 0..2
  static constexpr int N = 3 * CIB;
  static constexpr int XT = KT * CIB * 2;             // one kw tile
  static constexpr int XBytes = 3 * XT;
  static constexpr int DySet = kSlots * kSlotBytes;
  static constexpr int XStages = (XBytes > 24576) ? 2 : 4;  // (COB 32, CIB 64): 2 x 48 KB beside the 128 KB of dy sets
  static constexpr int OffDy = 0;
  static constexpr int OffX = 2 * DySet;
  static constexpr int OffBar = OffX + XStages * XBytes;
  static constexpr int NumBars = 2 * XStages + 4 + 1;
  static constexpr int OffTmem = OffBar + NumBars * 8;
  static constexpr int Total = OffTmem + 16;
  static constexpr int TmemCols = (NMMA * N <= 128) ? 128 : (NMMA * N <= 256 ? 256 : 512);
};

template <int COB, int CIB>
__global__ void __launch_bounds__(kThreads, 1)
conv3_wgrad_kernel(const __grid_constant__ CUtensorMap tmap_dy, const __grid_constant__ CUtensorMap tmap_x,
                   const WgradParams p) {
  using C = WgCfg<COB, CIB>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sDy = smem + C::OffDy;
  uint8_t* sX = smem + C::OffX;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::OffBar);
  uint64_t* xfull = bars;
  uint64_t* xempty = bars + C::XStages;
  uint64_t* dyfull = bars + 2 * C::XStages;
  uint64_t* dyempty = dyfull + 2;
  uint64_t* acc_full = dyempty + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + C::OffTmem);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int i = 0; i < C::XStages; ++i) {
      mbar_init(&xfull[i], 1);
      mbar_init(&xempty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&dyfull[i], 1);
      mbar_init(&dyempty[i], 1);
    }
    mbar_init(acc_full, 1);
    fence_mbar_init();
  }
  if (warp == 4 && lane == 0) {
    tma_prefetch_desc(&tmap_dy);
    tma_prefetch_desc(&tmap_x);
  }
  if (warp == 5) {
    tmem_alloc(tmem_slot, C::TmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // work item of this CTA
  const int split = blockIdx.x % p.ksplit;
  const int item = blockIdx.x / p.ksplit;
  const int cib = item % p.ncib;
  const int cob = (item / p.ncib) % p.ncob;
  const int kh = item / (p.ncib * p.ncob);
  const long long t_begin = p.ntiles * split / p.ksplit;
  const long long t_end = p.ntiles * (split + 1) / p.ksplit;
  const int rows = p.bw * p.bh;
  const uint32_t dy_tile_bytes = rows * COB * 2;
  const uint32_t x_tile_bytes = rows * CIB * 2;

  if (warp == 4) {
    // warp-uniform control flow, one elected lane issues (coordinates stay in uniform registers)
    const bool leader = elect_one() != 0;
    {
      int xs = 0, ds = 0;
      uint32_t xph = 0, dph = 0;
      for (long long t = t_begin; t < t_end; ++t) {
        const int tw = static_cast<int>(t % p.tiles_w);
        const int th = static_cast<int>((t / p.tiles_w) % p.tiles_h);
        const int n = static_cast<int>(t / (static_cast<long long>(p.tiles_w) * p.tiles_h));
        const int w0 = tw * p.bw, h0 = th * p.bh;
        for (int pg = 0; pg < p.ngroups; ++pg) {
          const int d0 = pg * p.G;
          const int dend = min(p.d, d0 + p.G);
          mbar_wait(&dyempty[ds], dph ^ 1);
          if (leader) {
            mbar_expect_tx(&dyfull[ds], (p.G + 3) * dy_tile_bytes);
            for (int sl = 0; sl < p.G + 3; ++sl)
              tma_load_5d(sDy + ds * C::DySet + sl * kSlotBytes, &tmap_dy, &dyfull[ds], cob * COB, w0, h0, d0 - 1 + sl,
                          n);
          }
          if (++ds == 2) {
            ds = 0;
            dph ^= 1;
          }
          for (int dp = d0; dp < dend; ++dp) {
            mbar_wait(&xempty[xs], xph ^ 1);
            if (leader) {
              mbar_expect_tx(&xfull[xs], 3 * x_tile_bytes);
#pragma unroll
              for (int kw = 0; kw < 3; ++kw)
                tma_load_5d(sX + xs * C::XBytes + kw * C::XT, &tmap_x, &xfull[xs], cib * CIB, w0 + kw - 1, h0 + kh - 1,
                            dp, n);
            }
            if (++xs == C::XStages) {
              xs = 0;
              xph ^= 1;
            }
          }
        }
      }
    }
  } else if (warp == 5) {
    const bool leader = elect_one() != 0;
    {
      constexpr uint32_t kSwzA = (COB == 64) ? kSwizzle128 : kSwizzle64;
      constexpr uint32_t kSwzB = (CIB == 64) ? kSwizzle128 : kSwizzle64;
      constexpr uint32_t kSboA = (COB == 64) ? 1024 : 512;
      constexpr uint32_t kSboB = (CIB == 64) ? 1024 : 512;
      // MN-major: LBO = stride between MN blocks (dy: next plane slot; x: next kw tile), SBO = 8 K rows
      const uint64_t adesc_hi = make_smem_desc_hi(kSlotBytes, kSboA, kSwzA);
      const uint64_t bdesc_hi = make_smem_desc_hi(C::XT, kSboB, kSwzB);
      const uint32_t idesc = make_idesc_bf16(128, C::N, 1, 1);
      // COB = 64: the second MMA of an x plane covers the dy planes d'+1 and d'+2, and only d'+1 is a tap (kd = 0). It
      // is issued with M = 64 - same tensor time, but a third less of the A operand crosses shared memory, which is
      // what bounds these MMAs - and not at all when d'+1 lies outside the volume (zero padding). An M = 64 accumulator
      // keeps row r in TMEM lane (r % 16) + 32 * (r / 16) (scripts/micro/umma_m64.cu); the reduction kernel reads it so.
      const uint32_t idesc_half = make_idesc_bf16(64, C::N, 1, 1);
      const int ksteps = rows / 16;
      // descriptors advance by additions on the (address >> 4) field (the smem window is < 256 KB)
      const uint64_t adesc0 = smem_desc(adesc_hi, smem_u32(sDy));
      const uint64_t bdesc0 = smem_desc(bdesc_hi, smem_u32(sX));
      int xs = 0, ds = 0;
      uint32_t xph = 0, dph = 0;
      uint32_t first = 1, first1 = 1;    // accumulator 0 / accumulator 1 not written yet
      for (long long t = t_begin; t < t_end; ++t) {
        for (int pg = 0; pg < p.ngroups; ++pg) {
          const int d0 = pg * p.G;
          const int dend = min(p.d, d0 + p.G);
          mbar_wait(&dyfull[ds], dph);
          tc_fence_after();
          const uint64_t dydesc = adesc0 + static_cast<uint64_t>(ds * (C::DySet >> 4));
          for (int dp = d0; dp < dend; ++dp) {
            mbar_wait(&xfull[xs], xph);
            tc_fence_after();
            const uint64_t xdesc = bdesc0 + static_cast<uint64_t>(xs * (C::XBytes >> 4));
#pragma unroll
            for (int i = 0; i < C::NMMA; ++i) {
              uint64_t ad = dydesc + static_cast<uint64_t>((dp - d0 + i * C::SP) * (kSlotBytes >> 4));
              uint64_t bd = xdesc;
              const bool half = (C::SP == 2) && (i == 1);
              if (half && dp + 1 >= p.d && !first1) continue;     // the only tap of this MMA reads zero padding
              const uint32_t id = half ? idesc_half : idesc;
              const uint32_t fresh = half ? first1 : first;
              for (int ks = 0; ks < ksteps; ++ks) {
                if (leader) umma_bf16(tmem_base + i * C::N, ad, bd, id, (fresh && ks == 0) ? 0u : 1u);
                ad += (2 * kSboA) >> 4;
                bd += (2 * kSboB) >> 4;
              }
              if (half) first1 = 0;
            }
            first = 0;
            if (leader) umma_commit(&xempty[xs]);
            if (++xs == C::XStages) {
              xs = 0;
              xph ^= 1;
            }
          }
          if (leader) umma_commit(&dyempty[ds]);
          if (++ds == 2) {
            ds = 0;
            dph ^= 1;
          }
        }
      }
      if (leader) umma_commit(acc_full);
    }
  } else {
    // epilogue: TMEM -> workspace partials [mma][lane][N]
    mbar_wait(acc_full, 0);
    tc_fence_after();
    float* out = p.partial + static_cast<size_t>(blockIdx.x) * C::NMMA * 128 * C::N;
    const int row = warp * 32 + lane;
#pragma unroll 1
    for (int i = 0; i < C::NMMA; ++i) {
#pragma unroll 1
      for (int c0 = 0; c0 < C::N; c0 += 32) {
        uint32_t v[32];
        tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + i * C::N + c0, v);
        tmem_ld_wait();
        float4* dst = reinterpret_cast<float4*>(out + (static_cast<size_t>(i) * 128 + row) * C::N + c0);
#pragma unroll
        for (int q = 0; q < 8; ++q)
          dst[q] = make_float4(__uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1]), __uint_as_float(v[4 * q + 2]),
                               __uint_as_float(v[4 * q + 3]));
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    tmem_dealloc(tmem_base, C::TmemCols);
  }
}

// dw[co][ci][kd][kh][kw] = beta*dw + sum over splits of the partial tiles (fixed order).
__global__ void conv3_wgrad_reduce_kernel(const float* __restrict__ partial, float* __restrict__ dw, int cout, int cin,
                                          int COB, int CIB, int ksplit, float beta) {
  const int SP = 128 / COB;
  const int NMMA = (3 + SP - 1) / SP;
  const int N = 3 * CIB;
  const int ncob = cout / COB, ncib = cin / CIB;
  const long long total = static_cast<long long>(cout) * cin * 27;
  for (long long e = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; e < total;
       e += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int kw = static_cast<int>(e % 3);
    const int kh = static_cast<int>((e / 3) % 3);
    const int kd = static_cast<int>((e / 9) % 3);
    const int ci = static_cast<int>((e / 27) % cin);
    const int co = static_cast<int>(e / (27LL * cin));
    const int cob = co / COB, col = co % COB, cib = ci / CIB, cic = ci % CIB;
    const int item = (kh * ncob + cob) * ncib + cib;
    const int o = 2 - kd;  // plane offset index: i*SP + b
    const int i = o / SP, b = o % SP;
    // COB = 64: accumulator 1 is written by M = 64 MMAs, whose row r lives in TMEM lane (r % 16) + 32 * (r / 16)
    const int lane = (SP == 2 && i == 1) ? (col % 16) + 32 * (col / 16) : b * COB + col;
    const int column = kw * CIB + cic;
    float acc = 0.f;
    for (int s = 0; s < ksplit; ++s)
      acc += partial[((static_cast<size_t>(item) * ksplit + s) * NMMA + i) * 128 * N + static_cast<size_t>(lane) * N +
                     column];
    dw[e] = (beta == 0.f) ? acc : fmaf(beta, dw[e], acc);
  }
}

// ---------------------------------------------------------------------------------------------
// 32 x 32 channel-block variant with all three kh taps in one CTA. With few channels the kernel above
// is bound by L2 -> SM traffic (ncu: 11.8 TB/s for 32 -> 32: every operand byte crosses 3 x for the kh
// work items, x another 3 x for the kw-shifted tiles). Here ONE x tile with a halo in h and w
// ((bh + 2) x (bw + 2) positions, bw = 8) serves all nine (kh, kw) taps of a plane: the kh-shifted operand
// is the tile read from row kh * (bw + 2), the three kw-shifted 32-channel blocks of N start one row
// (64 bytes) apart (LBO), and the 8-row K groups are the tile's image rows, bw + 2 rows apart (SBO) - the
// tensor core swizzles on the absolute address, so none of these offsets needs atom alignment. dy is
// loaded once per plane. (Round 2 before this: three separately loaded kw tiles, 3 % slower; debug key 9.)
// Accumulators: 3 (kh) x [128 x 96] fp32 = 288 TMEM columns. Work item = (cob, cib, split).
// ---------------------------------------------------------------------------------------------
struct WgKhCfg {
  static constexpr int COB = 32, CIB = 32, N = 96;
  static constexpr int XTMax = 160 * CIB * 2;            // (bh + 2) * bw <= 160 rows of 64 B; halo_w: ONE tile of (bh + 2) * (bw + 2) <= 180 rows per stage
  static constexpr int XBytes = 3 * XTMax;               // three kw tiles
  static constexpr int XStages = 3;
  static constexpr int DySet = kSlots * kSlotBytes;
  static constexpr int OffDy = 0;
  static constexpr int OffX = 2 * DySet;
  static constexpr int OffBar = OffX + XStages * XBytes;
  static constexpr int NumBars = 2 * XStages + 4 + 1;
  static constexpr int OffTmem = OffBar + NumBars * 8;
  static constexpr int Total = OffTmem + 16;
  static constexpr int TmemCols = 512;
};

__global__ void __launch_bounds__(kThreads, 1)
conv3_wgrad_kh_kernel(const __grid_constant__ CUtensorMap tmap_dy, const __grid_constant__ CUtensorMap tmap_x,
                      const WgradParams p) {
  using C = WgKhCfg;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sDy = smem + C::OffDy;
  uint8_t* sX = smem + C::OffX;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::OffBar);
  uint64_t* xfull = bars;
  uint64_t* xempty = bars + C::XStages;
  uint64_t* dyfull = bars + 2 * C::XStages;
  uint64_t* dyempty = dyfull + 2;
  uint64_t* acc_full = dyempty + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + C::OffTmem);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int i = 0; i < C::XStages; ++i) {
      mbar_init(&xfull[i], 1);
      mbar_init(&xempty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&dyfull[i], 1);
      mbar_init(&dyempty[i], 1);
    }
    mbar_init(acc_full, 1);
    fence_mbar_init();
  }
  if (warp == 4 && lane == 0) {
    tma_prefetch_desc(&tmap_dy);
    tma_prefetch_desc(&tmap_x);
  }
  if (warp == 5) {
    tmem_alloc(tmem_slot, C::TmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int split = blockIdx.x % p.ksplit;
  const int item = blockIdx.x / p.ksplit;
  const int cib = item % p.ncib;
  const int cob = item / p.ncib;
  const long long t_begin = p.ntiles * split / p.ksplit;
  const long long t_end = p.ntiles * (split + 1) / p.ksplit;
  const int rows = p.bw * p.bh;                    // dy tile rows (K)
  const int xpitch = p.halo_w ? p.bw + 2 : p.bw;   // positions per image row of the x tile
  const int xrows = xpitch * (p.bh + 2);           // x tile rows incl. the halo
  const uint32_t dy_tile_bytes = rows * C::COB * 2;
  const uint32_t x_tile_bytes = xrows * C::CIB * 2;

  if (warp == 4) {
    const bool leader = elect_one() != 0;
    int xs = 0, ds = 0;
    uint32_t xph = 0, dph = 0;
    for (long long t = t_begin; t < t_end; ++t) {
      const int tw = static_cast<int>(t % p.tiles_w);
      const int th = static_cast<int>((t / p.tiles_w) % p.tiles_h);
      const int n = static_cast<int>(t / (static_cast<long long>(p.tiles_w) * p.tiles_h));
      const int w0 = tw * p.bw, h0 = th * p.bh;
      for (int pg = 0; pg < p.ngroups; ++pg) {
        const int d0 = pg * p.G;
        const int dend = min(p.d, d0 + p.G);
        mbar_wait(&dyempty[ds], dph ^ 1);
        if (leader) {
          mbar_expect_tx(&dyfull[ds], (p.G + 3) * dy_tile_bytes);
          for (int sl = 0; sl < p.G + 3; ++sl)
            tma_load_5d(sDy + ds * C::DySet + sl * kSlotBytes, &tmap_dy, &dyfull[ds], cob * C::COB, w0, h0, d0 - 1 + sl, n);
        }
        if (++ds == 2) {
          ds = 0;
          dph ^= 1;
        }
        for (int dp = d0; dp < dend; ++dp) {
          mbar_wait(&xempty[xs], xph ^ 1);
          if (leader) {
            if (p.halo_w) {
              mbar_expect_tx(&xfull[xs], x_tile_bytes);
              tma_load_5d(sX + xs * C::XBytes, &tmap_x, &xfull[xs], cib * C::CIB, w0 - 1, h0 - 1, dp, n);
            } else {
              mbar_expect_tx(&xfull[xs], 3 * x_tile_bytes);
#pragma unroll
              for (int kw = 0; kw < 3; ++kw)
                tma_load_5d(sX + xs * C::XBytes + kw * C::XTMax, &tmap_x, &xfull[xs], cib * C::CIB, w0 + kw - 1, h0 - 1, dp, n);
            }
          }
          if (++xs == C::XStages) {
            xs = 0;
            xph ^= 1;
          }
        }
      }
    }
  } else if (warp == 5) {
    const bool leader = elect_one() != 0;
    // MN-major operands: LBO = stride between 32-channel blocks (dy: next plane slot; x: next kw tile), SBO = 8 K rows
    const uint64_t adesc_hi = make_smem_desc_hi(kSlotBytes, 512, kSwizzle64);
    // halo_w: the tensor core applies the swizzle to the absolute shared-memory address (scripts/micro/umma_rowshift.cu), so
    // the three kw-shifted operands are the SAME tile read one row (= one position) apart - the three 32-channel blocks
    // of N start 64 bytes from each other (LBO) - and an 8-row K group is one image row of the tile: SBO = bw + 2 rows.
    const uint32_t sbo_b = p.halo_w ? static_cast<uint32_t>(xpitch * C::CIB * 2) : 512u;
    const uint64_t bdesc_hi = make_smem_desc_hi(p.halo_w ? C::CIB * 2 : C::XTMax, sbo_b, kSwizzle64);
    const uint32_t idesc = make_idesc_bf16(128, C::N, 1, 1);
    const uint64_t adesc0 = smem_desc(adesc_hi, smem_u32(sDy));
    const uint64_t bdesc0 = smem_desc(bdesc_hi, smem_u32(sX));
    const int ksteps = rows / 16;
    const uint32_t kh_step = static_cast<uint32_t>(xpitch * C::CIB * 2) >> 4;   // one image row of the x tile
    const uint32_t b_step = (2 * sbo_b) >> 4;                                   // 16 K rows
    int xs = 0, ds = 0;
    uint32_t xph = 0, dph = 0;
    uint32_t first = 1;
    for (long long t = t_begin; t < t_end; ++t) {
      for (int pg = 0; pg < p.ngroups; ++pg) {
        const int d0 = pg * p.G;
        const int dend = min(p.d, d0 + p.G);
        mbar_wait(&dyfull[ds], dph);
        tc_fence_after();
        const uint64_t dydesc = adesc0 + static_cast<uint64_t>(ds * (C::DySet >> 4));
        for (int dp = d0; dp < dend; ++dp) {
          mbar_wait(&xfull[xs], xph);
          tc_fence_after();
          const uint64_t xdesc = bdesc0 + static_cast<uint64_t>(xs * (C::XBytes >> 4));
          const uint64_t ad0 = dydesc + static_cast<uint64_t>((dp - d0) * (kSlotBytes >> 4));
#pragma unroll
          for (int kh = 0; kh < 3; ++kh) {
            uint64_t ad = ad0;
            uint64_t bd = xdesc + static_cast<uint64_t>(kh * kh_step);
            for (int ks = 0; ks < ksteps; ++ks) {
              if (leader) umma_bf16(tmem_base + kh * C::N, ad, bd, idesc, (first && ks == 0) ? 0u : 1u);
              ad += (2 * 512) >> 4;
              bd += b_step;
            }
          }
          first = 0;
          if (leader) umma_commit(&xempty[xs]);
          if (++xs == C::XStages) {
            xs = 0;
            xph ^= 1;
          }
        }
        if (leader) umma_commit(&dyempty[ds]);
        if (++ds == 2) {
          ds = 0;
          dph ^= 1;
        }
      }
    }
    if (leader) umma_commit(acc_full);
  } else {
    // epilogue: TMEM -> workspace partials [block][kh][lane][N]
    mbar_wait(acc_full, 0);
    tc_fence_after();
    float* out = p.partial + static_cast<size_t>(blockIdx.x) * 3 * 128 * C::N;
    const int row = warp * 32 + lane;
#pragma unroll 1
    for (int kh = 0; kh < 3; ++kh) {
#pragma unroll 1
      for (int c0 = 0; c0 < C::N; c0 += 32) {
        uint32_t v[32];
        tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + kh * C::N + c0, v);
        tmem_ld_wait();
        float4* dst = reinterpret_cast<float4*>(out + (static_cast<size_t>(kh) * 128 + row) * C::N + c0);
#pragma unroll
        for (int q = 0; q < 8; ++q)
          dst[q] = make_float4(__uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1]), __uint_as_float(v[4 * q + 2]),
                               __uint_as_float(v[4 * q + 3]));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    tmem_dealloc(tmem_base, C::TmemCols);
  }
}

// reduction for the all-kh variant: partial[(cob*ncib + cib)*ksplit + split][kh][lane = b*32 + col][kw*32 + cic]
__global__ void conv3_wgrad_kh_reduce_kernel(const float* __restrict__ partial, float* __restrict__ dw, int cout, int cin,
                                             int ksplit, float beta) {
  const int ncib = cin / 32;
  const long long total = static_cast<long long>(cout) * cin * 27;
  for (long long e = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; e < total;
       e += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int kw = static_cast<int>(e % 3);
    const int kh = static_cast<int>((e / 3) % 3);
    const int kd = static_cast<int>((e / 9) % 3);
    const int ci = static_cast<int>((e / 27) % cin);
    const int co = static_cast<int>(e / (27LL * cin));
    const int item = (co / 32) * ncib + ci / 32;
    const int lane = (2 - kd) * 32 + co % 32;   // plane slot offset 2 - kd, as in the kernel above
    const int column = kw * 32 + ci % 32;
    float acc = 0.f;
    for (int s = 0; s < ksplit; ++s)
      acc += partial[((static_cast<size_t>(item) * ksplit + s) * 3 + kh) * 128 * 96 + static_cast<size_t>(lane) * 96 + column];
    dw[e] = (beta == 0.f) ? acc : fmaf(beta, dw[e], acc);
  }
}

struct WgPlan {
  int allkh;   // 1: conv3_wgrad_kh_kernel (32 x 32 channel blocks, all kh taps per CTA)
  int cob, cib, bw, bh, tiles_w, tiles_h, G, ngroups, ncob, ncib, ksplit, nmma, N;
  long long ntiles;
  int items;
  size_t ws_bytes;
};

WgPlan make_plan(int cin, int cout, spff_shape s) {
  WgPlan pl;
  pl.allkh = 0;
  // few channels (one side has a single 32-channel block): the all-kh variant, tiles of 16 x 8 positions
  if ((cin == 32 || cout == 32) && s.w % 8 == 0 && debug_flag(1) == 0) {
    pl.allkh = 1;
    pl.cob = pl.cib = 32;
    pl.bw = (debug_flag(9) == 0) ? 8 : (s.w < 16 ? s.w : 16);   // 8: one x tile with a w halo serves the three kw shifts (key 9: the three-tile form)
    int bh = 128 / pl.bw;
    if (bh > s.h) bh = s.h;
    while ((pl.bw * bh) % 16) ++bh;   // K steps of 16 rows (rows past H are zero filled)
    pl.bh = bh;
    pl.tiles_w = (s.w + pl.bw - 1) / pl.bw;
    pl.tiles_h = (s.h + pl.bh - 1) / pl.bh;
    pl.ntiles = static_cast<long long>(s.n) * pl.tiles_w * pl.tiles_h;
    pl.G = s.d < kMaxG ? s.d : kMaxG;
    pl.ngroups = (s.d + pl.G - 1) / pl.G;
    pl.ncob = cout / 32;
    pl.ncib = cin / 32;
    pl.items = pl.ncob * pl.ncib;
    int ks = num_sms() / pl.items;
    if (debug_ctas() > 0) ks = debug_ctas() / pl.items;
    if (ks < 1) ks = 1;
    if (ks > pl.ntiles) ks = static_cast<int>(pl.ntiles);
    pl.ksplit = ks;
    pl.nmma = 3;
    pl.N = 96;
    pl.ws_bytes = static_cast<size_t>(pl.items) * pl.ksplit * 3 * 128 * 96 * sizeof(float);
    return pl;
  }
  pl.cob = (cout % 64 == 0) ? 64 : 32;
  pl.cib = (cin % 64 == 0) ? 64 : 32;
  const int kt = kSlotBytes / (pl.cob * 2);
  // position tile = box (bw x bh) of one plane; rows = bw*bh must be a multiple of 16 (UMMA K) and fit a slot
  auto gcd16 = [](int v) { int g = 16; while (v % g) g >>= 1; return g; };
  pl.bw = s.w < kt ? s.w : kt;
  int step = 16 / gcd16(pl.bw);
  if (pl.bw * step > kt) {  // awkward widths: fall back to a 16-aligned strip of the row
    pl.bw = (pl.bw / 16) * 16;
    step = 1;
  }
  int bh = kt / pl.bw;
  if (bh > s.h) bh = s.h;
  bh = (bh / step) * step;
  if (bh == 0) bh = step;  // rows beyond H are out of bounds and zero filled
  pl.bh = bh;
  pl.tiles_w = (s.w + pl.bw - 1) / pl.bw;
  pl.tiles_h = (s.h + pl.bh - 1) / pl.bh;
  pl.ntiles = static_cast<long long>(s.n) * pl.tiles_w * pl.tiles_h;
  pl.G = s.d < kMaxG ? s.d : kMaxG;
  pl.ngroups = (s.d + pl.G - 1) / pl.G;
  pl.ncob = cout / pl.cob;
  pl.ncib = cin / pl.cib;
  pl.items = 3 * pl.ncob * pl.ncib;
  int ks = num_sms() / pl.items;
  if (debug_ctas() > 0) ks = debug_ctas() / pl.items;
  if (ks < 1) ks = 1;
  if (ks > pl.ntiles) ks = static_cast<int>(pl.ntiles);
  pl.ksplit = ks;
  const int sp = 128 / pl.cob;
  pl.nmma = (3 + sp - 1) / sp;
  pl.N = 3 * pl.cib;
  pl.ws_bytes = static_cast<size_t>(pl.items) * pl.ksplit * pl.nmma * 128 * pl.N * sizeof(float);
  return pl;
}

template <int COB, int CIB>
int launch_wgrad(const void* x, long long ldx, int cin, const void* dy, long long lddy, int cout, spff_shape s,
                 const WgPlan& pl, float* partial, cudaStream_t stream) {
  using C = WgCfg<COB, CIB>;
  WgradParams p;
  p.n = s.n; p.d = s.d; p.h = s.h; p.w = s.w;
  p.bw = pl.bw; p.bh = pl.bh; p.tiles_w = pl.tiles_w; p.tiles_h = pl.tiles_h; p.ntiles = pl.ntiles;
  p.G = pl.G; p.ngroups = pl.ngroups; p.ncob = pl.ncob; p.ncib = pl.ncib; p.ksplit = pl.ksplit;
  p.halo_w = 0;
  p.partial = partial;
  CUtensorMap tdy, tx;
  {
    uint64_t dims[5] = {static_cast<uint64_t>(cout), static_cast<uint64_t>(s.w), static_cast<uint64_t>(s.h),
                        static_cast<uint64_t>(s.d), static_cast<uint64_t>(s.n)};
    uint64_t str[4] = {static_cast<uint64_t>(lddy) * 2, static_cast<uint64_t>(lddy) * 2 * s.w,
                       static_cast<uint64_t>(lddy) * 2 * s.w * s.h, static_cast<uint64_t>(lddy) * 2 * s.w * s.h * s.d};
    uint32_t box[5] = {COB, static_cast<uint32_t>(pl.bw), static_cast<uint32_t>(pl.bh), 1, 1};
    int e = encode_tmap_bf16(&tdy, dy, 5, dims, str, box, COB * 2);
    if (e) return e;
  }
  {
    uint64_t dims[5] = {static_cast<uint64_t>(cin), static_cast<uint64_t>(s.w), static_cast<uint64_t>(s.h),
                        static_cast<uint64_t>(s.d), static_cast<uint64_t>(s.n)};
    uint64_t str[4] = {static_cast<uint64_t>(ldx) * 2, static_cast<uint64_t>(ldx) * 2 * s.w,
                       static_cast<uint64_t>(ldx) * 2 * s.w * s.h, static_cast<uint64_t>(ldx) * 2 * s.w * s.h * s.d};
    uint32_t box[5] = {CIB, static_cast<uint32_t>(pl.bw), static_cast<uint32_t>(pl.bh), 1, 1};
    int e = encode_tmap_bf16(&tx, x, 5, dims, str, box, CIB * 2);
    if (e) return e;
  }
  static bool attr_set_dev[kMaxDevices] = {};   // the opt-in is per device (and per template instance)
  bool& attr_set = attr_set_dev[current_device()];
  if (!attr_set) {
    SPFF_CUDA(cudaFuncSetAttribute(conv3_wgrad_kernel<COB, CIB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   C::Total + 1024));
    attr_set = true;
  }
  conv3_wgrad_kernel<COB, CIB><<<pl.items * pl.ksplit, kThreads, C::Total + 1024, stream>>>(tdy, tx, p);
  SPFF_CUDA(cudaGetLastError());
  return 0;
}


int launch_wgrad_kh(const void* x, long long ldx, int cin, const void* dy, long long lddy, int cout, spff_shape s,
                    const WgPlan& pl, float* partial, cudaStream_t stream) {
  using C = WgKhCfg;
  WgradParams p;
  p.n = s.n; p.d = s.d; p.h = s.h; p.w = s.w;
  p.bw = pl.bw; p.bh = pl.bh; p.tiles_w = pl.tiles_w; p.tiles_h = pl.tiles_h; p.ntiles = pl.ntiles;
  p.G = pl.G; p.ngroups = pl.ngroups; p.ncob = pl.ncob; p.ncib = pl.ncib; p.ksplit = pl.ksplit;
  p.partial = partial;
  p.halo_w = (pl.bw == 8) ? 1 : 0;
  CUtensorMap tdy, tx;
  {
    uint64_t dims[5] = {static_cast<uint64_t>(cout), static_cast<uint64_t>(s.w), static_cast<uint64_t>(s.h),
                        static_cast<uint64_t>(s.d), static_cast<uint64_t>(s.n)};
    uint64_t str[4] = {static_cast<uint64_t>(lddy) * 2, static_cast<uint64_t>(lddy) * 2 * s.w,
                       static_cast<uint64_t>(lddy) * 2 * s.w * s.h, static_cast<uint64_t>(lddy) * 2 * s.w * s.h * s.d};
    uint32_t box[5] = {32, static_cast<uint32_t>(pl.bw), static_cast<uint32_t>(pl.bh), 1, 1};
    int e = encode_tmap_bf16(&tdy, dy, 5, dims, str, box, 64);
    if (e) return e;
  }
  {
    uint64_t dims[5] = {static_cast<uint64_t>(cin), static_cast<uint64_t>(s.w), static_cast<uint64_t>(s.h),
                        static_cast<uint64_t>(s.d), static_cast<uint64_t>(s.n)};
    uint64_t str[4] = {static_cast<uint64_t>(ldx) * 2, static_cast<uint64_t>(ldx) * 2 * s.w,
                       static_cast<uint64_t>(ldx) * 2 * s.w * s.h, static_cast<uint64_t>(ldx) * 2 * s.w * s.h * s.d};
    uint32_t box[5] = {32, static_cast<uint32_t>(pl.bw + (p.halo_w ? 2 : 0)), static_cast<uint32_t>(pl.bh + 2), 1, 1};
    int e = encode_tmap_bf16(&tx, x, 5, dims, str, box, 64);
    if (e) return e;
  }
  static bool attr_set_dev[kMaxDevices] = {};   // the opt-in is per device (and per template instance)
  bool& attr_set = attr_set_dev[current_device()];
  if (!attr_set) {
    SPFF_CUDA(cudaFuncSetAttribute(conv3_wgrad_kh_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, C::Total + 1024));
    attr_set = true;
  }
  conv3_wgrad_kh_kernel<<<pl.items * pl.ksplit, kThreads, C::Total + 1024, stream>>>(tdy, tx, p);
  SPFF_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace
}  // namespace spff

extern "C" {

size_t spff_conv3d_k3_wgrad_workspace(int cin, int cout, spff_shape s) {
  if (cin % 32 || cout % 32 || cin <= 0 || cout <= 0 || s.n <= 0 || s.d <= 0 || s.h <= 0 || s.w <= 0) return 0;
  return spff::make_plan(cin, cout, s).ws_bytes;
}

int spff_conv3d_k3_wgrad(const void* x, long long ldx, int cin, const void* dy, long long lddy, int cout,
                         spff_shape s, float* dw, float beta, void* workspace, size_t workspace_bytes,
                         void* stream) {
  int e = spff_device_check();
  if (e) return e;
  SPFF_REQUIRE(x && dy && dw && workspace, "conv3d_k3_wgrad: null pointer");
  SPFF_REQUIRE(cin % 32 == 0 && cout % 32 == 0 && cin > 0 && cout > 0,
               "conv3d_k3_wgrad: channels (%d -> %d) must be multiples of 32", cin, cout);
  SPFF_REQUIRE(s.n > 0 && s.d > 0 && s.h > 0 && s.w > 0, "conv3d_k3_wgrad: empty shape");
  SPFF_REQUIRE(ldx >= cin && lddy >= cout && ldx % 8 == 0 && lddy % 8 == 0, "conv3d_k3_wgrad: bad channel pitch");
  const spff::WgPlan pl = spff::make_plan(cin, cout, s);
  if (workspace_bytes < pl.ws_bytes) {
    spff::set_error("conv3d_k3_wgrad: workspace %zu < %zu bytes", workspace_bytes, pl.ws_bytes);
    return SPFF_ERR_WORKSPACE;
  }
  SPFF_REQUIRE((pl.bw * pl.bh) % 16 == 0 && pl.bw * pl.bh * pl.cob * 2 <= spff::kSlotBytes,
               "conv3d_k3_wgrad: cannot tile a %dx%d plane", s.h, s.w);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  float* partial = static_cast<float*>(workspace);
  if (pl.allkh) {
    SPFF_REQUIRE(pl.bw * (pl.bh + 2) <= 160, "conv3d_k3_wgrad: halo tile too large");
    e = spff::launch_wgrad_kh(x, ldx, cin, dy, lddy, cout, s, pl, partial, st);
    if (e) return e;
    const long long tot = 27LL * cin * cout;
    spff::conv3_wgrad_kh_reduce_kernel<<<static_cast<int>((tot + 255) / 256), 256, 0, st>>>(partial, dw, cout, cin,
                                                                                           pl.ksplit, beta);
    SPFF_CUDA(cudaGetLastError());
    return 0;
  }
  if (pl.cob == 64 && pl.cib == 64)
    e = spff::launch_wgrad<64, 64>(x, ldx, cin, dy, lddy, cout, s, pl, partial, st);
  else if (pl.cob == 64)
    e = spff::launch_wgrad<64, 32>(x, ldx, cin, dy, lddy, cout, s, pl, partial, st);
  else if (pl.cib == 64)
    e = spff::launch_wgrad<32, 64>(x, ldx, cin, dy, lddy, cout, s, pl, partial, st);
  else
    e = spff::launch_wgrad<32, 32>(x, ldx, cin, dy, lddy, cout, s, pl, partial, st);
  if (e) return e;
  const long long total = 27LL * cin * cout;
  const int blocks = static_cast<int>((total + 255) / 256);
  spff::conv3_wgrad_reduce_kernel<<<blocks, 256, 0, st>>>(partial, dw, cout, cin, pl.cob, pl.cib, pl.ksplit, beta);
  SPFF_CUDA(cudaGetLastError());
  return 0;
}

}  // extern "C"
