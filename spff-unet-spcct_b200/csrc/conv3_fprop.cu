// 3x3x3 convolution (stride 1, zero pad 1, no bias) as an implicit GEMM on tcgen05 — forward and input
// gradient. Replaces the F.conv3d / convolution_backward(input) library calls behind
// `_conv3x3xk` (reference innovative3D/models.py:616-618).
//
// Formulation ("kw folded into N"). Positions of one energy plane are flattened q = h*W + w. For
// one (sample n, 128 flattened positions, block of 32 output channels) the kernel computes, for
// every output plane d of a plane group, three partial sums
//     T_kw[q][co] = sum_{kd,kh,ci} x[n][d+kd-1][q + (kh-1)*W][ci] * w[co][ci][kd][kh][kw]
// as ONE GEMM tile with M = 128 positions, N = 3*32 = (kw, co), K = 9*Cin. The (kd,kh) taps are
// whole-row shifts of the flattened index, so every A tile is a plain TMA box of a rank-4 tensor
// map (C, H*W, D, N) whose out-of-bounds rows are zero filled (the h and d padding). The w shift is
// applied in the epilogue:  y[q] = T_0[q-1] + T_1[q] + T_2[q+1]  (masked at w = 0 / w = W-1).
// Compared with one GEMM per tap this reads each A tile from shared memory once for three taps
// (N = 96 instead of 32) and an input plane tile loaded once feeds up to three output planes
// (kd), which is what keeps L2->SM traffic under the tensor pipe's needs for 32..64 channels.
//
// Warp roles (192 threads, one CTA per SM, persistent over work items):
//   warps 0-3  epilogue: TMEM -> registers (tcgen05.ld), w-shift add, bf16 pack, global store
//   warp  4    TMA producer (one elected lane)
//   warp  5    TMEM allocation + tcgen05.mma issue (one elected lane)
#include "common.h"
#include "ptx.cuh"

namespace spff {

namespace {

constexpr int kTileM = 128;
constexpr int kCoBlk = 32;            // output channels per work item
constexpr int kN = 3 * kCoBlk;        // UMMA N: (kw, co)
constexpr int kMaxPlanes = 5;         // output planes per group: 5 * 96 = 480 TMEM columns
constexpr int kThreads = 192;

struct FpropParams {
  int n, d, hw, w;
  int nkc;      // K chunks per tap (cin / KC)
  int ncb;      // output channel blocks (cout / 32)
  int mstep;    // output positions per tile: 128 (tile edges are row edges) or 126 (1-row halo each side)
  int halo;     // 0 or 1
  int qtiles;   // ceil(hw / mstep)
  int G;        // planes per group
  int ngroups;  // ceil(d / G)
  long long items;
  __nv_bfloat16* y;
  long long ldy;
  float* stat_partial;  // STATS: [n][qtiles*ngroups][2][cout] per-item {sum, sum of squares} of the outputs
  int cout;
};

template <int KC, bool RES>
struct SmemLayout {
  static constexpr int kStages = (KC == 64) ? 6 : 8;
  static constexpr int kABytes = kTileM * KC * 2;
  static constexpr int kWBlock = kN * KC * 2;     // one kd block: 96 rows
  static constexpr int kWBytes = 3 * kWBlock;     // three kd blocks per (kh, kc)
  // RES: all 27 taps of one 32-channel output block stay in smem (cin <= 64 -> one K chunk per tap);
  // otherwise the (kh, kc) weight blocks stream through a 2-deep ring.
  static constexpr int kWStages = RES ? 3 : 2;
  static constexpr int kOffA = 0;
  static constexpr int kOffW = kOffA + kStages * kABytes;
  static constexpr int kOffX = kOffW + kWStages * kWBytes;          // epilogue exchange rows
  static constexpr int kXBytes = 2 * 4 * 2 * kCoBlk * 4 + 2 * kCoBlk * 4;   // + one all-zero row pair (ALIGNED tile edges)
  static constexpr int kOffS = kOffX + kXBytes;                     // STATS: per-warp column sums [4][2][32]
  static constexpr int kSBytes = 4 * 2 * kCoBlk * 4;
  static constexpr int kOffBar = kOffS + kSBytes;
  static constexpr int kNumBars = 2 * kStages + 2 * 3 + 2 * kMaxPlanes;
  static constexpr int kOffTmem = kOffBar + kNumBars * 8;
  static constexpr int kTotal = kOffTmem + 16;
};

// Loop order. Streamed weights (RES = false): kh -> kc -> input plane; every output plane of the
// group completes at the end of the item. Resident weights (RES = true): input plane -> kh; output
// plane j is complete once input plane j+1 is consumed, so its epilogue overlaps the MMAs of the
// following planes and of the next item (per-plane full/empty barriers on the TMEM accumulators).
// Sum the 32 per-lane values of each of 32 columns across the warp: afterwards a[0] of lane L holds
// the total of column L (transpose-reduce: 31 shuffles instead of 32 x 5).
__device__ __forceinline__ void warp_column_sums(float (&a)[32], int lane) {
#pragma unroll
  for (int half = 16; half >= 1; half >>= 1) {
    const bool hi = (lane & half) != 0;
#pragma unroll
    for (int i = 0; i < half; ++i) {
      const float keep = hi ? a[i + half] : a[i];
      const float send = hi ? a[i] : a[i + half];
      a[i] = keep + __shfl_xor_sync(0xffffffffu, send, half);
    }
  }
}

// STATS: the epilogue also produces the InstanceNorm statistics of its output tile (sum and sum of
// squares per output channel, fp32 values before the bf16 rounding) as one partial row per work
// item - no atomics, reduced in a fixed order by spff_in_coeffs_from_partials. This removes the
// separate statistics pass over the freshly written conv output.
// ALIGNED: the tile is a whole number of image rows starting at a multiple of 32 positions (W in {32, 64, 128}, no halo)
// and every tile row is a real position. An image-row edge can then only fall between lane 31 and lane 0, so the
// per-element "has a left / right neighbour" selects and the row mask of the statistics disappear: the warp that
// hands a boundary value to its neighbour through shared memory writes zeros when that neighbour sits across an edge.
template <int KC, bool RES, bool STATS, bool ALIGNED>
__global__ void __launch_bounds__(kThreads, 1)
conv3_fprop_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w,
                   const FpropParams p) {
  using L = SmemLayout<KC, RES>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);  // swizzled tiles need 1024 B alignment
  uint8_t* sA = smem + L::kOffA;
  uint8_t* sW = smem + L::kOffW;
  float* sX = reinterpret_cast<float*>(smem + L::kOffX);
  float* sS = reinterpret_cast<float*>(smem + L::kOffS);
  float* sZero = sX + 2 * 4 * 2 * kCoBlk;   // 2 x kCoBlk zeros (visible after the set-up __syncthreads)
  if (threadIdx.x < 2 * kCoBlk) sZero[threadIdx.x] = 0.f;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::kOffBar);
  uint64_t* full = bars;
  uint64_t* empty = bars + L::kStages;
  uint64_t* wfull = bars + 2 * L::kStages;
  uint64_t* wempty = wfull + 3;
  uint64_t* acc_full = wempty + 3;
  uint64_t* acc_empty = acc_full + kMaxPlanes;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + L::kOffTmem);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int i = 0; i < L::kStages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 3; ++i) {
      mbar_init(&wfull[i], 1);
      mbar_init(&wempty[i], 1);
    }
    for (int i = 0; i < kMaxPlanes; ++i) {
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_empty[i], 128);
    }
    fence_mbar_init();
  }
  if (warp == 4 && lane == 0) {
    tma_prefetch_desc(&tmap_x);
    tma_prefetch_desc(&tmap_w);
  }
  if (warp == 5) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  constexpr uint32_t kSwz = (KC == 64) ? kSwizzle128 : kSwizzle64;
  constexpr uint32_t kSbo = (KC == 64) ? 1024 : 512;
  const uint64_t desc_hi = make_smem_desc_hi(16, kSbo, kSwz);
  const uint32_t idesc = make_idesc_bf16(kTileM, kN, 0, 0);

  // RES: a CTA keeps one output-channel block (cb = blockIdx % ncb) and strides over positions.
  const long long item0 = RES ? (blockIdx.x / p.ncb) : blockIdx.x;
  const long long istep = RES ? (gridDim.x / p.ncb) : gridDim.x;
  const long long nitems = RES ? p.items / p.ncb : p.items;
  auto decode = [&](long long item, int& n, int& qt, int& pg, int& cb) {
    long long r = item;
    if (RES) {
      cb = blockIdx.x % p.ncb;
    } else {
      cb = static_cast<int>(r % p.ncb);
      r /= p.ncb;
    }
    pg = static_cast<int>(r % p.ngroups);
    r /= p.ngroups;
    qt = static_cast<int>(r % p.qtiles);
    n = static_cast<int>(r / p.qtiles);
  };

  if (warp == 4) {
    // ------------------------------------------------------------------ TMA producer
    // All 32 lanes run the control flow, so every address / coordinate is provably warp-uniform
    // and stays in uniform registers; one elected lane issues the TMA instructions.
    const bool leader = elect_one() != 0;
    {
      int s = 0, ws = 0;
      uint32_t ph = 0, wph = 0;
      if (RES) {
        const int cb = blockIdx.x % p.ncb;
        for (int kh = 0; kh < 3; ++kh) {
          const int wrow = (cb * 3 + kh) * 3 * kN;
          if (leader) {
            mbar_expect_tx(&wfull[kh], L::kWBytes);
#pragma unroll
            for (int kd = 0; kd < 3; ++kd)
              tma_load_2d(sW + kh * L::kWBytes + kd * L::kWBlock, &tmap_w, &wfull[kh], 0, wrow + kd * kN);
          }
        }
      }
      for (long long item = item0; item < nitems; item += istep) {
        int n, qt, pg, cb;
        decode(item, n, qt, pg, cb);
        const int tq0 = qt * p.mstep - p.halo;
        const int d0 = pg * p.G;
        const int dlo = max(0, d0 - 1);
        const int dhi = min(p.d - 1, d0 + p.G);
        if (RES) {
          for (int dp = dlo; dp <= dhi; ++dp) {
            for (int kh = 0; kh < 3; ++kh) {
              mbar_wait(&empty[s], ph ^ 1);
              if (leader) {
                mbar_expect_tx(&full[s], L::kABytes);
                tma_load_4d(sA + s * L::kABytes, &tmap_x, &full[s], 0, tq0 + (kh - 1) * p.w, dp, n);
              }
              if (++s == L::kStages) {
                s = 0;
                ph ^= 1;
              }
            }
          }
        } else {
          for (int kh = 0; kh < 3; ++kh) {
            for (int kc = 0; kc < p.nkc; ++kc) {
              mbar_wait(&wempty[ws], wph ^ 1);
              const int wrow = ((cb * 3 + kh) * p.nkc + kc) * 3 * kN;
              if (leader) {
                mbar_expect_tx(&wfull[ws], L::kWBytes);
#pragma unroll
                for (int kd = 0; kd < 3; ++kd)
                  tma_load_2d(sW + ws * L::kWBytes + kd * L::kWBlock, &tmap_w, &wfull[ws], 0, wrow + kd * kN);
              }
              if (++ws == L::kWStages) {
                ws = 0;
                wph ^= 1;
              }
              for (int dp = dlo; dp <= dhi; ++dp) {
                mbar_wait(&empty[s], ph ^ 1);
                if (leader) {
                  mbar_expect_tx(&full[s], L::kABytes);
                  tma_load_4d(sA + s * L::kABytes, &tmap_x, &full[s], kc * KC, tq0 + (kh - 1) * p.w, dp, n);
                }
                if (++s == L::kStages) {
                  s = 0;
                  ph ^= 1;
                }
              }
            }
          }
        }
      }
    }
  } else if (warp == 5) {
    // ------------------------------------------------------------------ MMA issuer
    // Warp-uniform control flow (all 32 lanes), one elected lane issues tcgen05.mma / commit: the
    // descriptors, TMEM addresses and barrier addresses then live in uniform registers instead of
    // being moved there (R2UR) in front of every instruction by a single divergent thread.
    const bool leader = elect_one() != 0;
    {
      int s = 0, ws = 0;
      uint32_t ph = 0, wph = 0;
      uint32_t accpar = 0;  // per accumulator slot: parity of its next acc_empty wait
      if (RES) {
        for (int kh = 0; kh < 3; ++kh) mbar_wait(&wfull[kh], 0);
      }
      // one A stage against the kd blocks at wbase; returns the updated touched mask
      // descriptors advance by plain additions on the 14-bit (address >> 4) field: the whole dynamic
      // shared window is < 256 KB, so the field never overflows into its neighbours
      const uint64_t a_desc0 = smem_desc(desc_hi, smem_u32(sA));
      const uint64_t w_desc0 = smem_desc(desc_hi, smem_u32(sW));
      // One A stage (input plane dp, one kh) feeds the output planes dp-1, dp, dp+1 of the group through
      // the taps kd = 2, 1, 0. The kd blocks of the packed weights are stored in DESCENDING kd, i.e. in
      // ascending output plane, and plane j's accumulator sits at TMEM column j*96: two neighbouring
      // output planes are therefore ONE MMA with N = 192 (the A tile is read from shared memory once
      // for both), as long as their accumulate flags agree. Returns the updated touched mask.
      const uint32_t idesc2 = make_idesc_bf16(kTileM, 2 * kN, 0, 0);
      auto stage_mmas = [&](uint64_t adesc, uint64_t wdesc, int dp, int d0, int dend, uint32_t touched) -> uint32_t {
        const int jlo = max(dp - 1, d0) - d0, jhi = min(dp + 1, dend - 1) - d0;
        int j = jlo;
        while (j <= jhi) {
          const uint32_t tj = (touched >> j) & 1u;
          const bool pair = (j + 1 <= jhi) && (((touched >> (j + 1)) & 1u) == tj);
          const int cnt = pair ? 2 : 1;
          if (!tj) {  // first MMA into these accumulators: the epilogue must have drained them
            for (int q = j; q < j + cnt; ++q) {
              mbar_wait(&acc_empty[q], ((accpar >> q) & 1u) ^ 1u);
              accpar ^= 1u << q;
            }
            tc_fence_after();
          }
          const uint32_t dcol = tmem_base + j * kN;
          const int blk = j + d0 - dp + 1;   // weight block of the first plane of the run (= 2 - kd)
          const uint64_t bd0 = wdesc + static_cast<uint64_t>(blk * (L::kWBlock >> 4));
          const uint32_t id = pair ? idesc2 : idesc;
#pragma unroll
          for (int k = 0; k < KC / 16; ++k) {
            if (leader) umma_bf16(dcol, adesc + static_cast<uint64_t>(k * 2), bd0 + static_cast<uint64_t>(k * 2), id,
                                  tj | (k > 0 ? 1u : 0u));
          }
          touched |= (pair ? 3u : 1u) << j;
          j += cnt;
        }
        return touched;
      };
      for (long long item = item0; item < nitems; item += istep) {
        int n, qt, pg, cb;
        decode(item, n, qt, pg, cb);
        const int d0 = pg * p.G;
        const int dlo = max(0, d0 - 1);
        const int dhi = min(p.d - 1, d0 + p.G);
        const int dend = min(p.d, d0 + p.G);
        uint32_t touched = 0;
        if (RES) {
          for (int dp = dlo; dp <= dhi; ++dp) {
            for (int kh = 0; kh < 3; ++kh) {
              mbar_wait(&full[s], ph);
              tc_fence_after();
              touched = stage_mmas(a_desc0 + static_cast<uint64_t>(s * (L::kABytes >> 4)),
                                   w_desc0 + static_cast<uint64_t>(kh * (L::kWBytes >> 4)), dp, d0, dend, touched);
              if (leader) umma_commit(&empty[s]);
              if (++s == L::kStages) {
                s = 0;
                ph ^= 1;
              }
            }
            if (leader && dp - 1 >= d0 && dp - 1 < dend) umma_commit(&acc_full[dp - 1 - d0]);
            if (leader && dp == dhi && dhi < dend) umma_commit(&acc_full[dhi - d0]);
          }
        } else {
          for (int kh = 0; kh < 3; ++kh) {
            for (int kc = 0; kc < p.nkc; ++kc) {
              mbar_wait(&wfull[ws], wph);
              const uint64_t wbase = w_desc0 + static_cast<uint64_t>(ws * (L::kWBytes >> 4));
              for (int dp = dlo; dp <= dhi; ++dp) {
                mbar_wait(&full[s], ph);
                tc_fence_after();
                touched = stage_mmas(a_desc0 + static_cast<uint64_t>(s * (L::kABytes >> 4)), wbase, dp, d0, dend, touched);
                if (leader) umma_commit(&empty[s]);
                if (++s == L::kStages) {
                  s = 0;
                  ph ^= 1;
                }
              }
              if (leader) umma_commit(&wempty[ws]);
              if (++ws == L::kWStages) {
                ws = 0;
                wph ^= 1;
              }
            }
          }
          if (leader)
            for (int j = 0; j < dend - d0; ++j) umma_commit(&acc_full[j]);
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (warps 0-3)
    uint32_t accpar = 0;  // per accumulator slot: parity of its next acc_full wait
    uint32_t xpar = 0;
    const int m = warp * 32 + lane;  // tile row == TMEM lane
    for (long long item = item0; item < nitems; item += istep) {
      int n, qt, pg, cb;
      decode(item, n, qt, pg, cb);
      const int tq0 = qt * p.mstep - p.halo;
      const int d0 = pg * p.G;
      const int dend = min(p.d, d0 + p.G);
      const int q = tq0 + m;
      const int wq = (q >= 0) ? (q % p.w) : 0;
      const bool row_out = (m >= p.halo) && (m < p.halo + p.mstep) && (q < p.hw);
      const bool has_left = wq != 0;
      const bool has_right = wq != p.w - 1;
      float ssum[kCoBlk], ssq[kCoBlk];
      if constexpr (STATS) {
#pragma unroll
        for (int c = 0; c < kCoBlk; ++c) ssum[c] = ssq[c] = 0.f;
      }
      const float rowmask = row_out ? 1.f : 0.f;
      for (int d = d0; d < dend; ++d) {
        const int j = d - d0;
        mbar_wait(&acc_full[j], (accpar >> j) & 1u);
        accpar ^= 1u << j;
        tc_fence_after();
        const uint32_t tcol = tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + j * kN;
        uint32_t t0[32], t1[32], t2[32];
        tmem_ld_32x32(tcol, t0);
        tmem_ld_32x32(tcol + 32, t1);
        tmem_ld_32x32(tcol + 64, t2);
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(&acc_empty[j]);  // accumulator is in registers: the next item may overwrite it
        // rows m-1 / m+1 live in the neighbouring lanes; across warps they go through smem
        float4* xrow = reinterpret_cast<float4*>(sX + (xpar * 4 + warp) * 2 * kCoBlk);
        if (lane == 31) {
          const bool keep = !ALIGNED || has_right;   // ALIGNED: the receiver (row m+1) starts an image row -> zeros
#pragma unroll
          for (int v = 0; v < 8; ++v)
            xrow[v] = keep ? make_float4(__uint_as_float(t0[4 * v]), __uint_as_float(t0[4 * v + 1]),
                                         __uint_as_float(t0[4 * v + 2]), __uint_as_float(t0[4 * v + 3]))
                           : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        if (lane == 0) {
          const bool keep = !ALIGNED || has_left;    // ALIGNED: the receiver (row m-1) ends an image row -> zeros
#pragma unroll
          for (int v = 0; v < 8; ++v)
            xrow[8 + v] = keep ? make_float4(__uint_as_float(t2[4 * v]), __uint_as_float(t2[4 * v + 1]),
                                             __uint_as_float(t2[4 * v + 2]), __uint_as_float(t2[4 * v + 3]))
                               : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        named_bar_sync(1, 128);
        const float4* lrow = reinterpret_cast<const float4*>(sX + (xpar * 4 + (warp > 0 ? warp - 1 : 0)) * 2 * kCoBlk);
        const float4* rrow =
            reinterpret_cast<const float4*>(sX + (xpar * 4 + (warp < 3 ? warp + 1 : 3)) * 2 * kCoBlk + kCoBlk);
        if (ALIGNED) {   // the tile's first / last row has no neighbour inside the tile: an image-row edge
          if (warp == 0) lrow = reinterpret_cast<const float4*>(sZero);
          if (warp == 3) rrow = reinterpret_cast<const float4*>(sZero + kCoBlk);
        }
        uint32_t packed[16];
#pragma unroll
        for (int c4 = 0; c4 < 8; ++c4) {
          const float4 lb4 = lrow[c4];  // broadcast loads, used by lane 0 / lane 31 only
          const float4 rb4 = rrow[c4];
          const float lb[4] = {lb4.x, lb4.y, lb4.z, lb4.w};
          const float rb[4] = {rb4.x, rb4.y, rb4.z, rb4.w};
          float v[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int c = 4 * c4 + e;
            float l = __shfl_up_sync(0xffffffffu, __uint_as_float(t0[c]), 1);
            float r = __shfl_down_sync(0xffffffffu, __uint_as_float(t2[c]), 1);
            l = (lane == 0) ? lb[e] : l;
            r = (lane == 31) ? rb[e] : r;
            if (ALIGNED) {
              v[e] = __uint_as_float(t1[c]) + l + r;
              if constexpr (STATS) {
                ssum[c] += v[e];
                ssq[c] = fmaf(v[e], v[e], ssq[c]);
              }
            } else {
              v[e] = __uint_as_float(t1[c]) + (has_left ? l : 0.f) + (has_right ? r : 0.f);
              if constexpr (STATS) {
                const float vm = v[e] * rowmask;
                ssum[c] += vm;
                ssq[c] = fmaf(vm, v[e], ssq[c]);
              }
            }
          }
          packed[2 * c4] = pack_bf16x2(v[0], v[1]);
          packed[2 * c4 + 1] = pack_bf16x2(v[2], v[3]);
        }
        if (row_out) {
          __nv_bfloat16* dst = p.y + ((static_cast<long long>(n) * p.d + d) * p.hw + q) * p.ldy + cb * kCoBlk;
          if ((p.ldy & 15) == 0 && (reinterpret_cast<uintptr_t>(p.y) & 31) == 0) {   // 32-byte aligned rows: full-sector stores
            st_global_v8(dst, packed[0], packed[1], packed[2], packed[3], packed[4], packed[5], packed[6], packed[7]);
            st_global_v8(dst + 16, packed[8], packed[9], packed[10], packed[11], packed[12], packed[13], packed[14], packed[15]);
          } else {
            uint4* d4 = reinterpret_cast<uint4*>(dst);
#pragma unroll
            for (int v = 0; v < 4; ++v)
              d4[v] = make_uint4(packed[4 * v], packed[4 * v + 1], packed[4 * v + 2], packed[4 * v + 3]);
          }
        }
        xpar ^= 1;
      }
      if constexpr (STATS) {
        // item totals: columns across the 32 rows of each warp, then across the 4 warps through smem
        warp_column_sums(ssum, lane);
        warp_column_sums(ssq, lane);
        sS[(warp * 2 + 0) * kCoBlk + lane] = ssum[0];
        sS[(warp * 2 + 1) * kCoBlk + lane] = ssq[0];
        named_bar_sync(2, 128);
        if (threadIdx.x < 2 * kCoBlk) {
          const int which = threadIdx.x / kCoBlk, col = threadIdx.x % kCoBlk;
          const float t = sS[(0 * 2 + which) * kCoBlk + col] + sS[(1 * 2 + which) * kCoBlk + col] +
                          sS[(2 * 2 + which) * kCoBlk + col] + sS[(3 * 2 + which) * kCoBlk + col];
          const long long slot = static_cast<long long>(n) * (p.qtiles * p.ngroups) + qt * p.ngroups + pg;
          p.stat_partial[(slot * 2 + which) * p.cout + cb * kCoBlk + col] = t;
        }
        named_bar_sync(2, 128);   // sS is reused by the next item
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

template <int KC, bool RES, bool STATS, bool ALIGNED>
int launch_fprop_impl(const void* x, long long ldx, int cin, const void* wpk, void* y, long long ldy, int cout,
                 spff_shape s, float* stat_partial, cudaStream_t stream) {
  using L = SmemLayout<KC, RES>;
  FpropParams p;
  p.n = s.n;
  p.d = s.d;
  p.hw = s.h * s.w;
  p.w = s.w;
  p.nkc = cin / KC;
  p.ncb = cout / kCoBlk;
  if (kTileM % s.w == 0) {
    p.mstep = 128;
    p.halo = 0;
  } else {
    p.mstep = 126;
    p.halo = 1;
  }
  p.qtiles = (p.hw + p.mstep - 1) / p.mstep;
  // balanced plane groups: the fewest groups of <= kMaxPlanes planes, all (almost) equally deep — 16 planes run as
  // 4 x 4 (two N = 192 plane pairs per tap) instead of 5 + 5 + 5 + 1
  p.ngroups = (s.d + kMaxPlanes - 1) / kMaxPlanes;
  p.G = (s.d + p.ngroups - 1) / p.ngroups;
  p.items = static_cast<long long>(s.n) * p.qtiles * p.ngroups * p.ncb;
  p.y = static_cast<__nv_bfloat16*>(y);
  p.ldy = ldy;
  p.stat_partial = stat_partial;
  p.cout = cout;

  CUtensorMap tx, tw;
  {
    uint64_t dims[4] = {static_cast<uint64_t>(cin), static_cast<uint64_t>(p.hw), static_cast<uint64_t>(s.d),
                        static_cast<uint64_t>(s.n)};
    uint64_t str[3] = {static_cast<uint64_t>(ldx) * 2, static_cast<uint64_t>(ldx) * 2 * p.hw,
                       static_cast<uint64_t>(ldx) * 2 * p.hw * s.d};
    uint32_t box[4] = {KC, kTileM, 1, 1};
    int e = encode_tmap_bf16(&tx, x, 4, dims, str, box, KC * 2);
    if (e) return e;
  }
  {
    const uint64_t rows = static_cast<uint64_t>(p.ncb) * 3 * p.nkc * 3 * kN;
    uint64_t dims[2] = {KC, rows};
    uint64_t str[1] = {KC * 2};
    uint32_t box[2] = {KC, kN};
    int e = encode_tmap_bf16(&tw, wpk, 2, dims, str, box, KC * 2);
    if (e) return e;
  }
  static bool attr_set_dev[kMaxDevices] = {};   // the opt-in is per device (and per template instance)
  bool& attr_set = attr_set_dev[current_device()];
  if (!attr_set) {
    SPFF_CUDA(cudaFuncSetAttribute(conv3_fprop_kernel<KC, RES, STATS, ALIGNED>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   L::kTotal + 1024));
    attr_set = true;
  }
  int ctas = debug_ctas() > 0 ? debug_ctas() : num_sms();
  if (RES) {
    const long long pos_items = p.items / p.ncb;
    long long per_cb = ctas / p.ncb;
    if (per_cb < 1) per_cb = 1;
    if (per_cb > pos_items) per_cb = pos_items;
    ctas = static_cast<int>(per_cb * p.ncb);
  } else if (p.items < ctas) {
    ctas = static_cast<int>(p.items);
  }
  conv3_fprop_kernel<KC, RES, STATS, ALIGNED><<<ctas, kThreads, L::kTotal + 1024, stream>>>(tx, tw, p);
  SPFF_CUDA(cudaGetLastError());
  return 0;
}

template <int KC, bool RES, bool STATS>
int launch_fprop(const void* x, long long ldx, int cin, const void* wpk, void* y, long long ldy, int cout,
                 spff_shape s, float* stat_partial, cudaStream_t stream) {
  // whole image rows per tile, 32-aligned, no partial tile (see the ALIGNED note at the kernel)
  const long long hw = static_cast<long long>(s.h) * s.w;
  const bool aligned = (kTileM % s.w == 0) && (s.w % 32 == 0) && (hw % kTileM == 0) && !debug_flag(6);
  if (aligned) return launch_fprop_impl<KC, RES, STATS, true>(x, ldx, cin, wpk, y, ldy, cout, s, stat_partial, stream);
  return launch_fprop_impl<KC, RES, STATS, false>(x, ldx, cin, wpk, y, ldy, cout, s, stat_partial, stream);
}

// ---------------------------------------------------------------------------------------------
// Weight packing: nn.Conv3d weight [Cout][Cin][3][3][3] fp32 -> bf16 GEMM operand
//   P[cb][kh][kc][b][kw][co32][KC],  value = w[cb*32+co][kc*KC+k][2-b][kh][kw]          (forward)
//   P[cb][kh][kc][b][kw][ci32][KC],  value = w[kc*KC+k][cb*32+ci][b][2-kh][2-kw]       (dgrad)
// b = 2 - kd: blocks in descending tap order, so that block b serves output plane dp - 1 + b.
// ---------------------------------------------------------------------------------------------
__global__ void pack_conv3_weight_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out, int cout,
                                         int cin, int KC, int dgrad) {
  const int gout = dgrad ? cin : cout;  // GEMM N channels
  const int gin = dgrad ? cout : cin;   // GEMM K channels
  const int nkc = gin / KC;
  const long long total = static_cast<long long>(gout) * gin * 27;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    long long r = i;
    const int k = static_cast<int>(r % KC);
    r /= KC;
    const int co = static_cast<int>(r % kCoBlk);
    r /= kCoBlk;
    const int kw = static_cast<int>(r % 3);
    r /= 3;
    const int kd = static_cast<int>(r % 3);
    r /= 3;
    const int kc = static_cast<int>(r % nkc);
    r /= nkc;
    const int kh = static_cast<int>(r % 3);
    r /= 3;
    const int cb = static_cast<int>(r);
    const int go = cb * kCoBlk + co;
    const int gi = kc * KC + k;
    float v;
    const int kdt = 2 - kd;   // the kd blocks are stored in descending tap order (ascending output plane)
    if (!dgrad)
      v = w[((static_cast<long long>(go) * cin + gi) * 3 + kdt) * 9 + kh * 3 + kw];
    else
      v = w[((static_cast<long long>(gi) * cin + go) * 3 + (2 - kdt)) * 9 + (2 - kh) * 3 + (2 - kw)];
    out[i] = __float2bfloat16(v);
  }
}

}  // namespace

int conv3_kc(int gemm_k_channels) { return (gemm_k_channels % 64 == 0) ? 64 : 32; }

}  // namespace spff

extern "C" {

int spff_pack_conv3_weight(const float* w, void* w_fwd, void* w_dgrad, int cout, int cin, void* stream) {
  SPFF_REQUIRE(cout % 32 == 0 && cin % 32 == 0, "pack_conv3_weight: cin %d / cout %d must be multiples of 32", cin,
               cout);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long total = static_cast<long long>(cout) * cin * 27;
  const int blocks = static_cast<int>((total + 255) / 256 < 4096 ? (total + 255) / 256 : 4096);
  // each operand holds two layouts back to back: the flattened-row kernel's, then the halo kernel's (conv3_halo.cu);
  // which one a launch reads depends on the plane size, unknown here
  if (w_fwd) {
    spff::pack_conv3_weight_kernel<<<blocks, 256, 0, st>>>(w, static_cast<__nv_bfloat16*>(w_fwd), cout, cin,
                                                           spff::conv3_kc(cin), 0);
    int e = spff::conv3_halo_pack(w, static_cast<__nv_bfloat16*>(w_fwd) + total, cout, cin, 0, st);
    if (e) return e;
  }
  if (w_dgrad) {
    spff::pack_conv3_weight_kernel<<<blocks, 256, 0, st>>>(w, static_cast<__nv_bfloat16*>(w_dgrad), cout, cin,
                                                           spff::conv3_kc(cout), 1);
    int e = spff::conv3_halo_pack(w, static_cast<__nv_bfloat16*>(w_dgrad) + total, cout, cin, 1, st);
    if (e) return e;
  }
  SPFF_CUDA(cudaGetLastError());
  return 0;
}

long long spff_conv3_packed_elems(int cin, int cout) { return 2LL * 27 * cin * cout; }

// partial statistics slots per sample of the forward kernel for this shape: qtiles * plane groups
static int conv3_stat_slots(spff_shape s) {
  const int hw = s.h * s.w;
  const int mstep = (spff::kTileM % s.w == 0) ? 128 : 126;
  const int qtiles = (hw + mstep - 1) / mstep;
  return qtiles * ((s.d + spff::kMaxPlanes - 1) / spff::kMaxPlanes);
}

static int conv3_common(const void* x, long long ldx, int cin, const void* wpk, void* y, long long ldy, int cout,
                        spff_shape s, float* stat_partial, void* stream, const char* who) {
  int e = spff_device_check();
  if (e) return e;
  SPFF_REQUIRE(x && wpk && y, "%s: null pointer", who);
  SPFF_REQUIRE(cin % 32 == 0 && cout % 32 == 0 && cin > 0 && cout > 0, "%s: channels (%d -> %d) must be multiples of 32",
               who, cin, cout);
  SPFF_REQUIRE(ldx >= cin && ldy >= cout && ldx % 8 == 0 && ldy % 8 == 0, "%s: bad channel pitch %lld / %lld", who, ldx,
               ldy);
  SPFF_REQUIRE(s.n > 0 && s.d > 0 && s.h > 0 && s.w > 0, "%s: empty shape", who);
  SPFF_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(y) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(wpk) & 15) == 0,
               "%s: pointers must be 16-byte aligned", who);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (spff::conv3_halo_applicable(s, cin))   // the halo-tile kernel (conv3_halo.cu; second half of the packed operand)
    return spff::conv3_halo_launch(x, ldx, cin, static_cast<const __nv_bfloat16*>(wpk) + 27LL * cin * cout, y, ldy, cout, s,
                                   stat_partial, st);
  // cin <= 64: one K chunk per tap, all 27 taps of a 32-channel output block stay resident in smem
  if (stat_partial) {
    if (cin == 64) return spff::launch_fprop<64, true, true>(x, ldx, cin, wpk, y, ldy, cout, s, stat_partial, st);
    if (cin == 32) return spff::launch_fprop<32, true, true>(x, ldx, cin, wpk, y, ldy, cout, s, stat_partial, st);
    if (spff::conv3_kc(cin) == 64)
      return spff::launch_fprop<64, false, true>(x, ldx, cin, wpk, y, ldy, cout, s, stat_partial, st);
    return spff::launch_fprop<32, false, true>(x, ldx, cin, wpk, y, ldy, cout, s, stat_partial, st);
  }
  if (cin == 64) return spff::launch_fprop<64, true, false>(x, ldx, cin, wpk, y, ldy, cout, s, nullptr, st);
  if (cin == 32) return spff::launch_fprop<32, true, false>(x, ldx, cin, wpk, y, ldy, cout, s, nullptr, st);
  if (spff::conv3_kc(cin) == 64) return spff::launch_fprop<64, false, false>(x, ldx, cin, wpk, y, ldy, cout, s, nullptr, st);
  return spff::launch_fprop<32, false, false>(x, ldx, cin, wpk, y, ldy, cout, s, nullptr, st);
}

int spff_conv3d_k3_fwd(const void* x, long long ldx, int cin, const void* w_fwd, void* y, long long ldy, int cout,
                       spff_shape s, void* stream) {
  return conv3_common(x, ldx, cin, w_fwd, y, ldy, cout, s, nullptr, stream, "conv3d_k3_fwd");
}

int spff_conv3d_k3_stat_slots(spff_shape s) { return conv3_stat_slots(s); }

int spff_conv3d_k3_fwd_stats(const void* x, long long ldx, int cin, const void* w_fwd, void* y, long long ldy, int cout,
                             spff_shape s, float* stat_partial, void* stream) {
  SPFF_REQUIRE(stat_partial, "conv3d_k3_fwd_stats: null partial buffer");
  return conv3_common(x, ldx, cin, w_fwd, y, ldy, cout, s, stat_partial, stream, "conv3d_k3_fwd_stats");
}

int spff_conv3d_k3_dgrad(const void* dy, long long lddy, int cout, const void* w_dgrad, void* dx, long long lddx,
                         int cin, spff_shape s, void* stream) {
  return conv3_common(dy, lddy, cout, w_dgrad, dx, lddx, cin, s, nullptr, stream, "conv3d_k3_dgrad");
}

/* dgrad that also writes the per-item column statistics of dx ({sum, sum of squares} per input channel, the layout of
 * spff_conv3d_k3_fwd_stats): the column sums of a decoder block's input gradient are the bias gradient of the
 * transposed conv that produced that input (models.py:668-672), so no separate pass over dx is needed. */
int spff_conv3d_k3_dgrad_stats(const void* dy, long long lddy, int cout, const void* w_dgrad, void* dx, long long lddx,
                               int cin, spff_shape s, float* stat_partial, void* stream) {
  SPFF_REQUIRE(stat_partial, "conv3d_k3_dgrad_stats: null partial buffer");
  return conv3_common(dy, lddy, cout, w_dgrad, dx, lddx, cin, s, stat_partial, stream, "conv3d_k3_dgrad_stats");
}

}  // extern "C"

namespace spff {
int conv3_rows_stat_slots(spff_shape s) { return spff_conv3d_k3_stat_slots(s); }
}  // namespace spff
