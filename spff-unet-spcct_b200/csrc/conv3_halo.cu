// 3x3x3 convolution (stride 1, zero pad 1, no bias) on tcgen05 — "halo tile" formulation, forward and input gradient.
// Replaces F.conv3d / convolution_backward(input) behind `_conv3x3xk` (reference innovative3D/models.py:616-618) for
// planes whose height is a multiple of 16 and width a multiple of 8 (every level of the benchmark shapes); other
// shapes run the flattened-row kernel of conv3_fprop.cu.
//
// Why a second formulation. conv3_fprop.cu folds kw into the GEMM N dimension and adds the +-1-position shifted
// partial sums in the epilogue (TMEM lanes are positions, so the shift is a cross-lane exchange: 3x the TMEM reads,
// 64 shuffles and a CTA-wide barrier per plane). With 32 input channels that epilogue, not the MMAs, sets the pace
// (tensor pipe 35-38 % active, profiles/r01f_conv3_fprop_ncu_full.md). The tensor core, however, applies the shared
// memory swizzle to the ABSOLUTE address (scripts/micro/umma_rowshift.cu, profiles/r02_umma_rowshift_experiment.log):
// a K-major operand may start at any row of a TMA-written tile and step between its 8-row groups by any stride. So:
//
//   * one work item = (sample, 16 x 8 output positions, block of CO output channels, plane group);
//   * per input plane and 32-channel K chunk ONE TMA box [18 x 10 positions][32 ch] lands in shared memory (h and w halo,
//     out-of-bounds rows zero filled = the padding); the A operand of tap (kh, kw) is that same tile read from row
//     kh*10 + kw with a group stride of 10 rows. No per-kh reload (3x less L2 -> SM traffic for A), no kw fold;
//   * an input plane feeds up to three output planes (kd): their accumulators sit side by side in TMEM and the kd weight
//     blocks are packed in that order, so the three are ONE MMA with N = 3*CO;
//   * accumulators hold final sums: the epilogue is TMEM -> registers -> (statistics) -> bf16 -> global.
//
// Loop order: K chunk -> tap -> input plane. All input planes of the item for one K chunk stay resident (double
// buffered: the next chunk / next item loads while this one multiplies); weights stream through a ring in
// [tap][kd][CO][32] chunks, or stay resident for the whole launch when there is a single K chunk (cin = 32).
//
// Warp roles (192 threads, one CTA per SM, persistent): warps 0-3 epilogue, warp 4 TMA producer, warp 5 MMA issuer.
#include "common.h"
#include "ptx.cuh"

namespace spff {

namespace {

constexpr int kThreads = 192;
constexpr int kTH = 16, kTW = 8;                          // output tile (positions): M = 128
constexpr int kHaloW = kTW + 2, kHaloH = kTH + 2;
constexpr int kKC = 32;                                   // channels per K chunk: 64-byte rows, 64B swizzle
constexpr int kRowBytes = kKC * 2;
constexpr int kPlaneTile = kHaloW * kHaloH * kRowBytes;   // 11520 bytes per (input plane, K chunk)
constexpr int kPlaneStride = 11776;                       // next multiple of the 512-byte swizzle atom
constexpr int kMaxG = 5;                                  // output planes per group: the same grouping as conv3_fprop.cu (same statistics slots)
constexpr int kMaxB = 9;                                  // weight ring stages (9 = all taps of a single K chunk)

struct HaloParams {
  int n, d, h, w;
  int nkc, ncb;
  int tiles_w, tiles_h;
  int G, ngroups;
  int nplanes;        // planes per A set: min(d, G + 2)
  int nb;             // weight ring stages
  int acc_sets;       // TMEM accumulator sets (2 when 2*G*CO <= 512)
  long long items;    // n * tiles_h * tiles_w * ngroups * ncb
  __nv_bfloat16* y;
  long long ldy;
  float* stat_partial;   // STATS: [n][stat_slots][2][cout]; this kernel fills tiles_h*tiles_w*ngroups of them and zeroes the rest
  int stat_slots;        // slots per sample as spff_conv3d_k3_stat_slots reports them (the flattened-row kernel's count, >= ours)
  int cout;
  unsigned long long* prof;   // optional per-CTA cycle counters [grid][8] (spff_debug_set key 4), else null
  int dbg;               // timing experiments only (spff_debug_set key 5): 1 = aligned tap offsets, 2 = dense group stride (wrong results)
};

// Sum the 32 per-lane values of each of 32 columns across the warp: afterwards a[0] of lane L holds the total of
// column L (transpose-reduce: 31 shuffles instead of 32 x 5).
__device__ __forceinline__ void warp_column_sums32(float* a, int lane) {
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

// D5: depth 5 in one plane group (the five energy bins: every SPCT-family launch) - the per-plane tables of the MMA issuer
// are compile-time constants.
template <int CO, bool WRES, bool STATS, bool D5>
__global__ void __launch_bounds__(kThreads, 1)
conv3_halo_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w,
                  const HaloParams p) {
  constexpr int kChunk = 3 * CO * kRowBytes;   // one weight chunk: [kd block 0..2][CO rows][32 ch]
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sA = smem;
  uint8_t* sB = sA + 2 * p.nplanes * kPlaneStride;
  float* sS = reinterpret_cast<float*>(sB + p.nb * kChunk);              // STATS: [4 warps][2][CO]
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(sS) + 4 * 2 * CO * 4);
  uint64_t* afull = bars;                 // [2]
  uint64_t* aempty = bars + 2;            // [2]
  uint64_t* bfull = bars + 4;             // [kMaxB]
  uint64_t* bempty = bfull + kMaxB;       // [kMaxB]
  uint64_t* acc_full = bempty + kMaxB;    // [2]
  uint64_t* acc_empty = acc_full + 2;     // [2][kMaxG]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2 * kMaxG);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&afull[i], 1);
      mbar_init(&aempty[i], 1);
      mbar_init(&acc_full[i], 1);
    }
    for (int i = 0; i < kMaxB; ++i) {
      mbar_init(&bfull[i], 1);
      mbar_init(&bempty[i], 1);
    }
    for (int i = 0; i < 2 * kMaxG; ++i) mbar_init(&acc_empty[i], 128);
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
  // profiling: cycles spent in each kind of wait (only when p.prof is set)
  unsigned long long w0 = 0, w1 = 0, w2 = 0;
  const long long t_start = clock64();
#define SPFF_TIMED_WAIT(acc, bar, par)               \
  do {                                               \
    if (p.prof) {                                    \
      const long long t_ = clock64();                \
      mbar_wait(bar, par);                           \
      acc += static_cast<unsigned long long>(clock64() - t_); \
    } else {                                         \
      mbar_wait(bar, par);                           \
    }                                                \
  } while (0)

  // WRES: a CTA keeps one output-channel block (its weights never leave shared memory) and strides over positions.
  const long long item0 = WRES ? (blockIdx.x / p.ncb) : blockIdx.x;
  const long long istep = WRES ? (gridDim.x / p.ncb) : gridDim.x;
  const long long nitems = WRES ? p.items / p.ncb : p.items;
  auto decode = [&](long long item, int& n, int& th, int& tw, int& pg, int& cb) {
    long long r = item;
    if (WRES) {
      cb = blockIdx.x % p.ncb;
    } else {
      cb = static_cast<int>(r % p.ncb);
      r /= p.ncb;
    }
    pg = static_cast<int>(r % p.ngroups);
    r /= p.ngroups;
    tw = static_cast<int>(r % p.tiles_w);
    r /= p.tiles_w;
    th = static_cast<int>(r % p.tiles_h);
    n = static_cast<int>(r / p.tiles_h);
  };

  if (warp == 4) {
    // ------------------------------------------------------------------ TMA producer
    const bool leader = elect_one() != 0;
    int as = 0, bs = 0;
    uint32_t aph = 0, bph = 0;
    if (WRES) {
      const int cb = blockIdx.x % p.ncb;
      for (int tap = 0; tap < 9; ++tap) {
        if (leader) {
          mbar_expect_tx(&bfull[tap], kChunk);
          tma_load_2d(sB + tap * kChunk, &tmap_w, &bfull[tap], 0, (cb * 9 + tap) * 3 * CO);
        }
      }
    }
    for (long long item = item0; item < nitems; item += istep) {
      int n, th, tw, pg, cb;
      decode(item, n, th, tw, pg, cb);
      const int d0 = pg * p.G;
      const int dend = min(p.d, d0 + p.G);
      const int dlo = max(0, d0 - 1);
      const int dhi = min(p.d - 1, dend);
      const int np = dhi - dlo + 1;
      for (int kc = 0; kc < p.nkc; ++kc) {
        SPFF_TIMED_WAIT(w0, &aempty[as], aph ^ 1);
        if (leader) {
          mbar_expect_tx(&afull[as], np * kPlaneTile);
          for (int q = 0; q < np; ++q)
            tma_load_5d(sA + (as * p.nplanes + q) * kPlaneStride, &tmap_x, &afull[as], kc * kKC, tw * kTW - 1, th * kTH - 1,
                        dlo + q, n);
        }
        as ^= 1;
        if (as == 0) aph ^= 1;
        if (!WRES) {
          for (int tap = 0; tap < 9; ++tap) {
            SPFF_TIMED_WAIT(w1, &bempty[bs], bph ^ 1);
            if (leader) {
              mbar_expect_tx(&bfull[bs], kChunk);
              tma_load_2d(sB + bs * kChunk, &tmap_w, &bfull[bs], 0, ((cb * p.nkc + kc) * 9 + tap) * 3 * CO);
            }
            if (++bs == p.nb) {
              bs = 0;
              bph ^= 1;
            }
          }
        }
      }
    }
  } else if (warp == 5) {
    // ------------------------------------------------------------------ MMA issuer
    // Warp-uniform control flow, one elected lane issues. The inner (tap, plane) loop must cost less than the ~100 tensor
    // cycles of the two K = 16 MMAs it issues: everything that depends on the plane only (TMEM column, kd block, N) is
    // tabulated per item in registers (the plane loop is fully unrolled), descriptors advance by 32-bit additions on
    // the (address >> 4) field - the whole dynamic shared window is < 256 KB, so the field never carries out.
    const bool leader = elect_one() != 0;
    int as = 0, bs = 0;
    uint32_t aph = 0, bph = 0;
    uint32_t accpar0 = 0, accpar1 = 0;   // per set, per plane: parity of its next acc_empty wait
    int set = 0;
    if (WRES)
      for (int tap = 0; tap < 9; ++tap) mbar_wait(&bfull[tap], 0);
    const uint64_t a_hi = make_smem_desc_hi(16, (p.dbg & 2) ? 8 * kRowBytes : kHaloW * kRowBytes, kSwizzle64);   // 8-row groups are 10 rows apart
    const uint64_t b_hi = make_smem_desc_hi(16, 8 * kRowBytes, kSwizzle64);
    const uint32_t a_lo0 = (smem_u32(sA) >> 4) & 0x3FFF;
    const uint32_t b_lo0 = (smem_u32(sB) >> 4) & 0x3FFF;
    constexpr int kQ = kMaxG + 2;        // input planes per item, at most
    for (long long item = item0; item < nitems; item += istep) {
      int n, th, tw, pg, cb;
      decode(item, n, th, tw, pg, cb);
      const int d0 = pg * p.G;
      const int dend = min(p.d, d0 + p.G);
      const int dlo = max(0, d0 - 1);
      const int dhi = min(p.d - 1, dend);
      const int np = dhi - dlo + 1;
      const uint32_t acc_col = tmem_base + set * (p.G * CO);
      // per input plane q: first output plane, number of output planes, TMEM column, kd block offset, instruction descriptor
      uint32_t q_col[kQ], q_boff[kQ], q_id[kQ];
      int q_jlo[kQ], q_cnt[kQ];
#pragma unroll
      for (int q = 0; q < kQ; ++q) {
        const int dp = dlo + q;
        const int jlo = max(dp - 1, d0) - d0, jhi = min(dp + 1, dend - 1) - d0;
        q_jlo[q] = jlo;
        q_cnt[q] = jhi - jlo + 1;
        q_col[q] = acc_col + jlo * CO;
        q_boff[q] = static_cast<uint32_t>(((jlo + d0 - dp + 1) * CO * kRowBytes) >> 4);   // blocks are stored 2 - kd
        q_id[q] = make_idesc_bf16(128, q_cnt[q] * CO, 0, 0);
      }
      for (int kc = 0; kc < p.nkc; ++kc) {
        SPFF_TIMED_WAIT(w0, &afull[as], aph);
        tc_fence_after();
        const uint32_t a_set = a_lo0 + static_cast<uint32_t>((as * p.nplanes * kPlaneStride) >> 4);
        uint32_t a_tap = a_set;
        for (int kh = 0; kh < 3; ++kh) {
#pragma unroll
          for (int kw = 0; kw < 3; ++kw) {
            const int tap = kh * 3 + kw;
            uint32_t b_chunk;
            if (WRES) {
              b_chunk = b_lo0 + static_cast<uint32_t>((tap * kChunk) >> 4);
            } else {
              SPFF_TIMED_WAIT(w1, &bfull[bs], bph);
              tc_fence_after();
              b_chunk = b_lo0 + static_cast<uint32_t>((bs * kChunk) >> 4);
            }
            if (kc == 0 && tap == 0) {
              // first MMAs of the item: an accumulator that has not been written yet must have been drained by the
              // epilogue (acc_empty) and starts with accumulate = 0 - split each plane's run by that status
              uint32_t touched = 0;
#pragma unroll
              for (int q = 0; q < kQ; ++q) {
                if (q < np) {
                  const int jhi = q_jlo[q] + q_cnt[q] - 1;
                  int j = q_jlo[q];
                  while (j <= jhi) {
                    const uint32_t tj = (touched >> j) & 1u;
                    int cnt = 1;
                    while (j + cnt <= jhi && ((touched >> (j + cnt)) & 1u) == tj) ++cnt;
                    if (!tj) {
                      for (int r = j; r < j + cnt; ++r) {
                        const uint32_t par = set ? accpar1 : accpar0;
                        SPFF_TIMED_WAIT(w2, &acc_empty[set * kMaxG + r], ((par >> r) & 1u) ^ 1u);
                        if (set) accpar1 ^= 1u << r; else accpar0 ^= 1u << r;
                      }
                      tc_fence_after();
                    }
                    touched |= ((1u << cnt) - 1u) << j;
                    const uint64_t ad = a_hi | static_cast<uint64_t>(a_tap + static_cast<uint32_t>((q * kPlaneStride) >> 4));
                    const uint64_t bd = b_hi | static_cast<uint64_t>(b_chunk + q_boff[q] +
                                                                     static_cast<uint32_t>(((j - q_jlo[q]) * CO * kRowBytes) >> 4));
                    const uint32_t id = make_idesc_bf16(128, cnt * CO, 0, 0);
                    if (leader) {
                      umma_bf16(acc_col + j * CO, ad, bd, id, tj);
                      umma_bf16(acc_col + j * CO, ad + 2, bd + 2, id, 1u);
                    }
                    j += cnt;
                  }
                }
              }
            } else if constexpr (D5) {
              if (leader) {      // five input planes, immediates only: ~4 instructions per MMA
#pragma unroll
                for (int q = 0; q < 5; ++q) {
                  const int jlo = q > 0 ? q - 1 : 0, jhi = q < 4 ? q + 1 : 4;
                  const uint32_t id = make_idesc_bf16(128, (jhi - jlo + 1) * CO, 0, 0);
                  const uint64_t ad = a_hi | static_cast<uint64_t>(a_tap + static_cast<uint32_t>((q * kPlaneStride) >> 4));
                  const uint64_t bd = b_hi | static_cast<uint64_t>(b_chunk + static_cast<uint32_t>(((jlo - q + 1) * CO * kRowBytes) >> 4));
                  umma_bf16(acc_col + jlo * CO, ad, bd, id, 1u);
                  umma_bf16(acc_col + jlo * CO, ad + 2, bd + 2, id, 1u);
                }
              }
            } else {
#pragma unroll
              for (int q = 0; q < kQ; ++q) {
                if (q < np) {
                  const uint64_t ad = a_hi | static_cast<uint64_t>(a_tap + static_cast<uint32_t>((q * kPlaneStride) >> 4));
                  const uint64_t bd = b_hi | static_cast<uint64_t>(b_chunk + q_boff[q]);
                  if (leader) {
                    umma_bf16(q_col[q], ad, bd, q_id[q], 1u);
                    umma_bf16(q_col[q], ad + 2, bd + 2, q_id[q], 1u);
                  }
                }
              }
            }
            if (!WRES) {
              if (leader) umma_commit(&bempty[bs]);
              if (++bs == p.nb) {
                bs = 0;
                bph ^= 1;
              }
            }
            if (!(p.dbg & 1)) a_tap += kRowBytes >> 4;     // next kw: one row of the halo tile
          }
          if (!(p.dbg & 1)) a_tap += ((kHaloW - 3) * kRowBytes) >> 4;        // next kh: one image row of the halo tile
        }
        if (leader) umma_commit(&aempty[as]);
        as ^= 1;
        if (as == 0) aph ^= 1;
      }
      if (leader) umma_commit(&acc_full[set]);
      if (p.acc_sets == 2) set ^= 1;
    }
  } else {
    // ------------------------------------------------------------------ epilogue (warps 0-3)
    uint32_t fullpar0 = 0, fullpar1 = 0;
    int set = 0;
    const int m = warp * 32 + lane;              // tile row == TMEM lane: 8 consecutive w per image row
    const int hh = m >> 3, ww = m & 7;
    for (long long item = item0; item < nitems; item += istep) {
      int n, th, tw, pg, cb;
      decode(item, n, th, tw, pg, cb);
      const int d0 = pg * p.G;
      const int dend = min(p.d, d0 + p.G);
      const int h = th * kTH + hh, w = tw * kTW + ww;
      float ssum[STATS ? CO : 1], ssq[STATS ? CO : 1];
      if constexpr (STATS) {
#pragma unroll
        for (int c = 0; c < CO; ++c) ssum[c] = ssq[c] = 0.f;
      }
      SPFF_TIMED_WAIT(w0, &acc_full[set], set ? fullpar1 : fullpar0);
      if (set) fullpar1 ^= 1; else fullpar0 ^= 1;
      tc_fence_after();
      for (int d = d0; d < dend; ++d) {
        const int j = d - d0;
        const uint32_t tcol = tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + set * (p.G * CO) + j * CO;
        uint32_t v[CO];
#pragma unroll
        for (int c0 = 0; c0 < CO; c0 += 32) tmem_ld_32x32(tcol + c0, *reinterpret_cast<uint32_t(*)[32]>(&v[c0]));
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(&acc_empty[set * kMaxG + j]);   // the accumulator is in registers: the next item may overwrite it
        uint32_t packed[CO / 2];
#pragma unroll
        for (int c = 0; c < CO; c += 2) {
          const float a = __uint_as_float(v[c]), b = __uint_as_float(v[c + 1]);
          if constexpr (STATS) {
            ssum[c] += a;
            ssq[c] = fmaf(a, a, ssq[c]);
            ssum[c + 1] += b;
            ssq[c + 1] = fmaf(b, b, ssq[c + 1]);
          }
          packed[c / 2] = pack_bf16x2(a, b);
        }
        __nv_bfloat16* dst =
            p.y + (((static_cast<long long>(n) * p.d + d) * p.h + h) * p.w + w) * p.ldy + cb * CO;
        if ((p.ldy & 15) == 0 && (reinterpret_cast<uintptr_t>(p.y) & 31) == 0) {   // 32-byte aligned rows: one full sector per store instead of two halves
#pragma unroll
          for (int q = 0; q < CO / 16; ++q)
            st_global_v8(dst + 16 * q, packed[8 * q], packed[8 * q + 1], packed[8 * q + 2], packed[8 * q + 3], packed[8 * q + 4],
                         packed[8 * q + 5], packed[8 * q + 6], packed[8 * q + 7]);
        } else {
          uint4* d4 = reinterpret_cast<uint4*>(dst);
#pragma unroll
          for (int q = 0; q < CO / 8; ++q)
            d4[q] = make_uint4(packed[4 * q], packed[4 * q + 1], packed[4 * q + 2], packed[4 * q + 3]);
        }
      }
      if constexpr (STATS) {
        // item totals: columns across the 32 rows of each warp, then across the 4 warps through shared memory
#pragma unroll
        for (int c0 = 0; c0 < CO; c0 += 32) {
          warp_column_sums32(&ssum[c0], lane);
          warp_column_sums32(&ssq[c0], lane);
          sS[(warp * 2 + 0) * CO + c0 + lane] = ssum[c0];
          sS[(warp * 2 + 1) * CO + c0 + lane] = ssq[c0];
        }
        named_bar_sync(2, 128);
        for (int t = threadIdx.x; t < 2 * CO; t += 128) {
          const int which = t / CO, col = t % CO;
          const float tot = sS[(0 * 2 + which) * CO + col] + sS[(1 * 2 + which) * CO + col] +
                            sS[(2 * 2 + which) * CO + col] + sS[(3 * 2 + which) * CO + col];
          const int mine = p.tiles_h * p.tiles_w * p.ngroups;
          const int sidx = (th * p.tiles_w + tw) * p.ngroups + pg;
          float* row = p.stat_partial + (static_cast<long long>(n) * p.stat_slots * 2 + which) * p.cout + cb * CO + col;
          row[static_cast<long long>(sidx) * 2 * p.cout] = tot;
          for (int e = sidx + mine; e < p.stat_slots; e += mine) row[static_cast<long long>(e) * 2 * p.cout] = 0.f;
        }
        named_bar_sync(2, 128);   // sS is reused by the next item
      }
      if (p.acc_sets == 2) set ^= 1;
    }
  }

  if (p.prof && lane == 0 && (warp == 0 || warp >= 4)) {
    // [cta][role: 0 epilogue, 1 producer, 2 issuer][total, wait0, wait1, wait2]
    unsigned long long* o = p.prof + (static_cast<size_t>(blockIdx.x) * 3 + (warp == 0 ? 0 : warp - 3)) * 4;
    o[0] = static_cast<unsigned long long>(clock64() - t_start);
    o[1] = w0;
    o[2] = w1;
    o[3] = w2;
  }
#undef SPFF_TIMED_WAIT
  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// nn.Conv3d weight [Cout][Cin][3][3][3] fp32 -> bf16 operand of the halo kernel:
//   P[cb][kc][tap = kh*3 + kw][b][co < CO][k < 32],  value = w[cb*CO + co][kc*32 + k][2 - b][kh][kw]            (forward)
//   P[cb][kc][tap][b][ci < CO][k < 32],              value = w[kc*32 + k][cb*CO + ci][b][2 - kh][2 - kw]       (dgrad)
// b = 2 - kd: kd blocks in descending tap order, so that block b serves output plane dp - 1 + b.
__global__ void pack_conv3_halo_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out, int cout, int cin, int CO,
                                       int dgrad) {
  const int gout = dgrad ? cin : cout;   // GEMM N channels
  const int gin = dgrad ? cout : cin;    // GEMM K channels
  const int nkc = gin / kKC;
  const long long total = static_cast<long long>(gout) * gin * 27;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    long long r = i;
    const int k = static_cast<int>(r % kKC);
    r /= kKC;
    const int co = static_cast<int>(r % CO);
    r /= CO;
    const int b = static_cast<int>(r % 3);
    r /= 3;
    const int tap = static_cast<int>(r % 9);
    r /= 9;
    const int kc = static_cast<int>(r % nkc);
    r /= nkc;
    const int cb = static_cast<int>(r);
    const int kh = tap / 3, kw = tap % 3;
    const int go = cb * CO + co;
    const int gi = kc * kKC + k;
    float v;
    if (!dgrad)
      v = w[((static_cast<long long>(go) * cin + gi) * 3 + (2 - b)) * 9 + kh * 3 + kw];
    else
      v = w[((static_cast<long long>(gi) * cin + go) * 3 + b) * 9 + (2 - kh) * 3 + (2 - kw)];
    out[i] = __float2bfloat16(v);
  }
}

// dynamic shared memory of one CTA: two A sets, `nb` weight chunks, the statistics exchange, barriers, alignment slack
size_t halo_smem_bytes(int co, int nplanes, int nb) {
  return static_cast<size_t>(2) * nplanes * kPlaneStride + static_cast<size_t>(nb) * 3 * co * kRowBytes + 4 * 2 * co * 4 +
         (4 + 2 * kMaxB + 2 + 2 * kMaxG) * 8 + 16 + 1024;
}
constexpr size_t kSmemLimit = 227 * 1024;

struct HaloGeom {
  int ngroups, G, nplanes;
};
HaloGeom halo_geom(int d) {
  HaloGeom g;
  g.ngroups = (d + kMaxG - 1) / kMaxG;
  g.G = (d + g.ngroups - 1) / g.ngroups;
  g.nplanes = (g.G + 2 < d) ? g.G + 2 : d;
  return g;
}

template <int CO, bool WRES, bool STATS, bool D5>
int launch_halo(const void* x, long long ldx, int cin, const void* wpk, void* y, long long ldy, int cout, spff_shape s,
                float* stat_partial, cudaStream_t stream) {
  constexpr int kChunk = 3 * CO * kRowBytes;
  HaloParams p;
  p.n = s.n;
  p.d = s.d;
  p.h = s.h;
  p.w = s.w;
  p.nkc = cin / kKC;
  p.ncb = cout / CO;
  p.tiles_w = s.w / kTW;
  p.tiles_h = s.h / kTH;
  const HaloGeom g = halo_geom(s.d);
  p.ngroups = g.ngroups;
  p.G = g.G;
  p.nplanes = g.nplanes;
  p.acc_sets = (2 * p.G * CO <= 512) ? 2 : 1;
  p.items = static_cast<long long>(s.n) * p.tiles_h * p.tiles_w * p.ngroups * p.ncb;
  p.y = static_cast<__nv_bfloat16*>(y);
  p.ldy = ldy;
  p.stat_partial = stat_partial;
  p.stat_slots = conv3_rows_stat_slots(s);
  p.cout = cout;
  p.dbg = debug_flag(5);
  p.prof = reinterpret_cast<unsigned long long*>(debug_value(4));
  // the weight ring is as deep as shared memory allows (at most 9 chunks = every tap of one K chunk)
  int nb = kMaxB;
  while (nb > 2 && halo_smem_bytes(CO, p.nplanes, nb) > kSmemLimit) --nb;
  if (halo_smem_bytes(CO, p.nplanes, nb) > kSmemLimit || (WRES && nb < 9)) {
    set_error("conv3 halo: %d planes x %d output channels do not fit in shared memory", p.nplanes, CO);
    return SPFF_ERR_BAD_ARGUMENT;
  }
  p.nb = nb;
  const size_t smem = halo_smem_bytes(CO, p.nplanes, nb);
  static_assert(kChunk == 3 * CO * kRowBytes, "chunk size");

  CUtensorMap tx, tw;
  {
    uint64_t dims[5] = {static_cast<uint64_t>(cin), static_cast<uint64_t>(s.w), static_cast<uint64_t>(s.h),
                        static_cast<uint64_t>(s.d), static_cast<uint64_t>(s.n)};
    uint64_t str[4] = {static_cast<uint64_t>(ldx) * 2, static_cast<uint64_t>(ldx) * 2 * s.w,
                       static_cast<uint64_t>(ldx) * 2 * s.w * s.h, static_cast<uint64_t>(ldx) * 2 * s.w * s.h * s.d};
    uint32_t box[5] = {kKC, kHaloW, kHaloH, 1, 1};
    int e = encode_tmap_bf16(&tx, x, 5, dims, str, box, kRowBytes);
    if (e) return e;
  }
  {
    const uint64_t rows = static_cast<uint64_t>(p.ncb) * p.nkc * 9 * 3 * CO;
    uint64_t dims[2] = {kKC, rows};
    uint64_t str[1] = {kRowBytes};
    uint32_t box[2] = {kKC, 3 * CO};
    int e = encode_tmap_bf16(&tw, wpk, 2, dims, str, box, kRowBytes);
    if (e) return e;
  }
  static size_t attr_dev[kMaxDevices] = {};   // largest opt-in so far, per device (and per template instance)
  size_t& attr = attr_dev[current_device()];
  if (attr < smem) {
    SPFF_CUDA(cudaFuncSetAttribute(conv3_halo_kernel<CO, WRES, STATS, D5>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   static_cast<int>(smem)));
    attr = smem;
  }
  int ctas = debug_ctas() > 0 ? debug_ctas() : num_sms();
  if (WRES) {
    const long long pos_items = p.items / p.ncb;
    long long per_cb = ctas / p.ncb;
    if (per_cb < 1) per_cb = 1;
    if (per_cb > pos_items) per_cb = pos_items;
    ctas = static_cast<int>(per_cb * p.ncb);
  } else if (p.items < ctas) {
    ctas = static_cast<int>(p.items);
  }
  conv3_halo_kernel<CO, WRES, STATS, D5><<<ctas, kThreads, smem, stream>>>(tx, tw, p);
  SPFF_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace

int conv3_halo_co(int gemm_n_channels) { return (gemm_n_channels % 64 == 0) ? 64 : 32; }

// The plane must tile by 16 x 8 boxes (any depth fits: two A sets of at most 7 planes + a two-chunk weight ring). Which
// of the two kernels then runs is a measured choice (scripts/bench_conv.py, 256-slice groups on B200): this one is
// 6-25 % faster with one K chunk (gemm K = 32: weights resident) and 13-20 % faster from four K chunks on (gemm K >= 128:
// the lighter epilogue and the 3x smaller A traffic win), the flattened-row kernel is as fast or faster at gemm K = 64
// (its 64-channel K chunks need half as many MMA issues). Both produce the same number of statistics slots per sample
// (positions / 128 x plane groups), so the choice may depend on the channel counts.
bool conv3_halo_applicable(spff_shape s, int gemm_k) {
  const int force = debug_flag(7);                      // test hook: 1 = always the flattened-row kernel, 2 = always this one
  if (force == 1) return false;
  if (s.h % kTH || s.w % kTW) return false;
  if (halo_smem_bytes(64, halo_geom(s.d).nplanes, 2) > kSmemLimit) return false;
  return force == 2 || gemm_k == kKC || gemm_k >= 4 * kKC;
}

int conv3_halo_pack(const float* w, void* out, int cout, int cin, int dgrad, cudaStream_t st) {
  const long long total = static_cast<long long>(cout) * cin * 27;
  const int blocks = static_cast<int>((total + 255) / 256 < 4096 ? (total + 255) / 256 : 4096);
  pack_conv3_halo_kernel<<<blocks, 256, 0, st>>>(w, static_cast<__nv_bfloat16*>(out), cout, cin,
                                                 conv3_halo_co(dgrad ? cin : cout), dgrad);
  SPFF_CUDA(cudaGetLastError());
  return 0;
}

int conv3_halo_launch(const void* x, long long ldx, int cin, const void* wpk, void* y, long long ldy, int cout, spff_shape s,
                      float* stat_partial, cudaStream_t st) {
  const int co = conv3_halo_co(cout);
  // a single K chunk: all nine weight chunks of an output-channel block stay in shared memory for the whole launch
  const bool wres = (cin == kKC) && halo_smem_bytes(co, halo_geom(s.d).nplanes, 9) <= kSmemLimit;
  const bool d5 = (s.d == 5);
#define SPFF_HALO(CO_, W_, S_)                                                                            \
  do {                                                                                                    \
    if (d5) return launch_halo<CO_, W_, S_, true>(x, ldx, cin, wpk, y, ldy, cout, s, stat_partial, st);  \
    return launch_halo<CO_, W_, S_, false>(x, ldx, cin, wpk, y, ldy, cout, s, stat_partial, st);         \
  } while (0)
  if (co == 64) {
    if (wres) {
      if (stat_partial) SPFF_HALO(64, true, true);
      SPFF_HALO(64, true, false);
    }
    if (stat_partial) SPFF_HALO(64, false, true);
    SPFF_HALO(64, false, false);
  }
  if (wres) {
    if (stat_partial) SPFF_HALO(32, true, true);
    SPFF_HALO(32, true, false);
  }
  if (stat_partial) SPFF_HALO(32, false, true);
  SPFF_HALO(32, false, false);
#undef SPFF_HALO
}

}  // namespace spff
