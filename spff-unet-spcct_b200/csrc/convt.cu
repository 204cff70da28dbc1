// ConvTranspose3d with kernel = stride = (1,2,2) (+bias), the decoder up-sampling of
// UNet3D_SpectralCore (reference innovative3D/models.py:668-672, applied at :698-700).
// The windows do not overlap, so every quadrant (i,j) of the output is a plain GEMM:
//   forward  y[n,d,2h+i,2w+j][co] = b[co] + sum_ci x[n,d,h,w][ci] * W[ci][co][i][j]
//   dgrad    dx[n,d,h,w][ci]      = sum_{i,j,co} dy[n,d,2h+i,2w+j][co] * W[ci][co][i][j]
//   wgrad    dW[ci][co][i][j]     = sum_pos x[pos][ci] * dy[quadrant(i,j) of pos][co]
// Quadrants of the fine grid are expressed as rank-5 TMA tensor maps with doubled h/w strides, so
// the same position tile (a box of the coarse grid) addresses x and all four dy/y quadrants.
//   * forward / dgrad: K-major tcgen05 GEMM, M = 128 positions, accumulators double-buffered in
//     TMEM, epilogue adds the bias and scatters rows straight into the destination view (the `up`
//     half of the skip-concat buffer — the torch.cat of models.py:687-691 never materialises).
//   * wgrad: MN-major operands as in conv3_wgrad.cu; the four dy quadrant tiles are stacked in M.
#include "common.h"
#include "ptx.cuh"

namespace spff {
namespace {

constexpr int kThreads = 192;

struct PwParams {
  int n, d, h, w;        // coarse grid
  int bw, bh, tiles_w, tiles_h;
  long long ntiles;      // n*d*tiles_h*tiles_w
  int nkc;               // K chunks per tap
  int ntap;              // taps reduced over K (1 forward, 4 dgrad)
  int nquad;             // output quadrants (4 forward, 1 dgrad)
  int N;                 // output channels (UMMA N)
  __nv_bfloat16* out;
  long long ldo;
  int ho, wo;            // output grid extents
  int scale;             // 2 forward (rows scatter to 2h+i, 2w+j), 1 dgrad
  int dscale;            // output planes per input plane: 2 for the (2,2,2) forward (plane 2d + (q >> 2)), else 1
  int nfull, noff;       // rows of one packed weight block and the first row this launch reads (N-split of wide outputs)
  int nfold;             // forward: output quadrants folded into one MMA tile (N_mma = nfold * N <= 256): the position tile
                         // is loaded once for all of them instead of once per quadrant
  const float* bias;     // [N] or null
};

struct PwMaps {
  CUtensorMap a[8];
  CUtensorMap b;
};

template <int KC>
struct PwSmem {
  static constexpr int kStages = 4;
  static constexpr int kABytes = 128 * KC * 2;
  static constexpr int kBBytes = 256 * KC * 2;
  static constexpr int kStage = kABytes + kBBytes;
  static constexpr int kOffBar = kStages * kStage;
  static constexpr int kNumBars = 2 * kStages + 4;
  static constexpr int kOffTmem = kOffBar + kNumBars * 8;
  static constexpr int kOffBias = kOffTmem + 16;          // fp32 bias of the launch's N <= 256 output channels
  static constexpr int kTotal = kOffBias + 256 * 4;
};

template <int KC>
__global__ void __launch_bounds__(kThreads, 1) pw_gemm_kernel(const __grid_constant__ PwMaps maps, const PwParams p) {
  using L = PwSmem<KC>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::kOffBar);
  uint64_t* full = bars;
  uint64_t* empty = bars + L::kStages;
  uint64_t* acc_full = bars + 2 * L::kStages;
  uint64_t* acc_empty = acc_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + L::kOffTmem);
  float* sbias = reinterpret_cast<float*>(smem + L::kOffBias);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int i = 0; i < L::kStages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_empty[i], 128);
    }
    fence_mbar_init();
  }
  for (int i = threadIdx.x; i < p.N; i += kThreads) sbias[i] = p.bias ? __ldg(p.bias + p.noff + i) : 0.f;
  if (warp == 4 && lane == 0) {
    for (int t = 0; t < p.ntap; ++t) tma_prefetch_desc(&maps.a[t]);
    tma_prefetch_desc(&maps.b);
  }
  if (warp == 5) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int ngroups = p.nquad / p.nfold;            // quadrant groups per position tile
  const long long items = p.ntiles * ngroups;
  const int rows = p.bw * p.bh;
  auto decode = [&](long long item, int& qo, int& w0, int& h0, int& dd, int& n) {
    qo = static_cast<int>(item % ngroups) * p.nfold;   // first quadrant of the group
    long long t = item / ngroups;
    w0 = static_cast<int>(t % p.tiles_w) * p.bw;
    t /= p.tiles_w;
    h0 = static_cast<int>(t % p.tiles_h) * p.bh;
    t /= p.tiles_h;
    dd = static_cast<int>(t % p.d);
    n = static_cast<int>(t / p.d);
  };

  if (warp == 4) {
    const bool leader = elect_one() != 0;   // warp-uniform control flow, one elected lane issues
    {
      int s = 0;
      uint32_t ph = 0;
      const uint32_t bytes = rows * KC * 2 + p.nfold * p.N * KC * 2;
      for (long long item = blockIdx.x; item < items; item += gridDim.x) {
        int qo, w0, h0, dd, n;
        decode(item, qo, w0, h0, dd, n);
        for (int t = 0; t < p.ntap; ++t) {
          for (int kc = 0; kc < p.nkc; ++kc) {
            mbar_wait(&empty[s], ph ^ 1);
            uint8_t* st = smem + s * L::kStage;
            if (leader) {
              mbar_expect_tx(&full[s], bytes);
              tma_load_5d(st, &maps.a[t], &full[s], kc * KC, w0, h0, dd, n);
              for (int f = 0; f < p.nfold; ++f) {
                const int blk = (p.nquad > 1 ? qo + f : t) * p.nkc + kc;
                tma_load_2d(st + L::kABytes + f * p.N * KC * 2, &maps.b, &full[s], 0, blk * p.nfull + p.noff);
              }
            }
            if (++s == L::kStages) {
              s = 0;
              ph ^= 1;
            }
          }
        }
      }
    }
  } else if (warp == 5) {
    const bool leader = elect_one() != 0;
    {
      constexpr uint32_t kSwz = (KC == 64) ? kSwizzle128 : kSwizzle64;
      constexpr uint32_t kSbo = (KC == 64) ? 1024 : 512;
      const uint64_t desc_hi = make_smem_desc_hi(16, kSbo, kSwz);
      const uint32_t idesc = make_idesc_bf16(128, p.nfold * p.N, 0, 0);
      const uint64_t desc0 = smem_desc(desc_hi, smem_u32(smem));
      int s = 0;
      uint32_t ph = 0, accph = 0;
      int buf = 0;
      for (long long item = blockIdx.x; item < items; item += gridDim.x) {
        mbar_wait(&acc_empty[buf], ((accph >> buf) & 1u) ^ 1u);
        accph ^= 1u << buf;
        tc_fence_after();
        const uint32_t dcol = tmem_base + buf * 256;
        uint32_t first = 1;
        for (int it = 0; it < p.ntap * p.nkc; ++it) {
          mbar_wait(&full[s], ph);
          tc_fence_after();
          const uint64_t ad0 = desc0 + static_cast<uint64_t>(s * (L::kStage >> 4));
          const uint64_t bd0 = ad0 + (L::kABytes >> 4);
#pragma unroll
          for (int k = 0; k < KC / 16; ++k) {
            if (leader) umma_bf16(dcol, ad0 + k * 2, bd0 + k * 2, idesc, (first && k == 0) ? 0u : 1u);
          }
          first = 0;
          if (leader) umma_commit(&empty[s]);
          if (++s == L::kStages) {
            s = 0;
            ph ^= 1;
          }
        }
        if (leader) umma_commit(&acc_full[buf]);
        buf ^= 1;
      }
    }
  } else {
    uint32_t accph = 0;
    int buf = 0;
    const int m = warp * 32 + lane;
    for (long long item = blockIdx.x; item < items; item += gridDim.x) {
      int qo, w0, h0, dd, n;
      decode(item, qo, w0, h0, dd, n);
      const int ww = w0 + m % p.bw, hh = h0 + m / p.bw;
      const bool ok = m < rows && ww < p.w && hh < p.h;
      mbar_wait(&acc_full[buf], (accph >> buf) & 1u);
      accph ^= 1u << buf;
      tc_fence_after();
      for (int f = 0; f < p.nfold; ++f) {
        const int q = qo + f;
        const int oh = hh * p.scale + (p.nquad > 1 ? ((q >> 1) & 1) : 0);
        const int ow = ww * p.scale + (p.nquad > 1 ? (q & 1) : 0);
        const int od = dd * p.dscale + (p.nquad > 4 ? (q >> 2) : 0);
        __nv_bfloat16* dst =
            p.out + (((static_cast<long long>(n) * p.d * p.dscale + od) * p.ho + oh) * p.wo + ow) * p.ldo;
        for (int c0 = 0; c0 < p.N; c0 += 32) {
          uint32_t v[32];
          tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + buf * 256 + f * p.N + c0, v);
          tmem_ld_wait();
          if (f + 1 == p.nfold && c0 + 32 >= p.N) {  // last chunk read: the MMA warp may reuse this accumulator
            tc_fence_before();
            mbar_arrive(&acc_empty[buf]);
          }
          if (ok) {
            uint32_t pk[16];
#pragma unroll
            for (int c = 0; c < 32; c += 4) {      // bias: one broadcast 16-byte shared load per 4 channels
              const float4 bv = *reinterpret_cast<const float4*>(sbias + c0 + c);
              pk[c >> 1] = pack_bf16x2(__uint_as_float(v[c]) + bv.x, __uint_as_float(v[c + 1]) + bv.y);
              pk[(c >> 1) + 1] = pack_bf16x2(__uint_as_float(v[c + 2]) + bv.z, __uint_as_float(v[c + 3]) + bv.w);
            }
            if ((p.ldo & 15) == 0 && (reinterpret_cast<uintptr_t>(p.out) & 31) == 0) {   // 32-byte aligned rows: full-sector stores
              st_global_v8(dst + c0, pk[0], pk[1], pk[2], pk[3], pk[4], pk[5], pk[6], pk[7]);
              st_global_v8(dst + c0 + 16, pk[8], pk[9], pk[10], pk[11], pk[12], pk[13], pk[14], pk[15]);
            } else {
              uint4* d4 = reinterpret_cast<uint4*>(dst + c0);
#pragma unroll
              for (int qq = 0; qq < 4; ++qq) d4[qq] = make_uint4(pk[4 * qq], pk[4 * qq + 1], pk[4 * qq + 2], pk[4 * qq + 3]);
            }
          }
        }
      }
      buf ^= 1;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// position tile of the coarse grid: box (bw x bh), rows = bw*bh <= max_rows, multiple of 16 if k16
void pick_tile(int h, int w, int max_rows, bool k16, int* bw_out, int* bh_out) {
  auto gcd16 = [](int v) { int g = 16; while (v % g) g >>= 1; return g; };
  int bw = w < max_rows ? w : max_rows;
  int step = k16 ? 16 / gcd16(bw) : 1;
  if (bw * step > max_rows) {
    bw = (bw / 16) * 16;
    if (bw == 0) bw = w < 16 ? w : 16;
    step = k16 ? 16 / gcd16(bw) : 1;
  }
  int bh = max_rows / bw;
  if (bh > h) bh = h;
  bh = (bh / step) * step;
  if (bh == 0) bh = step;
  *bw_out = bw;
  *bh_out = bh;
}

// rank-5 map (c, w, h, d, n) of a position-major view whose (h, w) grid is sub-sampled by `scale`
// starting at (oh, ow) — the quadrant views of the fine grid — and whose planes are sub-sampled by
// `dscale` starting at `od` (the octant views of the (2,2,2) transposed conv; dfull = planes of the view).
int encode_grid_map(CUtensorMap* m, const void* base, long long ld, int c, int n, int dfull, int hfull, int wfull, int scale,
                    int oh, int ow, int box_c, int bw, int bh, int swizzle, int dscale = 1, int od = 0) {
  const __nv_bfloat16* b = static_cast<const __nv_bfloat16*>(base) +
                           ((static_cast<long long>(od) * hfull + oh) * wfull + ow) * ld;
  uint64_t dims[5] = {static_cast<uint64_t>(c), static_cast<uint64_t>(wfull / scale), static_cast<uint64_t>(hfull / scale),
                      static_cast<uint64_t>(dfull / dscale), static_cast<uint64_t>(n)};
  uint64_t str[4] = {static_cast<uint64_t>(ld) * 2 * scale, static_cast<uint64_t>(ld) * 2 * wfull * scale,
                     static_cast<uint64_t>(ld) * 2 * wfull * hfull * dscale,
                     static_cast<uint64_t>(ld) * 2 * wfull * hfull * dfull};
  uint32_t box[5] = {static_cast<uint32_t>(box_c), static_cast<uint32_t>(bw), static_cast<uint32_t>(bh), 1, 1};
  return encode_tmap_bf16(m, b, 5, dims, str, box, swizzle);
}

template <int KC>
int launch_pw(const PwMaps& maps, const PwParams& p, cudaStream_t st) {
  using L = PwSmem<KC>;
  static bool attr_set_dev[kMaxDevices] = {};   // the opt-in is per device (and per template instance)
  bool& attr_set = attr_set_dev[current_device()];
  if (!attr_set) {
    SPFF_CUDA(cudaFuncSetAttribute(pw_gemm_kernel<KC>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal + 1024));
    attr_set = true;
  }
  long long items = p.ntiles * (p.nquad / p.nfold);
  int ctas = debug_ctas() > 0 ? debug_ctas() : num_sms();
  if (items < ctas) ctas = static_cast<int>(items);
  pw_gemm_kernel<KC><<<ctas, kThreads, L::kTotal + 1024, st>>>(maps, p);
  SPFF_CUDA(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------------------------
// weight packing: W[cin][cout][1][2][2] fp32 ->
//   fwd   B[q][kc][co][k] = W[kc*KC+k][co][q]      (N = cout, K = cin,  KC  = 64 / 32)
//   dgrad B[q][kc][ci][k] = W[ci][kc*KC'+k][q]     (N = cin,  K = cout, KC' = 64 / 32)
// ---------------------------------------------------------------------------------------------
__global__ void pack_convt_weight_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out, int cin, int cout,
                                         int KC, int dgrad, int nq) {
  const int N = dgrad ? cin : cout, K = dgrad ? cout : cin;
  const int nkc = K / KC;
  const long long total = static_cast<long long>(nq) * cin * cout;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    long long r = i;
    const int k = static_cast<int>(r % KC); r /= KC;
    const int nn = static_cast<int>(r % N); r /= N;
    const int kc = static_cast<int>(r % nkc); r /= nkc;
    const int q = static_cast<int>(r);
    const int kk = kc * KC + k;
    const int ci = dgrad ? nn : kk, co = dgrad ? kk : nn;
    out[i] = __float2bfloat16(w[(static_cast<long long>(ci) * cout + co) * nq + q]);
  }
}

// ---------------------------------------------------------------------------------------------
// wgrad: D[(quadrant, co)][ci] += sum_pos dy_q[pos][co] * x[pos][ci]
// ---------------------------------------------------------------------------------------------
struct PwWgMaps {
  CUtensorMap dy[4];
  CUtensorMap x;
};
struct PwWgParams {
  int n, d, h, w, bw, bh, tiles_w, tiles_h;
  long long ntiles;
  int ncob, ncib, ksplit;
  float* partial;  // [item][split][mma][128][CIB]
};

template <int COB, int CIB>
struct PwWgCfg {
  static constexpr int KT = 8192 / (COB * 2);     // rows per tile: 128 / 64
  static constexpr int SP = 128 / COB;            // quadrants per MMA
  static constexpr int NMMA = 4 / SP;
  static constexpr int NSUB = CIB / 64;           // 64-channel sub tiles of x
  static constexpr int XSub = KT * 128;
  static constexpr int Stage = 4 * 8192 + NSUB * XSub;
  static constexpr int Stages = 3;
  static constexpr int OffBar = Stages * Stage;
  static constexpr int OffTmem = OffBar + (2 * Stages + 1) * 8;
  static constexpr int Total = OffTmem + 16;
  static constexpr int TmemCols = (NMMA * CIB <= 128) ? 128 : 256;
};

template <int COB, int CIB>
__global__ void __launch_bounds__(kThreads, 1) pw_wgrad_kernel(const __grid_constant__ PwWgMaps maps, const PwWgParams p) {
  using C = PwWgCfg<COB, CIB>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::OffBar);
  uint64_t* full = bars;
  uint64_t* empty = bars + C::Stages;
  uint64_t* acc_full = bars + 2 * C::Stages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + C::OffTmem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < C::Stages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    mbar_init(acc_full, 1);
    fence_mbar_init();
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
  const int cib = item % p.ncib, cob = item / p.ncib;
  const long long t_begin = p.ntiles * split / p.ksplit;
  const long long t_end = p.ntiles * (split + 1) / p.ksplit;
  const int rows = p.bw * p.bh;

  if (warp == 4) {
    const bool leader = elect_one() != 0;
    {
      int s = 0;
      uint32_t ph = 0;
      const uint32_t bytes = 4 * rows * COB * 2 + rows * CIB * 2;
      for (long long t = t_begin; t < t_end; ++t) {
        long long r = t;
        const int w0 = static_cast<int>(r % p.tiles_w) * p.bw; r /= p.tiles_w;
        const int h0 = static_cast<int>(r % p.tiles_h) * p.bh; r /= p.tiles_h;
        const int dd = static_cast<int>(r % p.d);
        const int n = static_cast<int>(r / p.d);
        mbar_wait(&empty[s], ph ^ 1);
        uint8_t* st = smem + s * C::Stage;
        if (leader) {
          mbar_expect_tx(&full[s], bytes);
#pragma unroll
          for (int q = 0; q < 4; ++q) tma_load_5d(st + q * 8192, &maps.dy[q], &full[s], cob * COB, w0, h0, dd, n);
#pragma unroll
          for (int u = 0; u < C::NSUB; ++u)
            tma_load_5d(st + 4 * 8192 + u * C::XSub, &maps.x, &full[s], cib * CIB + u * 64, w0, h0, dd, n);
        }
        if (++s == C::Stages) {
          s = 0;
          ph ^= 1;
        }
      }
    }
  } else if (warp == 5) {
    const bool leader = elect_one() != 0;
    {
      constexpr uint32_t kSwzA = (COB == 64) ? kSwizzle128 : kSwizzle64;
      constexpr uint32_t kSboA = (COB == 64) ? 1024 : 512;
      const uint64_t adesc_hi = make_smem_desc_hi(8192, kSboA, kSwzA);
      const uint64_t bdesc_hi = make_smem_desc_hi(C::XSub, 1024, kSwizzle128);
      const uint32_t idesc = make_idesc_bf16(128, CIB, 1, 1);
      const int ksteps = rows / 16;
      int s = 0;
      uint32_t ph = 0, first = 1;
      for (long long t = t_begin; t < t_end; ++t) {
        mbar_wait(&full[s], ph);
        tc_fence_after();
        const uint32_t base = smem_u32(smem + s * C::Stage);
#pragma unroll
        for (int i = 0; i < C::NMMA; ++i) {
          for (int ks = 0; ks < ksteps; ++ks) {
            if (leader)
              umma_bf16(tmem_base + i * CIB, smem_desc(adesc_hi, base + i * C::SP * 8192 + ks * 2 * kSboA),
                        smem_desc(bdesc_hi, base + 4 * 8192 + ks * 2048), idesc, (first && ks == 0) ? 0u : 1u);
          }
        }
        first = 0;
        if (leader) umma_commit(&empty[s]);
        if (++s == C::Stages) {
          s = 0;
          ph ^= 1;
        }
      }
      if (leader) umma_commit(acc_full);
    }
  } else {
    mbar_wait(acc_full, 0);
    tc_fence_after();
    float* out = p.partial + static_cast<size_t>(blockIdx.x) * C::NMMA * 128 * CIB;
    const int row = warp * 32 + lane;
#pragma unroll 1
    for (int i = 0; i < C::NMMA; ++i) {
#pragma unroll 1
      for (int c0 = 0; c0 < CIB; c0 += 32) {
        uint32_t v[32];
        tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + i * CIB + c0, v);
        tmem_ld_wait();
        float4* dst = reinterpret_cast<float4*>(out + (static_cast<size_t>(i) * 128 + row) * CIB + c0);
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

// dW[ci][co][q] = beta*dW + sum_split partial
// (nq, qoff): dW holds nq taps per (ci, co) and this launch produced taps [qoff, qoff + 4) — the (2,2,2) kernel
// runs the four-quadrant contraction once per depth tap.
__global__ void pw_wgrad_reduce_kernel(const float* __restrict__ partial, float* __restrict__ dw, int cin, int cout,
                                       int COB, int CIB, int ksplit, float beta, int nq, int qoff) {
  const int SP = 128 / COB, NMMA = 4 / SP;
  const int ncib = cin / CIB;
  const long long total = 4LL * cin * cout;
  for (long long e = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; e < total;
       e += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int q = static_cast<int>(e % 4);
    const int co = static_cast<int>((e / 4) % cout);
    const int ci = static_cast<int>(e / (4LL * cout));
    const int cob = co / COB, col = co % COB, cib = ci / CIB, cic = ci % CIB;
    const int item = cob * ncib + cib;
    const int i = q / SP, b = q % SP;
    const int lane = b * COB + col;
    float acc = 0.f;
    for (int s = 0; s < ksplit; ++s)
      acc += partial[((static_cast<size_t>(item) * ksplit + s) * NMMA + i) * 128 * CIB + static_cast<size_t>(lane) * CIB + cic];
    float* dst = dw + (static_cast<long long>(ci) * cout + co) * nq + qoff + q;
    *dst = (beta == 0.f) ? acc : fmaf(beta, *dst, acc);
  }
}

struct PwWgPlan {
  int cob, cib, bw, bh, tiles_w, tiles_h, ncob, ncib, ksplit, nmma;
  long long ntiles;
  size_t ws_bytes;
};

PwWgPlan make_pw_plan(int cin, int cout, spff_shape s) {
  PwWgPlan pl;
  pl.cob = (cout % 64 == 0) ? 64 : 32;
  pl.cib = (cin % 128 == 0) ? 128 : 64;
  pick_tile(s.h, s.w, 8192 / (pl.cob * 2), true, &pl.bw, &pl.bh);
  pl.tiles_w = (s.w + pl.bw - 1) / pl.bw;
  pl.tiles_h = (s.h + pl.bh - 1) / pl.bh;
  pl.ntiles = static_cast<long long>(s.n) * s.d * pl.tiles_w * pl.tiles_h;
  pl.ncob = cout / pl.cob;
  pl.ncib = cin / pl.cib;
  const int items = pl.ncob * pl.ncib;
  int ks = (debug_ctas() > 0 ? debug_ctas() : num_sms()) / items;
  if (ks < 1) ks = 1;
  if (ks > pl.ntiles) ks = static_cast<int>(pl.ntiles);
  pl.ksplit = ks;
  pl.nmma = 4 / (128 / pl.cob);
  pl.ws_bytes = static_cast<size_t>(items) * ks * pl.nmma * 128 * pl.cib * sizeof(float);
  return pl;
}

template <int COB, int CIB>
int launch_pw_wgrad(const PwWgMaps& maps, const PwWgParams& p, int ctas, cudaStream_t st) {
  using C = PwWgCfg<COB, CIB>;
  static bool attr_set_dev[kMaxDevices] = {};   // the opt-in is per device (and per template instance)
  bool& attr_set = attr_set_dev[current_device()];
  if (!attr_set) {
    SPFF_CUDA(cudaFuncSetAttribute(pw_wgrad_kernel<COB, CIB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   C::Total + 1024));
    attr_set = true;
  }
  pw_wgrad_kernel<COB, CIB><<<ctas, kThreads, C::Total + 1024, st>>>(maps, p);
  SPFF_CUDA(cudaGetLastError());
  return 0;
}

// ---- host-side launch plans shared by the (1,2,2) and (2,2,2) entry points --------------------------------
// nd = 1: kernel = stride = (1,2,2), 4 taps; nd = 2: kernel = stride = (2,2,2), 8 taps (tap = (i*2 + j)*2 + k).
int pack_convt_impl(const float* w, void* w_fwd, void* w_dgrad, int cin, int cout, int nq, cudaStream_t st) {
  const long long total = static_cast<long long>(nq) * cin * cout;
  const int blocks = static_cast<int>((total + 255) / 256 < 2048 ? (total + 255) / 256 : 2048);
  if (w_fwd)
    pack_convt_weight_kernel<<<blocks, 256, 0, st>>>(w, static_cast<__nv_bfloat16*>(w_fwd), cin, cout, conv3_kc(cin), 0, nq);
  if (w_dgrad)
    pack_convt_weight_kernel<<<blocks, 256, 0, st>>>(w, static_cast<__nv_bfloat16*>(w_dgrad), cin, cout, conv3_kc(cout), 1,
                                                     nq);
  SPFF_CUDA(cudaGetLastError());
  return 0;
}

void fill_tiles(PwParams* p, spff_shape s) {
  p->n = s.n; p->d = s.d; p->h = s.h; p->w = s.w;
  pick_tile(s.h, s.w, 128, false, &p->bw, &p->bh);
  p->tiles_w = (s.w + p->bw - 1) / p->bw;
  p->tiles_h = (s.h + p->bh - 1) / p->bh;
  p->ntiles = static_cast<long long>(s.n) * s.d * p->tiles_w * p->tiles_h;
}

int encode_b_map(CUtensorMap* m, const void* w, int KC, long long rows, int box_rows) {
  uint64_t dims[2] = {static_cast<uint64_t>(KC), static_cast<uint64_t>(rows)};
  uint64_t str[1] = {static_cast<uint64_t>(KC) * 2};
  uint32_t box[2] = {static_cast<uint32_t>(KC), static_cast<uint32_t>(box_rows)};
  return encode_tmap_bf16(m, w, 2, dims, str, box, KC * 2);
}

// an output wider than 256 channels (one UMMA N) runs as several launches over 256-wide column blocks
int pick_nchunk(int n) { return n <= 256 ? n : (n % 256 == 0 ? 256 : (n % 128 == 0 ? 128 : (n % 64 == 0 ? 64 : 32))); }

int convt_fwd_impl(const void* x, long long ldx, int cin, const void* w_fwd, const float* bias, void* y, long long ldy,
                   int cout, spff_shape s, int nd, cudaStream_t st) {
  const int KC = conv3_kc(cin), nq = 4 * nd;
  PwParams p{};
  fill_tiles(&p, s);
  p.nkc = cin / KC; p.ntap = 1; p.nquad = nq; p.nfull = cout;
  p.ldo = ldy; p.ho = 2 * s.h; p.wo = 2 * s.w; p.scale = 2; p.dscale = nd; p.bias = bias;
  PwMaps maps;
  int e = encode_grid_map(&maps.a[0], x, ldx, cin, s.n, s.d, s.h, s.w, 1, 0, 0, KC, p.bw, p.bh, KC * 2);
  if (e) return e;
  const int nc = pick_nchunk(cout);
  e = encode_b_map(&maps.b, w_fwd, KC, static_cast<long long>(nq) * p.nkc * cout, nc);
  if (e) return e;
  p.nfold = 1;
  while (p.nfold * 2 <= nq && p.nfold * 2 * nc <= 256 && !debug_flag(5)) p.nfold *= 2;   // test hook: key 5 disables folding
  for (int noff = 0; noff < cout; noff += nc) {
    p.N = nc; p.noff = noff;
    p.out = static_cast<__nv_bfloat16*>(y) + noff;
    e = KC == 64 ? launch_pw<64>(maps, p, st) : launch_pw<32>(maps, p, st);
    if (e) return e;
  }
  return 0;
}

int convt_dgrad_impl(const void* dy, long long lddy, int cout, const void* w_dgrad, void* dx, long long lddx, int cin,
                     spff_shape s, int nd, cudaStream_t st) {
  const int KC = conv3_kc(cout), nq = 4 * nd;
  PwParams p{};
  fill_tiles(&p, s);
  p.nkc = cout / KC; p.ntap = nq; p.nquad = 1; p.nfull = cin; p.nfold = 1;
  p.ldo = lddx; p.ho = s.h; p.wo = s.w; p.scale = 1; p.dscale = 1; p.bias = nullptr;
  PwMaps maps;
  int e;
  for (int q = 0; q < nq; ++q) {
    e = encode_grid_map(&maps.a[q], dy, lddy, cout, s.n, nd * s.d, 2 * s.h, 2 * s.w, 2, (q >> 1) & 1, q & 1, KC, p.bw, p.bh,
                        KC * 2, nd, q >> 2);
    if (e) return e;
  }
  const int nc = pick_nchunk(cin);
  e = encode_b_map(&maps.b, w_dgrad, KC, static_cast<long long>(nq) * p.nkc * cin, nc);
  if (e) return e;
  for (int noff = 0; noff < cin; noff += nc) {
    p.N = nc; p.noff = noff;
    p.out = static_cast<__nv_bfloat16*>(dx) + noff;
    e = KC == 64 ? launch_pw<64>(maps, p, st) : launch_pw<32>(maps, p, st);
    if (e) return e;
  }
  return 0;
}

int convt_wgrad_impl(const void* x, long long ldx, int cin, const void* dy, long long lddy, int cout, spff_shape s,
                     float* dw, float beta, void* workspace, size_t workspace_bytes, int nd, cudaStream_t st) {
  const PwWgPlan pl = make_pw_plan(cin, cout, s);
  if (workspace_bytes < pl.ws_bytes) {
    set_error("convt wgrad: workspace %zu < %zu bytes", workspace_bytes, pl.ws_bytes);
    return SPFF_ERR_WORKSPACE;
  }
  SPFF_REQUIRE((pl.bw * pl.bh) % 16 == 0 && pl.bw * pl.bh * pl.cob * 2 <= 8192, "convt wgrad: cannot tile %dx%d", s.h, s.w);
  PwWgParams p{};
  p.n = s.n; p.d = s.d; p.h = s.h; p.w = s.w; p.bw = pl.bw; p.bh = pl.bh; p.tiles_w = pl.tiles_w; p.tiles_h = pl.tiles_h;
  p.ntiles = pl.ntiles; p.ncob = pl.ncob; p.ncib = pl.ncib; p.ksplit = pl.ksplit;
  p.partial = static_cast<float*>(workspace);
  PwWgMaps maps;
  int e = encode_grid_map(&maps.x, x, ldx, cin, s.n, s.d, s.h, s.w, 1, 0, 0, 64, pl.bw, pl.bh, 128);
  if (e) return e;
  const int ctas = pl.ncob * pl.ncib * pl.ksplit;
  const long long total = 4LL * cin * cout;
  for (int i = 0; i < nd; ++i) {   // depth tap
    for (int q = 0; q < 4; ++q) {
      e = encode_grid_map(&maps.dy[q], dy, lddy, cout, s.n, nd * s.d, 2 * s.h, 2 * s.w, 2, q >> 1, q & 1, pl.cob, pl.bw,
                          pl.bh, pl.cob * 2, nd, i);
      if (e) return e;
    }
    if (pl.cob == 64 && pl.cib == 128) e = launch_pw_wgrad<64, 128>(maps, p, ctas, st);
    else if (pl.cob == 64) e = launch_pw_wgrad<64, 64>(maps, p, ctas, st);
    else if (pl.cib == 128) e = launch_pw_wgrad<32, 128>(maps, p, ctas, st);
    else e = launch_pw_wgrad<32, 64>(maps, p, ctas, st);
    if (e) return e;
    pw_wgrad_reduce_kernel<<<static_cast<int>((total + 255) / 256), 256, 0, st>>>(p.partial, dw, cin, cout, pl.cob, pl.cib,
                                                                                  pl.ksplit, beta, 4 * nd, 4 * i);
  }
  SPFF_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace
}  // namespace spff

extern "C" {

int spff_pack_convt_weight(const float* w, void* w_fwd, void* w_dgrad, int cin, int cout, void* stream) {
  SPFF_REQUIRE(w && cin % 32 == 0 && cout % 32 == 0, "pack_convt_weight: channels must be multiples of 32");
  return spff::pack_convt_impl(w, w_fwd, w_dgrad, cin, cout, 4, static_cast<cudaStream_t>(stream));
}

int spff_pack_convt_weight_k222(const float* w, void* w_fwd, void* w_dgrad, int cin, int cout, void* stream) {
  SPFF_REQUIRE(w && cin % 32 == 0 && cout % 32 == 0, "pack_convt_weight_k222: channels must be multiples of 32");
  return spff::pack_convt_impl(w, w_fwd, w_dgrad, cin, cout, 8, static_cast<cudaStream_t>(stream));
}

/* `s` is the INPUT (coarse) grid; y is [n, d, 2h, 2w, cout]. */
int spff_convt_k122_fwd(const void* x, long long ldx, int cin, const void* w_fwd, const float* bias, void* y,
                        long long ldy, int cout, spff_shape s, void* stream) {
  int e = spff_device_check();
  if (e) return e;
  SPFF_REQUIRE(x && w_fwd && y, "convt_k122_fwd: null pointer");
  SPFF_REQUIRE(cin % 32 == 0 && cout % 32 == 0 && cin > 0 && cout > 0, "convt_k122_fwd: bad channels %d -> %d", cin, cout);
  return spff::convt_fwd_impl(x, ldx, cin, w_fwd, bias, y, ldy, cout, s, 1, static_cast<cudaStream_t>(stream));
}

/* `s` is the coarse grid (that of dx); dy is [n, d, 2h, 2w, cout]. */
int spff_convt_k122_dgrad(const void* dy, long long lddy, int cout, const void* w_dgrad, void* dx, long long lddx,
                          int cin, spff_shape s, void* stream) {
  int e = spff_device_check();
  if (e) return e;
  SPFF_REQUIRE(dy && w_dgrad && dx, "convt_k122_dgrad: null pointer");
  SPFF_REQUIRE(cin % 32 == 0 && cout % 32 == 0 && cin > 0 && cout > 0, "convt_k122_dgrad: bad channels %d -> %d", cin, cout);
  return spff::convt_dgrad_impl(dy, lddy, cout, w_dgrad, dx, lddx, cin, s, 1, static_cast<cudaStream_t>(stream));
}

size_t spff_convt_k122_wgrad_workspace(int cin, int cout, spff_shape s) {
  if (cin % 64 || cout % 32 || s.n <= 0 || s.d <= 0 || s.h <= 0 || s.w <= 0) return 0;
  return spff::make_pw_plan(cin, cout, s).ws_bytes;
}

/* dw[cin][cout][1][2][2] (fp32) = beta*dw + gradient. The bias gradient is the column sum of dy
 * (spff_in_stats gives it). `s` is the coarse grid. */
int spff_convt_k122_wgrad(const void* x, long long ldx, int cin, const void* dy, long long lddy, int cout,
                          spff_shape s, float* dw, float beta, void* workspace, size_t workspace_bytes,
                          void* stream) {
  int e = spff_device_check();
  if (e) return e;
  SPFF_REQUIRE(x && dy && dw && workspace, "convt_k122_wgrad: null pointer");
  SPFF_REQUIRE(cin % 64 == 0 && cout % 32 == 0, "convt_k122_wgrad: needs cin %% 64 == 0 and cout %% 32 == 0 (%d -> %d)", cin,
               cout);
  return spff::convt_wgrad_impl(x, ldx, cin, dy, lddy, cout, s, dw, beta, workspace, workspace_bytes, 1,
                                static_cast<cudaStream_t>(stream));
}

/* ---- ConvTranspose3d kernel = stride = (2,2,2) (Cicek3DUNet up4..up1, models.py:733-739) -------------------
 * `s` is the coarse grid; the fine tensors are [n, 2d, 2h, 2w, c]. Weight [cin][cout][2][2][2]. */
int spff_convt_k222_fwd(const void* x, long long ldx, int cin, const void* w_fwd, const float* bias, void* y,
                        long long ldy, int cout, spff_shape s, void* stream) {
  int e = spff_device_check();
  if (e) return e;
  SPFF_REQUIRE(x && w_fwd && y, "convt_k222_fwd: null pointer");
  SPFF_REQUIRE(cin % 32 == 0 && cout % 32 == 0 && cin > 0 && cout > 0, "convt_k222_fwd: bad channels %d -> %d", cin, cout);
  return spff::convt_fwd_impl(x, ldx, cin, w_fwd, bias, y, ldy, cout, s, 2, static_cast<cudaStream_t>(stream));
}

int spff_convt_k222_dgrad(const void* dy, long long lddy, int cout, const void* w_dgrad, void* dx, long long lddx,
                          int cin, spff_shape s, void* stream) {
  int e = spff_device_check();
  if (e) return e;
  SPFF_REQUIRE(dy && w_dgrad && dx, "convt_k222_dgrad: null pointer");
  SPFF_REQUIRE(cin % 32 == 0 && cout % 32 == 0 && cin > 0 && cout > 0, "convt_k222_dgrad: bad channels %d -> %d", cin, cout);
  return spff::convt_dgrad_impl(dy, lddy, cout, w_dgrad, dx, lddx, cin, s, 2, static_cast<cudaStream_t>(stream));
}

size_t spff_convt_k222_wgrad_workspace(int cin, int cout, spff_shape s) {
  return spff_convt_k122_wgrad_workspace(cin, cout, s);
}

/* dw[cin][cout][2][2][2] (fp32) = beta*dw + gradient: the four-quadrant contraction once per depth tap. */
int spff_convt_k222_wgrad(const void* x, long long ldx, int cin, const void* dy, long long lddy, int cout,
                          spff_shape s, float* dw, float beta, void* workspace, size_t workspace_bytes,
                          void* stream) {
  int e = spff_device_check();
  if (e) return e;
  SPFF_REQUIRE(x && dy && dw && workspace, "convt_k222_wgrad: null pointer");
  SPFF_REQUIRE(cin % 64 == 0 && cout % 32 == 0, "convt_k222_wgrad: needs cin %% 64 == 0 and cout %% 32 == 0 (%d -> %d)", cin,
               cout);
  return spff::convt_wgrad_impl(x, ldx, cin, dy, lddy, cout, s, dw, beta, workspace, workspace_bytes, 2,
                                static_cast<cudaStream_t>(stream));
}

}  // extern "C"
