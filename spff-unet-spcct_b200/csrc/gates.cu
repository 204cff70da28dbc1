// The per-sample "micro" part of the SPFF tail. After the second conv of a block
//   a = lrelu(IN(x2)),  out = fgate(efilm(a)) [-> SpectralSE -> ChannelSE on encoder stages]
// (reference innovative3D/models.py:1473-1478, 684-685) every gate is a function of the plane sums
// S[n][d][c] = sum_hw a only, so  out = a * P[n][d][c] + Q[n][d][c]  (SURVEY.md §7.3):
//   EnergyFiLM3D  (models.py:1505-1512)  e = a*g1[c][d] + bt[c][d],  g1 = 1 + tanh(gamma), bt = beta
//   FourierGate3D (models.py:1527-1544)  f = e * w1[d],  w1 = sigmoid(circular_conv(mean_{c,hw} e, kfg))
//                                         kfg = irfft(freq_mask*mag_scale) (real symmetric kernel)
//   _SpectralSE   (models.py:611-614)    g = f * w2[d],  w2 = sigmoid(mean_{c,hw} f)
//   _SEChannelLite(models.py:600-609)    o = g * w3[c],  w3 = sigmoid(W2 relu(W1 mean_{d,hw} g + b1) + b2)
// The tables g1/bt/kfg depend on parameters only; the host computes them (and back-propagates
// through them) with a handful of tiny tensor ops, these kernels take them as inputs and return
// their gradients. One CTA per sample; the backward kernel also folds the InstanceNorm backward
// coefficients, since they need the same per-plane sums.
#include "common.h"

namespace spff {
namespace {

constexpr int kT = 256;
constexpr int kMaxD = 16;
constexpr int kMaxHid = 64;

struct GateArgs {
  const float* S;   // [n][d][c]
  const float* g1;  // [c][d] or null
  const float* bt;  // [c][d] or null
  const float* kfg; // [d] or null
  const float* w1; const float* b1; const float* w2; const float* b2;  // SE fc: [hid][c],[hid],[c][hid],[c]
  int hid, flags, c, d;
  float hw;
};

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + __expf(-x)); }

// Sum each of part[0..d) over the block; afterwards every thread holds the totals.
template <int MD>
__device__ __forceinline__ void block_sum_vec(float (&part)[MD], int d, float* red /* [kT/32][kMaxD] */) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < MD; ++i) {
    if (i < d) {
      float v = part[i];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0) red[warp * kMaxD + i] = v;
    }
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < MD; ++i) {
    if (i < d) {
      float v = 0.f;
#pragma unroll
      for (int w = 0; w < kT / 32; ++w) v += red[w * kMaxD + i];
      part[i] = v;
    }
  }
  __syncthreads();
}

struct GateSmem {
  float* Se;   // [d][c]
  float* t;    // [c]
  float* w3;   // [c]
  float* dv;   // [c]
  float* s1;   // [kMaxD] each below
  float* w1;
  float* w2;
  float* vec;  // 4*kMaxD scratch
  float* hdn;  // [kMaxHid]
  float* dpre; // [kMaxHid]
  float* red;  // [kT/32][kMaxD]
};

__device__ __forceinline__ GateSmem carve(float* sm, int c, int d) {
  GateSmem g;
  g.Se = sm; sm += d * c;
  g.t = sm; sm += c;
  g.w3 = sm; sm += c;
  g.dv = sm; sm += c;
  g.s1 = sm; sm += kMaxD;
  g.w1 = sm; sm += kMaxD;
  g.w2 = sm; sm += kMaxD;
  g.vec = sm; sm += 4 * kMaxD;
  g.hdn = sm; sm += kMaxHid;
  g.dpre = sm; sm += kMaxHid;
  g.red = sm;
  return g;
}
__host__ __device__ inline size_t gate_smem_bytes(int c, int d) {
  return sizeof(float) * (static_cast<size_t>(d) * c + 3 * c + 7 * kMaxD + 2 * kMaxHid + (kT / 32) * kMaxD);
}

// Forward gates of sample n: fills sm.Se, s1, w1, w2, t, hdn, w3.
// MD = compile-time bound on the number of energy bins (8 or 16): every per-bin vector lives in
// registers with static indexing, so the unrolled code size grows with MD (quadratically in the
// circular convolutions) - with MD = 16 the backward kernel was 19k instructions and spent its time
// in instruction fetch.
// DX > 0: the number of bins is the compile-time constant DX (= MD): guards fold away and the
// circular index arithmetic `% d` becomes a constant modulo. The SPCCT data has 5 bins
// (config.py:22), which gets its own instantiation.
template <int MD, int DX>
__device__ void gate_forward(const GateArgs& a, int n, GateSmem& sm) {
  const int tid = threadIdx.x, c = a.c, d = DX > 0 ? DX : a.d;
  const float* S = a.S + static_cast<size_t>(n) * d * c;
  float part[MD];
#pragma unroll
  for (int i = 0; i < MD; ++i) part[i] = 0.f;
#pragma unroll
  for (int dd = 0; dd < MD; ++dd) {
    if (dd < d) {
      #pragma unroll 1
      for (int ch = tid; ch < c; ch += kT) {
        float se = a.flags ? S[dd * c + ch] : 0.f;
        if (a.flags & SPFF_GATE_EFILM) se = fmaf(se, a.g1[ch * d + dd], a.bt[ch * d + dd] * a.hw);
        sm.Se[dd * c + ch] = se;
        part[dd] += se;
      }
    }
  }
  block_sum_vec(part, d, sm.red);
  const float inv_chw = 1.f / (static_cast<float>(c) * a.hw);
  if (tid == 0) {
#pragma unroll
    for (int i = 0; i < MD; ++i)
      if (i < d) sm.s1[i] = part[i] * inv_chw;
  }
  __syncthreads();
  if (tid < d) {
    float w1 = 1.f;
    if (a.flags & SPFF_GATE_FOURIER) {
      float u = 0.f;
      for (int e = 0; e < d; ++e) u = fmaf(a.kfg[(tid - e + d) % d], sm.s1[e], u);
      w1 = sigmoidf_(u);
    }
    float w2 = 1.f;
    if (a.flags & SPFF_GATE_SPECSE) w2 = sigmoidf_(sm.s1[tid] * w1);  // mean_{c,hw} f = s1 * w1
    sm.w1[tid] = w1;
    sm.w2[tid] = w2;
  }
  __syncthreads();
  if (a.flags & SPFF_GATE_CHANSE) {
    const float inv_dhw = 1.f / (static_cast<float>(d) * a.hw);
    #pragma unroll 1
    for (int ch = tid; ch < c; ch += kT) {
      float s = 0.f;
      for (int dd = 0; dd < d; ++dd) s = fmaf(sm.Se[dd * c + ch], sm.w1[dd] * sm.w2[dd], s);
      sm.t[ch] = s * inv_dhw;
    }
    __syncthreads();
    const int lane = tid & 31, warp = tid >> 5;
    #pragma unroll 1
    for (int j = warp; j < a.hid; j += kT / 32) {
      float s = 0.f;
      #pragma unroll 1
      for (int ch = lane; ch < c; ch += 32) s = fmaf(a.w1[j * c + ch], sm.t[ch], s);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (lane == 0) sm.hdn[j] = fmaxf(s + a.b1[j], 0.f);
    }
    __syncthreads();
    #pragma unroll 1
    for (int ch = tid; ch < c; ch += kT) {
      float v = a.b2[ch];
      #pragma unroll 1
      for (int j = 0; j < a.hid; ++j) v = fmaf(a.w2[ch * a.hid + j], sm.hdn[j], v);
      sm.w3[ch] = sigmoidf_(v);
    }
  } else {
    #pragma unroll 1
    for (int ch = tid; ch < c; ch += kT) sm.w3[ch] = 1.f;
  }
  __syncthreads();
}

template <int MD, int DX>
__global__ void __launch_bounds__(kT) gate_fwd_kernel(GateArgs a, float* __restrict__ P, float* __restrict__ Q) {
  extern __shared__ float smem_f[];
  GateSmem sm = carve(smem_f, a.c, a.d);
  const int n = blockIdx.x;
  gate_forward<MD, DX>(a, n, sm);
  const int c = a.c, d = DX > 0 ? DX : a.d;
  #pragma unroll 1
  for (int idx = threadIdx.x; idx < d * c; idx += kT) {
    const int dd = idx / c, ch = idx % c;
    const float G = sm.w1[dd] * sm.w2[dd] * sm.w3[ch];
    const float g1 = (a.flags & SPFF_GATE_EFILM) ? a.g1[ch * d + dd] : 1.f;
    const float bt = (a.flags & SPFF_GATE_EFILM) ? a.bt[ch * d + dd] : 0.f;
    P[static_cast<size_t>(n) * d * c + idx] = g1 * G;
    Q[static_cast<size_t>(n) * d * c + idx] = bt * G;
  }
}

struct GateBwdOut {
  float* bcoef;  // [n][c][4]
  float* dSa;    // [n][d][c]
  float* Pout;   // [n][d][c]
  float* dgamma; float* dbeta;            // [c]   (+=)
  float* dg1; float* dbt;                 // [c][d] (+=)
  float* dkfg;                            // [d]   (+=)
  float* dw1; float* db1; float* dw2; float* db2;  // SE fc (+=)
};

template <int MD, int DX>
__global__ void __launch_bounds__(kT)
gate_bwd_kernel(GateArgs a, const float* __restrict__ R, const float* __restrict__ coef,
                const float* __restrict__ gamma, GateBwdOut o) {
  extern __shared__ float smem_f[];
  GateSmem sm = carve(smem_f, a.c, a.d);
  const int n = blockIdx.x, tid = threadIdx.x, c = a.c, d = DX > 0 ? DX : a.d;
  gate_forward<MD, DX>(a, n, sm);
  const float* S = a.S + static_cast<size_t>(n) * d * c;
  const float* Rn = R + static_cast<size_t>(n) * d * c * 6;
  const bool efilm = a.flags & SPFF_GATE_EFILM;
  const float inv_chw = 1.f / (static_cast<float>(c) * a.hw);
  const float inv_dhw = 1.f / (static_cast<float>(d) * a.hw);

  // (a) dG = dP*g1 + dQ*bt ; direct table grads ; dw3[c] ; dw12[d]
  float dw12[MD];
#pragma unroll
  for (int i = 0; i < MD; ++i) dw12[i] = 0.f;
  #pragma unroll 1
  for (int ch = tid; ch < c; ch += kT) {
    float dw3 = 0.f;
#pragma unroll
    for (int dd = 0; dd < MD; ++dd) {
      if (dd < d) {
        const float dP = Rn[(dd * c + ch) * 6 + 0], dQ = Rn[(dd * c + ch) * 6 + 1];
        const float g1 = efilm ? a.g1[ch * d + dd] : 1.f;
        const float bt = efilm ? a.bt[ch * d + dd] : 0.f;
        const float w12 = sm.w1[dd] * sm.w2[dd];
        const float dG = dP * g1 + dQ * bt;
        if (efilm) {
          atomicAdd(o.dg1 + ch * d + dd, dP * w12 * sm.w3[ch]);
          atomicAdd(o.dbt + ch * d + dd, dQ * w12 * sm.w3[ch]);
        }
        dw3 = fmaf(dG, w12, dw3);
        dw12[dd] = fmaf(dG, sm.w3[ch], dw12[dd]);
      }
    }
    sm.dv[ch] = dw3 * sm.w3[ch] * (1.f - sm.w3[ch]);  // d(pre-sigmoid) of the channel gate
  }
  block_sum_vec(dw12, d, sm.red);

  // (b) channel SE backward -> dt[c] (kept in sm.dv after use of dv)
  if (a.flags & SPFF_GATE_CHANSE) {
    const int lane = tid & 31, warp = tid >> 5;
    #pragma unroll 1
    for (int j = warp; j < a.hid; j += kT / 32) {
      float s = 0.f;
      #pragma unroll 1
      for (int ch = lane; ch < c; ch += 32) s = fmaf(a.w2[ch * a.hid + j], sm.dv[ch], s);
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
      if (lane == 0) {
        const float dp = sm.hdn[j] > 0.f ? s : 0.f;
        sm.dpre[j] = dp;
        atomicAdd(o.db1 + j, dp);
      }
    }
    __syncthreads();
    #pragma unroll 1
    for (int ch = tid; ch < c; ch += kT) {
      const float dv = sm.dv[ch];
      atomicAdd(o.db2 + ch, dv);
      float dt = 0.f;
      #pragma unroll 1
      for (int j = 0; j < a.hid; ++j) {
        atomicAdd(o.dw2 + ch * a.hid + j, dv * sm.hdn[j]);
        atomicAdd(o.dw1 + j * c + ch, sm.dpre[j] * sm.t[ch]);
        dt = fmaf(a.w1[j * c + ch], sm.dpre[j], dt);
      }
      sm.dv[ch] = dt;  // reuse: dt[c]
    }
  } else {
    #pragma unroll 1
    for (int ch = tid; ch < c; ch += kT) sm.dv[ch] = 0.f;
  }
  __syncthreads();

  // (c) dSg = dt/(D*hw);  dw2[d] = dw12*w1 + sum_c dSg*Sf
  float acc[MD];
#pragma unroll
  for (int i = 0; i < MD; ++i) acc[i] = 0.f;
  #pragma unroll 1
  for (int ch = tid; ch < c; ch += kT) {
    const float dsg = sm.dv[ch] * inv_dhw;
#pragma unroll
    for (int dd = 0; dd < MD; ++dd)
      if (dd < d) acc[dd] = fmaf(dsg, sm.Se[dd * c + ch] * sm.w1[dd], acc[dd]);
  }
  block_sum_vec(acc, d, sm.red);
  // (d) ds2 -> uniform part of dSf ; (e) dw1[d] = dw12*w2 + sum_c dSf*Se
  float ds2c[MD];  // ds2[d] / (c*hw)
#pragma unroll
  for (int i = 0; i < MD; ++i) {
    ds2c[i] = 0.f;
    if (i < d) {
      const float dw2 = dw12[i] * sm.w1[i] + acc[i];
      const float w2 = sm.w2[i];
      ds2c[i] = (a.flags & SPFF_GATE_SPECSE) ? dw2 * w2 * (1.f - w2) * inv_chw : 0.f;
    }
  }
#pragma unroll
  for (int i = 0; i < MD; ++i) acc[i] = 0.f;
  #pragma unroll 1
  for (int ch = tid; ch < c; ch += kT) {
    const float dsg = sm.dv[ch] * inv_dhw;
#pragma unroll
    for (int dd = 0; dd < MD; ++dd)
      if (dd < d) acc[dd] = fmaf(dsg * sm.w2[dd] + ds2c[dd], sm.Se[dd * c + ch], acc[dd]);
  }
  block_sum_vec(acc, d, sm.red);
  // (f) Fourier gate backward: du, ds1, dkfg
  // dw1[d] lives in registers (acc/dw12 are block-uniform): every thread computes the small vectors
  float du[MD], ds1c[MD];
#pragma unroll
  for (int i = 0; i < MD; ++i) {
    du[i] = 0.f;
    if (i < d) {
      const float dw1 = dw12[i] * sm.w2[i] + acc[i];
      const float w1 = sm.w1[i];
      du[i] = (a.flags & SPFF_GATE_FOURIER) ? dw1 * w1 * (1.f - w1) : 0.f;
    }
  }
#pragma unroll
  for (int e = 0; e < MD; ++e) {
    ds1c[e] = 0.f;
    if (e < d && (a.flags & SPFF_GATE_FOURIER)) {
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < MD; ++i)
        if (i < d) s = fmaf(du[i], a.kfg[(i - e + d) % d], s);
      ds1c[e] = s * inv_chw;
    }
  }
  if ((a.flags & SPFF_GATE_FOURIER) && tid < d) {
    // dkfg[r] += sum_i du[i] * s1[(i - r) mod d],  r = tid
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < MD; ++i)
      if (i < d) s = fmaf(du[i], sm.s1[(i - tid + d) % d], s);
    atomicAdd(o.dkfg + tid, s);
  }
  // (g,h,i,j) per element: dSe, table grads through S_e, dS; IN backward coefficients
  #pragma unroll 1
  for (int ch = tid; ch < c; ch += kT) {
    const float dsg = sm.dv[ch] * inv_dhw;
    float sum_dz = 0.f, sum_dzx = 0.f;
#pragma unroll
    for (int dd = 0; dd < MD; ++dd) {
      if (dd < d) {
        const float dSf = dsg * sm.w2[dd] + ds2c[dd];
        const float dSe = dSf * sm.w1[dd] + ds1c[dd];
        float dS = dSe;
        const float g1 = efilm ? a.g1[ch * d + dd] : 1.f;
        if (efilm) {
          dS = dSe * g1;
          atomicAdd(o.dg1 + ch * d + dd, dSe * S[dd * c + ch]);
          atomicAdd(o.dbt + ch * d + dd, dSe * a.hw);
        }
        const float Pv = g1 * sm.w1[dd] * sm.w2[dd] * sm.w3[ch];
        const size_t idx = static_cast<size_t>(n) * d * c + dd * c + ch;
        if (o.dSa) o.dSa[idx] = dS;
        if (o.Pout) o.Pout[idx] = Pv;
        const float* r = Rn + (dd * c + ch) * 6;
        sum_dz += Pv * r[2] + dS * r[3];
        sum_dzx += Pv * r[4] + dS * r[5];
      }
    }
    const float4 cf = reinterpret_cast<const float4*>(coef)[static_cast<size_t>(n) * c + ch];
    const float ga = gamma ? gamma[ch] : 1.f;
    float4 bc;
    bc.x = ga * cf.w;
    bc.y = sum_dz * inv_dhw;
    bc.z = sum_dzx * inv_dhw;
    bc.w = 0.f;
    reinterpret_cast<float4*>(o.bcoef)[static_cast<size_t>(n) * c + ch] = bc;
    if (o.dgamma) atomicAdd(o.dgamma + ch, sum_dzx);
    if (o.dbeta) atomicAdd(o.dbeta + ch, sum_dz);
  }
}

// Gate-free case (plain InstanceNorm + LeakyReLU backward): per (n, c)
//   bcoef = {gamma*rstd, mean(dz), mean(dz*xhat), 0},  dgamma += sum dz*xhat,  dbeta += sum dz
// from the slots 2 and 4 of R. One thread per (sample, channel); no shared memory, no phases.
__global__ void in_bwd_coeffs_kernel(const float* __restrict__ R, const float* __restrict__ coef,
                                     const float* __restrict__ gamma, int n, int c, int d, float inv_dhw,
                                     float* __restrict__ bcoef, float* __restrict__ dgamma, float* __restrict__ dbeta) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * c) return;
  const int s = i / c, ch = i % c;
  float sum_dz = 0.f, sum_dzx = 0.f;
  for (int dd = 0; dd < d; ++dd) {
    const float* r = R + ((static_cast<size_t>(s) * d + dd) * c + ch) * 6;
    sum_dz += r[2];
    sum_dzx += r[4];
  }
  const float4 cf = reinterpret_cast<const float4*>(coef)[i];
  float4 bc;
  bc.x = (gamma ? gamma[ch] : 1.f) * cf.w;
  bc.y = sum_dz * inv_dhw;
  bc.z = sum_dzx * inv_dhw;
  bc.w = 0.f;
  reinterpret_cast<float4*>(bcoef)[i] = bc;
  if (dgamma) atomicAdd(dgamma + ch, sum_dzx);
  if (dbeta) atomicAdd(dbeta + ch, sum_dz);
}

int check_gate_args(const GateArgs& a, int n) {
  if (a.c <= 0 || a.d <= 0 || a.d > kMaxD || n <= 0) {
    set_error("gate_micro: bad shape (c %d, d %d <= %d, n %d)", a.c, a.d, kMaxD, n);
    return SPFF_ERR_BAD_ARGUMENT;
  }
  if ((a.flags & SPFF_GATE_EFILM) && (!a.g1 || !a.bt)) {
    set_error("gate_micro: EFILM needs g1 and bt");
    return SPFF_ERR_BAD_ARGUMENT;
  }
  if ((a.flags & SPFF_GATE_FOURIER) && !a.kfg) {
    set_error("gate_micro: FOURIER needs kfg");
    return SPFF_ERR_BAD_ARGUMENT;
  }
  if ((a.flags & SPFF_GATE_CHANSE) && (!a.w1 || !a.b1 || !a.w2 || !a.b2 || a.hid <= 0 || a.hid > kMaxHid)) {
    set_error("gate_micro: CHANSE needs the fc parameters and 0 < hid <= %d", kMaxHid);
    return SPFF_ERR_BAD_ARGUMENT;
  }
  return 0;
}


// ---------------------------------------------------------------------------------------------
// Parameter-only tables of the gates, and their backward (one CTA per block; replaces ~40 tiny tensor ops + an
// autograd pass per block and step):
//   EnergyFiLM3D (models.py:1494-1512): pe[j][f] sinusoidal code of the bin index (16 rows), h = relu(W0 pe + b0) [32][F],
//     gb = W2 h + b2 [2C][F], g1 = 1 + tanh(gb[:C]), bt = gb[C:]
//   FourierGate3D (models.py:1537-1542): kfg = irfft(freq_mask * mag_scale, n = F)  (real spectrum of L = F/2 + 1 bins)
// ---------------------------------------------------------------------------------------------
constexpr int kPe = 16, kHidE = 32, kMaxF = 16;

__device__ __forceinline__ float pe_value(int j, int f) {
  const int half = kPe / 2;
  const int i = j < half ? j : j - half;
  const float denom = expf(static_cast<float>(i) * (-logf(10000.f) / static_cast<float>(half)));
  const float a = static_cast<float>(f) * denom;
  return j < half ? sinf(a) : cosf(a);
}
// weight of spectral bin l in the inverse real FFT of length F: x[t] = (1/F) sum_l w_l M_l cos(2 pi l t / F)
__device__ __forceinline__ float irfft_weight(int l, int F) { return (l == 0 || (F % 2 == 0 && l == F / 2)) ? 1.f : 2.f; }

__global__ void gate_tables_fwd_kernel(const float* __restrict__ w0, const float* __restrict__ b0, const float* __restrict__ w2,
                                       const float* __restrict__ b2, const float* __restrict__ mask,
                                       const float* __restrict__ scale, int C, int F, float* __restrict__ g1,
                                       float* __restrict__ bt, float* __restrict__ kfg) {
  __shared__ float pe[kPe][kMaxF];
  __shared__ float h[kHidE][kMaxF];
  const int tid = threadIdx.x;
  if (w0) {
    for (int i = tid; i < kPe * F; i += blockDim.x) pe[i / F][i % F] = pe_value(i / F, i % F);
    __syncthreads();
    for (int i = tid; i < kHidE * F; i += blockDim.x) {
      const int k = i / F, f = i % F;
      float a = b0[k];
      for (int j = 0; j < kPe; ++j) a = fmaf(w0[k * kPe + j], pe[j][f], a);
      h[k][f] = fmaxf(a, 0.f);
    }
    __syncthreads();
    for (int i = tid; i < 2 * C * F; i += blockDim.x) {
      const int m = i / F, f = i % F;
      float a = b2[m];
      for (int k = 0; k < kHidE; ++k) a = fmaf(w2[m * kHidE + k], h[k][f], a);
      if (m < C) g1[m * F + f] = 1.f + tanhf(a);
      else bt[(m - C) * F + f] = a;
    }
  }
  if (mask && tid < F) {
    const int L = F / 2 + 1;
    const float sc = scale[0];
    float a = 0.f;
    for (int l = 0; l < L; ++l)
      a += irfft_weight(l, F) * mask[l] * sc * cospif(2.f * static_cast<float>((l * tid) % F) / static_cast<float>(F));
    kfg[tid] = a / static_cast<float>(F);
  }
}

// Accumulates (+=) the parameter gradients from the table gradients dg1 / dbt [C][F] and dkfg [F]. One CTA; every
// output element has one writer, sums run in a fixed order.
__global__ void gate_tables_bwd_kernel(const float* __restrict__ w0, const float* __restrict__ b0, const float* __restrict__ w2,
                                       const float* __restrict__ b2, const float* __restrict__ mask,
                                       const float* __restrict__ scale, int C, int F, const float* __restrict__ dg1,
                                       const float* __restrict__ dbt, const float* __restrict__ dkfg, float* __restrict__ dw0,
                                       float* __restrict__ db0, float* __restrict__ dw2, float* __restrict__ db2,
                                       float* __restrict__ dmask, float* __restrict__ dscale) {
  extern __shared__ float sm[];
  const int tid = threadIdx.x;
  if (w0) {
    float* pe = sm;                      // [kPe][F]
    float* pre = pe + kPe * F;           // [kHidE][F]  pre-activation of the hidden layer
    float* dgb = pre + kHidE * F;        // [2C][F]
    float* dh = dgb + 2 * C * F;         // [kHidE][F]
    for (int i = tid; i < kPe * F; i += blockDim.x) pe[i] = pe_value(i / F, i % F);
    __syncthreads();
    for (int i = tid; i < kHidE * F; i += blockDim.x) {
      const int k = i / F, f = i % F;
      float a = b0[k];
      for (int j = 0; j < kPe; ++j) a = fmaf(w0[k * kPe + j], pe[j * F + f], a);
      pre[i] = a;
    }
    __syncthreads();
    for (int i = tid; i < 2 * C * F; i += blockDim.x) {
      const int m = i / F, f = i % F;
      if (m < C) {
        float a = b2[m];
        for (int k = 0; k < kHidE; ++k) a = fmaf(w2[m * kHidE + k], fmaxf(pre[k * F + f], 0.f), a);
        const float th = tanhf(a);
        dgb[i] = dg1[m * F + f] * (1.f - th * th);
      } else {
        dgb[i] = dbt[(m - C) * F + f];
      }
    }
    __syncthreads();
    for (int i = tid; i < 2 * C * (kHidE + 1); i += blockDim.x) {     // dW2 [2C][32] and db2 [2C]
      const int m = i / (kHidE + 1), k = i % (kHidE + 1);
      float a = 0.f;
      if (k < kHidE) {
        for (int f = 0; f < F; ++f) a = fmaf(dgb[m * F + f], fmaxf(pre[k * F + f], 0.f), a);
        dw2[m * kHidE + k] += a;
      } else {
        for (int f = 0; f < F; ++f) a += dgb[m * F + f];
        db2[m] += a;
      }
    }
    for (int i = tid; i < kHidE * F; i += blockDim.x) {
      const int k = i / F, f = i % F;
      float a = 0.f;
      for (int m = 0; m < 2 * C; ++m) a = fmaf(w2[m * kHidE + k], dgb[m * F + f], a);
      dh[i] = pre[i] > 0.f ? a : 0.f;
    }
    __syncthreads();
    for (int i = tid; i < kHidE * (kPe + 1); i += blockDim.x) {       // dW0 [32][16] and db0 [32]
      const int k = i / (kPe + 1), j = i % (kPe + 1);
      float a = 0.f;
      if (j < kPe) {
        for (int f = 0; f < F; ++f) a = fmaf(dh[k * F + f], pe[j * F + f], a);
        dw0[k * kPe + j] += a;
      } else {
        for (int f = 0; f < F; ++f) a += dh[k * F + f];
        db0[k] += a;
      }
    }
  }
  if (mask && tid == 0) {
    const int L = F / 2 + 1;
    const float sc = scale[0];
    float ds = 0.f;
    for (int l = 0; l < L; ++l) {
      float dm = 0.f;
      for (int t = 0; t < F; ++t)
        dm += dkfg[t] * cospif(2.f * static_cast<float>((l * t) % F) / static_cast<float>(F));
      dm *= irfft_weight(l, F) / static_cast<float>(F);
      dmask[l] += dm * sc;
      ds += dm * mask[l];
    }
    dscale[0] += ds;
  }
}

}  // namespace
}  // namespace spff

extern "C" {

int spff_gate_micro_fwd(const float* S, const float* g1, const float* bt, const float* kfg, const float* se_w1,
                        const float* se_b1, const float* se_w2, const float* se_b2, int hid, int flags, int c,
                        spff_shape s, float* P, float* Q, void* stream) {
  int e = spff_device_check();
  if (e) return e;
  spff::GateArgs a{S, g1, bt, kfg, se_w1, se_b1, se_w2, se_b2, hid, flags, c, s.d, static_cast<float>(s.h) * s.w};
  e = spff::check_gate_args(a, s.n);
  if (e) return e;
  SPFF_REQUIRE(S && P && Q, "gate_micro_fwd: null pointer");
  const size_t smem = spff::gate_smem_bytes(c, s.d);
  auto kern = s.d == 5 ? spff::gate_fwd_kernel<5, 5> : (s.d <= 8 ? spff::gate_fwd_kernel<8, 0> : spff::gate_fwd_kernel<16, 0>);
  if (smem > 48 * 1024)
    SPFF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  kern<<<s.n, spff::kT, smem, static_cast<cudaStream_t>(stream)>>>(a, P, Q);
  SPFF_CUDA(cudaGetLastError());
  return 0;
}

int spff_gate_micro_bwd(const float* R, const float* S, const float* coef, const float* gamma, const float* g1,
                        const float* bt, const float* kfg, const float* se_w1, const float* se_b1,
                        const float* se_w2, const float* se_b2, int hid, int flags, int c, spff_shape s,
                        float* bcoef, float* dSa, float* Pout, float* dgamma, float* dbeta, float* dg1, float* dbt,
                        float* dkfg, float* dse_w1, float* dse_b1, float* dse_w2, float* dse_b2, void* stream) {
  int e = spff_device_check();
  if (e) return e;
  spff::GateArgs a{S, g1, bt, kfg, se_w1, se_b1, se_w2, se_b2, hid, flags, c, s.d, static_cast<float>(s.h) * s.w};
  e = spff::check_gate_args(a, s.n);
  if (e) return e;
  SPFF_REQUIRE(R && coef && bcoef && (flags == 0 || (S && dSa && Pout)), "gate_micro_bwd: null pointer");
  SPFF_REQUIRE(!(flags & SPFF_GATE_EFILM) || (dg1 && dbt), "gate_micro_bwd: EFILM needs dg1/dbt");
  SPFF_REQUIRE(!(flags & SPFF_GATE_FOURIER) || dkfg, "gate_micro_bwd: FOURIER needs dkfg");
  SPFF_REQUIRE(!(flags & SPFF_GATE_CHANSE) || (dse_w1 && dse_b1 && dse_w2 && dse_b2), "gate_micro_bwd: CHANSE needs d(fc)");
  if (flags == 0 && !dSa && !Pout) {
    const int total = s.n * c;
    spff::in_bwd_coeffs_kernel<<<(total + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(
        R, coef, gamma, s.n, c, s.d, 1.f / (static_cast<float>(s.d) * s.h * s.w), bcoef, dgamma, dbeta);
    SPFF_CUDA(cudaGetLastError());
    return 0;
  }
  spff::GateBwdOut o{bcoef, dSa, Pout, dgamma, dbeta, dg1, dbt, dkfg, dse_w1, dse_b1, dse_w2, dse_b2};
  const size_t smem = spff::gate_smem_bytes(c, s.d);
  auto kern = s.d == 5 ? spff::gate_bwd_kernel<5, 5> : (s.d <= 8 ? spff::gate_bwd_kernel<8, 0> : spff::gate_bwd_kernel<16, 0>);
  if (smem > 48 * 1024)
    SPFF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  kern<<<s.n, spff::kT, smem, static_cast<cudaStream_t>(stream)>>>(a, R, coef, gamma, o);
  SPFF_CUDA(cudaGetLastError());
  return 0;
}

/* Gate tables from the parameters (EnergyFiLM3D MLP over the sinusoidal bin code, models.py:1494-1512; FourierGate3D
 * kernel irfft(freq_mask * mag_scale), models.py:1537-1542). Either group of pointers may be NULL (variant without that gate). */
int spff_gate_tables_fwd(const float* w0, const float* b0, const float* w2, const float* b2, const float* freq_mask,
                         const float* mag_scale, int c, int frames, float* g1, float* bt, float* kfg, void* stream) {
  int e = spff_device_check();
  if (e) return e;
  SPFF_REQUIRE(c > 0 && frames > 0 && frames <= spff::kMaxF, "gate_tables_fwd: 0 < frames <= %d, got %d", spff::kMaxF, frames);
  SPFF_REQUIRE(!w0 || (b0 && w2 && b2 && g1 && bt), "gate_tables_fwd: EFiLM needs all of w0 b0 w2 b2 g1 bt");
  SPFF_REQUIRE(!freq_mask || (mag_scale && kfg), "gate_tables_fwd: FourierGate needs freq_mask, mag_scale, kfg");
  if (!w0 && !freq_mask) return 0;
  spff::gate_tables_fwd_kernel<<<1, 256, 0, static_cast<cudaStream_t>(stream)>>>(w0, b0, w2, b2, freq_mask, mag_scale, c, frames,
                                                                                  g1, bt, kfg);
  SPFF_CUDA(cudaGetLastError());
  return 0;
}

/* Their backward: the parameter gradients are ACCUMULATED (+=) from the table gradients the gate kernels summed. */
int spff_gate_tables_bwd(const float* w0, const float* b0, const float* w2, const float* b2, const float* freq_mask,
                         const float* mag_scale, int c, int frames, const float* dg1, const float* dbt, const float* dkfg,
                         float* dw0, float* db0, float* dw2, float* db2, float* dfreq_mask, float* dmag_scale, void* stream) {
  int e = spff_device_check();
  if (e) return e;
  SPFF_REQUIRE(c > 0 && frames > 0 && frames <= spff::kMaxF, "gate_tables_bwd: 0 < frames <= %d, got %d", spff::kMaxF, frames);
  SPFF_REQUIRE(!w0 || (b0 && w2 && b2 && dg1 && dbt && dw0 && db0 && dw2 && db2), "gate_tables_bwd: EFiLM pointers incomplete");
  SPFF_REQUIRE(!freq_mask || (mag_scale && dkfg && dfreq_mask && dmag_scale), "gate_tables_bwd: FourierGate pointers incomplete");
  if (!w0 && !freq_mask) return 0;
  const size_t smem = sizeof(float) * (static_cast<size_t>(spff::kPe) * frames + 2 * spff::kHidE * frames + 2 * static_cast<size_t>(c) * frames);
  auto kern = spff::gate_tables_bwd_kernel;
  if (smem > 48 * 1024) SPFF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  kern<<<1, 256, smem, static_cast<cudaStream_t>(stream)>>>(w0, b0, w2, b2, freq_mask, mag_scale, c, frames, dg1, dbt, dkfg, dw0, db0,
                                                           dw2, db2, dfreq_mask, dmag_scale);
  SPFF_CUDA(cudaGetLastError());
  return 0;
}

}  // extern "C"
