/*
 * spff_b200.h — C ABI of libspff_b200.so: the B200 (sm_100a) kernels behind the SPFF-UNet hot path.
 *
 * The reference (NF-91/spff-unet-spcct) is pure Python: its hot path is the sequence of ATen /
 * cuDNN / cuFFT library calls that `innovative3D/models.py` and `innovative3D/helpers.py` dispatch.
 * Each entry point below replaces one of those call sites (cited per function as file:line of the
 * reference). The host layer that binds them is `spff-unet-spcct_b200/spff_b200/_lib.py` (ctypes);
 * INTEGRATION.md shows the binding a maintainer of the reference would add.
 *
 * Conventions
 *  - Plain C types only. Every pointer is a DEVICE pointer unless the name ends in `_host`.
 *  - Activations are bf16, position-major ("NDHWC"): element (n,d,h,w,c) of a view lives at
 *    base + (((n*D + d)*H + h)*W + w)*ld + c, `ld` = channel pitch in elements (ld >= C, ld % 8 == 0,
 *    base 16-byte aligned). A channel slice of a wider buffer (skip-concat halves) is expressed by
 *    offsetting `base` and keeping the buffer's `ld`.
 *  - The five energy bins are the D axis (reference: models.py:1551, datasets.py:228-233).
 *  - All work is enqueued on `stream` (a cudaStream_t passed as void*); no call synchronises the host.
 *  - Return 0 on success, a negative SPFF_ERR_* otherwise; spff_last_error() holds the text
 *    (thread local). There is no CPU fallback: on a device that is not sm_100 every compute entry
 *    point returns SPFF_ERR_UNSUPPORTED_ARCH.
 *  - The library allocates no persistent device memory; scratch is caller-provided `workspace`.
 */
#ifndef SPFF_B200_H_
#define SPFF_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SPFF_OK 0
#define SPFF_ERR_BAD_ARGUMENT (-1)
#define SPFF_ERR_UNSUPPORTED_ARCH (-2)
#define SPFF_ERR_CUDA (-3)
#define SPFF_ERR_WORKSPACE (-4)

/* gate flags for the collapsed SPFF tail (spff_gate_micro_*) */
#define SPFF_GATE_EFILM 1   /* EnergyFiLM3D      models.py:1479-1512 */
#define SPFF_GATE_FOURIER 2 /* FourierGate3D     models.py:1515-1544 */
#define SPFF_GATE_SPECSE 4  /* _SpectralSE       models.py:611-614   */
#define SPFF_GATE_CHANSE 8  /* _SEChannelLite    models.py:600-609   */

typedef struct spff_shape {
  int n, d, h, w; /* samples, energy bins (depth), height, width of the position grid */
} spff_shape;

/* ---- library ------------------------------------------------------------------------------ */
int spff_version(void);
const char* spff_last_error(void);
/* 0 when the current CUDA device is sm_100 (B200); SPFF_ERR_UNSUPPORTED_ARCH otherwise. */
int spff_device_check(void);
/* test hook. key 0: number of CTAs for persistent kernels (0 = one per SM). */
int spff_debug_set(int key, long long value);

/* ---- 3x3x3 convolution, stride 1, zero pad 1, no bias ----------------------------------------
 * Replaces nn.Conv3d(cin, cout, (3,3,3), padding 1, bias=False) built by `_conv3x3xk`
 * (models.py:616-618) — forward, and the input / weight halves of its convolution_backward.
 * cin, cout multiples of 32 (the Cin = 1 stem has its own entry points below). */
/* Elements of one packed weight operand: 27*cin*cout bf16. */
/* nn.Conv3d weight [cout][cin][3][3][3] fp32 -> bf16 GEMM operands (either may be NULL):
 *   w_fwd   [cout/32][kh][cin/KC][kd][kw][32][KC]          KC = 64 if cin % 64 == 0 else 32
 *   w_dgrad [cin/32][kh][cout/KC'][kd][kw][32][KC']        taps flipped, in/out transposed */
int spff_pack_conv3_weight(const float* w, void* w_fwd, void* w_dgrad, int cout, int cin, void* stream);
/* y[n,d,h,w,0:cout] = conv3d(x)[...]  (F.conv3d at models.py:616-618 via nn.Sequential :1459-1469) */
int spff_conv3d_k3_fwd(const void* x, long long ldx, int cin, const void* w_fwd, void* y, long long ldy, int cout,
                       spff_shape s, void* stream);
/* dx = input gradient of the same convolution (ATen convolution_backward, grad_input). */
int spff_conv3d_k3_dgrad(const void* dy, long long lddy, int cout, const void* w_dgrad, void* dx, long long lddx,
                         int cin, spff_shape s, void* stream);

/* dw[cout][cin][3][3][3] (fp32) = beta*dw + weight gradient (ATen convolution_backward, grad_weight).
 * Split over positions; fp32 partial tiles go to `workspace` (size from the query below, which
 * depends on the SM count of the current device) and are reduced in a fixed order. */
size_t spff_conv3d_k3_wgrad_workspace(int cin, int cout, spff_shape s);
int spff_conv3d_k3_wgrad(const void* x, long long ldx, int cin, const void* dy, long long lddy, int cout,
                         spff_shape s, float* dw, float beta, void* workspace, size_t workspace_bytes,
                         void* stream);

/* @@ENTRY_POINTS@@ */

#ifdef __cplusplus
}
#endif
#endif /* SPFF_B200_H_ */
