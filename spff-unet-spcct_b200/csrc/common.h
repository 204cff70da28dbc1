// Host-side plumbing shared by every translation unit of libspff_b200.so: error reporting across
// the C ABI (include/spff_b200.h), the driver entry point for TMA descriptors, tile geometry.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/spff_b200.h"

namespace spff {

// thread-local last error text, returned by spff_last_error()
void set_error(const char* fmt, ...);
int check_cuda(cudaError_t e, const char* what);
int num_sms();
// index of the current CUDA device, clamped to [0, kMaxDevices) - key of the per-device caches
constexpr int kMaxDevices = 64;
int current_device();
// test hook (spff_debug_set key 0): CTA count override for persistent kernels, 0 = one per SM
int debug_ctas();
// test hook (spff_debug_set key >= 1): generic integer flags, 0 by default. key 1: disable the all-kh wgrad variant
int debug_flag(int key);
long long debug_value(int key);   // the same, 64 bit (key 4: device pointer of a cycle-counter buffer, profiling builds of the conv kernels)
// conv3_halo.cu: the halo-tile formulation of the 3x3x3 convolution (planes tiled by 16 x 8 boxes)
bool conv3_halo_applicable(spff_shape s, int gemm_k);
// statistics slots per sample (conv3_fprop.cu; what spff_conv3d_k3_stat_slots reports, whichever kernel runs)
int conv3_rows_stat_slots(spff_shape s);
int conv3_halo_pack(const float* w, void* out, int cout, int cin, int dgrad, cudaStream_t st);
int conv3_halo_launch(const void* x, long long ldx, int cin, const void* wpk, void* y, long long ldy, int cout, spff_shape s,
                      float* stat_partial, cudaStream_t st);
// K chunk (channels per TMA box) the conv3 kernels use for a GEMM-K channel count: 64 or 32
int conv3_kc(int gemm_k_channels);

#define SPFF_REQUIRE(cond, ...)          \
  do {                                   \
    if (!(cond)) {                       \
      ::spff::set_error(__VA_ARGS__);    \
      return SPFF_ERR_BAD_ARGUMENT;      \
    }                                    \
  } while (0)

#define SPFF_CUDA(expr)                                  \
  do {                                                   \
    int _e = ::spff::check_cuda((expr), #expr);          \
    if (_e) return _e;                                   \
  } while (0)

// Encode a tiled TMA descriptor over a bf16 tensor of rank `rank` (dim 0 innermost, contiguous).
// strides_bytes[i] is the byte stride of dim i+1. swizzle_bytes in {0, 64, 128}.
int encode_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                     const uint64_t* strides_bytes, const uint32_t* box, int swizzle_bytes);

// A 128-row (or 64-row) tile of "positions": a box over the four outer dims of a
// [c, x1, x2, x3, x4] tensor (for an NDHWC activation: x1=W, x2=H, x3=D, x4=N).
struct TileGeom {
  int ext[4];    // extents of x1..x4
  int box[4];    // box size per dim (powers of two, product = rows)
  int tiles[4];  // ceil(ext / box)
  int ntiles;    // product of tiles
};
// rows must be a power of two. Fills boxes greedily from the innermost dim.
TileGeom make_tile_geom(const int ext[4], int rows);

}  // namespace spff
