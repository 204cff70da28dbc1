// Library plumbing: error text, device check, TMA descriptor encoding, tile geometry.
#include "common.h"

#include <stdarg.h>
#include <string.h>

#include <mutex>

namespace spff {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_cuda(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return 0;
  set_error("CUDA error %s (%d) at %s", cudaGetErrorString(e), static_cast<int>(e), what);
  return SPFF_ERR_CUDA;
}

int current_device() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0) return 0;
  return dev < kMaxDevices ? dev : kMaxDevices - 1;
}

int num_sms() {
  static int cached[kMaxDevices] = {};   // per device: one process may drive several GPUs
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  const int slot = (dev >= 0 && dev < kMaxDevices) ? dev : kMaxDevices - 1;
  if (cached[slot]) return cached[slot];
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) return 148;
  cached[slot] = n;
  return n;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess) {
      fn = reinterpret_cast<EncodeTiledFn>(p);
    }
  });
  return fn;
}

int encode_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                     const uint64_t* strides_bytes, const uint32_t* box, int swizzle_bytes) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled is not available from the CUDA driver");
    return SPFF_ERR_CUDA;
  }
  cuuint64_t gdim[5];
  cuuint64_t gstr[4];
  cuuint32_t bx[5];
  cuuint32_t es[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
    if (i > 0) gstr[i - 1] = strides_bytes[i - 1];
  }
  CUtensorMapSwizzle sw = swizzle_bytes == 128  ? CU_TENSOR_MAP_SWIZZLE_128B
                          : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                          : swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B
                                                : CU_TENSOR_MAP_SWIZZLE_NONE;
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, static_cast<cuuint32_t>(rank), const_cast<void*>(base),
                  gdim, gstr, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d): rank %d dims [%llu %llu %llu %llu %llu] box [%u %u %u %u %u]",
              static_cast<int>(r), rank, (unsigned long long)gdim[0], (unsigned long long)(rank > 1 ? gdim[1] : 0),
              (unsigned long long)(rank > 2 ? gdim[2] : 0), (unsigned long long)(rank > 3 ? gdim[3] : 0),
              (unsigned long long)(rank > 4 ? gdim[4] : 0), bx[0], rank > 1 ? bx[1] : 0, rank > 2 ? bx[2] : 0,
              rank > 3 ? bx[3] : 0, rank > 4 ? bx[4] : 0);
    return SPFF_ERR_CUDA;
  }
  return 0;
}

static int pow2_floor(int v) {
  int p = 1;
  while (p * 2 <= v) p *= 2;
  return p;
}
static int pow2_ceil(int v) {
  int p = 1;
  while (p < v) p *= 2;
  return p;
}

TileGeom make_tile_geom(const int ext[4], int rows) {
  TileGeom g;
  int remaining = rows;
  for (int i = 0; i < 4; ++i) {
    g.ext[i] = ext[i];
    int b;
    if (i == 0) {
      // innermost (W): at most 16 wide so that a tile stays compact in (h, w)
      b = pow2_floor(ext[i] < 16 ? ext[i] : 16);
      if (b > remaining) b = remaining;
    } else if (i == 3) {
      b = remaining;  // whatever is left goes to the outermost dim (overhang is masked)
    } else {
      b = pow2_ceil(ext[i]);
      if (i == 1) b = pow2_floor(ext[i]);
      if (b > remaining) b = remaining;
    }
    if (b < 1) b = 1;
    g.box[i] = b;
    remaining /= b;
  }
  g.ntiles = 1;
  for (int i = 0; i < 4; ++i) {
    g.tiles[i] = (g.ext[i] + g.box[i] - 1) / g.box[i];
    g.ntiles *= g.tiles[i];
  }
  return g;
}

static int g_debug_ctas = 0;
int debug_ctas() { return g_debug_ctas; }
static long long g_debug_flags[16] = {0};
int debug_flag(int key) { return (key >= 0 && key < 16) ? static_cast<int>(g_debug_flags[key]) : 0; }
long long debug_value(int key) { return (key >= 0 && key < 16) ? g_debug_flags[key] : 0; }

}  // namespace spff

extern "C" {

int spff_version(void) { return 100; }

const char* spff_last_error(void) { return spff::g_err; }

int spff_device_check(void) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) {
    spff::set_error("no CUDA device: %s", cudaGetErrorString(e));
    return SPFF_ERR_UNSUPPORTED_ARCH;
  }
  int major = 0, minor = 0;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
  if (major != 10) {
    spff::set_error("libspff_b200 needs an sm_100 (B200) device, found sm_%d%d; there is no fallback path", major,
                    minor);
    return SPFF_ERR_UNSUPPORTED_ARCH;
  }
  return 0;
}

int spff_debug_set(int key, long long value) {
  if (key == 0) {
    spff::g_debug_ctas = static_cast<int>(value);
    return 0;
  }
  if (key > 0 && key < 16) {
    spff::g_debug_flags[key] = value;
    return 0;
  }
  spff::set_error("unknown debug key %d", key);
  return SPFF_ERR_BAD_ARGUMENT;
}

}  // extern "C"
