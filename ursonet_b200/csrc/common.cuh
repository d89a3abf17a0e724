// Host-side helpers shared by the C-ABI translation units: error reporting, tensor-map encoding.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/urso_b200.h"

namespace urso {

void set_error(const char* fmt, ...);
int num_sms();
bool pdl_enabled();   // urso_set_pdl: launch the tcgen05 engines with programmatic stream serialization (default on)
bool residual_mma_enabled();   // urso_set_residual_mma: Engine F accumulates the addend on the tensor core (default on)
bool wgrad_halo_enabled();   // urso_set_wgrad_halo: Engine W reads the taps of a 3x3 / stem filter from one box + halo (default on)
bool tail_split_enabled();   // urso_set_tail_split: Engine F cuts the tiles of a partial last wave along N (default on)
bool dry_run();   // urso_set_dry_run(1): create-calls plan only (CPU-side tests of the planners)
int max_ctas();   // num_sms() or the urso_set_max_ctas() limit: grid size of the persistent Engine-F kernels

// cuTensorMapEncodeTiled resolved through the runtime (no link-time libcuda dependency, so the
// library loads -- and its symbols can be checked -- on a machine without a driver).
typedef CUresult (*encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
encode_tiled_fn get_encode_tiled();

// bf16 NHWC (possibly strided) view -> 4-D tensor map {C, W, H, N}, box {64, bw, bh, 1}, SWIZZLE_128B, zero OOB fill.
int make_view_map(CUtensorMap* out, const urso_view4& v, int box_w, int box_h);
// bf16 row-major [rows, k] -> 2-D tensor map {k, rows}, box {64, box_rows}, SWIZZLE_128B.
int make_mat_map(CUtensorMap* out, const void* base, int64_t rows, int64_t k, int box_rows);

// dense.cu: weight-streaming kernels of the Dense heads
int dense_fwd2(const float* x, const float* w, float* y, int B, int K, int N, cudaStream_t s);
int dense_dgrad2(const float* dy, const float* w, float* dx, int B, int K, int N, cudaStream_t s);
int dense_wgrad2(const float* x, const float* dy, float* dw, int B, int K, int N, cudaStream_t s);

#define URSO_CUDA_OK(expr)                                                                 \
  do {                                                                                     \
    cudaError_t _e = (expr);                                                               \
    if (_e != cudaSuccess) {                                                               \
      urso::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return 1;                                                                            \
    }                                                                                      \
  } while (0)

#define URSO_REQUIRE(cond, ...)      \
  do {                               \
    if (!(cond)) {                   \
      urso::set_error(__VA_ARGS__);  \
      return 2;                      \
    }                                \
  } while (0)

}  // namespace urso
