#include "common.cuh"

#include <mutex>
#include <string>

namespace urso {

static thread_local std::string g_err;
static int g_dry_run = 0;    // urso_set_dry_run: plan only (no tensor-map encoding, no device access)

void set_error(const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_err = buf;
}

int num_sms() {
  static int n = 0;
  if (g_dry_run) return 148;      // planning without a device: assume a B200
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) n = 0;
  }
  return n;
}

static int g_pdl = 1;
static int g_residual_mma = 1;
static int g_wgrad_halo = 1;
static int g_tail_split = 1;
bool pdl_enabled() { return g_pdl != 0; }
bool wgrad_halo_enabled() { return g_wgrad_halo != 0; }
bool tail_split_enabled() { return g_tail_split != 0; }
bool residual_mma_enabled() { return g_residual_mma != 0; }
static int g_max_ctas = 0;
bool dry_run() { return g_dry_run != 0; }
int max_ctas() {
  const int n = num_sms();
  return (g_max_ctas > 0 && g_max_ctas < n) ? g_max_ctas : n;
}

encode_tiled_fn get_encode_tiled() {
  static encode_tiled_fn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<encode_tiled_fn>(p);
  });
  return fn;
}

int make_view_map(CUtensorMap* out, const urso_view4& v, int box_w, int box_h) {
  if (dry_run()) return 0;
  encode_tiled_fn enc = get_encode_tiled();
  URSO_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled not available (no CUDA driver?)");
  URSO_REQUIRE(v.base != nullptr && (reinterpret_cast<uintptr_t>(v.base) & 15) == 0, "view base must be 16B aligned");
  URSO_REQUIRE(v.C >= 64 && v.C % 8 == 0, "view C=%d must be >= 64 and a multiple of 8", v.C);
  URSO_REQUIRE(v.stride_w % 8 == 0 && v.stride_h % 8 == 0 && v.stride_n % 8 == 0,
               "view strides must be multiples of 8 elements (16 bytes)");
  URSO_REQUIRE(box_w >= 1 && box_w <= 256 && box_h >= 1 && box_h <= 256, "bad box %dx%d", box_w, box_h);
  cuuint64_t dims[4] = {(cuuint64_t)v.C, (cuuint64_t)v.W, (cuuint64_t)v.H, (cuuint64_t)v.N};
  cuuint64_t strides[3] = {(cuuint64_t)v.stride_w * 2, (cuuint64_t)v.stride_h * 2, (cuuint64_t)v.stride_n * 2};
  cuuint32_t box[4] = {64, (cuuint32_t)box_w, (cuuint32_t)box_h, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(v.base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  URSO_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(view C=%d W=%d H=%d N=%d) failed: %d", v.C, v.W, v.H, v.N,
               (int)r);
  return 0;
}

int make_mat_map(CUtensorMap* out, const void* base, int64_t rows, int64_t k, int box_rows) {
  if (dry_run()) return 0;
  encode_tiled_fn enc = get_encode_tiled();
  URSO_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled not available (no CUDA driver?)");
  URSO_REQUIRE(base != nullptr && (reinterpret_cast<uintptr_t>(base) & 15) == 0, "matrix base must be 16B aligned");
  URSO_REQUIRE(k % 64 == 0, "matrix K=%lld must be a multiple of 64", (long long)k);
  cuuint64_t dims[2] = {(cuuint64_t)k, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)k * 2};
  cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  URSO_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(matrix rows=%lld k=%lld) failed: %d", (long long)rows,
               (long long)k, (int)r);
  return 0;
}

}  // namespace urso

extern "C" {
int urso_version(void) { return 100; }
const char* urso_last_error(void) { return urso::g_err.c_str(); }
int urso_num_sms(void) { return urso::num_sms(); }
void urso_set_max_ctas(int n) { urso::g_max_ctas = n; }
void urso_set_dry_run(int on) { urso::g_dry_run = on; }
void urso_set_pdl(int on) { urso::g_pdl = on; }
void urso_set_residual_mma(int on) { urso::g_residual_mma = on; }
void urso_set_wgrad_halo(int on) { urso::g_wgrad_halo = on; }
void urso_set_tail_split(int on) { urso::g_tail_split = on; }
int urso_sizeof_convgemm_desc(void) { return (int)sizeof(urso_convgemm_desc); }
int urso_sizeof_wgrad_desc(void) { return (int)sizeof(urso_wgrad_desc); }
}
