#include "../../include/db1_sm100.h"
#include "common.cuh"

#include <stdarg.h>
#include <stdlib.h>

namespace db1 {

static thread_local char g_err[512] = "";

char* err_buf() { return g_err; }

int set_err(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// libcuda is resolved at run time through the runtime's driver entry point so that the library still loads
// (for symbol checks) on a machine without a driver.
static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

int make_tmap_f16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                  const uint32_t* box, int swizzle_bytes) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return set_err(-2, "cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
  if (((uintptr_t)base & 15) != 0) return set_err(-1, "tensor map base %p not 16-byte aligned", base);
  cuuint64_t gdim[5];
  cuuint64_t gstr[5];
  cuuint32_t bx[5];
  cuuint32_t es[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
    if (i > 0) {
      gstr[i - 1] = strides_bytes[i - 1];
      if (gstr[i - 1] % 16 != 0) return set_err(-1, "tensor map stride %llu not a multiple of 16 bytes",
                                                 (unsigned long long)gstr[i - 1]);
    }
  }
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bx, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE,
                  swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_err(-3, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return 0;
}

bool pdl_enabled() {
  static int v = -1;
  if (v < 0) v = getenv("DB1_NO_PDL") ? 0 : 1;
  return v != 0;
}

static int g_sm_budget = 0;  // 0 = every SM; else the number of SMs persistent grids may cover (db1_set_sm_budget)

int sm_count_physical() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

int sm_count() {
  const int n = sm_count_physical();
  return (g_sm_budget > 0 && g_sm_budget < n) ? g_sm_budget : n;
}

void set_sm_budget(int n) { g_sm_budget = n; }

}  // namespace db1

extern "C" const char* db1_last_error() { return db1::err_buf(); }
extern "C" int db1_abi_version() { return 3; }

/* Persistent kernels of this library size their grids to one CTA (or CTA pair) per SM. While a communication kernel
 * (NCCL all-reduce) is co-resident it occupies some SMs for its whole duration; a 148-CTA grid then needs a second wave
 * for the CTAs that found no SM, i.e. up to twice the kernel time. The caller that knows a collective is in flight
 * (DB1Engine during backward) lowers the budget to physical SMs minus the collective's CTAs, and resets it with 0. */
extern "C" int db1_set_sm_budget(int n_sms) {
  DB1_CHECK_ARG(n_sms >= 0, "db1_set_sm_budget: negative budget");
  db1::set_sm_budget(n_sms >= 8 || n_sms == 0 ? n_sms : 8);
  return 0;
}
extern "C" int db1_sm_count(void) { return db1::sm_count_physical(); }
