// Host-side helpers shared by all translation units of libdb1_sm100.so.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/db1_sm100.h"

namespace db1 {

// thread-local last-error string returned by db1_last_error()
char* err_buf();
int set_err(int code, const char* fmt, ...);

#define DB1_CHECK_ARG(cond, ...)                    \
  do {                                              \
    if (!(cond)) return db1::set_err(-1, __VA_ARGS__); \
  } while (0)

#define DB1_CUDA(expr)                                                                  \
  do {                                                                                  \
    cudaError_t _e = (expr);                                                            \
    if (_e != cudaSuccess)                                                              \
      return db1::set_err((int)_e, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
  } while (0)

// Encode a 2-D/3-D fp16 tensor map with 128-byte swizzle. dims/strides innermost first; strides in BYTES for
// dims 1.. (dim 0 is contiguous). Returns 0 or a negative error after set_err().
int make_tmap_f16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                  const uint32_t* box, int swizzle_bytes = 128);

int sm_count();           // SMs persistent grids may cover (physical count unless db1_set_sm_budget lowered it)
int sm_count_physical();
bool pdl_enabled();  // programmatic dependent launch on (default) unless DB1_NO_PDL is set

// cudaLaunchKernelEx with (optionally) a cluster dimension and programmatic stream serialization.
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                     int cluster_x, Args&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  int n = 0;
  if (cluster_x > 1) {
    attr[n].id = cudaLaunchAttributeClusterDimension;
    attr[n].val.clusterDim.x = cluster_x;
    attr[n].val.clusterDim.y = 1;
    attr[n].val.clusterDim.z = 1;
    ++n;
  }
  if (pdl_enabled()) {
    attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  cfg.attrs = attr;
  cfg.numAttrs = n;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

static inline int cdiv(int a, int b) { return (a + b - 1) / b; }

// csrc/skinny.cu: the few-row (M <= 8) GEMM of the decode step; db1_gemm_f16 routes to it when it applies
bool skinny_gemm_applies(const db1_gemm_desc* d);
int skinny_gemm(const db1_gemm_desc* d, cudaStream_t stream);

}  // namespace db1
