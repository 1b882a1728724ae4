// Parameters of the persistent tcgen05 GEMM (see gemm.cu).
#pragma once
#include <cuda_fp16.h>
#include <stdint.h>

namespace db1 {

enum GemmEpilogue {
  EPI_PLAIN = 0,   // C = [resid +] dropout(alpha*acc [+ bias]) [+ C]
  EPI_QKV = 1,     // acc columns < d_model are written twice (+u, +v); output is [Qu | Qv | K | V]
  EPI_GEGLU = 2,   // acc = [a | g] (tile pairs column n with n+F); writes H = [a|g] + b1 and C = a*gelu(g)
  EPI_DGEGLU = 3,  // acc = dY; reads H = [a|g]; writes C = [dY*gelu(g) | dY*a*gelu'(g)]
};

struct GemmParams {
  int M, N, K;  // C[M,N] = A[M,K] * B[N,K]^T in MMA terms; for EPI_GEGLU N = 2F rows of B are consumed
  int a_mn, b_mn;  // 1: operand is stored MN-contiguous ([K][MN] in memory) instead of K-contiguous
  float alpha;
  __half* C;
  int ldc;
  const __half* bias;   // [N] or null
  const __half* resid;  // [M, ldr] or null
  int ldr;
  int accumulate;  // C += (gradient accumulation)
  uint32_t drop_thr16;  // dropout threshold in 1/65536 units; 0 = no dropout
  float drop_scale;
  uint64_t seed;
  const __half* u;  // EPI_QKV: r_w_bias flattened [d_model]
  const __half* v;  // EPI_QKV: r_r_bias flattened [d_model]
  int d_model;
  __half* H;  // EPI_GEGLU: out [M, ldh] ; EPI_DGEGLU: in
  int ldh;
  int F;  // GeGLU half width
};

}  // namespace db1
