// Few-row GEMM for the decode step: C[M,N] = epilogue(alpha * A[M,K] . B[N,K]^T) with M <= 8.
//
// With one to eight rows of activations (evaluate_rl.py:157-266 feeds one action token at a time) the product is a
// stream over the weight matrix B: 2*N*K bytes read once, M*N*K*2 FLOP - HBM-bound by a factor of > 100. The tensor-core
// kernel's 128-row tiles make such a call cost a full 128 x 256 x K tile per 256 output columns on N/256 CTA pairs
// (10-20 us whatever M is); here every warp of a persistent grid owns groups of four output columns, streams their
// weight rows with 16-byte non-allocating loads (sixteen of them in flight per lane), multiplies against the activation
// rows held in shared memory, reduces with shuffles and applies the same fused epilogues db1_gemm_f16 offers on this path (bias, residual, the QKV split with +u / +v, GeGLU).
// db1_gemm_f16 routes here by itself (include/db1_sm100.h); DB1_NO_SKINNY=1 keeps the tensor-core kernel.
#include <stdlib.h>

#include "../../include/db1_sm100.h"
#include "common.cuh"
#include "ptx.cuh"

namespace db1 {

constexpr int SK_THREADS = 128;  // 4 warps: small N (2048 columns = 512 groups) still spreads over 128 CTAs
// Output columns (weight rows) per warp pass: R in {1, 2, 4}, chosen per launch so that the 16 loads a lane keeps in
// flight cover as much of K as possible in ONE round trip (K = 2048: 8 chunks per lane and row -> R = 2; K >= 4096: R = 1).

struct SkinnyParams {
  const __half* A;
  long long lda;
  const __half* B;
  long long ldb;
  __half* C;
  long long ldc;
  const __half* bias;
  const __half* resid;
  long long ldr;
  const __half* u;
  const __half* v;
  __half* H;
  long long ldh;
  int M, N, K, F, d_model;
  float alpha;
};

DEVI uint4 ldg_stream(const __half* p) {
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "l"(p));
  return v;
}

DEVI float dot8(const uint4& a, const uint4& b, float acc) {
  const float2 a0 = unpack_half2(a.x), a1 = unpack_half2(a.y), a2 = unpack_half2(a.z), a3 = unpack_half2(a.w);
  const float2 b0 = unpack_half2(b.x), b1 = unpack_half2(b.y), b2 = unpack_half2(b.z), b3 = unpack_half2(b.w);
  acc = fmaf(a0.x, b0.x, acc); acc = fmaf(a0.y, b0.y, acc);
  acc = fmaf(a1.x, b1.x, acc); acc = fmaf(a1.y, b1.y, acc);
  acc = fmaf(a2.x, b2.x, acc); acc = fmaf(a2.y, b2.y, acc);
  acc = fmaf(a3.x, b3.x, acc); acc = fmaf(a3.y, b3.y, acc);
  return acc;
}

// MT: compile-time bound on the activation rows (1, 2, 4, 8). NR weight rows per pass: SK_ROWS, twice that for GeGLU
// (column n of the a half and column n of the g half are finished by the same warp).
template <int MT, int EPI, int SK_ROWS>
__global__ void __launch_bounds__(SK_THREADS) skinny_gemm_kernel(const SkinnyParams p) {
  constexpr int NR = (EPI == DB1_EPI_GEGLU) ? 2 * SK_ROWS : SK_ROWS;
  extern __shared__ uint4 sA[];  // [MT][K / 8]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int kc = p.K >> 3;
  pdl_launch_dependents();
  pdl_wait();
  for (int idx = tid; idx < MT * kc; idx += SK_THREADS) {
    const int m = idx / kc, c = idx - m * kc;
    sA[idx] = (m < p.M) ? *reinterpret_cast<const uint4*>(p.A + (long long)m * p.lda + c * 8) : make_uint4(0u, 0u, 0u, 0u);
  }
  __syncthreads();
  const int NO = (EPI == DB1_EPI_GEGLU) ? p.F : p.N;  // output column groups run over [0, NO)
  const int ngroups = (NO + SK_ROWS - 1) / SK_ROWS;
  const int nwarps = gridDim.x * (SK_THREADS / 32);
  for (int g = blockIdx.x * (SK_THREADS / 32) + warp; g < ngroups; g += nwarps) {
    const int n0 = g * SK_ROWS;
    const __half* brow[NR];
#pragma unroll
    for (int r = 0; r < NR; ++r) {
      int n = n0 + (r % SK_ROWS);
      if (n >= NO) n = NO - 1;  // clamped duplicate: computed, never written
      if (EPI == DB1_EPI_GEGLU && r >= SK_ROWS) n += p.F;
      brow[r] = p.B + (long long)n * p.ldb;
    }
    float acc[NR][MT];
#pragma unroll
    for (int r = 0; r < NR; ++r)
#pragma unroll
      for (int m = 0; m < MT; ++m) acc[r][m] = 0.f;
    // U chunks per row and iteration: U * NR = 16 independent 16-byte loads per lane in flight
    constexpr int U = 16 / NR;
    int c = lane;
    for (; c + 32 * (U - 1) < kc; c += 32 * U) {
      uint4 bb[U][NR];
#pragma unroll
      for (int x = 0; x < U; ++x)
#pragma unroll
        for (int r = 0; r < NR; ++r) bb[x][r] = ldg_stream(brow[r] + (c + 32 * x) * 8);
#pragma unroll
      for (int x = 0; x < U; ++x)
#pragma unroll
        for (int m = 0; m < MT; ++m) {
          const uint4 a0 = sA[m * kc + c + 32 * x];
#pragma unroll
          for (int r = 0; r < NR; ++r) acc[r][m] = dot8(a0, bb[x][r], acc[r][m]);
        }
    }
    for (; c < kc; c += 32) {
      uint4 b0[NR];
#pragma unroll
      for (int r = 0; r < NR; ++r) b0[r] = ldg_stream(brow[r] + c * 8);
#pragma unroll
      for (int m = 0; m < MT; ++m) {
        const uint4 a0 = sA[m * kc + c];
#pragma unroll
        for (int r = 0; r < NR; ++r) acc[r][m] = dot8(a0, b0[r], acc[r][m]);
      }
    }
#pragma unroll
    for (int r = 0; r < NR; ++r)
#pragma unroll
      for (int m = 0; m < MT; ++m) {
        float v = acc[r][m];
#pragma unroll
        for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        acc[r][m] = v * p.alpha;
      }
    // lane l < SK_ROWS * MT finishes output (row m = l % MT, column n0 + l / MT)
    float mine = 0.f, mine_g = 0.f;
#pragma unroll
    for (int r = 0; r < SK_ROWS; ++r)
#pragma unroll
      for (int m = 0; m < MT; ++m)
        if (lane == r * MT + m) {
          mine = acc[r][m];
          if (EPI == DB1_EPI_GEGLU) mine_g = acc[(EPI == DB1_EPI_GEGLU) ? r + SK_ROWS : r][m];
        }
    const int m = lane % MT, n = n0 + lane / MT;
    if (lane < SK_ROWS * MT && m < p.M && n < NO) {
      if (EPI == DB1_EPI_PLAIN) {
        float f = mine;
        if (p.bias) f += __half2float(p.bias[n]);
        if (p.resid) f = __half2float(p.resid[(long long)m * p.ldr + n]) + f;
        p.C[(long long)m * p.ldc + n] = __float2half_rn(f);
      } else if (EPI == DB1_EPI_QKV) {
        __half* crow = p.C + (long long)m * p.ldc;
        if (n < p.d_model) {
          crow[n] = __float2half_rn(mine + __half2float(p.u[n]));
          crow[p.d_model + n] = __float2half_rn(mine + __half2float(p.v[n]));
        } else {
          crow[p.d_model + n] = __float2half_rn(mine);
        }
      } else {
        float a = mine, gg = mine_g;
        if (p.bias) {
          a += __half2float(p.bias[n]);
          gg += __half2float(p.bias[p.F + n]);
        }
        const __half ha = __float2half_rn(a), hg = __float2half_rn(gg);  // as the tensor-core epilogue: fp16 pre-activations
        if (p.H) {
          p.H[(long long)m * p.ldh + n] = ha;
          p.H[(long long)m * p.ldh + p.F + n] = hg;
        }
        p.C[(long long)m * p.ldc + n] = __float2half_rn(__half2float(ha) * gelu_erf(__half2float(hg)));
      }
    }
  }
}

template <int MT, int EPI, int R>
static int launch_skinny_mer(const SkinnyParams& p, cudaStream_t stream) {
  const size_t smem = (size_t)MT * p.K * 2;
  static size_t configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    DB1_CUDA(cudaFuncSetAttribute(skinny_gemm_kernel<MT, EPI, R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  const int NO = (EPI == DB1_EPI_GEGLU) ? p.F : p.N;
  const int ngroups = cdiv(NO, R);
  int grid = cdiv(ngroups, SK_THREADS / 32);
  const int cap = 4 * sm_count();  // persistent: every warp walks several column groups (amortises the A fill)
  if (grid > cap) grid = cap;
  DB1_CUDA(launch_pdl(skinny_gemm_kernel<MT, EPI, R>, dim3(grid), dim3(SK_THREADS), smem, stream, 1, p));
  return 0;
}

template <int MT, int EPI>
static int launch_skinny_me(const SkinnyParams& p, cudaStream_t stream) {
  // chunks per lane and weight row: K / 256. With 16 loads in flight per lane, R * (GeGLU ? 2 : 1) * chunks <= 16 keeps a
  // column group to one memory round trip.
  const int per_lane = cdiv(p.K, 256);
  const int rows16 = 16 / ((EPI == DB1_EPI_GEGLU ? 2 : 1) * (per_lane < 1 ? 1 : per_lane));
  if (rows16 >= 4) return launch_skinny_mer<MT, EPI, 4>(p, stream);
  if (rows16 >= 2) return launch_skinny_mer<MT, EPI, 2>(p, stream);
  return launch_skinny_mer<MT, EPI, 1>(p, stream);
}

template <int EPI>
static int launch_skinny_e(const SkinnyParams& p, cudaStream_t stream) {
  if (p.M <= 1) return launch_skinny_me<1, EPI>(p, stream);
  if (p.M <= 2) return launch_skinny_me<2, EPI>(p, stream);
  if (p.M <= 4) return launch_skinny_me<4, EPI>(p, stream);
  return launch_skinny_me<8, EPI>(p, stream);
}

bool skinny_gemm_applies(const db1_gemm_desc* d) {
  static int off = -1;
  if (off < 0) off = getenv("DB1_NO_SKINNY") ? 1 : 0;
  if (off) return false;
  const bool batched = d->Z1 > 1 || d->Z2 > 1;
  if (d->M > 8 || d->M < 1 || batched || d->a_mn || d->b_mn || d->k_mode != DB1_K_FULL || d->skip_upper) return false;
  if (d->accumulate || d->drop_p != 0.f || d->dot_out || d->dot_with) return false;
  if (d->K % 8 != 0 || d->lda % 8 != 0 || d->ldb % 8 != 0 || (size_t)8 * d->K * 2 > 200 * 1024) return false;
  if ((((uintptr_t)d->A) | ((uintptr_t)d->B)) & 15) return false;
  if (d->epilogue == DB1_EPI_PLAIN) return true;
  if (d->epilogue == DB1_EPI_QKV) return d->u && d->v && d->d_model > 0 && d->N == 3 * d->d_model && !d->bias && !d->resid;
  if (d->epilogue == DB1_EPI_GEGLU) return d->F > 0 && d->N == 2 * d->F && !d->resid;
  return false;
}

int skinny_gemm(const db1_gemm_desc* d, cudaStream_t stream) {
  SkinnyParams p;
  p.A = (const __half*)d->A; p.lda = d->lda; p.B = (const __half*)d->B; p.ldb = d->ldb;
  p.C = (__half*)d->C; p.ldc = d->ldc; p.bias = (const __half*)d->bias; p.resid = (const __half*)d->resid; p.ldr = d->ldr;
  p.u = (const __half*)d->u; p.v = (const __half*)d->v; p.H = (__half*)d->H; p.ldh = d->ldh;
  p.M = d->M; p.N = d->N; p.K = d->K; p.F = d->F; p.d_model = d->d_model; p.alpha = d->alpha;
  switch (d->epilogue) {
    case DB1_EPI_PLAIN: return launch_skinny_e<DB1_EPI_PLAIN>(p, stream);
    case DB1_EPI_QKV: return launch_skinny_e<DB1_EPI_QKV>(p, stream);
    case DB1_EPI_GEGLU: return launch_skinny_e<DB1_EPI_GEGLU>(p, stream);
  }
  return set_err(-1, "skinny_gemm: unsupported epilogue");
}

}  // namespace db1
