// Few-row GEMM for the decode step: C[M,N] = epilogue(alpha * A[M,K] . B[N,K]^T) with M <= 8.
//
// With one to eight rows of activations (evaluate_rl.py:157-266 feeds one action token at a time) the product is a
// stream over the weight matrix B: 2*N*K bytes read once, M*N*K*2 FLOP - HBM-bound by a factor of > 100. The tensor-core
// kernel's 128-row tiles make such a call cost a full 128 x 256 x K tile per 256 output columns on N/256 CTA pairs
// (10-20 us whatever M is); here every warp of a persistent grid owns groups of four output columns, streams their
// weight rows with 16-byte non-allocating loads (sixteen of them in flight per lane), multiplies against the activation
// rows held in shared memory, reduces with shuffles and applies the same fused epilogues db1_gemm_f16 offers on this path (bias, residual, the QKV split with +u / +v, GeGLU).
// With db1_gemm_desc.b_static (weights nobody is writing: inference) the first batch of every warp's loads is issued
// BEFORE griddepcontrol.wait, so the weight stream of a launch starts under the previous kernel's tail.
// A second kernel stages the weight rows in shared memory with 1-D bulk copies (cp.async.bulk, a 3- to 6-stage ring of
// 32 KB filled by one producer thread). Measured on the decode step (profiles/README.md): 1.38 - 1.41 ms per step against
// 1.27 ms for the register-streaming kernel - a launch is a 10 - 35 MB stream, so its fixed cost (launch, fill, first
// round trip, drain) decides, and 128-thread CTAs with 4 KB of shared memory overlap consecutive launches better than
// one 100 - 200 KB CTA per SM. Kept as an opt-in: DB1_SKINNY_BULK=1.
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
  int b_static;  // weights: may be requested before griddepcontrol.wait
  const __half* ln_gamma;  // LayerNorm of the activation rows on load (NULL: off)
  const __half* ln_beta;
  float ln_eps;
  __half* ln_out;  // [M, K] normalised rows (optional)
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

// LayerNorm of the staged activation rows, in place (db1_gemm_desc.ln_*): all threads of the CTA on one row at a time,
// two-pass statistics and the fp16 rounding of db1_layernorm_fwd (csrc/elementwise.cu: ln_fwd_warp_kernel), every CTA for
// itself (M * K <= 8 * 8192 elements); the CTA with `write` also stores the rows to ln_out. The first two gamma / beta
// chunks of every thread (all of them for K <= 2048) arrive as arguments: they were requested before
// griddepcontrol.wait, LayerNorm weights being as static as the GEMM's own.
template <int NT>
DEVI float block_sum_nt(float v, float* red, int lane, int warp) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float t = 0.f;
#pragma unroll
  for (int w = 0; w < NT / 32; ++w) t += red[w];
  return t;
}

template <int NT>
DEVI void layernorm_rows_in_smem(uint4* sA, const SkinnyParams& p, int kc, int tid, bool write, const uint4 (&pg)[2],
                                 const uint4 (&pb)[2], float (*red)[NT / 32]) {
  const int lane = tid & 31, warp = tid >> 5;
  for (int m = 0; m < p.M; ++m) {
    uint4* row = sA + (size_t)m * kc;
    float sum = 0.f;
    for (int c = tid; c < kc; c += NT) {
      const uint4 a = row[c];
      const float2 x0 = unpack_half2(a.x), x1 = unpack_half2(a.y), x2 = unpack_half2(a.z), x3 = unpack_half2(a.w);
      sum += ((x0.x + x0.y) + (x1.x + x1.y)) + ((x2.x + x2.y) + (x3.x + x3.y));
    }
    const float mean = block_sum_nt<NT>(sum, red[0], lane, warp) / (float)p.K;
    float var = 0.f;
    for (int c = tid; c < kc; c += NT) {
      const uint4 a = row[c];
      const uint32_t w[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 x = unpack_half2(w[i]);
        var = fmaf(x.x - mean, x.x - mean, var);
        var = fmaf(x.y - mean, x.y - mean, var);
      }
    }
    const float rstd = rsqrtf(block_sum_nt<NT>(var, red[1], lane, warp) / (float)p.K + p.ln_eps);
    int it = 0;
    for (int c = tid; c < kc; c += NT, ++it) {
      const uint4 a = row[c];
      const uint4 g = it < 2 ? pg[it < 2 ? it : 0] : *reinterpret_cast<const uint4*>(p.ln_gamma + c * 8);
      const uint4 b = it < 2 ? pb[it < 2 ? it : 0] : *reinterpret_cast<const uint4*>(p.ln_beta + c * 8);
      const uint32_t wa[4] = {a.x, a.y, a.z, a.w}, wg[4] = {g.x, g.y, g.z, g.w}, wb[4] = {b.x, b.y, b.z, b.w};
      uint32_t wo[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 x = unpack_half2(wa[i]), gg = unpack_half2(wg[i]), bb = unpack_half2(wb[i]);
        wo[i] = pack_half2((x.x - mean) * rstd * gg.x + bb.x, (x.y - mean) * rstd * gg.y + bb.y);
      }
      const uint4 o4 = make_uint4(wo[0], wo[1], wo[2], wo[3]);
      row[c] = o4;
      if (write && p.ln_out != nullptr) *reinterpret_cast<uint4*>(p.ln_out + (size_t)m * p.K + c * 8) = o4;
    }
    __syncthreads();  // red[] is reused by the next row; the normalised row is visible to every warp
  }
}

// MT: compile-time bound on the activation rows (1, 2, 4, 8). NR weight rows per pass: SK_ROWS, twice that for GeGLU
// (column n of the a half and column n of the g half are finished by the same warp).
template <int MT, int EPI, int SK_ROWS>
__global__ void __launch_bounds__(SK_THREADS) skinny_gemm_kernel(const SkinnyParams p) {
  constexpr int NR = (EPI == DB1_EPI_GEGLU) ? 2 * SK_ROWS : SK_ROWS;
  extern __shared__ uint4 sA[];  // [MT][K / 8]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int kc = p.K >> 3;
  constexpr int U = 16 / NR;  // chunks per row and batch: U * NR = 16 independent 16-byte loads per lane in flight
  const int NO = (EPI == DB1_EPI_GEGLU) ? p.F : p.N;  // output column groups run over [0, NO)
  const int ngroups = (NO + SK_ROWS - 1) / SK_ROWS;
  const int nwarps = gridDim.x * (SK_THREADS / 32);
  const int g_first = blockIdx.x * (SK_THREADS / 32) + warp;
  pdl_launch_dependents();
  // b_static (weights nobody is writing): the first batch of the warp's first column group - for K <= 4096 all of it - is
  // requested BEFORE waiting for the previous kernel, so the weight stream runs under that kernel's tail
  uint4 pre[U][NR];
  const bool preloaded = p.b_static && g_first < ngroups && lane + 32 * (U - 1) < kc;
  if (preloaded) {
#pragma unroll
    for (int r = 0; r < NR; ++r) {
      int n = g_first * SK_ROWS + (r % SK_ROWS);
      if (n >= NO) n = NO - 1;
      if (EPI == DB1_EPI_GEGLU && r >= SK_ROWS) n += p.F;
      const __half* row = p.B + (long long)n * p.ldb;
#pragma unroll
      for (int x = 0; x < U; ++x) pre[x][r] = ldg_stream(row + (lane + 32 * x) * 8);
    }
  }
  __shared__ float ln_red[2][SK_THREADS / 32];
  uint4 pg[2], pb[2];  // LayerNorm weights of this thread's first two chunks (static like B: requested before the wait)
  if (p.ln_gamma != nullptr) {
#pragma unroll
    for (int it = 0; it < 2; ++it) {
      const int c = tid + it * SK_THREADS;
      pg[it] = pb[it] = make_uint4(0u, 0u, 0u, 0u);
      if (c < kc) {
        pg[it] = *reinterpret_cast<const uint4*>(p.ln_gamma + c * 8);
        pb[it] = *reinterpret_cast<const uint4*>(p.ln_beta + c * 8);
      }
    }
  }
  pdl_wait();
  for (int idx = tid; idx < MT * kc; idx += SK_THREADS) {
    const int m = idx / kc, c = idx - m * kc;
    sA[idx] = (m < p.M) ? *reinterpret_cast<const uint4*>(p.A + (long long)m * p.lda + c * 8) : make_uint4(0u, 0u, 0u, 0u);
  }
  __syncthreads();
  if (p.ln_gamma != nullptr) layernorm_rows_in_smem<SK_THREADS>(sA, p, kc, tid, blockIdx.x == 0, pg, pb, ln_red);
  for (int g = g_first; g < ngroups; g += nwarps) {
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
    int c = lane;
    for (; c + 32 * (U - 1) < kc; c += 32 * U) {
      uint4 bb[U][NR];
      if (preloaded && g == g_first && c == lane) {
#pragma unroll
        for (int x = 0; x < U; ++x)
#pragma unroll
          for (int r = 0; r < NR; ++r) bb[x][r] = pre[x][r];
      } else {
#pragma unroll
        for (int x = 0; x < U; ++x)
#pragma unroll
          for (int r = 0; r < NR; ++r) bb[x][r] = ldg_stream(brow[r] + (c + 32 * x) * 8);
      }
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

// ------------------------------------------------------------------------------------------------------------------
// Bulk-copy staged variant. A stage holds RB weight rows of K halves (GeGLU: RB/2 rows of the a half followed by the
// matching RB/2 rows of the g half). Warp 4 = producer (one thread), warps 0-3 = consumers: warp w finishes the output
// columns w, w + 4, ... of the stage (lanes stride over the 16-byte chunks of K, activations from shared memory).
// ------------------------------------------------------------------------------------------------------------------
constexpr int SKB_CONSUMERS = 128;
constexpr int SKB_THREADS = SKB_CONSUMERS + 32;
constexpr int SKB_MAX_STAGES = 6;

DEVI void bulk_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

struct SkinnyBulkCfg {
  int rb;       // weight rows per stage (GeGLU: a rows + g rows)
  int stages;
  int a_bytes;  // MT * K * 2, rounded up to 128
  int stage_bytes;
};

template <int MT, int EPI>
__global__ void __launch_bounds__(SKB_THREADS) skinny_bulk_kernel(const SkinnyParams p, const SkinnyBulkCfg cfg) {
  extern __shared__ __align__(128) uint8_t sk_smem[];
  uint4* sA = reinterpret_cast<uint4*>(sk_smem);
  uint8_t* stage0 = sk_smem + cfg.a_bytes;
  uint64_t* full = reinterpret_cast<uint64_t*>(stage0 + (size_t)cfg.stages * cfg.stage_bytes);
  uint64_t* empty = full + SKB_MAX_STAGES;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int kc = p.K >> 3;
  const uint32_t row_bytes = (uint32_t)p.K * 2u;
  const int NO = (EPI == DB1_EPI_GEGLU) ? p.F : p.N;            // output columns
  const int cols = (EPI == DB1_EPI_GEGLU) ? cfg.rb / 2 : cfg.rb;  // output columns per stage
  const int nblocks = (NO + cols - 1) / cols;
  if (tid == 0) {
    for (int i = 0; i < cfg.stages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], SKB_CONSUMERS / 32);
    }
    mbar_fence_init();
  }
  __syncthreads();
  pdl_launch_dependents();
  if (warp == SKB_CONSUMERS / 32) {
    // ---- producer: with b_static the weights are parameters, not outputs of the previous kernel, and are requested
    // without waiting for it (their stream overlaps its tail)
    if (lane == 0) {
      if (!p.b_static) pdl_wait();
      int s = 0;
      uint32_t ph = 0;
      for (int blk = blockIdx.x; blk < nblocks; blk += gridDim.x) {
        mbar_wait(&empty[s], ph ^ 1);
        const int n0 = blk * cols;
        const int nv = (NO - n0 < cols) ? NO - n0 : cols;
        mbar_expect_tx(&full[s], (uint32_t)nv * row_bytes * (EPI == DB1_EPI_GEGLU ? 2u : 1u));
        uint8_t* dst = stage0 + (size_t)s * cfg.stage_bytes;
        for (int r = 0; r < nv; ++r) {
          bulk_load_1d(dst + (size_t)r * row_bytes, p.B + (long long)(n0 + r) * p.ldb, row_bytes, &full[s]);
          if (EPI == DB1_EPI_GEGLU)
            bulk_load_1d(dst + (size_t)(cols + r) * row_bytes, p.B + (long long)(p.F + n0 + r) * p.ldb, row_bytes, &full[s]);
        }
        if (++s == cfg.stages) {
          s = 0;
          ph ^= 1;
        }
      }
    }
    return;
  }
  // ---- consumers
  pdl_wait();  // the activations come from the previous kernel
  for (int idx = tid; idx < MT * kc; idx += SKB_CONSUMERS) {
    const int m = idx / kc, c = idx - m * kc;
    sA[idx] = (m < p.M) ? *reinterpret_cast<const uint4*>(p.A + (long long)m * p.lda + c * 8) : make_uint4(0u, 0u, 0u, 0u);
  }
  named_bar_sync(1, SKB_CONSUMERS);
  int s = 0;
  uint32_t ph = 0;
  for (int blk = blockIdx.x; blk < nblocks; blk += gridDim.x) {
    mbar_wait(&full[s], ph);
    const uint8_t* st = stage0 + (size_t)s * cfg.stage_bytes;
    const int n0 = blk * cols;
    for (int r = warp; r < cols && n0 + r < NO; r += SKB_CONSUMERS / 32) {
      const uint4* ba = reinterpret_cast<const uint4*>(st + (size_t)r * row_bytes);
      const uint4* bg = reinterpret_cast<const uint4*>(st + (size_t)(cols + r) * row_bytes);
      float acc[MT], accg[MT];
#pragma unroll
      for (int m = 0; m < MT; ++m) acc[m] = accg[m] = 0.f;
#pragma unroll 4
      for (int c = lane; c < kc; c += 32) {
        const uint4 b0 = ba[c];
        uint4 b1 = make_uint4(0u, 0u, 0u, 0u);
        if (EPI == DB1_EPI_GEGLU) b1 = bg[c];
#pragma unroll
        for (int m = 0; m < MT; ++m) {
          const uint4 a0 = sA[m * kc + c];
          acc[m] = dot8(a0, b0, acc[m]);
          if (EPI == DB1_EPI_GEGLU) accg[m] = dot8(a0, b1, accg[m]);
        }
      }
#pragma unroll
      for (int m = 0; m < MT; ++m) {
#pragma unroll
        for (int o = 16; o; o >>= 1) {
          acc[m] += __shfl_xor_sync(0xffffffffu, acc[m], o);
          if (EPI == DB1_EPI_GEGLU) accg[m] += __shfl_xor_sync(0xffffffffu, accg[m], o);
        }
      }
      float mine = 0.f, mine_g = 0.f;
#pragma unroll
      for (int m = 0; m < MT; ++m)
        if (lane == m) {
          mine = acc[m] * p.alpha;
          mine_g = accg[m] * p.alpha;
        }
      const int m = lane, n = n0 + r;
      if (m < p.M) {
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
          const __half ha = __float2half_rn(a), hg = __float2half_rn(gg);
          if (p.H) {
            p.H[(long long)m * p.ldh + n] = ha;
            p.H[(long long)m * p.ldh + p.F + n] = hg;
          }
          p.C[(long long)m * p.ldc + n] = __float2half_rn(__half2float(ha) * gelu_erf(__half2float(hg)));
        }
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[s]);
    if (++s == cfg.stages) {
      s = 0;
      ph ^= 1;
    }
  }
}

// false: the shape does not fit (the caller falls back to the register-streaming kernel)
template <int MT, int EPI>
static bool launch_skinny_bulk(const SkinnyParams& p, cudaStream_t stream, int* rc) {
  static int on = -1;
  if (on < 0) on = getenv("DB1_SKINNY_BULK") ? 1 : 0;
  if (!on || p.ln_gamma != nullptr) return false;
  const long long row_bytes = (long long)p.K * 2;
  if (row_bytes % 16 != 0 || (p.ldb * 2) % 16 != 0 || row_bytes > 32768) return false;
  SkinnyBulkCfg cfg;
  cfg.a_bytes = (int)(((long long)MT * row_bytes + 127) / 128 * 128);
  int rb = (int)(32768 / row_bytes);  // rows per 32 KB stage
  if (EPI == DB1_EPI_GEGLU) rb &= ~1;
  if (rb > 32) rb = 32;
  if (rb < (EPI == DB1_EPI_GEGLU ? 2 : 1)) return false;
  cfg.rb = rb;
  cfg.stage_bytes = (int)(rb * row_bytes);
  // at most ~half of an SM's shared memory: the next launch's CTAs (programmatic dependent launch) must fit beside this
  // one's for its weight stream to start under this kernel's tail
  static int half = -1;
  if (half < 0) half = getenv("DB1_SKINNY_FULL_SMEM") ? 0 : 1;
  const int budget = (half ? 110 : 227) * 1024 - cfg.a_bytes - 2 * SKB_MAX_STAGES * 8 - 128;
  int stages = budget / cfg.stage_bytes;
  if (stages > SKB_MAX_STAGES) stages = SKB_MAX_STAGES;
  if (stages < 2) return false;
  cfg.stages = stages;
  const size_t smem = (size_t)cfg.a_bytes + (size_t)stages * cfg.stage_bytes + 2 * SKB_MAX_STAGES * 8;
  static size_t configured = 0;
  if (smem > configured) {
    if (cudaFuncSetAttribute(skinny_bulk_kernel<MT, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
      cudaGetLastError();
      return false;
    }
    configured = smem;
  }
  const int NO = (EPI == DB1_EPI_GEGLU) ? p.F : p.N;
  const int cols = (EPI == DB1_EPI_GEGLU) ? rb / 2 : rb;
  const int nblocks = cdiv(NO, cols);
  int grid = nblocks < sm_count() ? nblocks : sm_count();
  cudaError_t e = launch_pdl(skinny_bulk_kernel<MT, EPI>, dim3(grid), dim3(SKB_THREADS), smem, stream, 1, p, cfg);
  *rc = (e == cudaSuccess) ? 0 : set_err((int)e, "skinny_bulk_kernel launch failed: %s", cudaGetErrorString(e));
  return true;
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
  int rc = 0;
  if (launch_skinny_bulk<MT, EPI>(p, stream, &rc)) return rc;
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
  p.M = d->M; p.N = d->N; p.K = d->K; p.F = d->F; p.d_model = d->d_model; p.alpha = d->alpha; p.b_static = d->b_static;
  p.ln_gamma = (const __half*)d->ln_gamma; p.ln_beta = (const __half*)d->ln_beta; p.ln_eps = d->ln_eps; p.ln_out = (__half*)d->ln_out;
  if (p.ln_gamma != nullptr) {
    DB1_CHECK_ARG(p.ln_beta != nullptr && ((((uintptr_t)p.ln_gamma) | ((uintptr_t)p.ln_beta) | ((uintptr_t)p.ln_out)) & 15) == 0,
                  "gemm(ln): ln_beta missing or ln_gamma / ln_beta / ln_out not 16-byte aligned");
  }
  switch (d->epilogue) {
    case DB1_EPI_PLAIN: return launch_skinny_e<DB1_EPI_PLAIN>(p, stream);
    case DB1_EPI_QKV: return launch_skinny_e<DB1_EPI_QKV>(p, stream);
    case DB1_EPI_GEGLU: return launch_skinny_e<DB1_EPI_GEGLU>(p, stream);
  }
  return set_err(-1, "skinny_gemm: unsupported epilogue");
}

}  // namespace db1
