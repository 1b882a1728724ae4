// Persistent warp-specialised tcgen05 GEMM for sm_100a.
//
//   C[M,N] = epilogue( A[M,K] * B[N,K]^T ),  fp16 operands, fp32 accumulation in TMEM.
//
// One CTA per SM walks output tiles (128 x BN) round-robin. Roles:
//   warp 0     TMA producer : global -> smem ring (STAGES x [A 128x64 | B BNx64], 128B swizzle)
//   warp 1     MMA issuer   : one elected thread issues tcgen05.mma (M=128, N=BN, K=16), accumulators in TMEM;
//                             the accumulator is double-buffered (2 x BN columns) so tile i+1's main loop
//                             overlaps tile i's epilogue
//   warps 2..5 epilogue     : tcgen05.ld -> registers -> fused epilogue -> 16-byte global stores
//
// Either operand may be K-contiguous ("K-major", e.g. activations x weights^T in the forward pass) or
// MN-contiguous ("MN-major": weights in dgrad, both operands in wgrad) - the UMMA descriptors transpose for free,
// so dgrad/wgrad never materialise a transposed copy.
//
// Replaces the cuBLAS calls behind nn.Linear / F.linear in the reference
// (src/model/transformer_xl.py:138-139, 228, 265-268, 595) and their autograd backward.
#include "common.cuh"
#include "gemm.cuh"
#include "ptx.cuh"

namespace db1 {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int GEMM_THREADS = 192;
constexpr int A_STAGE_BYTES = BM * BK * 2;  // 16 KB

template <int BN>
struct GemmCfg {
  static constexpr int STAGES = (BN == 256) ? 4 : 6;
  static constexpr int B_STAGE_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
  static constexpr int TMEM_COLS = 2 * BN;  // 512 or 256
};

DEVI float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }
DEVI float gelu_erf_grad(float x) {
  return 0.5f * (1.0f + erff(x * 0.70710678118654752f)) + x * 0.3989422804014327f * __expf(-0.5f * x * x);
}

struct alignas(16) Half8 {
  __half2 h[4];
};
DEVI Half8 ld_half8(const __half* p) { return *reinterpret_cast<const Half8*>(p); }
DEVI void st_half8(__half* p, const Half8& v) { *reinterpret_cast<Half8*>(p) = v; }
DEVI void half8_to_float(const Half8& v, float (&f)[8]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 t = __half22float2(v.h[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
DEVI Half8 float_to_half8(const float (&f)[8]) {
  Half8 v;
#pragma unroll
  for (int i = 0; i < 4; ++i) v.h[i] = __floats2half2_rn(f[2 * i], f[2 * i + 1]);
  return v;
}

template <int BN, int EPI>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmParams p) {
  using Cfg = GemmCfg<BN>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
  uint64_t* full = bars;
  uint64_t* empty = bars + STAGES;
  uint64_t* tfull = bars + 2 * STAGES;
  uint64_t* tempty = bars + 2 * STAGES + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int MT = (p.M + BM - 1) / BM;
  const int NT = (EPI == EPI_GEGLU) ? (p.F / (BN / 2)) : (p.N + BN - 1) / BN;
  const int num_tiles = MT * NT;
  const int KB = (p.K + BK - 1) / BK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int i = 0; i < STAGES; ++i) {
        mbar_init(&full[i], 1);
        mbar_init(&empty[i], 1);
      }
      for (int i = 0; i < 2; ++i) {
        mbar_init(&tfull[i], 1);
        mbar_init(&tempty[i], 4);
      }
      mbar_fence_init();
    }
    __syncwarp();
    tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int mt = tile % MT, nt = tile / MT;
        const int m0 = mt * BM;
        for (int kb = 0; kb < KB; ++kb) {
          mbar_wait(&empty[s], ph ^ 1);
          uint8_t* sa = smem + s * Cfg::STAGE_BYTES;
          uint8_t* sb = sa + A_STAGE_BYTES;
          mbar_expect_tx(&full[s], Cfg::STAGE_BYTES);
          const int k0 = kb * BK;
          if (!p.a_mn) {
            tma_load_2d(sa, &tmA, &full[s], k0, m0);
          } else {
            tma_load_2d(sa, &tmA, &full[s], m0, k0);
            tma_load_2d(sa + 8192, &tmA, &full[s], m0 + 64, k0);
          }
          if (!p.b_mn) {
#pragma unroll
            for (int j = 0; j < BN / 128; ++j) {
              int row0;
              if (EPI == EPI_GEGLU) row0 = j * p.F + nt * (BN / 2);
              else row0 = nt * BN + j * 128;
              tma_load_2d(sb + j * 16384, &tmB, &full[s], k0, row0);
            }
          } else {
#pragma unroll
            for (int j = 0; j < BN / 64; ++j) tma_load_2d(sb + j * 8192, &tmB, &full[s], nt * BN + j * 64, k0);
          }
          if (++s == STAGES) {
            s = 0;
            ph ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      const uint32_t idesc = umma_idesc(BM, BN, p.a_mn, p.b_mn, 0);
      const uint32_t a_lbo = p.a_mn ? 8192u : 16u, b_lbo = p.b_mn ? 8192u : 16u;
      const uint32_t a_kadv = p.a_mn ? (2048u >> 4) : (32u >> 4);  // descriptor start-address units of 16 B
      const uint32_t b_kadv = p.b_mn ? (2048u >> 4) : (32u >> 4);
      int s = 0;
      uint32_t ph = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        const int as = it & 1;
        const uint32_t aph = (it >> 1) & 1;
        mbar_wait(&tempty[as], aph ^ 1);
        tc_fence_after();
        const uint32_t tacc = tmem_base + as * BN;
        for (int kb = 0; kb < KB; ++kb) {
          mbar_wait(&full[s], ph);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + s * Cfg::STAGE_BYTES);
          const uint64_t adesc = umma_smem_desc(sa, a_lbo, 1024);
          const uint64_t bdesc = umma_smem_desc(sa + A_STAGE_BYTES, b_lbo, 1024);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k)
            umma_ss(tacc, adesc + (uint64_t)(k * a_kadv), bdesc + (uint64_t)(k * b_kadv), idesc, (kb | k) ? 1u : 0u);
          umma_commit(&empty[s]);
          if (++s == STAGES) {
            s = 0;
            ph ^= 1;
          }
        }
        umma_commit(&tfull[as]);
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue warps (2..5)
    const int quad = warp & 3;  // TMEM lane quadrant this warp may access
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int mt = tile % MT, nt = tile / MT;
      const int as = it & 1;
      const uint32_t aph = (it >> 1) & 1;
      const int row = mt * BM + quad * 32 + lane;
      const bool row_ok = row < p.M;
      mbar_wait(&tfull[as], aph);
      tc_fence_after();
      const uint32_t tacc = tmem_base + as * BN + ((uint32_t)(quad * 32) << 16);

      if (EPI == EPI_PLAIN || EPI == EPI_QKV) {
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
          uint32_t r[32];
          tmem_ld32(tacc + c * 32, r);
          tmem_ld_wait();
          const int col0 = nt * BN + c * 32;
          if (row_ok) {
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              const int col = col0 + g * 8;
              if (col >= p.N) break;
              float f[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) f[i] = __uint_as_float(r[g * 8 + i]) * p.alpha;
              if (EPI == EPI_QKV) {
                const size_t orow = (size_t)row * p.ldc;
                if (col < p.d_model) {
                  float uu[8], vv[8], o[8];
                  half8_to_float(ld_half8(p.u + col), uu);
                  half8_to_float(ld_half8(p.v + col), vv);
#pragma unroll
                  for (int i = 0; i < 8; ++i) o[i] = f[i] + uu[i];
                  st_half8(p.C + orow + col, float_to_half8(o));
#pragma unroll
                  for (int i = 0; i < 8; ++i) o[i] = f[i] + vv[i];
                  st_half8(p.C + orow + p.d_model + col, float_to_half8(o));
                } else {
                  st_half8(p.C + orow + p.d_model + col, float_to_half8(f));
                }
              } else {
                if (p.bias) {
                  float b[8];
                  half8_to_float(ld_half8(p.bias + col), b);
#pragma unroll
                  for (int i = 0; i < 8; ++i) f[i] += b[i];
                }
                if (p.drop_thr16) {
                  const uint64_t e = (uint64_t)row * (uint64_t)p.N + (uint64_t)col;
                  const uint64_t b0 = rng64(p.seed, e >> 2), b1 = rng64(p.seed, (e >> 2) + 1);
#pragma unroll
                  for (int i = 0; i < 4; ++i) {
                    f[i] = dropout_keep(b0, i, p.drop_thr16) ? f[i] * p.drop_scale : 0.f;
                    f[4 + i] = dropout_keep(b1, i, p.drop_thr16) ? f[4 + i] * p.drop_scale : 0.f;
                  }
                }
                if (p.resid) {
                  float b[8];
                  half8_to_float(ld_half8(p.resid + (size_t)row * p.ldr + col), b);
#pragma unroll
                  for (int i = 0; i < 8; ++i) f[i] += b[i];
                }
                __half* dst = p.C + (size_t)row * p.ldc + col;
                if (p.accumulate) {
                  float b[8];
                  half8_to_float(ld_half8(dst), b);
#pragma unroll
                  for (int i = 0; i < 8; ++i) f[i] += b[i];
                }
                st_half8(dst, float_to_half8(f));
              }
            }
          }
        }
      } else {
        // GeGLU forward / backward: accumulator columns [0,BN/2) pair with [BN/2,BN) (forward) or the tile's
        // BN output columns pair with saved a|g (backward).
        constexpr int HALF = BN / 2;
#pragma unroll 1
        for (int c = 0; c < (EPI == EPI_GEGLU ? HALF : BN) / 32; ++c) {
          uint32_t ra[32];
          tmem_ld32(tacc + c * 32, ra);
          if (EPI == EPI_GEGLU) {
            uint32_t rg[32];
            tmem_ld32(tacc + HALF + c * 32, rg);
            tmem_ld_wait();
            if (row_ok) {
              const int n0 = nt * HALF + c * 32;  // column within [0,F)
#pragma unroll
              for (int g = 0; g < 4; ++g) {
                const int n = n0 + g * 8;
                float a[8], gg[8], y[8], b[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                  a[i] = __uint_as_float(ra[g * 8 + i]);
                  gg[i] = __uint_as_float(rg[g * 8 + i]);
                }
                if (p.bias) {
                  half8_to_float(ld_half8(p.bias + n), b);
#pragma unroll
                  for (int i = 0; i < 8; ++i) a[i] += b[i];
                  half8_to_float(ld_half8(p.bias + p.F + n), b);
#pragma unroll
                  for (int i = 0; i < 8; ++i) gg[i] += b[i];
                }
#pragma unroll
                for (int i = 0; i < 8; ++i) y[i] = a[i] * gelu_erf(gg[i]);
                st_half8(p.H + (size_t)row * p.ldh + n, float_to_half8(a));
                st_half8(p.H + (size_t)row * p.ldh + p.F + n, float_to_half8(gg));
                st_half8(p.C + (size_t)row * p.ldc + n, float_to_half8(y));
              }
            }
          } else {
            tmem_ld_wait();
            if (row_ok) {
              const int n0 = nt * BN + c * 32;
#pragma unroll
              for (int g = 0; g < 4; ++g) {
                const int n = n0 + g * 8;
                if (n >= p.N) break;
                float dy[8], a[8], gg[8], da[8], dg[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) dy[i] = __uint_as_float(ra[g * 8 + i]) * p.alpha;
                half8_to_float(ld_half8(p.H + (size_t)row * p.ldh + n), a);
                half8_to_float(ld_half8(p.H + (size_t)row * p.ldh + p.F + n), gg);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                  da[i] = dy[i] * gelu_erf(gg[i]);
                  dg[i] = dy[i] * a[i] * gelu_erf_grad(gg[i]);
                }
                st_half8(p.C + (size_t)row * p.ldc + n, float_to_half8(da));
                st_half8(p.C + (size_t)row * p.ldc + p.F + n, float_to_half8(dg));
              }
            }
          }
        }
      }
      // release this accumulator stage back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[as]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
  }
}

template <int BN, int EPI>
static int launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmParams& p, cudaStream_t stream) {
  using Cfg = GemmCfg<BN>;
  static bool configured = false;
  if (!configured) {
    DB1_CUDA(cudaFuncSetAttribute(gemm_kernel<BN, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    configured = true;
  }
  const int MT = cdiv(p.M, BM);
  const int NT = (EPI == EPI_GEGLU) ? p.F / (BN / 2) : cdiv(p.N, BN);
  int grid = MT * NT;
  if (grid > sm_count()) grid = sm_count();
  gemm_kernel<BN, EPI><<<grid, GEMM_THREADS, Cfg::SMEM_BYTES, stream>>>(tmA, tmB, p);
  DB1_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace db1

using namespace db1;

// See include/db1_sm100.h for the contract.
extern "C" int db1_gemm_f16(int epilogue, const void* A, int lda, int a_mn, const void* B, int ldb, int b_mn, void* C,
                            int ldc, int M, int N, int K, float alpha, int accumulate, const void* bias,
                            const void* resid, int ldr, float drop_p, uint64_t seed, const void* u, const void* v,
                            int d_model, void* H, int ldh, int F, int bn_hint, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  DB1_CHECK_ARG(M > 0 && N > 0 && K > 0, "gemm: bad shape M=%d N=%d K=%d", M, N, K);
  DB1_CHECK_ARG(lda % 8 == 0 && ldb % 8 == 0 && ldc % 8 == 0, "gemm: leading dims must be multiples of 8 (%d %d %d)",
                lda, ldb, ldc);
  DB1_CHECK_ARG(epilogue >= 0 && epilogue <= 3, "gemm: unknown epilogue %d", epilogue);
  DB1_CHECK_ARG(drop_p >= 0.f && drop_p < 1.f, "gemm: dropout p=%f out of range", drop_p);
  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.M = M; p.N = N; p.K = K; p.a_mn = a_mn ? 1 : 0; p.b_mn = b_mn ? 1 : 0;
  p.alpha = alpha; p.C = (__half*)C; p.ldc = ldc; p.bias = (const __half*)bias;
  p.resid = (const __half*)resid; p.ldr = ldr; p.accumulate = accumulate;
  p.drop_thr16 = (uint32_t)(drop_p * 65536.0f + 0.5f);
  p.drop_scale = p.drop_thr16 ? 65536.0f / (65536.0f - (float)p.drop_thr16) : 1.0f;
  p.seed = seed; p.u = (const __half*)u; p.v = (const __half*)v; p.d_model = d_model;
  p.H = (__half*)H; p.ldh = ldh; p.F = F;

  int BNsel = 256;
  if (epilogue == EPI_PLAIN || epilogue == EPI_DGEGLU) {
    if (bn_hint == 128 || (bn_hint == 0 && N <= 128)) BNsel = 128;
  }
  if (epilogue == EPI_QKV) {
    DB1_CHECK_ARG(u && v && d_model > 0 && N == 3 * d_model, "gemm(qkv): need u, v and N == 3*d_model");
    DB1_CHECK_ARG(d_model % 128 == 0, "gemm(qkv): d_model %d must be a multiple of 128", d_model);
    DB1_CHECK_ARG(!accumulate && !bias && !resid && drop_p == 0.f, "gemm(qkv): unsupported fused option");
    BNsel = (d_model % 256 == 0 && bn_hint != 128) ? 256 : 128;
  }
  if (epilogue == EPI_GEGLU) {
    DB1_CHECK_ARG(H && F > 0 && N == 2 * F && F % 128 == 0 && !b_mn, "gemm(geglu): need H, N == 2F, F %% 128 == 0");
  }
  if (epilogue == EPI_DGEGLU) {
    DB1_CHECK_ARG(H && F > 0 && N == F && F % 8 == 0, "gemm(dgeglu): need H and N == F");
  }
  if (epilogue == EPI_PLAIN && (bias || resid || accumulate || p.drop_thr16))
    DB1_CHECK_ARG(N % 8 == 0, "gemm: fused bias/residual/dropout/accumulate need N %% 8 == 0 (N=%d)", N);
  if (p.drop_thr16) DB1_CHECK_ARG(N % 4 == 0, "gemm: dropout needs N %% 4 == 0");

  CUtensorMap tmA, tmB;
  {
    uint64_t dims[2], str[1];
    uint32_t box[2];
    if (!p.a_mn) { dims[0] = (uint64_t)K; dims[1] = (uint64_t)M; box[0] = 64; box[1] = 128; }
    else         { dims[0] = (uint64_t)M; dims[1] = (uint64_t)K; box[0] = 64; box[1] = 64; }
    str[0] = (uint64_t)lda * 2;
    int e = make_tmap_f16(&tmA, A, 2, dims, str, box);
    if (e) return e;
    if (!p.b_mn) { dims[0] = (uint64_t)K; dims[1] = (uint64_t)N; box[0] = 64; box[1] = 128; }
    else         { dims[0] = (uint64_t)N; dims[1] = (uint64_t)K; box[0] = 64; box[1] = 64; }
    str[0] = (uint64_t)ldb * 2;
    e = make_tmap_f16(&tmB, B, 2, dims, str, box);
    if (e) return e;
  }
  switch (epilogue) {
    case EPI_PLAIN:
      return BNsel == 256 ? launch_gemm<256, EPI_PLAIN>(tmA, tmB, p, stream)
                          : launch_gemm<128, EPI_PLAIN>(tmA, tmB, p, stream);
    case EPI_QKV:
      return BNsel == 256 ? launch_gemm<256, EPI_QKV>(tmA, tmB, p, stream)
                          : launch_gemm<128, EPI_QKV>(tmA, tmB, p, stream);
    case EPI_GEGLU:
      return launch_gemm<256, EPI_GEGLU>(tmA, tmB, p, stream);
    case EPI_DGEGLU:
      return BNsel == 256 ? launch_gemm<256, EPI_DGEGLU>(tmA, tmB, p, stream)
                          : launch_gemm<128, EPI_DGEGLU>(tmA, tmB, p, stream);
  }
  return set_err(-1, "gemm: unreachable");
}
