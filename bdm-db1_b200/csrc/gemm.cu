// Persistent warp-specialised tcgen05 GEMM for sm_100a.
//
//   for every batch (z1, z2):  C[M,N] (+)= epilogue( alpha * A[M,K] * B[N,K]^T ),  fp16 operands, fp32 accumulation in TMEM.
//
// One CTA per SM walks output tiles (128 x BN); large un-batched problems run as CTA PAIRS (cta_group::2: a 256 x 256
// MMA over two vertically adjacent tiles, each CTA staging its 128 rows of A and half of the B tile). Roles:
//   warp 0     TMA producer : global -> smem ring (STAGES x [A 128x64 | B], 128B swizzle, 4-D tensor maps
//                             (inner, rows, z1, z2) so batched / strided operands need no gather)
//   warp 1     MMA issuer   : one elected thread issues tcgen05.mma (K = 16 per instruction), accumulators in TMEM;
//                             the accumulator is double-buffered (2 x BN columns) so tile i+1's main loop
//                             overlaps tile i's epilogue
//   warps 2..9 epilogue     : two warps per TMEM lane quadrant, a thread owns one output row. Three styles:
//                             TMA epilogue (PLAIN on CTA pairs: double-buffered tcgen05.ld -> compile-time specialised
//                             math -> swizzled smem tile -> cp.async.bulk.tensor store, addend fetched by TMA),
//                             register-direct (GeGLU backward on pairs: 256-bit row loads / stores, no smem), and
//                             staged (all others: per-warp smem transpose so every store covers 8 rows x 64 bytes)
// Tile order: data parallel (pair, pair + G, ...); causal k-ranges: cost order, heaviest first, alternating direction;
// opt-in: stream-K tail (DB1_GEMM_SK) and 4-CTA clusters with B multicast (DB1_GEMM_CL=4) - see DESIGN.md 3.1 for why
// both stay off.
//
// Either operand may be K-contiguous ("K-major", e.g. activations x weights^T in the forward pass) or
// MN-contiguous ("MN-major": weights in dgrad, both operands in wgrad) - the UMMA descriptors transpose for free,
// so dgrad/wgrad never materialise a transposed copy.
//
// Replaces the cuBLAS calls behind nn.Linear / F.linear / einsum in the reference
// (src/model/transformer_xl.py:138-139, 163-170, 220, 228, 265-268, 595) and their autograd backward.
#include "../../include/db1_sm100.h"
#include "common.cuh"
#include "ptx.cuh"

#include <stdlib.h>

namespace db1 {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int GEMM_THREADS = 320;  // TMA warp + MMA warp + 8 epilogue warps
constexpr int A_STAGE_BYTES = BM * BK * 2;  // 16 KB

struct GemmParams {
  int M, N, K;
  int a_mn, b_mn;
  int Z1, Z2;
  int a_z1on, a_z2on, b_z1on, b_z2on;
  int reduce_z2;
  int k_mode;
  int skip_upper;
  float alpha;
  __half* C;
  long long ldc, c_z1, c_z2;
  const __half* bias;
  const __half* resid;
  long long ldr;
  int accumulate;
  uint32_t drop_thr16;
  float drop_scale;
  uint64_t seed;
  const __half* u;
  const __half* v;
  int d_model;
  __half* H;
  long long ldh;
  int F;
  const __half* P;
  __half* C2;
  const float* Drow;
  int window;
  int dbg;  // development switches (DB1_GEMM_DBG): 1 = epilogue skips global stores, 2 = epilogue skips TMEM loads too
  // stream-K tail (CTA-pair launches only; 0 = off): the last sk_tiles pair-tiles are cut along K into sk_pairs
  // contiguous, equally long k-block ranges, one per CTA pair; see SkSched below
  int sk_tiles, sk_pairs;
  float* sk_ws;         // partial accumulators, one slot per contributing CTA pair: [pair][crank][128 x BN] fp32
  unsigned int* sk_cnt; // [2 * pairs]: arrivals per owner pair | owner warps that consumed them (self-resetting)
  // TMA epilogue (PLAIN, CTA pairs): C leaves through shared memory and cp.async.bulk.tensor stores; the optional
  // addend (residual, or C itself when accumulating) arrives the same way
  int tma_epi;
  int tma_in;  // 0 = no addend, 1 = residual / old C through tmIn
  int snake;   // causal k-ranges: tiles are visited heaviest first in alternating direction (see sk_item)
  // row-dot side output of the TMA epilogue: dot_out[(b * dot_H + h) * dot_L + i] = sum over head h's 128 columns of
  // acc[row = b * dot_L + i, :] * X[row, :], X arriving through tmIn (attention backward: D = rowsum(dO * O))
  float* dot_out;
  int dot_L, dot_H;
};

// CL == 2: the CTA pair of a cluster runs cta_group::2 MMAs (M = 256 across the pair); each CTA stages its own 128
// rows of A and HALF of the B tile, so a stage is 32 KB instead of 48 KB and the ring gets deeper.
template <int BN, int CL = 1>
struct GemmCfg {
  static constexpr int STAGES = (BN == 256) ? (CL >= 2 ? 5 : 4) : 6;
  static constexpr int B_STAGE_BYTES = BN / (CL >= 2 ? 2 : 1) * BK * 2;  // CTA pairs: half of the B tile per CTA
  static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
  // per-warp 32 rows x (64 B + 16 B pad) transpose areas; CTA-pair kernels: four 8 KB in/out tiles of the TMA epilogue
  static constexpr int EPI_NB = 4;  // in/out tiles per column half of the TMA epilogue
  static constexpr int STAGING_BYTES = (CL >= 2) ? 2 * EPI_NB * 8192 : 8 * 32 * 80;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + STAGING_BYTES + 1024 /*align*/ + 384 /*barriers*/;
  static constexpr int TMEM_COLS = 2 * BN;  // 512 or 256
};

// Store a warp's 32 x 32 fp16 tile (thread = row, hv = its 32 columns) with full-sector writes: the tile is transposed
// through a per-warp staging buffer so that every store instruction covers 8 rows x 64 contiguous bytes. (16-byte
// row-strided stores straight from the accumulator layout write half sectors, which the L2 turns into
// read-modify-write fills: measured 2.5x slower end to end on the K = 2048 GEMMs.)
DEVI void warp_store_tile(uint8_t* stg, int lane, const Half8 (&hv)[4], __half* base, long long ld, int rows_valid,
                          int cols_valid, bool accumulate) {
  const uint32_t sbase = smem_u32(stg);
#pragma unroll
  for (int g = 0; g < 4; ++g) sts_half8(sbase + lane * 80 + g * 16, hv[g]);
  __syncwarp();
  const int piece = lane & 3;
  const int col = piece * 8;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int rr = k * 8 + (lane >> 2);
    if (rr < rows_valid && col < cols_valid) {
      Half8 v = lds_half8(sbase + rr * 80 + piece * 16);
      __half* dst = base + (long long)rr * ld + col;
      if (col + 8 <= cols_valid) {
        if (accumulate) {
          float a[8], b[8];
          half8_to_float(v, a);
          half8_to_float(ld_half8(dst), b);
#pragma unroll
          for (int i = 0; i < 8; ++i) a[i] += b[i];
          v = float_to_half8(a);
        }
        st_half8(dst, v);
      } else {
        // ragged last group (N not a multiple of 8): scalar tail, kept in registers (taking the address of `v`
        // would push every tile through local memory)
        const uint32_t w[4] = {v.u.x, v.u.y, v.u.z, v.u.w};
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          if (i < cols_valid - col) {
            const __half hv = __ushort_as_half((unsigned short)((i & 1) ? (w[i >> 1] >> 16) : (w[i >> 1] & 0xffffu)));
            dst[i] = accumulate ? __float2half_rn(__half2float(hv) + __half2float(dst[i])) : hv;
          }
        }
      }
    }
  }
  __syncwarp();
}

// Coalesced load of a warp's 32 x 32 fp16 tile (the mirror image of warp_store_tile): every load instruction covers
// 8 rows x 64 contiguous bytes; the tile is handed to the row-owning threads through the staging buffer.
// (Row-strided loads straight into the accumulator layout touch 32 different 128-byte lines per instruction.)
// Split so the global loads of the NEXT chunk can be in flight while the current one is processed:
// tile_issue = 4 coalesced LDG.128 into registers; tile_finish = the transpose through the staging buffer.
DEVI void tile_issue(int lane, Half8 (&v)[4], const __half* base, long long ld, int rows_valid, int cols_valid) {
  const int col = (lane & 3) * 8;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int rr = k * 8 + (lane >> 2);
    v[k] = half8_zero();
    if (rr < rows_valid && col < cols_valid) v[k] = ld_half8(base + (long long)rr * ld + col);
  }
}
DEVI void tile_finish(uint8_t* stg, int lane, const Half8 (&v)[4], Half8 (&hv)[4]) {
  const uint32_t sbase = smem_u32(stg);
  const int piece = lane & 3;
#pragma unroll
  for (int k = 0; k < 4; ++k) sts_half8(sbase + (k * 8 + (lane >> 2)) * 80 + piece * 16, v[k]);
  __syncwarp();
#pragma unroll
  for (int g = 0; g < 4; ++g) hv[g] = lds_half8(sbase + lane * 80 + g * 16);
  __syncwarp();
}
DEVI void warp_load_tile(uint8_t* stg, int lane, Half8 (&hv)[4], const __half* base, long long ld, int rows_valid,
                         int cols_valid) {
  Half8 v[4];
  tile_issue(lane, v, base, ld, rows_valid, cols_valid);
  tile_finish(stg, lane, v, hv);
}

struct TileCoord {
  int mt, nt, z1, z2;
  int kb0, kb1;  // k-block range
  bool skip;
};

template <int BN, int EPI, int CL>
DEVI TileCoord decode_tile(const GemmParams& p, int tile, int MT, int NT, int KB, int crank) {
  TileCoord t;
  int r;
  if (CL >= 2) {
    // `tile` indexes groups of CL vertically adjacent tiles: CL == 2: one cta_group::2 MMA tile of 256 rows;
    // CL == 4: two such pairs (512 rows) that share - and multicast - the B tile
    const int MTG = (MT + CL - 1) / CL;
    t.mt = CL * (tile % MTG) + crank;
    r = tile / MTG;
  } else if (p.snake) {
    // position in cost order: all tiles of the heaviest row block first. Rows get heavier with mt for the k-ranges that
    // END at the row (or begin K-(mt+1)*128 from the end), lighter for the one that BEGINS at the row.
    const int per = NT * p.Z1 * (p.reduce_z2 ? 1 : p.Z2);
    const int g = tile / per;
    t.mt = (p.k_mode == DB1_K_BEGIN_BY_ROW) ? g : MT - 1 - g;
    r = tile % per;
  } else {
    t.mt = tile % MT;
    r = tile / MT;
  }
  t.nt = r % NT;
  r /= NT;
  t.z1 = r % p.Z1;
  t.z2 = r / p.Z1;
  t.kb0 = 0;
  t.kb1 = KB;
  if (p.k_mode == DB1_K_END_BY_ROW) {
    int e = ((t.mt + 1) * BM + BK - 1) / BK;
    t.kb1 = e < KB ? e : KB;
  } else if (p.k_mode == DB1_K_BEGIN_BY_ROW) {
    t.kb0 = (t.mt * BM) / BK;
  } else if (p.k_mode == DB1_K_BEGIN_REV) {
    int b = p.K - (t.mt + 1) * BM;
    t.kb0 = b > 0 ? b / BK : 0;
  }
  t.skip = p.skip_upper && (t.nt * BN > t.mt * BM + BM - 1);
  if (t.kb0 >= t.kb1) t.skip = true;
  return t;
}

// ---------------------------------------------------------------------------------------------------------------------
// Work list of one CTA (pair). Without stream-K: tiles pair, pair + G, ... (data parallel, "DP"). With stream-K the
// tile count T is not a multiple of the G pairs, so instead of a last wave that keeps only T mod G pairs busy, the
// R = sk_tiles trailing tiles are treated as one strip of R * KB k-blocks that is cut into sk_pairs equal ranges
// [b(g), b(g+1)), b(g) = g * R * KB / sk_pairs. A range is shorter than a tile, so it touches at most two tiles:
//   * the tail [kb0, KB) of a tile whose head belongs to a lower-numbered pair  -> CONTRIB: the fp32 accumulator goes to
//     this pair's workspace slot and the owner's arrival counter is bumped;
//   * the head [0, kb1) of the next tile                                        -> OWNER: waits for the contributors
//     (pairs g+1 ...), adds their partials in pair order (deterministic) and runs the normal fused epilogue.
// Every pair works through its stream-K range FIRST (contribution before owned head), then its DP tiles: a
// contribution an owner waits for was the first thing its producer did, and the fix-up hides under the DP tiles.
// ---------------------------------------------------------------------------------------------------------------------
enum { SK_FULL = 0, SK_CONTRIB = 1, SK_OWNER = 2 };
// One CTA's work list, computed once by one thread and kept in shared memory (the role loops only keep a counter live:
// the fused epilogues have no registers to spare).
struct SkPlan {
  int nseg;   // stream-K segments of this pair (0, 1 or 2), processed before the DP tiles
  int T_dp;   // tiles [0, T_dp) are data parallel: pair, pair + G, ...
  int tile[2], kb0[2], kb1[2];
  int role[2];   // SK_*
  int peer[2];   // CONTRIB: owner pair; OWNER: first contributing pair
  int npeer[2];  // OWNER: number of contributing pairs (consecutive)
};
__host__ __device__ inline void sk_plan_build(SkPlan* pl, int R, int GS, int pair, int num_tiles, int KB) {
  auto bound = [&](int g) { return (int)(((long long)g * R * KB) / GS); };
  pl->nseg = 0;
  pl->T_dp = num_tiles - R;
  if (R <= 0 || pair >= GS) return;
  int u = bound(pair);
  const int u1 = bound(pair + 1);
  int n = 0;
  while (u < u1 && n < 2) {
    const int s = u / KB;
    const int e = u1 < (s + 1) * KB ? u1 : (s + 1) * KB;
    const int kb0 = u - s * KB, kb1 = e - s * KB;
    pl->tile[n] = pl->T_dp + s;
    pl->kb0[n] = kb0;
    pl->kb1[n] = kb1;
    pl->peer[n] = 0;
    pl->npeer[n] = 0;
    if (kb0 == 0 && kb1 == KB) {
      pl->role[n] = SK_FULL;
    } else if (kb0 == 0) {
      int g = pair + 1;
      while (g < GS && bound(g) < (s + 1) * KB) ++g;
      pl->role[n] = SK_OWNER;
      pl->peer[n] = pair + 1;
      pl->npeer[n] = g - (pair + 1);
    } else {
      int g = pair - 1;
      while (g > 0 && bound(g) > s * KB) --g;
      pl->role[n] = SK_CONTRIB;
      pl->peer[n] = g;
    }
    ++n;
    u = e;
  }
  pl->nseg = n;
}
// n-th work item of this pair: returns false when the list is exhausted. kb1 < 0 = the tile's own k-range.
// snake != 0 (causal k-ranges: tile cost falls with the tile index, decode_tile sorts heaviest first): passes alternate
// direction - pass 0: CTA c takes position c, pass 1: 2G-1-c, ... - so every CTA gets the same load within one tile;
// a position past the end is reported as tile = -1 (skip).
DEVI bool sk_item(const volatile SkPlan* pl, int n, int pair, int G, int snake, int& tile, int& kb0, int& kb1) {
  const int nseg = pl->nseg;
  if (n < nseg) {
    tile = pl->tile[n];
    kb0 = pl->kb0[n];
    kb1 = pl->kb1[n];
    return true;
  }
  kb0 = 0;
  kb1 = -1;
  const int T = pl->T_dp;
  if (snake) {
    if (n * G >= T) return false;
    tile = (n & 1) ? (n + 1) * G - 1 - pair : n * G + pair;
    if (tile >= T) tile = -1;
    return true;
  }
  tile = pair + (n - nseg) * G;
  return tile < T;
}

DEVI unsigned int ld_acquire_gpu(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// ---- stream-K fix-up. Every epilogue warp owns the same 32 lanes x BN/2 accumulator columns in both roles, so a
// partial is stored / fetched in "register order": [chunk][8 x (32 lanes x float4)] - fully coalesced 512-byte rows.
// Kept out of line: they run at most twice per CTA and must not add register pressure to the fused epilogues.
template <int BN>
DEVI void sk_store_partial(const GemmParams& p, uint32_t tacc, int pair, int owner, int crank, int ew,
                                              int half, int lane) {
  constexpr int CPW = BN / 64;
  float* ws = p.sk_ws + (size_t)pair * (2 * BM * BN) + (size_t)crank * (BM * BN) + (size_t)ew * (CPW * 8 * 128) +
              (size_t)lane * 4;
#pragma unroll 1
  for (int ci = 0; ci < CPW; ++ci) {
    uint32_t r[32];
    tmem_ld32(tacc + (half * CPW + ci) * 32, r);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 8; ++j)
      __stcg(reinterpret_cast<float4*>(ws + (ci * 8 + j) * 128),
             make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]), __uint_as_float(r[4 * j + 2]),
                         __uint_as_float(r[4 * j + 3])));
  }
  __threadfence();  // the partial is visible device-wide before the arrival is
  __syncwarp();
  if (lane == 0) atomicAdd(p.sk_cnt + owner, 1u);
}

template <int BN>
DEVI void sk_reduce_partials(const GemmParams& p, uint32_t tacc, int pair, int npairs, int first,
                                                int npeer, int crank, int ew, int half, int lane) {
  constexpr int CPW = BN / 64;
  // all 16 epilogue warps (8 per CTA of the pair) of every contributing pair have arrived?
  if (lane == 0) {
    const unsigned int want = (unsigned int)(16 * npeer);
    while (ld_acquire_gpu(p.sk_cnt + pair) < want) __nanosleep(40);
  }
  __syncwarp();
  __threadfence();
  const size_t my_off = (size_t)crank * (BM * BN) + (size_t)ew * (CPW * 8 * 128) + (size_t)lane * 4;
#pragma unroll 1
  for (int ci = 0; ci < CPW; ++ci) {
    uint32_t r[32];
    tmem_ld32(tacc + (half * CPW + ci) * 32, r);
    tmem_ld_wait();
#pragma unroll 1
    for (int q = 0; q < npeer; ++q) {  // fixed order: deterministic sums
      const float* ws = p.sk_ws + (size_t)(first + q) * (2 * BM * BN) + my_off;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 v = __ldcg(reinterpret_cast<const float4*>(ws + (ci * 8 + j) * 128));
        r[4 * j] = __float_as_uint(__uint_as_float(r[4 * j]) + v.x);
        r[4 * j + 1] = __float_as_uint(__uint_as_float(r[4 * j + 1]) + v.y);
        r[4 * j + 2] = __float_as_uint(__uint_as_float(r[4 * j + 2]) + v.z);
        r[4 * j + 3] = __float_as_uint(__uint_as_float(r[4 * j + 3]) + v.w);
      }
    }
    tmem_st32(tacc + (half * CPW + ci) * 32, r);
  }
  tmem_st_wait();
  tc_fence_before();
  __syncwarp();
  if (lane == 0) {
    // the last of the pair's 16 owner warps to get here re-arms both counters for the next launch
    const unsigned int done = atomicAdd(p.sk_cnt + npairs + pair, 1u);
    if (done == 15u) {
      p.sk_cnt[pair] = 0u;
      p.sk_cnt[npairs + pair] = 0u;
    }
  }
}

// One 32-column chunk of the TMA epilogue for one thread (= one output row): accumulator registers -> fused math ->
// four 16-byte pieces of the row in the [128][32] fp16 shared-memory tile (64-byte swizzle). Specialised at compile time:
// with run-time flags every piece carries ~60 predicated-off or branch instructions and a handful of constant-bank
// reloads, which with two epilogue warps per scheduler were most of the epilogue's time.
template <bool FULL, int IN>  // IN: 0 = no second operand, 1 = add it (residual / old C), 2 = dot it with the result (dacc)
DEVI void epi_chunk(const uint32_t (&r)[32], uint32_t rowp, uint32_t swz, int row, int col0, int N, float alpha,
                    const __half* bias, uint32_t thr, float dscale, uint64_t seed, float& dacc) {
  Half8 bv[4];
  if (FULL) {
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      bv[g] = half8_zero();
      if (bias && col0 + g * 8 < N) bv[g] = ld_half8(bias + col0 + g * 8);
    }
  }
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    float f[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) f[i] = __uint_as_float(r[g * 8 + i]);
    if (FULL) {
      float bb[8];
      half8_to_float(bv[g], bb);
#pragma unroll
      for (int i = 0; i < 8; ++i) f[i] = fmaf(f[i], alpha, bb[i]);
      if (thr) {
        const uint64_t e = (uint64_t)row * (uint64_t)N + (uint64_t)(col0 + g * 8);
        const uint64_t b0 = rng64(seed, e >> 2), b1 = rng64(seed, (e >> 2) + 1);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          f[i] = dropout_keep(b0, i, thr) ? f[i] * dscale : 0.f;
          f[4 + i] = dropout_keep(b1, i, thr) ? f[4 + i] * dscale : 0.f;
        }
      }
    }
    const uint32_t sa = rowp + (((uint32_t)g ^ swz) << 4);
    if (IN == 1) {
      float bb[8];
      half8_to_float(lds_half8(sa), bb);
#pragma unroll
      for (int i = 0; i < 8; ++i) f[i] += bb[i];
    }
    if (IN == 2) {
      float bb[8];
      half8_to_float(lds_half8(sa), bb);
#pragma unroll
      for (int i = 0; i < 8; ++i) dacc = fmaf(f[i], bb[i], dacc);
    }
    sts_half8(sa, float_to_half8(f));
  }
}

template <int BN, int EPI, int CL>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
            const __grid_constant__ CUtensorMap tmB64, const __grid_constant__ CUtensorMap tmC,
            const __grid_constant__ CUtensorMap tmIn, const GemmParams p) {
  using Cfg = GemmCfg<BN, CL>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* staging = smem + STAGES * Cfg::STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(staging + Cfg::STAGING_BYTES);
  uint64_t* full = bars;
  uint64_t* empty = bars + STAGES;
  uint64_t* tfull = bars + 2 * STAGES;
  uint64_t* tempty = bars + 2 * STAGES + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);
  uint64_t* ebar = bars + 2 * STAGES + 6;  // TMA epilogue: "in/out tile ready" per (column half, buffer)
  SkPlan* plan = reinterpret_cast<SkPlan*>(bars + 2 * STAGES + 6 + 2 * Cfg::EPI_NB);
  static_assert((2 * STAGES + 6 + 2 * Cfg::EPI_NB) * 8 + sizeof(SkPlan) <= 384, "barrier area too small");

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int MT = (p.M + BM - 1) / BM;
  const int NT = (EPI == DB1_EPI_GEGLU) ? (p.F / (BN / 2)) : (p.N + BN - 1) / BN;
  const int ZO = p.reduce_z2 ? 1 : p.Z2;
  const int num_tiles = (CL >= 2 ? (MT + CL - 1) / CL : MT) * NT * p.Z1 * ZO;  // tile groups when clustered
  const int KB = (p.K + BK - 1) / BK;
  const int crank = (CL >= 2) ? (int)cluster_ctarank() : 0;
  const int prank = crank & 1;   // rank within the CTA pair (0 = leader: expects the bytes, issues the MMAs)
  const int pairi = crank >> 1;  // pair within the cluster (CL == 4)
  const int tile0 = (CL >= 2) ? (int)(blockIdx.x / CL) : (int)blockIdx.x;
  const int tstep = (CL >= 2) ? (int)(gridDim.x / CL) : (int)gridDim.x;
  const int KZ = p.reduce_z2 ? p.Z2 : 1;  // extra contraction loop over z2

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (p.tma_epi) {
      tma_prefetch_desc(&tmC);
      if (p.tma_in) tma_prefetch_desc(&tmIn);
    }
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int i = 0; i < STAGES; ++i) {
        mbar_init(&full[i], 1);   // CL == 2: only the leader's is used; it counts the bytes of both CTAs' loads
        mbar_init(&empty[i], CL == 4 ? 2 : 1);  // CL >= 2: released in every CTA of the cluster by each pair leader's multicast commit
      }
      for (int i = 0; i < 2; ++i) {
        mbar_init(&tfull[i], 1);
        mbar_init(&tempty[i], CL >= 2 ? 16 : 8);  // CL >= 2: the epilogue warps of BOTH CTAs of a pair arrive on its leader's barrier
      }
      for (int i = 0; i < 2 * Cfg::EPI_NB; ++i) mbar_init(&ebar[i], 1);
      mbar_fence_init();
      sk_plan_build(plan, p.sk_tiles, p.sk_pairs, tile0, num_tiles, KB);
    }
    __syncwarp();
    if (CL >= 2) tmem_alloc2<Cfg::TMEM_COLS>(tmem_slot);
    else tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
  }
  tc_fence_before();
  __syncthreads();
  if (CL >= 2) cluster_sync_all();  // barriers of all CTAs are initialised before any remote arrive / multicast
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // everything above overlapped the previous kernel's tail; from here on global memory is touched
  pdl_launch_dependents();
  pdl_wait();
  // development timeline (DB1_GEMM_DBG & 8): clock64 stamps per CTA into the registered workspace
  unsigned long long* tl = (p.dbg & 8) ? reinterpret_cast<unsigned long long*>(p.sk_ws) + (size_t)blockIdx.x * 64 : nullptr;
  if (tl && threadIdx.x == 64) tl[0] = clock64();

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int wn = 0;; ++wn) {
        int tile, skb0, skb1;
        if (!sk_item(plan, wn, tile0, tstep, p.snake, tile, skb0, skb1)) break;
        if (tile < 0) continue;
        TileCoord t = decode_tile<BN, EPI, CL>(p, tile, MT, NT, KB, crank);
        if (t.skip) continue;
        if (skb1 >= 0) { t.kb0 = skb0; t.kb1 = skb1; }
        const int m0 = t.mt * BM;
        for (int kz = 0; kz < KZ; ++kz) {
          const int z2 = p.reduce_z2 ? kz : t.z2;
          const int az1 = p.a_z1on ? t.z1 : 0, az2 = p.a_z2on ? z2 : 0;
          const int bz1 = p.b_z1on ? t.z1 : 0, bz2 = p.b_z2on ? z2 : 0;
          for (int kb = t.kb0; kb < t.kb1; ++kb) {
            mbar_wait(&empty[s], ph ^ 1);
            uint8_t* sa = smem + s * Cfg::STAGE_BYTES;
            uint8_t* sb = sa + A_STAGE_BYTES;
            const int k0 = kb * BK;
            if (CL >= 2) {
              // both CTAs' boxes are counted on the pair leader's barrier (the MMA issuer lives there)
              // (dbg & 4: experiment - every other k-block reuses the stale B tile, i.e. 25 % less L2 -> SM traffic)
              const bool skip_b = (p.dbg & 4) && (kb & 1);
              if (prank == 0) mbar_expect_tx(&full[s], skip_b ? 2 * A_STAGE_BYTES : 2 * Cfg::STAGE_BYTES);
              if (!p.a_mn) {
                tma_load_4d_2sm(sa, &tmA, &full[s], k0, m0, az1, az2);
              } else {
                tma_load_4d_2sm(sa, &tmA, &full[s], m0, k0, az1, az2);
                tma_load_4d_2sm(sa + 8192, &tmA, &full[s], m0 + 64, k0, az1, az2);
              }
              // this CTA's half of the B tile: rows (K-major) / columns (MN-major) [prank * BN/2, +BN/2)
              if (skip_b) {
              } else if (CL == 4) {
                // The two pairs of the cluster need the same B halves: the half is two 8 KB boxes (64 rows / columns
                // each); pair i fetches box i and multicasts it to the CTAs of equal pair rank in both pairs, so the
                // cluster reads every B byte from L2 once instead of twice.
                const uint16_t mask = (uint16_t)(0x5u << prank);
                uint8_t* dst = sb + pairi * 8192;
                const int n0 = t.nt * BN + prank * (BN / 2) + pairi * 64;
                if (!p.b_mn) tma_load_4d_2sm_mc(dst, &tmB64, &full[s], k0, n0, bz1, bz2, mask);
                else tma_load_4d_2sm_mc(dst, &tmB, &full[s], n0, k0, bz1, bz2, mask);
              } else if (!p.b_mn) {
#pragma unroll
                for (int j = 0; j < BN / 256; ++j) {
                  int row0;
                  if (EPI == DB1_EPI_GEGLU) row0 = prank * p.F + t.nt * (BN / 2);
                  else row0 = t.nt * BN + prank * (BN / 2);
                  tma_load_4d_2sm(sb + j * 16384, &tmB, &full[s], k0, row0 + j * 128, bz1, bz2);
                }
              } else {
#pragma unroll
                for (int j = 0; j < BN / 128; ++j)
                  tma_load_4d_2sm(sb + j * 8192, &tmB, &full[s], t.nt * BN + prank * (BN / 2) + j * 64, k0, bz1, bz2);
              }
            } else {
              mbar_expect_tx(&full[s], Cfg::STAGE_BYTES);
              if (!p.a_mn) {
                tma_load_4d(sa, &tmA, &full[s], k0, m0, az1, az2);
              } else {
                tma_load_4d(sa, &tmA, &full[s], m0, k0, az1, az2);
                tma_load_4d(sa + 8192, &tmA, &full[s], m0 + 64, k0, az1, az2);
              }
              if (!p.b_mn) {
#pragma unroll
                for (int j = 0; j < BN / 128; ++j) {
                  int row0;
                  if (EPI == DB1_EPI_GEGLU) row0 = j * p.F + t.nt * (BN / 2);
                  else row0 = t.nt * BN + j * 128;
                  tma_load_4d(sb + j * 16384, &tmB, &full[s], k0, row0, bz1, bz2);
                }
              } else {
#pragma unroll
                for (int j = 0; j < BN / 64; ++j)
                  tma_load_4d(sb + j * 8192, &tmB, &full[s], t.nt * BN + j * 64, k0, bz1, bz2);
              }
            }
            if (++s == STAGES) {
              s = 0;
              ph ^= 1;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if (lane == 0 && prank == 0) {  // CL >= 2: the pair leader issues for its pair
      const uint32_t idesc = umma_idesc(CL >= 2 ? 2 * BM : BM, BN, p.a_mn, p.b_mn, 0);
      const uint32_t a_lbo = p.a_mn ? 8192u : 16u, b_lbo = p.b_mn ? 8192u : 16u;
      const uint32_t a_kadv = p.a_mn ? (2048u >> 4) : (32u >> 4);  // descriptor start-address units of 16 B
      const uint32_t b_kadv = p.b_mn ? (2048u >> 4) : (32u >> 4);
      int s = 0;
      uint32_t ph = 0;
      int it = 0;
      for (int wn = 0;; ++wn) {
        int tile, skb0, skb1;
        if (!sk_item(plan, wn, tile0, tstep, p.snake, tile, skb0, skb1)) break;
        if (tile < 0) continue;
        TileCoord t = decode_tile<BN, EPI, CL>(p, tile, MT, NT, KB, crank);
        if (t.skip) continue;
        if (skb1 >= 0) { t.kb0 = skb0; t.kb1 = skb1; }
        const int as = it & 1;
        const uint32_t aph = (it >> 1) & 1;
        ++it;
        mbar_wait(&tempty[as], aph ^ 1);
        tc_fence_after();
        const uint32_t tacc = tmem_base + as * BN;
        uint32_t acc = 0;
        for (int kz = 0; kz < KZ; ++kz) {
          for (int kb = t.kb0; kb < t.kb1; ++kb) {
            mbar_wait(&full[s], ph);
            tc_fence_after();
            if (tl && kb == t.kb0 && it <= 15) tl[4 + 4 * (it - 1)] = clock64();
            const uint32_t sa = smem_u32(smem + s * Cfg::STAGE_BYTES);
            const uint64_t adesc = umma_smem_desc(sa, a_lbo, 1024);
            const uint64_t bdesc = umma_smem_desc(sa + A_STAGE_BYTES, b_lbo, 1024);
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) {
              if (CL >= 2) umma_ss2(tacc, adesc + (uint64_t)(k * a_kadv), bdesc + (uint64_t)(k * b_kadv), idesc, acc);
              else umma_ss(tacc, adesc + (uint64_t)(k * a_kadv), bdesc + (uint64_t)(k * b_kadv), idesc, acc);
              acc = 1;
            }
            if (CL >= 2) umma_commit2_mc(&empty[s], (uint16_t)(CL == 4 ? 0xF : 3));
            else umma_commit(&empty[s]);
            if (++s == STAGES) {
              s = 0;
              ph ^= 1;
            }
          }
        }
        if (CL >= 2) umma_commit2_mc(&tfull[as], (uint16_t)(3u << (crank & 2)));
        else umma_commit(&tfull[as]);
        if (tl && it <= 15) tl[4 + 4 * (it - 1) + 1] = clock64();
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue warps (2..9)
    // Two warps per TMEM lane quadrant (a warp may only touch lanes 32*(warp%4)..+31); each takes half of the
    // tile's 32-column chunks. Global operands of the next chunk (residual / saved activations / P) are fetched
    // before waiting on the current chunk's TMEM load so their latency overlaps.
    const int ew = warp - 2;
    const int quad = warp & 3;
    const int half = ew >> 2;
    int it = 0;
    unsigned int eseq = 0;  // TMA epilogue: chunks this column half has processed (buffer = eseq % NB)
    // Epilogue parameters as locals: every inline-asm statement here clobbers "memory", after which the compiler
    // re-reads kernel parameters from the constant bank (LDC / LDCU + a dependent uniform branch, ~100 cycles each);
    // measured: 1000 of the 1450 cycles per 32-column chunk were spent on those reloads.
    const float e_alpha = p.alpha;
    const __half* const e_bias = p.bias;
    const uint32_t e_thr = p.drop_thr16;
    const float e_dscale = p.drop_scale;
    const uint64_t e_seed = p.seed;
    const int e_in = p.tma_in;
    const int e_N = p.N;
    const bool e_full = e_alpha != 1.0f || e_bias != nullptr || e_thr != 0u;
    const bool e_dot = p.dot_out != nullptr;
    // register-direct epilogues: 256-bit row accesses need 32-byte aligned rows of C and H
    const bool e_al32 = (((uintptr_t)p.C | (uintptr_t)p.H) & 31) == 0 && ((p.ldc | p.ldh | p.F) & 15) == 0;
    for (int wn = 0;; ++wn) {
      int tile, skb0, skb1;
      if (!sk_item(plan, wn, tile0, tstep, p.snake, tile, skb0, skb1)) break;
        if (tile < 0) continue;
      const TileCoord t = decode_tile<BN, EPI, CL>(p, tile, MT, NT, KB, crank);
      if (t.skip) continue;
      const int mt = t.mt, nt = t.nt;
      const int as = it & 1;
      const uint32_t aph = (it >> 1) & 1;
      ++it;
      const int row = mt * BM + quad * 32 + lane;
      const bool row_ok = row < p.M;
      const long long zoff = (long long)t.z1 * p.c_z1 + (long long)t.z2 * p.c_z2;
      const uint32_t tacc = tmem_base + as * BN + ((uint32_t)(quad * 32) << 16);
      uint8_t* stg = staging + ew * (32 * 80);
      const int row_base = mt * BM + quad * 32;

      if (CL == 2 && wn < 2 && skb1 >= 0) {
        const volatile SkPlan* vp = plan;
        const int role = vp->role[wn];
        if (role != SK_FULL) {
          mbar_wait_backoff(&tfull[as], aph, p.dbg & 16 ? 0u : 100u);
          tc_fence_after();
          if (role == SK_CONTRIB) {
            sk_store_partial<BN>(p, tacc, tile0, vp->peer[wn], prank, ew, half, lane);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_leader(&tempty[as]);
            continue;
          }
          sk_reduce_partials<BN>(p, tacc, tile0, tstep, vp->peer[wn], vp->npeer[wn], prank, ew, half, lane);
          // the regular epilogue below may read columns that the other warp of this lane quadrant has just patched
          named_bar_sync(1 + quad, 64);
          tc_fence_after();
        }
      }

      if constexpr (CL >= 2 && EPI == DB1_EPI_PLAIN) {
        {
          // ---- TMA epilogue. The four warps of a column half (one per TMEM lane quadrant) own 128 rows x 128 columns
          // and work through them in 32-column chunks: tcgen05.ld (next chunk's load in flight) -> fused math -> the
          // chunk as a [128][32] fp16 tile in shared memory (64-byte swizzle, conflict-free 16-byte accesses) -> one
          // cp.async.bulk.tensor store by the half's leader thread. The addend (residual / old C) is fetched into
          // the same buffer by TMA one chunk ahead and overwritten in place. Two buffers per half; nothing in this
          // path waits on a global load or store round trip except through the mbarrier / bulk-group machinery.
          constexpr int CPW = BN / 64;
          constexpr int NB = Cfg::EPI_NB;
          static_assert(NB == 4 && CPW <= NB, "buffer ring sized for one tile's chunks");
          const int c_first = half * CPW;
          uint8_t* ebuf = staging + half * (NB * 8192);
          uint64_t* ebh = ebar + half * NB;
          const bool eleader = ((ew & 3) == 0) && lane == 0;
          const int trow = quad * 32 + lane;
          const uint32_t swz = (uint32_t)((trow >> 1) & 3);
          const int row0 = mt * BM;
          const int colh = nt * BN + c_first * 32;  // first column of this half
          int nch = (e_N - colh + 31) / 32;         // chunks of this half that hold valid columns (ragged N)
          nch = nch < 0 ? 0 : (nch > CPW ? CPW : nch);
          if (p.dbg & 2) nch = 0;
          const int nch_arm = (p.dbg & 64) ? 0 : nch;
          // leader: make buffer seq % NB ready for chunk `seq` (its previous user, chunk seq - NB, has been stored:
          // callers keep at most ONE store group pending before arming)
          auto arm = [&](unsigned int seq, int col0) {
            const unsigned int b = seq & (NB - 1);
            if (e_in) {
              mbar_expect_tx(&ebh[b], 8192);
              tma_load_2d(ebuf + b * 8192, &tmIn, &ebh[b], col0, row0);
            } else {
              mbar_arrive(&ebh[b]);
            }
          };
          if (eleader && nch_arm > 0) {  // under the tail of this tile's main loop: the first NB - 1 chunks
            bulk_wait_group_read<1>();
            for (int k = 0; k < nch && k < NB - 1; ++k) arm(eseq + k, colh + k * 32);
          }
          mbar_wait_backoff(&tfull[as], aph, p.dbg & 16 ? 0u : 100u);
          tc_fence_after();
          if (tl && ew == 0 && lane == 0 && it <= 15) tl[4 + 4 * (it - 1) + 2] = clock64();
          uint32_t r[2][32];
          float dacc = 0.f;
          if (p.dbg & 32) {
#pragma unroll
            for (int k = 0; k < 32; ++k) r[0][k] = r[1][k] = 0u;
          }
          if (p.dbg & 64) {  // experiment: TMEM loads only
            for (int ci = 0; ci < nch; ++ci) {
              tmem_ld32(tacc + (c_first + ci) * 32, r[0]);
              tmem_ld_wait();
            }
            nch = 0;
          }
          if (nch > 0 && !(p.dbg & 32)) tmem_ld32(tacc + c_first * 32, r[0]);
#pragma unroll
          for (int ci = 0; ci < CPW; ++ci) {
            if (ci < nch) {
              tmem_ld_wait();
              if (ci + 1 < nch && !(p.dbg & 32)) tmem_ld32(tacc + (c_first + ci + 1) * 32, r[(ci + 1) & 1]);
              const int col0 = colh + ci * 32;
              const unsigned int b = eseq & (NB - 1);
              mbar_wait(&ebh[b], (eseq / NB) & 1u);
              const uint32_t rowp = smem_u32(ebuf + b * 8192 + trow * 64);
              if (e_dot) {
                epi_chunk<false, 2>(r[ci & 1], rowp, swz, row, col0, e_N, 1.f, nullptr, 0u, 1.f, 0ull, dacc);
              } else if (e_full) {
                if (e_in) epi_chunk<true, 1>(r[ci & 1], rowp, swz, row, col0, e_N, e_alpha, e_bias, e_thr, e_dscale, e_seed, dacc);
                else epi_chunk<true, 0>(r[ci & 1], rowp, swz, row, col0, e_N, e_alpha, e_bias, e_thr, e_dscale, e_seed, dacc);
              } else {
                if (e_in) epi_chunk<false, 1>(r[ci & 1], rowp, swz, row, col0, e_N, 1.f, nullptr, 0u, 1.f, 0ull, dacc);
                else epi_chunk<false, 0>(r[ci & 1], rowp, swz, row, col0, e_N, 1.f, nullptr, 0u, 1.f, 0ull, dacc);
              }
              fence_proxy_async_smem();
              named_bar_sync(5 + half, 128);
              if (eleader) {
                if (!(p.dbg & 1)) tma_store_2d(&tmC, ebuf + b * 8192, col0, row0);
                bulk_commit_group();
                if (ci + NB - 1 < nch) {  // the ring's last buffer: free once the previous chunk's store has read it
                  bulk_wait_group_read<1>();
                  arm(eseq + NB - 1, col0 + (NB - 1) * 32);
                }
              }
              ++eseq;
            }
          }
          if (e_dot && row_ok && nch == CPW) {
            // this column half = one 128-wide head: dacc is the complete row-dot for (row, head)
            const int hh = (colh >> 7), bb = row / p.dot_L, ii = row - bb * p.dot_L;
            p.dot_out[((size_t)bb * p.dot_H + hh) * p.dot_L + ii] = dacc;
          }
          if (tl && ew == 0 && lane == 0 && it <= 15) tl[4 + 4 * (it - 1) + 3] = clock64();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_leader(&tempty[as]);
          continue;
        }
      }
      if constexpr (CL >= 2 && EPI == DB1_EPI_DGEGLU) {
        // ---- GeGLU backward, register-direct: a thread owns one output row; per 32-column chunk it pulls its 64 bytes
        // of a and of g with 256-bit loads (next chunk's in flight), turns the accumulator chunk into dY * gelu(g) and
        // dY * a * gelu'(g) and writes both with 256-bit stores (full sectors). No shared memory: the staged variants
        // of this epilogue (transposes or TMA tiles) cost the main loop 25-45 % through the shared-memory port.
        constexpr int CPW = BN / 64;
        const int c_first = half * CPW;
        const int colh = nt * BN + c_first * 32;
        const __half* hrow = p.H + (size_t)row * p.ldh;
        __half* crow = p.C + (size_t)row * p.ldc;
        const bool al32 = e_al32;
        {  // the saved pre-activations were written a forward pass ago: start them towards L2 under the main loop
          if (row_ok && colh < e_N) {
#pragma unroll
            for (int j = 0; j < 2; ++j) {
              if (colh + j * 64 < e_N) {
                asm volatile("prefetch.global.L2 [%0];" ::"l"(hrow + colh + j * 64));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(hrow + p.F + colh + j * 64));
              }
            }
          }
        }
        uint32_t va[2][16], vg[2][16];
        auto fetch = [&](int ci, uint32_t (&da)[16], uint32_t (&dg)[16]) {
          const int col0 = colh + ci * 32;
          int n = e_N - col0;
          n = n > 32 ? 32 : n;
          if (row_ok && n > 0) {
            ldg_row32(hrow + col0, da, n, al32);
            ldg_row32(hrow + p.F + col0, dg, n, al32);
          }
        };
        fetch(0, va[0], vg[0]);
        mbar_wait_backoff(&tfull[as], aph, p.dbg & 16 ? 0u : 100u);
        tc_fence_after();
        if (tl && ew == 0 && lane == 0 && it <= 15) tl[4 + 4 * (it - 1) + 2] = clock64();
#pragma unroll
        for (int ci = 0; ci < CPW; ++ci) {
          const int col0 = colh + ci * 32;
          if (col0 < e_N) {  // warp-uniform
            uint32_t r[32];
            tmem_ld32(tacc + (c_first + ci) * 32, r);
            if (ci + 1 < CPW) fetch(ci + 1, va[(ci + 1) & 1], vg[(ci + 1) & 1]);
            tmem_ld_wait();
            uint32_t oa[16], og[16];
#pragma unroll
            for (int k = 0; k < 16; ++k) {
              const float2 a2 = unpack_half2(va[ci & 1][k]);
              const float2 g2 = unpack_half2(vg[ci & 1][k]);
              const float dy0 = __uint_as_float(r[2 * k]) * e_alpha, dy1 = __uint_as_float(r[2 * k + 1]) * e_alpha;
              float gl0, dgl0, gl1, dgl1;
              gelu_erf_both(g2.x, gl0, dgl0);
              gelu_erf_both(g2.y, gl1, dgl1);
              oa[k] = pack_half2(dy0 * gl0, dy1 * gl1);
              og[k] = pack_half2(dy0 * a2.x * dgl0, dy1 * a2.y * dgl1);
            }
            if (row_ok) {
              int n = e_N - col0;
              n = n > 32 ? 32 : n;
              stg_row32(crow + col0, oa, n, al32);
              stg_row32(crow + p.F + col0, og, n, al32);
            }
          }
        }
        if (tl && ew == 0 && lane == 0 && it <= 15) tl[4 + 4 * (it - 1) + 3] = clock64();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_leader(&tempty[as]);
        continue;
      }
      if constexpr ((EPI == DB1_EPI_PLAIN && CL == 1) || EPI == DB1_EPI_QKV) {
        constexpr int CPW = BN / 64;  // chunks per warp
        const int c_first = half * CPW;
        Half8 rs[CPW][4];
        if (EPI == DB1_EPI_PLAIN && p.resid != nullptr) {  // independent of the accumulator: fetched before the wait
#pragma unroll
          for (int ci = 0; ci < CPW; ++ci) {
            const int col0 = nt * BN + (c_first + ci) * 32;
            if (col0 < p.N)
              warp_load_tile(stg, lane, rs[ci], p.resid + (size_t)row_base * p.ldr + col0, p.ldr, p.M - row_base,
                             p.N - col0);
          }
        }
        mbar_wait_backoff(&tfull[as], aph, p.dbg & 16 ? 0u : 100u);
        tc_fence_after();
        if (tl && ew == 0 && lane == 0 && it <= 15) tl[4 + 4 * (it - 1) + 2] = clock64();
#pragma unroll
        for (int ci = 0; ci < CPW; ++ci) {
          const int c = c_first + ci;
          uint32_t r[32];
          if (p.dbg & 2) continue;
          tmem_ld32(tacc + c * 32, r);
          tmem_ld_wait();
          const int col0 = nt * BN + c * 32;
          if (col0 >= p.N) break;  // warp-uniform
          const int rows_valid = p.M - row_base;
          Half8 h0[4], h1[4];
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const int col = col0 + g * 8;
            float f[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) f[i] = __uint_as_float(r[g * 8 + i]) * p.alpha;
            if (EPI == DB1_EPI_QKV) {
              if (col < p.d_model) {
                float uu[8], vv[8], o[8];
                half8_to_float(ld_half8(p.u + col), uu);
                half8_to_float(ld_half8(p.v + col), vv);
#pragma unroll
                for (int i = 0; i < 8; ++i) o[i] = f[i] + uu[i];
                h0[g] = float_to_half8(o);
#pragma unroll
                for (int i = 0; i < 8; ++i) o[i] = f[i] + vv[i];
                h1[g] = float_to_half8(o);
              } else {
                h1[g] = float_to_half8(f);
              }
            } else {
              if (col < p.N) {
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
                if (p.resid && row_ok) {
                  float b[8];
                  half8_to_float(rs[ci][g], b);
#pragma unroll
                  for (int i = 0; i < 8; ++i) f[i] += b[i];
                }
              }
              h0[g] = float_to_half8(f);
            }
          }
          if (p.dbg & 1) continue;
          if (EPI == DB1_EPI_QKV) {
            // columns < d_model are written twice (q+u at col, q+v at d_model+col); k, v shift right by d_model
            __half* crow = p.C + (size_t)row_base * p.ldc;
            if (col0 < p.d_model) warp_store_tile(stg, lane, h0, crow + col0, p.ldc, rows_valid, 32, false);
            warp_store_tile(stg, lane, h1, crow + p.d_model + col0, p.ldc, rows_valid, 32, false);
          } else {
            warp_store_tile(stg, lane, h0, p.C + zoff + (size_t)row_base * p.ldc + col0, p.ldc, rows_valid, p.N - col0,
                            p.accumulate != 0);
          }
        }
      } else if (EPI == DB1_EPI_DS) {
        // dS = P * (dP - Drow) * alpha on the causal / windowed region, 0 elsewhere; second copy in
        // relative-position order (the adjoint of _rel_shift, transformer_xl.py:98-110).
        constexpr int CPW = BN / 64;
        const int c_first = half * CPW;
        const long long zlin = (long long)t.z2 * p.Z1 + t.z1;
        const float drow = row_ok ? p.Drow[zlin * p.M + row] : 0.f;
        Half8 pp[2][4];
        auto load_p = [&](int c, Half8 (&dst)[4]) {
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const int col = nt * BN + c * 32 + g * 8;
            if (row_ok && col <= row && col < p.N) dst[g] = ld_half8(p.P + zoff + (size_t)row * p.ldc + col);
          }
        };
        load_p(c_first, pp[0]);
        mbar_wait_backoff(&tfull[as], aph, p.dbg & 16 ? 0u : 100u);
        tc_fence_after();
#pragma unroll
        for (int ci = 0; ci < CPW; ++ci) {
          const int c = c_first + ci;
          const int col0 = nt * BN + c * 32;
          if (col0 > row_base + 31 || col0 >= p.N) break;  // warp-uniform: chunk entirely above the diagonal
          uint32_t r[32];
          tmem_ld32(tacc + c * 32, r);
          if (ci + 1 < CPW) load_p(c + 1, pp[(ci + 1) & 1]);
          tmem_ld_wait();
          Half8 hv[4];
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const int col = col0 + g * 8;
            float pv[8], o[8];
            const bool any = row_ok && col <= row && col < p.N;
            half8_to_float(pp[ci & 1][g], pv);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int j = col + i;
              const bool ok = any && j <= row && (row - j) < p.window;
              o[i] = ok ? pv[i] * (__uint_as_float(r[g * 8 + i]) - drow) * p.alpha : 0.f;
            }
            hv[g] = float_to_half8(o);
          }
          warp_store_tile(stg, lane, hv, p.C + zoff + (size_t)row_base * p.ldc + col0, p.ldc, p.M - row_base,
                          p.N - col0, false);
        }
      } else {
        // GeGLU forward / backward: accumulator columns [0,BN/2) pair with [BN/2,BN) (forward) or the tile's
        // BN output columns pair with saved a|g (backward).
        constexpr int HALF = BN / 2;
        constexpr int NCH = (EPI == DB1_EPI_GEGLU ? HALF : BN) / 32;
        constexpr int CPW = NCH / 2;
        const int c_first = half * CPW;
        Half8 la[2][4], lg[2][4];  // saved pre-activations a | g of the chunk in flight (coalesced pieces)
        auto issue_h = [&](int c, Half8 (&da)[4], Half8 (&dg)[4]) {
          if (EPI == DB1_EPI_DGEGLU) {
            const int n0 = nt * BN + c * 32;
            const __half* hrow = p.H + (size_t)row_base * p.ldh + n0;
            tile_issue(lane, da, hrow, p.ldh, p.M - row_base, p.N - n0);
            tile_issue(lane, dg, hrow + p.F, p.ldh, p.M - row_base, p.N - n0);
          }
        };
        if (EPI == DB1_EPI_DGEGLU) {
          // The saved pre-activations were written a whole forward pass ago (DRAM-resident): pull this warp's
          // 32 rows x 128 columns of a and g into L2 while the tile's main loop is still running, so the LDGs below
          // (issued only one chunk ahead - registers) see L2 latency instead of DRAM latency.
          const int r = row_base + lane;
          const int n0 = nt * BN + c_first * 32;
          if (r < p.M && n0 < p.N) {
            const __half* hp = p.H + (size_t)r * p.ldh + n0;
#pragma unroll
            for (int j = 0; j < 2; ++j) {
              if (n0 + j * 64 < p.N) {
                asm volatile("prefetch.global.L2 [%0];" ::"l"(hp + j * 64));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(hp + p.F + j * 64));
              }
            }
          }
        }
        issue_h(c_first, la[0], lg[0]);
        mbar_wait_backoff(&tfull[as], aph, p.dbg & 16 ? 0u : 100u);
        tc_fence_after();
#pragma unroll
        for (int ci = 0; ci < CPW; ++ci) {
          const int c = c_first + ci;
          uint32_t ra[32];
          tmem_ld32(tacc + c * 32, ra);
          if (EPI == DB1_EPI_GEGLU) {
            uint32_t rg[32];
            tmem_ld32(tacc + HALF + c * 32, rg);
            tmem_ld_wait();
            const int n0 = nt * HALF + c * 32;  // column within [0,F)
            Half8 xa[4], xg[4], xy[4];
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
              // round the pre-activations to fp16 first so that forward and backward see the same values
              xa[g] = float_to_half8(a);
              xg[g] = float_to_half8(gg);
              half8_to_float(xa[g], a);
              half8_to_float(xg[g], gg);
#pragma unroll
              for (int i = 0; i < 8; ++i) y[i] = a[i] * gelu_erf(gg[i]);
              xy[g] = float_to_half8(y);
            }
            const int rows_valid = p.M - row_base;
            warp_store_tile(stg, lane, xa, p.H + (size_t)row_base * p.ldh + n0, p.ldh, rows_valid, 32, false);
            warp_store_tile(stg, lane, xg, p.H + (size_t)row_base * p.ldh + p.F + n0, p.ldh, rows_valid, 32, false);
            warp_store_tile(stg, lane, xy, p.C + (size_t)row_base * p.ldc + n0, p.ldc, rows_valid, 32, false);
          } else {
            if (ci + 1 < CPW) issue_h(c + 1, la[(ci + 1) & 1], lg[(ci + 1) & 1]);
            Half8 ha[4], hg[4];
            tile_finish(stg, lane, la[ci & 1], ha);
            tile_finish(stg, lane, lg[ci & 1], hg);
            tmem_ld_wait();
            const int n0 = nt * BN + c * 32;
            if (n0 >= p.N) break;  // warp-uniform
            Half8 xda[4], xdg[4];
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              float dy[8], a[8], gg[8], da[8], dg[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) dy[i] = __uint_as_float(ra[g * 8 + i]) * p.alpha;
              half8_to_float(ha[g], a);
              half8_to_float(hg[g], gg);
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                float gl, dgl;
                gelu_erf_both(gg[i], gl, dgl);
                da[i] = dy[i] * gl;
                dg[i] = dy[i] * a[i] * dgl;
              }
              xda[g] = float_to_half8(da);
              xdg[g] = float_to_half8(dg);
            }
            const int rows_valid = p.M - row_base;
            warp_store_tile(stg, lane, xda, p.C + (size_t)row_base * p.ldc + n0, p.ldc, rows_valid, p.N - n0, false);
            warp_store_tile(stg, lane, xdg, p.C + (size_t)row_base * p.ldc + p.F + n0, p.ldc, rows_valid, p.N - n0, false);
          }
        }
      }
      // release this accumulator stage back to the MMA warp
      if (tl && ew == 0 && lane == 0 && it <= 15) tl[4 + 4 * (it - 1) + 3] = clock64();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (CL >= 2) mbar_arrive_leader(&tempty[as]);
        else mbar_arrive(&tempty[as]);
      }
    }
  }

  if (CL >= 2 && p.tma_epi && warp >= 2 && ((warp - 2) & 3) == 0 && lane == 0)
    bulk_wait_group_read<0>();  // shared memory must outlive the last tile stores
  if (tl && threadIdx.x == 64) tl[1] = clock64();
  tc_fence_before();
  __syncthreads();
  if (CL >= 2) cluster_sync_all();  // peers may still multicast into / arrive on this CTA's shared memory
  if (warp == 1) {
    tc_fence_after();
    if (CL >= 2) tmem_dealloc2<Cfg::TMEM_COLS>(tmem_base);
    else tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
  }
}

// Stream-K workspace registered by the caller (db1_gemm_set_workspace): counters first, then the partial-tile slots.
struct SkWorkspace {
  void* base = nullptr;
  long long bytes = 0;
  int device = -1;
};
static SkWorkspace g_sk_ws;
constexpr long long SK_CNT_BYTES = 4096;
constexpr int SK_MAX_SPLIT = 4;     // a tile is cut into at most this many k-ranges ...
constexpr int SK_MIN_KB = 4;        // ... of at least this many 64-wide k-blocks

// 2-D map of a row-major [rows, cols] fp16 matrix for the TMA epilogue: box = 32 columns x 128 rows, 64-byte swizzle.
static int make_tile_map(CUtensorMap* tm, const void* base, long long rows, long long cols, long long ld) {
  uint64_t dims[2] = {(uint64_t)cols, (uint64_t)rows};
  uint64_t str[1] = {(uint64_t)ld * 2};
  uint32_t box[2] = {32, 128};
  return make_tmap_f16(tm, base, 2, dims, str, box, 64);
}

// 3-D map (column, part, row) over a row-major [rows, 2F] matrix seen as two [rows, F] halves (GeGLU's [a | g]): the
// column dimension is clipped at F, so a ragged last chunk never spills into the other half.
static int make_part_map(CUtensorMap* tm, const void* base, long long rows, long long F, long long ld) {
  uint64_t dims[3] = {(uint64_t)F, 2, (uint64_t)rows};
  uint64_t str[2] = {(uint64_t)F * 2, (uint64_t)ld * 2};
  uint32_t box[3] = {32, 1, 128};
  return make_tmap_f16(tm, base, 3, dims, str, box, 64);
}

// Stream-K tail: worth it when the last data-parallel wave would leave a visible share of the pairs idle.
static bool sk_choose(long long tiles, int slots, int KB, int* R_out, int* GS_out) {
  const int R = (int)(tiles % slots);
  if (R == 0) return false;
  const long long waves = (tiles + slots - 1) / slots;
  const double idle = 1.0 - (double)tiles / (double)(waves * slots);  // share of the data-parallel schedule wasted
  int split = KB / SK_MIN_KB;
  if (split > SK_MAX_SPLIT) split = SK_MAX_SPLIT;
  if (idle < 0.04 || split < 2) return false;
  const long long gs = (long long)R * split;
  *R_out = R;
  *GS_out = (int)(gs < slots ? gs : slots);
  return true;
}

template <int BN, int EPI, int CL>
static int launch_gemm_cl(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmB64, const GemmParams& p_in,
                          cudaStream_t stream, int max_groups) {
  using Cfg = GemmCfg<BN, CL>;
  static bool configured = false;
  if (!configured) {
    DB1_CUDA(cudaFuncSetAttribute(gemm_kernel<BN, EPI, CL>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  Cfg::SMEM_BYTES));
    configured = true;
  }
  GemmParams p = p_in;
  if (CL == 1 && p.dot_out != nullptr)
    return set_err(-1, "gemm(dot): the row-dot side output needs the CTA-pair kernel (unset DB1_GEMM_NO_CLUSTER)");
  const long long MT = cdiv(p.M, BM);
  const long long NT = (EPI == DB1_EPI_GEGLU) ? p.F / (BN / 2) : cdiv(p.N, BN);
  long long tiles = (CL >= 2 ? (MT + CL - 1) / CL : MT) * NT * p.Z1 * (p.reduce_z2 ? 1 : p.Z2);
  const int slots = max_groups > 0 ? max_groups : sm_count() / CL;
  int grid = (int)(tiles > slots ? slots : tiles) * CL;
  if (CL == 2 && g_sk_ws.base != nullptr && getenv("DB1_GEMM_SK")) {
    int dev = -1;
    cudaGetDevice(&dev);
    const long long need = SK_CNT_BYTES + (long long)slots * CL * BM * BN * 4;
    int R = 0, GS = 0;
    if (dev == g_sk_ws.device && g_sk_ws.bytes >= need && 2 * slots * (long long)sizeof(unsigned int) <= SK_CNT_BYTES &&
        sk_choose(tiles, slots, cdiv(p.K, BK), &R, &GS)) {
      p.sk_tiles = R;
      p.sk_pairs = GS;
      p.sk_cnt = reinterpret_cast<unsigned int*>(g_sk_ws.base);
      p.sk_ws = reinterpret_cast<float*>(reinterpret_cast<char*>(g_sk_ws.base) + SK_CNT_BYTES);
      grid = slots * CL;  // every pair takes part, also when there are fewer tiles than pairs
    }
  }
  CUtensorMap tmC = tmA, tmIn = tmA;  // placeholders unless the TMA epilogue is used
  if (CL >= 2 && EPI == DB1_EPI_PLAIN) {  // launch_gemm has checked tma_epilogue_ok()
    const void* in = p.resid ? (const void*)p.resid : (p.accumulate ? (const void*)p.C : nullptr);  // dot: resid = X
    const long long ldin = p.resid ? p.ldr : p.ldc;
    int e = make_tile_map(&tmC, p.C, p.M, p.N, p.ldc);
    if (e) return e;
    if (in && (e = make_tile_map(&tmIn, in, p.M, p.N, ldin))) return e;
    p.tma_epi = 1;
    p.tma_in = in ? 1 : 0;
  }
  if ((p.dbg & 8) && g_sk_ws.base != nullptr && p.sk_tiles == 0)
    p.sk_ws = reinterpret_cast<float*>(reinterpret_cast<char*>(g_sk_ws.base) + SK_CNT_BYTES);
  else
    p.dbg &= ~8;
  DB1_CUDA(launch_pdl(gemm_kernel<BN, EPI, CL>, dim3(grid), dim3(GEMM_THREADS), Cfg::SMEM_BYTES, stream, CL, tmA, tmB, tmB64,
                      tmC, tmIn, p));
  return 0;
}

// The CTA-pair PLAIN kernel has only the TMA epilogue: C (and the addend) must be addressable by a tensor map, and
// "residual + accumulate" (two addends) stays on the single-CTA kernel.
static bool tma_epilogue_ok(const GemmParams& p, int epi) {
  if (epi != DB1_EPI_PLAIN) return true;
  if (p.resid && p.accumulate) return false;
  if (((uintptr_t)p.C & 15) != 0) return false;
  if (p.resid && (((uintptr_t)p.resid & 15) != 0 || p.ldr % 8 != 0)) return false;
  return true;
}

// How many 4-CTA clusters of this kernel the device can hold at once (GPC granularity: fewer than SMs / 4).
template <int BN, int EPI>
static int max_clusters4() {
  static int n = -1;
  if (n < 0) {
    using Cfg = GemmCfg<BN, 4>;
    n = 0;
    if (cudaFuncSetAttribute(gemm_kernel<BN, EPI, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES) ==
        cudaSuccess) {
      cudaLaunchConfig_t cfg;
      memset(&cfg, 0, sizeof(cfg));
      cfg.gridDim = dim3(sm_count() / 4 * 4);
      cfg.blockDim = dim3(GEMM_THREADS);
      cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = 4;
      attr[0].val.clusterDim.y = 1;
      attr[0].val.clusterDim.z = 1;
      cfg.attrs = attr;
      cfg.numAttrs = 1;
      int k = 0;
      if (cudaOccupancyMaxActiveClusters(&k, gemm_kernel<BN, EPI, 4>, &cfg) == cudaSuccess) n = k;
    }
    cudaGetLastError();
    if (getenv("DB1_GEMM_VERBOSE")) fprintf(stderr, "[db1] gemm<%d,%d>: %d resident 4-CTA clusters\n", BN, EPI, n);
  }
  return n;
}

// CTA pairs (cta_group::2: one 256 x BN MMA over two vertically adjacent 128-row tiles, each CTA staging half of the B
// tile) cut both the L2 -> SM and the shared-memory operand traffic per tile from 48 KB to 32 KB per k-block; used
// whenever both CTAs see the same k-range. Two pairs in one 4-CTA cluster (512 x BN) additionally share the B tile by
// TMA multicast (24 KB per CTA and k-block): the main loops are L2-throughput bound (lts sectors / cycle within 15 % of
// the chip cap, ncu), so fewer L2 reads per flop is the remaining lever.
template <int BN, int EPI>
static int launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmB64, const GemmParams& p,
                       cudaStream_t stream) {
  const bool batched = p.Z1 * p.Z2 > 1;
  if (BN == 256 && EPI != DB1_EPI_DS && !batched && p.k_mode == DB1_K_FULL && !p.skip_upper && p.M > BM &&
      ((EPI != DB1_EPI_PLAIN && EPI != DB1_EPI_DGEGLU) || tma_epilogue_ok(p, EPI)) && !getenv("DB1_GEMM_NO_CLUSTER")) {
    constexpr bool HAS4 = (BN == 256 && (EPI == DB1_EPI_PLAIN || EPI == DB1_EPI_DGEGLU));
    if (HAS4 && p.M > 3 * BM) {
      const char* e = getenv("DB1_GEMM_CL");
      // opt-in: measured equal to CTA pairs for both epilogues (33 clusters = 132 of 148 SMs; multicast saves L2 reads,
      // not the bytes each SM receives - and the SM's inbound port is what the main loop saturates)
      const int want = e ? atoi(e) : 2;
      const int nc = max_clusters4<BN, HAS4 ? EPI : DB1_EPI_PLAIN>();
      if (want == 4 && nc >= 30)
        return launch_gemm_cl<BN, EPI, HAS4 ? 4 : 2>(tmA, tmB, tmB64, p, stream, nc);
    }
    return launch_gemm_cl<BN, EPI, (BN == 256 && EPI != DB1_EPI_DS) ? 2 : 1>(tmA, tmB, tmB64, p, stream, 0);
  }
  return launch_gemm_cl<BN, EPI, 1>(tmA, tmB, tmB64, p, stream, 0);
}

// 4-D map (inner, rows, z1, z2). A broadcast batch dim (stride 0) is encoded as a dim of size 1.
static int make_operand_map(CUtensorMap* tm, const void* base, int mn_major, long long mn, long long k, long long ld,
                            int Z1, long long z1s, int Z2, long long z2s, int mn_box_rows) {
  uint64_t dims[4], str[3];
  uint32_t box[4];
  if (!mn_major) {
    dims[0] = (uint64_t)k; dims[1] = (uint64_t)mn; box[0] = 64; box[1] = (uint32_t)mn_box_rows;
  } else {
    dims[0] = (uint64_t)mn; dims[1] = (uint64_t)k; box[0] = 64; box[1] = 64;
  }
  dims[2] = z1s ? (uint64_t)Z1 : 1; dims[3] = z2s ? (uint64_t)Z2 : 1;
  box[2] = 1; box[3] = 1;
  str[0] = (uint64_t)ld * 2;
  str[1] = z1s ? (uint64_t)z1s * 2 : str[0] * dims[1];
  str[2] = z2s ? (uint64_t)z2s * 2 : (z1s ? str[1] * dims[2] : str[0] * dims[1]);
  return make_tmap_f16(tm, base, 4, dims, str, box);
}

}  // namespace db1

using namespace db1;

extern "C" int db1_gemm_set_workspace(void* ws, long long bytes) {
  if (ws == nullptr || bytes <= 0) {
    g_sk_ws = SkWorkspace();
    return 0;
  }
  DB1_CHECK_ARG(((uintptr_t)ws & 15) == 0, "gemm workspace %p not 16-byte aligned", ws);
  int dev = -1;
  DB1_CUDA(cudaGetDevice(&dev));
  g_sk_ws.base = ws;
  g_sk_ws.bytes = bytes;
  g_sk_ws.device = dev;
  return 0;
}

// Host-side view of the stream-K schedule (tests / tooling): the decision for a tile count and the work list of one pair.
extern "C" int db1_gemm_sk_choose(long long tiles, int pairs, int KB, int* sk_tiles, int* sk_pairs) {
  DB1_CHECK_ARG(tiles > 0 && pairs > 0 && KB > 0 && sk_tiles && sk_pairs, "sk_choose: bad argument");
  *sk_tiles = 0;
  *sk_pairs = 0;
  return sk_choose(tiles, pairs, KB, sk_tiles, sk_pairs) ? 1 : 0;
}
/* out[14] = nseg, T_dp, tile[2], kb0[2], kb1[2], role[2], peer[2], npeer[2] */
extern "C" int db1_gemm_sk_plan(int pair, long long tiles, int KB, int sk_tiles, int sk_pairs, int* out) {
  DB1_CHECK_ARG(out && pair >= 0 && tiles > 0 && KB > 0, "sk_plan: bad argument");
  SkPlan pl;
  memset(&pl, 0, sizeof(pl));
  sk_plan_build(&pl, sk_tiles, sk_pairs, pair, (int)tiles, KB);
  memcpy(out, &pl, sizeof(pl));
  return 0;
}

extern "C" long long db1_gemm_workspace_bytes(void) {
  return SK_CNT_BYTES + (long long)(sm_count_physical() / 2) * 2 * BM * 256 * 4;
}

extern "C" int db1_gemm_f16(const db1_gemm_desc* d, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  DB1_CHECK_ARG(d != nullptr, "gemm: null descriptor");
  const int M = d->M, N = d->N, K = d->K, epilogue = d->epilogue;
  DB1_CHECK_ARG(M > 0 && N > 0 && K > 0, "gemm: bad shape M=%d N=%d K=%d", M, N, K);
  DB1_CHECK_ARG(d->A && d->B && d->C, "gemm: null operand");
  DB1_CHECK_ARG(d->lda % 8 == 0 && d->ldb % 8 == 0 && d->ldc % 8 == 0,
                "gemm: leading dims must be multiples of 8 (%lld %lld %lld)", (long long)d->lda, (long long)d->ldb,
                (long long)d->ldc);
  DB1_CHECK_ARG(epilogue >= 0 && epilogue <= 4, "gemm: unknown epilogue %d", epilogue);
  DB1_CHECK_ARG(d->drop_p >= 0.f && d->drop_p < 1.f, "gemm: dropout p=%f out of range", d->drop_p);
  DB1_CHECK_ARG(d->k_mode >= 0 && d->k_mode <= 3, "gemm: unknown k_mode %d", d->k_mode);
  const int Z1 = d->Z1 > 0 ? d->Z1 : 1, Z2 = d->Z2 > 0 ? d->Z2 : 1;
  const bool batched = Z1 * Z2 > 1;
  // one to eight activation rows (the decode step): a stream over B on the CUDA cores, HBM-bound (csrc/skinny.cu)
  if (skinny_gemm_applies(d)) return skinny_gemm(d, stream);
  DB1_CHECK_ARG(d->ln_gamma == nullptr, "gemm: LayerNorm-on-load (ln_gamma) exists on the few-row path only (M <= 8)");
  DB1_CHECK_ARG((d->a_z1 % 8 == 0) && (d->a_z2 % 8 == 0) && (d->b_z1 % 8 == 0) && (d->b_z2 % 8 == 0) &&
                    (d->c_z1 % 8 == 0) && (d->c_z2 % 8 == 0),
                "gemm: batch strides must be multiples of 8 elements");
  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.M = M; p.N = N; p.K = K; p.a_mn = d->a_mn ? 1 : 0; p.b_mn = d->b_mn ? 1 : 0;
  p.Z1 = Z1; p.Z2 = Z2;
  p.a_z1on = d->a_z1 != 0; p.a_z2on = d->a_z2 != 0; p.b_z1on = d->b_z1 != 0; p.b_z2on = d->b_z2 != 0;
  p.reduce_z2 = d->reduce_z2 ? 1 : 0; p.k_mode = d->k_mode; p.skip_upper = d->skip_upper ? 1 : 0;
  p.alpha = d->alpha; p.C = (__half*)d->C; p.ldc = d->ldc; p.c_z1 = d->c_z1; p.c_z2 = d->c_z2;
  p.bias = (const __half*)d->bias; p.resid = (const __half*)d->resid; p.ldr = d->ldr; p.accumulate = d->accumulate;
  p.drop_thr16 = (uint32_t)(d->drop_p * 65536.0f + 0.5f);
  p.drop_scale = p.drop_thr16 ? 65536.0f / (65536.0f - (float)p.drop_thr16) : 1.0f;
  p.seed = d->seed; p.u = (const __half*)d->u; p.v = (const __half*)d->v; p.d_model = d->d_model;
  p.H = (__half*)d->H; p.ldh = d->ldh; p.F = d->F;
  p.P = (const __half*)d->P; p.C2 = (__half*)d->C2; p.Drow = d->Drow; p.window = d->window;
  p.snake = (d->k_mode != DB1_K_FULL && !d->skip_upper && !getenv("DB1_GEMM_NO_SNAKE")) ? 1 : 0;
  if (d->dot_out != nullptr) {
    DB1_CHECK_ARG(epilogue == DB1_EPI_PLAIN && d->dot_with && !d->resid && !d->accumulate && !d->bias && d->drop_p == 0.f &&
                      d->alpha == 1.0f && !batched && d->dot_H > 0 && d->dot_L > 0 && N == d->dot_H * 128 &&
                      M % d->dot_L == 0 && d->ld_dot % 8 == 0 && M > 3 * 128 && d->bn_hint != 128 &&
                      2LL * cdiv(M, BM) * cdiv(N, 256) > sm_count(),
                  "gemm(dot): plain un-batched GEMM at CTA-pair size with N == dot_H * 128 (head dim 128) required");
    p.resid = (const __half*)d->dot_with;  // travels through the epilogue's TMA addend path
    p.ldr = d->ld_dot;
    p.dot_out = d->dot_out;
    p.dot_L = d->dot_L;
    p.dot_H = d->dot_H;
  }
  { const char* e = getenv("DB1_GEMM_DBG"); p.dbg = e ? atoi(e) : 0; }
  if (p.reduce_z2) DB1_CHECK_ARG(d->c_z2 == 0, "gemm: reduce_z2 needs c_z2 == 0");

  int BNsel = 256;
  if (epilogue == DB1_EPI_PLAIN || epilogue == DB1_EPI_DGEGLU) {
    if (d->bn_hint == 128 || (d->bn_hint == 0 && N <= 128)) BNsel = 128;
    // small problems: if 128 x 256 tiles leave at least half of the SMs idle, halve the tile width (twice the CTAs)
    if (d->bn_hint == 0 && !batched && 2LL * cdiv(M, BM) * cdiv(N, 256) <= sm_count()) BNsel = 128;
  }
  if (epilogue == DB1_EPI_QKV) {
    DB1_CHECK_ARG(d->u && d->v && d->d_model > 0 && N == 3 * d->d_model, "gemm(qkv): need u, v and N == 3*d_model");
    DB1_CHECK_ARG(d->d_model % 128 == 0, "gemm(qkv): d_model %d must be a multiple of 128", d->d_model);
    DB1_CHECK_ARG(!d->accumulate && !d->bias && !d->resid && d->drop_p == 0.f && !batched,
                  "gemm(qkv): unsupported fused option");
    BNsel = (d->d_model % 256 == 0 && d->bn_hint != 128) ? 256 : 128;
  }
  if (epilogue == DB1_EPI_GEGLU) {
    DB1_CHECK_ARG(d->H && d->F > 0 && N == 2 * d->F && d->F % 128 == 0 && !d->b_mn && !batched,
                  "gemm(geglu): need H, N == 2F, F %% 128 == 0, K-major B, no batching");
  }
  if (epilogue == DB1_EPI_DGEGLU) {
    DB1_CHECK_ARG(d->H && d->F > 0 && N == d->F && d->F % 8 == 0 && !batched, "gemm(dgeglu): need H and N == F");
  }
  if (epilogue == DB1_EPI_DS) {
    DB1_CHECK_ARG(d->P && d->Drow && M == N && N % 8 == 0 && d->window > 0,
                  "gemm(ds): need P, Drow, square M == N (multiple of 8) and window > 0");
    BNsel = 128;
  }
  if (epilogue == DB1_EPI_PLAIN && (d->bias || d->resid || p.drop_thr16))
    DB1_CHECK_ARG(N % 8 == 0, "gemm: fused bias/residual/dropout need N %% 8 == 0 (N=%d)", N);
  if (d->resid) DB1_CHECK_ARG(!batched, "gemm: residual only for un-batched calls");

  CUtensorMap tmA, tmB;
  int e = make_operand_map(&tmA, d->A, p.a_mn, M, K, d->lda, Z1, d->a_z1, Z2, d->a_z2, 128);
  if (e) return e;
  // EPI_GEGLU addresses rows up to 2F; the tensor's row count is N in every case
  e = make_operand_map(&tmB, d->B, p.b_mn, N, K, d->ldb, Z1, d->b_z1, Z2, d->b_z2, 128);
  if (e) return e;
  CUtensorMap tmB64 = tmB;  // K-major B in 64-row boxes: the quarter tiles the 4-CTA clusters multicast
  if (!p.b_mn && (e = make_operand_map(&tmB64, d->B, 0, N, K, d->ldb, Z1, d->b_z1, Z2, d->b_z2, 64))) return e;

  switch (epilogue) {
    case DB1_EPI_PLAIN:
      return BNsel == 256 ? launch_gemm<256, DB1_EPI_PLAIN>(tmA, tmB, tmB64, p, stream)
                          : launch_gemm<128, DB1_EPI_PLAIN>(tmA, tmB, tmB64, p, stream);
    case DB1_EPI_QKV:
      return BNsel == 256 ? launch_gemm<256, DB1_EPI_QKV>(tmA, tmB, tmB64, p, stream)
                          : launch_gemm<128, DB1_EPI_QKV>(tmA, tmB, tmB64, p, stream);
    case DB1_EPI_GEGLU:
      return launch_gemm<256, DB1_EPI_GEGLU>(tmA, tmB, tmB64, p, stream);
    case DB1_EPI_DGEGLU:
      return BNsel == 256 ? launch_gemm<256, DB1_EPI_DGEGLU>(tmA, tmB, tmB64, p, stream)
                          : launch_gemm<128, DB1_EPI_DGEGLU>(tmA, tmB, tmB64, p, stream);
    case DB1_EPI_DS:
      return launch_gemm<128, DB1_EPI_DS>(tmA, tmB, tmB64, p, stream);
  }
  return set_err(-1, "gemm: unreachable");
}
