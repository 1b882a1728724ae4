// Fused relative-position ("Transformer-XL") causal attention forward for sm_100a.
//
//   S[i,j] = scale * ( (q_i+u).k_j + (q_i+v).r[j + L-1-i] )      (transformer_xl.py:161-174; the index on r IS
//   P      = softmax_j(S) on  0 <= i-j < window                    _rel_shift, :98-110, for qlen == klen == L)
//   O[i]   = sum_j P[i,j] v_j                                      (:209-225)
//
// Persistent: one CTA per SM walks a static, balanced list of (128-query tile, head, sequence) items. Key tiles are
// walked from the diagonal outwards; per step three tcgen05 MMAs run on TMEM accumulators:
//     S   = Qu . K_j^T                     (128x128, fp32)
//     BDc = Qv . Rchunk^T                  (128x128: the 128 new relative positions this step needs; the other
//                                           128 are the previous step's chunk, kept in a 2-slot TMEM ring)
//     O  += P . V_j                        (128xD)            [mode 0]   /   dP = dO . V_j^T   [mode 2]
// The per-row shift of the position term cannot be expressed by tcgen05.ld (lanes share the column address): each
// softmax thread loads the two 32-column blocks its row needs and runs a 5-stage barrel shifter on registers
// (shift = 31 - lane). Softmax is online (fp32, exp2 with folded scale, lazy rescale of O), no mask tensor exists:
// the causal / sliding-window predicate is computed from indices.
//
// Roles (320 threads): warp 0 = TMA producer, warp 1 = MMA issuer (one elected thread), warps 2..9 = softmax (two per
// TMEM lane quadrant; a thread owns one query row x half of the 128-key tile).
//
// MODE (template parameter):
//   0: writes O [B*L, ldo] fp16 and LSE2 [B,H,L] fp32 (log2 domain).
//   1: reads LSE2 and writes the normalised probabilities P [B,H,L,L] fp16 (recompute for the backward pass).
//   2: as 1, plus dP = dO . V^T on the tensor cores and dS = P * (dP - D) * scale (D given, or formed here from dO and O).
#include "../../include/db1_sm100.h"
#include "common.cuh"
#include "ptx.cuh"

namespace db1 {

constexpr int AT_THREADS = 320;  // TMA warp + MMA warp + 8 softmax warps
constexpr float NEG_BIG = -1e30f;
constexpr int STG_PITCH = 68;  // floats per staged row (64 + 4): 16-byte stores and odd-stride reads are conflict-free

struct AttnParams {
  int L, H, B, dh;
  int window;
  float scale_log2;
  __half* O;
  long long ldo;
  float* lse2;
  __half* P;
  int mode;
  __half* dS;          // mode 2
  const float* Drow;   // mode 2: rowsum(dO * O) [B,H,L], or NULL: computed here from dOp / Op at the start of every item
  const __half* dOp;   // mode 2 without Drow: dO and O as plain row-major [B*L, ld] matrices (head h at column h*dh)
  const __half* Op;
  long long lddo, ldop;
  float scale;         // mode 2: softmax scale (natural domain) applied to dS
  int q_start;         // first query row that is computed (memory-augmented inference: rows < q_start are memory)
  int tiled;           // modes 1 / 2: P / dS are written as contiguous 128 x 128 tiles (see db1_relattn_bwd_ds_tiled)
};

template <int D>
struct AttnSmem {
  static constexpr int TILE = 128 * D * 2;  // bytes of one [128][D] fp16 tile
  static constexpr int QU = 0;
  static constexpr int QV = QU + TILE;
  static constexpr int KT = QV + TILE;
  static constexpr int VT = KT + TILE;
  static constexpr int RT = VT + TILE;
  static constexpr int PT = RT + TILE;            // [128][128] fp16 = 32 KB
  static constexpr int STG = PT + 128 * 128 * 2;  // 4 warps x 32 rows x 68 floats
  static constexpr int BARS = STG + 4 * 32 * STG_PITCH * 4;
  static constexpr int TOTAL = BARS + 128;
};

template <int D, int MODE>
__global__ void __launch_bounds__(AT_THREADS, 1)
relattn_fwd_kernel(const __grid_constant__ CUtensorMap tmQu, const __grid_constant__ CUtensorMap tmQv,
                   const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
                   const __grid_constant__ CUtensorMap tmR, const __grid_constant__ CUtensorMap tmDO,
                   const AttnParams p) {
  using SM = AttnSmem<D>;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SM::BARS);
  uint64_t* bar_q = bars + 0;
  uint64_t* bar_k = bars + 1;
  uint64_t* bar_v = bars + 2;
  uint64_t* bar_kfree = bars + 3;
  uint64_t* bar_s = bars + 4;
  uint64_t* bar_p = bars + 5;
  uint64_t* bar_o = bars + 6;
  uint64_t* bar_s1 = bars + 7;
  uint64_t* bar_qfree = bars + 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // Persistent: one CTA per SM walks a static, balanced list of (query tile, head, sequence) items. Items are ordered
  // heaviest first (descending query tile = descending number of key steps) and dealt to the CTAs in snake order
  // (pass 0: CTA c takes item c, pass 1: item 2G-1-c, ...), so every CTA gets the same number of key steps +-1 tile.
  // TMEM, barriers and the smem ring live across items; barrier phases are tracked by global step / item counters.
  const int nq = (p.L + 127) / 128;
  const int HB = p.H * p.B;
  const int n_items = (nq - p.q_start / 128) * HB;  // query tiles that contain rows >= q_start
  const int G = gridDim.x;
  auto item_of = [&](int pass) -> int {
    const int c = (pass & 1) ? (G - 1 - (int)blockIdx.x) : (int)blockIdx.x;
    const int k = pass * G + c;
    return k < n_items ? k : -1;
  };
  const int n_pass = (n_items + G - 1) / G;
  struct Item {
    int I, h, b, I0, nsteps;
  };
  auto decode = [&](int k) -> Item {
    Item t;
    t.I = nq - 1 - k / HB;
    const int hb = k % HB;
    t.h = hb % p.H;
    t.b = hb / p.H;
    t.I0 = t.I * 128;
    int jlo = t.I0 - p.window + 1;  // oldest key any row of this tile may attend
    if (jlo < 0) jlo = 0;
    t.nsteps = t.I - jlo / 128 + 1;
    return t;
  };

  if ((smem_u32(smem) & 1023u) != 0) __trap();

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQu);
    tma_prefetch_desc(&tmQv);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    tma_prefetch_desc(&tmR);
    if (MODE == 2) tma_prefetch_desc(&tmDO);
  }
  if (warp == 1) {
    if (lane == 0) {
      mbar_init(bar_q, 1);
      mbar_init(bar_k, 1);
      mbar_init(bar_v, 1);
      mbar_init(bar_kfree, 1);
      mbar_init(bar_s, 1);
      mbar_init(bar_p, 8);
      mbar_init(bar_o, 1);
      mbar_init(bar_s1, 8);
      mbar_init(bar_qfree, 1);
      mbar_fence_init();
    }
    __syncwarp();
    tmem_alloc<512>(tmem_slot);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();  // prologue done: from here on global memory is read / written
  pdl_wait();
  const uint32_t T_S = tmem_base;          // 128 columns
  const uint32_t T_BD = tmem_base + 128;   // 2 x 128 columns
  const uint32_t T_O = tmem_base + 384;    // D columns

  constexpr int NSLAB = D / 64;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int gs = 0, ti = 0;  // global key-step / item counters (barrier phases)
      for (int pass = 0; pass < n_pass; ++pass) {
        const int k = item_of(pass);
        if (k < 0) continue;
        const Item t = decode(k);
        if (ti > 0) mbar_wait(bar_qfree, (ti - 1) & 1);  // the previous item's last score (and dP) MMAs have read Q / dO
        mbar_expect_tx(bar_q, (MODE == 2 ? 3 : 2) * SM::TILE);
#pragma unroll
        for (int s = 0; s < NSLAB; ++s) {
          tma_load_4d(smem + SM::QU + s * 16384, &tmQu, bar_q, s * 64, t.I0, t.h, t.b);
          tma_load_4d(smem + SM::QV + s * 16384, &tmQv, bar_q, s * 64, t.I0, t.h, t.b);
          if (MODE == 2)  // dO_I lives where P would
            tma_load_4d(smem + SM::PT + s * 16384, &tmDO, bar_q, s * 64, t.I0, t.h, t.b);
        }
        for (int st = 0; st < t.nsteps; ++st, ++gs) {
          const int J0 = (t.I - st) * 128;
          const int cb = p.L - 128 - t.I0 + J0;  // first row of the new chunk of r
          if (gs > 0) mbar_wait(bar_kfree, (gs - 1) & 1);
          mbar_expect_tx(bar_k, (MODE == 2 ? 3 : 2) * SM::TILE);
#pragma unroll
          for (int s = 0; s < NSLAB; ++s) {
            tma_load_4d(smem + SM::KT + s * 16384, &tmK, bar_k, s * 64, J0, t.h, t.b);
            tma_load_4d(smem + SM::RT + s * 16384, &tmR, bar_k, s * 64, cb, t.h, 0);
            if (MODE == 2) tma_load_4d(smem + SM::VT + s * 16384, &tmV, bar_k, s * 64, J0, t.h, t.b);
          }
          if (MODE == 0) {
            if (gs > 0) mbar_wait(bar_o, (gs - 1) & 1);
            mbar_expect_tx(bar_v, SM::TILE);
#pragma unroll
            for (int s = 0; s < NSLAB; ++s) tma_load_4d(smem + SM::VT + s * 16384, &tmV, bar_v, s * 64, J0, t.h, t.b);
          }
        }
        ++ti;
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    // Step st's scores are issued as soon as the softmax warps have pulled step st-1's S / band out of TMEM (bar_s1),
    // i.e. they run under step st-1's exponentials; P.V (or dP) follows when P has been published (bar_p).
    if (lane == 0) {
      const uint32_t idesc_s = umma_idesc(128, 128, 0, 0, 0);
      const uint32_t idesc_o = umma_idesc(128, D, 0, 1, 0);
      const uint32_t qu = smem_u32(smem + SM::QU), qv = smem_u32(smem + SM::QV), kt = smem_u32(smem + SM::KT),
                     vt = smem_u32(smem + SM::VT), rt = smem_u32(smem + SM::RT), pt = smem_u32(smem + SM::PT);
      auto issue_scores = [&](int gs) {
        mbar_wait(bar_k, gs & 1);
        tc_fence_after();
        const uint32_t tbd = T_BD + (gs & 1) * 128;
#pragma unroll
        for (int k = 0; k < D / 16; ++k) {
          const uint32_t off = (k >> 2) * 16384 + (k & 3) * 32;
          umma_ss(T_S, umma_smem_desc(qu + off, 16, 1024), umma_smem_desc(kt + off, 16, 1024), idesc_s, k ? 1u : 0u);
        }
#pragma unroll
        for (int k = 0; k < D / 16; ++k) {
          const uint32_t off = (k >> 2) * 16384 + (k & 3) * 32;
          umma_ss(tbd, umma_smem_desc(qv + off, 16, 1024), umma_smem_desc(rt + off, 16, 1024), idesc_s, k ? 1u : 0u);
        }
        if (MODE != 2) umma_commit(bar_kfree);  // mode 2: V_J (same load group) is still needed by dP
        umma_commit(bar_s);
      };
      auto issue_dp = [&]() {
        // dP = dO_I . V_J^T into the columns the forward uses for O (same K-major x K-major form as the scores)
#pragma unroll
        for (int k = 0; k < D / 16; ++k) {
          const uint32_t off = (k >> 2) * 16384 + (k & 3) * 32;
          umma_ss(T_O, umma_smem_desc(pt + off, 16, 1024), umma_smem_desc(vt + off, 16, 1024), idesc_s, k ? 1u : 0u);
        }
        umma_commit(bar_kfree);
        umma_commit(bar_o);
      };
      int gs = 0, ti = 0;
      for (int pass = 0; pass < n_pass; ++pass) {
        const int kk = item_of(pass);
        if (kk < 0) continue;
        const Item t = decode(kk);
        mbar_wait(bar_q, ti & 1);
        tc_fence_after();
        // T_S / the band slot are free: bar_s1 of the previous item's last step was waited for below
        issue_scores(gs);
        if (MODE == 2) issue_dp();
        if (t.nsteps == 1) umma_commit(bar_qfree);
        for (int st = 0; st < t.nsteps; ++st, ++gs) {
          mbar_wait(bar_s1, gs & 1);
          tc_fence_after();
          if (st + 1 < t.nsteps) {
            issue_scores(gs + 1);
            if (MODE != 2 && st + 2 == t.nsteps) umma_commit(bar_qfree);  // last scores of this item issued
          }
          mbar_wait(bar_p, gs & 1);
          tc_fence_after();
          if (MODE == 0) {
            mbar_wait(bar_v, gs & 1);
            tc_fence_after();
#pragma unroll
            for (int k = 0; k < 8; ++k) {  // 128 keys / 16
              const uint64_t adesc = umma_smem_desc(pt + (k >> 2) * 16384 + (k & 3) * 32, 16, 1024);
              const uint64_t bdesc = umma_smem_desc(vt + k * 2048, 16384, 1024);
              umma_ss(T_O, adesc, bdesc, idesc_o, (st | k) ? 1u : 0u);
            }
            umma_commit(bar_o);
          } else if (MODE == 2) {
            if (st + 1 < t.nsteps) {
              issue_dp();  // the softmax warps are done reading dP(st)
              if (st + 2 == t.nsteps) umma_commit(bar_qfree);  // last dP of this item issued: Q / dO may be replaced
            }
          }
        }
        ++ti;
      }
    }
  } else {
    // ------------------------------------------------------------------ softmax warps
    // Eight warps, two per TMEM lane quadrant: thread = one query row x HALF of the tile's 128 key columns. The
    // per-row shift of the position band (_rel_shift) is a 5-stage barrel shifter on registers (shift = 31 - lane),
    // fed by two 32-column TMEM loads; the two halves of a row exchange their row maximum through shared memory.
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const int r = q * 32 + lane;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    uint8_t* wstg = smem + SM::STG + (warp - 2) * 2560;                     // per-warp 32 x 80 B transpose area for P / dS stores
    float* xmax = reinterpret_cast<float*>(smem + SM::STG + 8 * 2560);      // [2 parity][2 half][128]
    float* xsum = xmax + 2 * 2 * 128;                                       // [2 half][128]
    uint8_t* prow = smem + SM::PT + r * 128;
    const int sh = 31 - lane;
    const bool b16 = sh & 16, b8 = sh & 8, b4 = sh & 4, b2 = sh & 2, b1 = sh & 1;
    int gs = 0;
    for (int pass = 0; pass < n_pass; ++pass) {
    const int kk = item_of(pass);
    if (kk < 0) continue;
    const Item it = decode(kk);
    const int I = it.I, I0 = it.I0, h = it.h, b = it.b, nsteps = it.nsteps;
    const int i = I0 + r;
    float m_run = NEG_BIG, l_run = 0.f;
    float lse_row = 0.f;
    float d_row = 0.f;
    if (MODE >= 1) lse_row = (i < p.L) ? p.lse2[((long long)b * p.H + h) * p.L + i] : 0.f;
    if (MODE == 2) {
      if (p.Drow != nullptr) {
        d_row = (i < p.L) ? p.Drow[((long long)b * p.H + h) * p.L + i] : 0.f;
      } else {
        // D_i = sum_d dO[i,d] * O[i,d] over this head (the softmax-backward row term): each half of the row's thread
        // pair takes half of the head dimension; the loads fly under the wait for the item's first scores
        float acc = 0.f;
        if (i < p.L) {
          const long long ro = (long long)b * p.L + i;
          const __half* pa = p.dOp + ro * p.lddo + (long long)h * p.dh + half * (D / 2);
          const __half* po = p.Op + ro * p.ldop + (long long)h * p.dh + half * (D / 2);
#pragma unroll
          for (int c = 0; c < D / 2; c += 8) {
            if (half * (D / 2) + c < p.dh) {
              float xa[8], xo[8];
              half8_to_float(ld_half8(pa + c), xa);
              half8_to_float(ld_half8(po + c), xo);
#pragma unroll
              for (int e = 0; e < 8; ++e) acc = fmaf(xa[e], xo[e], acc);
            }
          }
        }
        xsum[half * 128 + r] = acc;
        named_bar_sync(1 + q, 64);
        d_row = acc + xsum[(half ^ 1) * 128 + r];
        named_bar_sync(1 + q, 64);  // xsum is written again by the next item
      }
    }
    const long long zrow0 = (((long long)b * p.H + h) * p.L + I0 + q * 32) * p.L;  // first row of this warp in P / dS

    for (int st = 0; st < nsteps; ++st, ++gs) {
      const int J0 = (I - st) * 128;
      const uint32_t tnew = T_BD + (gs & 1) * 128, tprev = T_BD + ((gs + 1) & 1) * 128;
      // interior tiles need no predicate: every (i, j) is causal, inside the window and inside the sequence
      const bool need_mask = !(J0 + 127 <= I0 && I0 + 127 - J0 < p.window && I0 + 127 < p.L);
      mbar_wait(bar_s, gs & 1);
      tc_fence_after();
      // ---- pass 1: content + shifted position scores -> registers (raw, unscaled), row max over this half
      // scores are kept as packed fp32 pairs: the add of the position term, the scale-and-shift before the exponential
      // and the row sum issue as FADD2 / FFMA2 (one slot per two elements)
      float2 sc[2][16];
      float mx = NEG_BIG;
#pragma unroll
      for (int c2 = 0; c2 < 2; ++c2) {
        const int cc = half * 2 + c2;
        uint32_t s[32], w0[32], w1[32];
        const int blk = cc - q + 3;  // 32-column block of the 256-wide [new | prev] window
        tmem_ld32(((blk < 4) ? tnew + blk * 32 : tprev + (blk - 4) * 32) + lane_off, w0);
        tmem_ld32(((blk + 1 < 4) ? tnew + (blk + 1) * 32 : tprev + (blk - 3) * 32) + lane_off, w1);
        tmem_ld32(T_S + lane_off + cc * 32, s);
        tmem_ld_wait();
        // x[t] <- window[t + sh], t < 32: stages 16, 8, 4, 2, 1 (in place, ascending t reads not-yet-written slots)
        float x[64];
#pragma unroll
        for (int t = 0; t < 32; ++t) {
          x[t] = __uint_as_float(w0[t]);
          x[32 + t] = __uint_as_float(w1[t]);
        }
#pragma unroll
        for (int t = 0; t < 47; ++t) x[t] = b16 ? x[t + 16] : x[t];
#pragma unroll
        for (int t = 0; t < 39; ++t) x[t] = b8 ? x[t + 8] : x[t];
#pragma unroll
        for (int t = 0; t < 35; ++t) x[t] = b4 ? x[t + 4] : x[t];
#pragma unroll
        for (int t = 0; t < 33; ++t) x[t] = b2 ? x[t + 2] : x[t];
#pragma unroll
        for (int t = 0; t < 32; ++t) x[t] = b1 ? x[t + 1] : x[t];
#pragma unroll
        for (int t = 0; t < 16; ++t)
          sc[c2][t] = __fadd2_rn(make_float2(__uint_as_float(s[2 * t]), __uint_as_float(s[2 * t + 1])),
                                 make_float2(x[2 * t], x[2 * t + 1]));
        if (need_mask) {
#pragma unroll
          for (int t = 0; t < 16; ++t) {
            const int j = J0 + cc * 32 + 2 * t;
            const bool ok0 = (j <= i) && (i - j < p.window) && (j < p.L);
            const bool ok1 = (j + 1 <= i) && (i - j - 1 < p.window) && (j + 1 < p.L);
            sc[c2][t].x = ok0 ? sc[c2][t].x : NEG_BIG;
            sc[c2][t].y = ok1 ? sc[c2][t].y : NEG_BIG;
          }
        }
#pragma unroll
        for (int t = 0; t < 16; ++t) mx = fmaxf(mx, fmaxf(sc[c2][t].x, sc[c2][t].y));
      }
      // S and the previous band chunk are consumed: the MMA warp may overwrite them with step st+1's scores
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_s1);
      if (MODE == 0) {
        // row maximum over both halves (double-buffered slot, one 64-thread named barrier per quadrant and step);
        // the recompute modes take the saved log-sum-exp instead and need neither the maximum nor the exchange
        xmax[((gs & 1) * 2 + half) * 128 + r] = mx;
        named_bar_sync(1 + q, 64);
        mx = fmaxf(mx, xmax[((gs & 1) * 2 + (half ^ 1)) * 128 + r]);
        mx *= p.scale_log2;  // scale > 0: max commutes with the scaling; masked entries stay hugely negative
      }

      if (MODE == 0) {
        // ---- online softmax bookkeeping with lazy rescale (rescale only when the max grew by more than 2^8)
        if (st > 0) {
          mbar_wait(bar_o, (gs - 1) & 1);  // O accumulated, P smem free
          tc_fence_after();
        }
        const bool grow = (st == 0) || (mx > m_run + 8.0f);
        if (__any_sync(0xffffffffu, grow)) {
          const float m_new = grow ? mx : m_run;
          const float f = (st == 0) ? 0.f : exp2f(m_run - m_new);
          l_run *= f;
          m_run = m_new;
          if (st > 0) {
#pragma unroll 1
            for (int c = half * (D / 64); c < (half + 1) * (D / 64); ++c) {  // this half's columns of O
              uint32_t o[32];
              tmem_ld32(T_O + lane_off + c * 32, o);
              tmem_ld_wait();
#pragma unroll
              for (int t = 0; t < 32; ++t) o[t] = __float_as_uint(__uint_as_float(o[t]) * f);
              tmem_st32(T_O + lane_off + c * 32, o);
            }
            tmem_st_wait();
          }
        }
      } else if (MODE == 2) {
        mbar_wait(bar_o, gs & 1);  // dP(st) is in TMEM
        tc_fence_after();
      }
      const float m_use = (MODE == 0) ? m_run : lse_row;
      // ---- pass 2: probabilities from the registers, p = exp2(raw * scale_log2 - m)
      float2 sum2 = make_float2(0.f, 0.f);
      const float2 sl2 = make_float2(p.scale_log2, p.scale_log2), nm2 = make_float2(-m_use, -m_use);
#pragma unroll
      for (int c2 = 0; c2 < 2; ++c2) {
        const int cc = half * 2 + c2;
        uint32_t pk[16];
#pragma unroll
        for (int t = 0; t < 16; ++t) {
          const float2 a = __ffma2_rn(sc[c2][t], sl2, nm2);
          const float2 pr = make_float2(ex2_approx(a.x), ex2_approx(a.y));
          sum2 = __fadd2_rn(sum2, pr);
          pk[t] = pack_half2(pr.x, pr.y);
        }
        if (MODE == 0) {
          // K-major SW128 tile: slab = cc/2 (64 keys each), 16-byte chunk index within the row = (cc&1)*4 + g
          uint8_t* dst = prow + (cc >> 1) * 16384;
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const int chunk = ((cc & 1) * 4 + g) ^ (r & 7);
            *reinterpret_cast<uint4*>(dst + chunk * 16) = make_uint4(pk[4 * g], pk[4 * g + 1], pk[4 * g + 2], pk[4 * g + 3]);
          }
        } else {
          // P (and dS) leave through the staging area so that every global store covers 8 rows x 64 contiguous bytes
          const int rows_valid = p.L - (I0 + q * 32);
          const int cols_valid = p.L - (J0 + cc * 32);
          const uint32_t sbase = smem_u32(wstg);
          const int piece = lane & 3;
          auto store_tile = [&](const uint32_t (&v)[16], __half* base) {
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              Half8 h8;
              h8.u = make_uint4(v[4 * g], v[4 * g + 1], v[4 * g + 2], v[4 * g + 3]);
              sts_half8(sbase + lane * 80 + g * 16, h8);
            }
            __syncwarp();
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const int rr = k * 8 + (lane >> 2);
              if (rr < rows_valid && piece * 8 < cols_valid)
                st_half8(base + (long long)rr * p.L + piece * 8, lds_half8(sbase + rr * 80 + piece * 16));
            }
            __syncwarp();
          };
          // Tiled layout (what the backward's consumer kernels read by TMA): tile (b, h, I, J) = 32 KB contiguous, two
          // slabs of [128 rows][64 columns]; every element of a visited tile is written (rows beyond the sequence as
          // zeros), so a consumer's 16 KB box is one contiguous read and needs no bounds.
          const long long tme = ((((long long)b * p.H + h) * nq + I) * nq + (I - st)) * 16384 + (cc >> 1) * 8192 +
                                (long long)r * 64 + (cc & 1) * 32;
          const bool direct = p.tiled || (p.L & 15) == 0;  // rows 32-byte aligned: 256-bit stores from the row's owner
          const long long zme = p.tiled ? tme : zrow0 + (long long)lane * p.L + J0 + cc * 32;
          const int nvalid = p.tiled ? 32 : ((lane < rows_valid) ? (cols_valid > 32 ? 32 : (cols_valid < 0 ? 0 : cols_valid)) : 0);
          if (p.tiled && i >= p.L) {
#pragma unroll
            for (int t = 0; t < 16; ++t) pk[t] = 0u;
          }
          if (direct) stg_row32(p.P + zme, pk, nvalid, true);
          else store_tile(pk, p.P + zrow0 + J0 + cc * 32);
          if (MODE == 2) {
            // dS = P * (dP - D) * scale; masked entries have P == 0 exactly
            uint32_t dp[32], dk[16];
            tmem_ld32(T_O + lane_off + cc * 32, dp);
            tmem_ld_wait();
            const float2 nd2 = make_float2(-d_row, -d_row), s2 = make_float2(p.scale, p.scale);
#pragma unroll
            for (int t = 0; t < 16; ++t) {
              const float2 pp = unpack_half2(pk[t]);
              const float2 dd = __fadd2_rn(make_float2(__uint_as_float(dp[2 * t]), __uint_as_float(dp[2 * t + 1])), nd2);
              const float2 ds2 = __fmul2_rn(__fmul2_rn(pp, dd), s2);
              dk[t] = pack_half2(ds2.x, ds2.y);
            }
            if (p.tiled && i >= p.L) {
#pragma unroll
              for (int t = 0; t < 16; ++t) dk[t] = 0u;
            }
            if (direct) stg_row32(p.dS + zme, dk, nvalid, true);
            else store_tile(dk, p.dS + zrow0 + J0 + cc * 32);
          }
        }
      }
      l_run += sum2.x + sum2.y;
      // publish: P is in smem (generic proxy -> async proxy) / dP has been read
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_p);
    }

    if (MODE == 0) {
      // total row sum over both halves, then each half normalises and writes its D/2 columns of O
      xsum[half * 128 + r] = l_run;
      mbar_wait(bar_o, (gs - 1) & 1);
      tc_fence_after();
      named_bar_sync(1 + q, 64);
      const float l_tot = l_run + xsum[(half ^ 1) * 128 + r];
      const float inv = 1.0f / l_tot;
      if (half == 0 && i < p.L && i >= p.q_start) p.lse2[((long long)b * p.H + h) * p.L + i] = m_run + log2f(l_tot);
      __half* orow = p.O + ((long long)b * p.L + i) * p.ldo + (long long)h * p.dh;
#pragma unroll 1
      for (int c = half * (D / 64); c < (half + 1) * (D / 64); ++c) {
        uint32_t o[32];
        tmem_ld32(T_O + lane_off + c * 32, o);
        tmem_ld_wait();
        if (i < p.L && i >= p.q_start) {
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const int col = c * 32 + g * 8;
            if (col < p.dh) {
              uint4 vv;
              vv.x = pack_half2(__uint_as_float(o[g * 8 + 0]) * inv, __uint_as_float(o[g * 8 + 1]) * inv);
              vv.y = pack_half2(__uint_as_float(o[g * 8 + 2]) * inv, __uint_as_float(o[g * 8 + 3]) * inv);
              vv.z = pack_half2(__uint_as_float(o[g * 8 + 4]) * inv, __uint_as_float(o[g * 8 + 5]) * inv);
              vv.w = pack_half2(__uint_as_float(o[g * 8 + 6]) * inv, __uint_as_float(o[g * 8 + 7]) * inv);
              *reinterpret_cast<uint4*>(orow + col) = vv;
            }
          }
        }
      }
      // the next item's first P.V (accumulate = 0) must not start before every row of O has been read: it is issued
      // only after bar_p of that step, which these warps arrive on after this point. The second named barrier keeps
      // a fast half from overwriting xsum (next item) before its partner has read it.
      tc_fence_before();
      named_bar_sync(1 + q, 64);
    }
    }  // items
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

// 4-D map over a [B, L, H, dh] view of a row-major buffer: dims (dh, L, H, B).
static int make_head_map(CUtensorMap* tm, const void* base, int dh, int L, int H, int B, long long ld) {
  uint64_t dims[4] = {(uint64_t)dh, (uint64_t)L, (uint64_t)H, (uint64_t)(B > 0 ? B : 1)};
  uint64_t str[3] = {(uint64_t)ld * 2, (uint64_t)dh * 2, (uint64_t)L * (uint64_t)ld * 2};
  uint32_t box[4] = {64, 128, 1, 1};
  return make_tmap_f16(tm, base, 4, dims, str, box);
}

template <int D, int MODE>
static int launch_attn_m(const CUtensorMap* tm, const AttnParams& p, cudaStream_t stream) {
  using SM = AttnSmem<D>;
  static bool configured = false;
  if (!configured) {
    DB1_CUDA(cudaFuncSetAttribute(relattn_fwd_kernel<D, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM::TOTAL));
    configured = true;
  }
  const int n_items = ((p.L + 127) / 128 - p.q_start / 128) * p.H * p.B;
  const int grid = n_items < sm_count() ? n_items : sm_count();
  DB1_CUDA(launch_pdl(relattn_fwd_kernel<D, MODE>, dim3(grid), dim3(AT_THREADS), SM::TOTAL, stream, 1, tm[0], tm[1], tm[2],
                      tm[3], tm[4], tm[5], p));
  return 0;
}

// the mode (0: O + LSE, 1: P, 2: P + dS) is a compile-time parameter: every per-step branch on it disappears
template <int D>
static int launch_attn(const CUtensorMap* tm, const AttnParams& p, cudaStream_t stream) {
  if (p.mode == 0) return launch_attn_m<D, 0>(tm, p, stream);
  if (p.mode == 1) return launch_attn_m<D, 1>(tm, p, stream);
  return launch_attn_m<D, 2>(tm, p, stream);
}

}  // namespace db1

using namespace db1;

static int relattn_launch(const void* qu, const void* qv, const void* k, const void* v, long long ld_qkv,
                          const void* r, long long ld_r, void* out, long long ld_out, float* lse2, void* probs,
                          const void* dout, long long ld_do, const float* drow, void* ds, int B, int L, int H, int dh,
                          int window, float scale, int mode, int q_start, cudaStream_t stream,
                          const void* o_for_d = nullptr, long long ld_o_for_d = 0, int tiled = 0) {
  DB1_CHECK_ARG(qu && qv && k && r && lse2, "relattn: null pointer");
  DB1_CHECK_ARG(q_start >= 0 && q_start < L && (q_start == 0 || mode == 0), "relattn: bad q_start %d", q_start);
  DB1_CHECK_ARG(B > 0 && L > 0 && H > 0, "relattn: bad shape B=%d L=%d H=%d", B, L, H);
  DB1_CHECK_ARG(dh % 8 == 0 && dh >= 8 && dh <= 128, "relattn: head dim %d unsupported (multiple of 8, <= 128)", dh);
  DB1_CHECK_ARG(mode == 0 || tiled || L % 8 == 0, "relattn: sequence length %d must be a multiple of 8 for the P / dS outputs", L);
  DB1_CHECK_ARG(ld_qkv % 8 == 0 && ld_r % 8 == 0 && ld_out % 8 == 0 && ld_do % 8 == 0,
                "relattn: row strides must be multiples of 8");
  DB1_CHECK_ARG(window > 0, "relattn: window (mem_len) must be > 0; mem_len == 0 masks every key");
  AttnParams p;
  p.L = L; p.H = H; p.B = B; p.dh = dh; p.window = window;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.O = (__half*)out; p.ldo = ld_out; p.lse2 = lse2; p.P = (__half*)probs; p.mode = mode;
  p.dS = (__half*)ds; p.Drow = drow; p.scale = scale; p.q_start = q_start; p.tiled = tiled;
  p.dOp = (const __half*)dout; p.lddo = ld_do; p.Op = (const __half*)o_for_d; p.ldop = ld_o_for_d;
  CUtensorMap tm[6];
  int e;
  if ((e = make_head_map(&tm[0], qu, dh, L, H, B, ld_qkv))) return e;
  if ((e = make_head_map(&tm[1], qv, dh, L, H, B, ld_qkv))) return e;
  if ((e = make_head_map(&tm[2], k, dh, L, H, B, ld_qkv))) return e;
  if ((e = make_head_map(&tm[3], mode == 1 ? k : v, dh, L, H, B, ld_qkv))) return e;
  if ((e = make_head_map(&tm[4], r, dh, L, H, 1, ld_r))) return e;
  if (mode == 2) {
    if ((e = make_head_map(&tm[5], dout, dh, L, H, B, ld_do))) return e;
  } else {
    tm[5] = tm[2];
  }
  if (dh <= 64) return launch_attn<64>(tm, p, stream);
  return launch_attn<128>(tm, p, stream);
}

extern "C" int db1_relattn_fwd(const void* qu, const void* qv, const void* k, const void* v, long long ld_qkv,
                               const void* r, long long ld_r, void* out, long long ld_out, float* lse2, void* probs,
                               int B, int L, int H, int dh, int window, float scale, int mode, void* stream_) {
  DB1_CHECK_ARG(mode == 0 || mode == 1, "relattn: mode must be 0 (O, LSE) or 1 (P)");
  DB1_CHECK_ARG((mode == 0 && v && out) || (mode == 1 && probs), "relattn: missing output for mode %d", mode);
  return relattn_launch(qu, qv, k, v, ld_qkv, r, ld_r, out, ld_out, lse2, probs, nullptr, 0, nullptr, nullptr, B, L, H,
                        dh, window, scale, mode, 0, (cudaStream_t)stream_);
}

extern "C" int db1_relattn_mem_fwd(const void* qu, const void* qv, const void* k, const void* v, long long ld_qkv,
                                   const void* r, long long ld_r, void* out, long long ld_out, float* lse2, int B, int K,
                                   int H, int dh, int window, float scale, int mlen, void* stream_) {
  DB1_CHECK_ARG(v && out, "relattn_mem_fwd: null pointer");
  return relattn_launch(qu, qv, k, v, ld_qkv, r, ld_r, out, ld_out, lse2, nullptr, nullptr, 0, nullptr, nullptr, B, K, H,
                        dh, window, scale, 0, mlen, (cudaStream_t)stream_);
}

extern "C" int db1_relattn_bwd_ds(const void* qu, const void* qv, const void* k, const void* v, long long ld_qkv,
                                  const void* r, long long ld_r, const void* dout, long long ld_do, const float* lse2,
                                  const float* drow, void* probs, void* ds, int B, int L, int H, int dh, int window,
                                  float scale, void* stream_) {
  DB1_CHECK_ARG(v && dout && drow && probs && ds, "relattn_bwd_ds: null pointer");
  return relattn_launch(qu, qv, k, v, ld_qkv, r, ld_r, nullptr, 0, const_cast<float*>(lse2), probs, dout, ld_do, drow,
                        ds, B, L, H, dh, window, scale, 2, 0, (cudaStream_t)stream_);
}

/* Same, with D = rowsum(dO * O) computed inside the kernel from `o` (the forward output, [B*L, ld_o], head h at column
 * h*dh) instead of being read from a db1_rowdot result: one launch and one pass over dO / O less per layer. */
extern "C" int db1_relattn_bwd_ds_o(const void* qu, const void* qv, const void* k, const void* v, long long ld_qkv,
                                    const void* r, long long ld_r, const void* dout, long long ld_do, const void* o,
                                    long long ld_o, const float* lse2, void* probs, void* ds, int B, int L, int H, int dh,
                                    int window, float scale, void* stream_) {
  DB1_CHECK_ARG(v && dout && o && probs && ds, "relattn_bwd_ds_o: null pointer");
  DB1_CHECK_ARG(ld_o % 8 == 0 && (((uintptr_t)o | (uintptr_t)dout) & 15) == 0, "relattn_bwd_ds_o: o / dout must be 16-byte aligned rows");
  return relattn_launch(qu, qv, k, v, ld_qkv, r, ld_r, nullptr, 0, const_cast<float*>(lse2), probs, dout, ld_do, nullptr,
                        ds, B, L, H, dh, window, scale, 2, 0, (cudaStream_t)stream_, o, ld_o);
}

/* Same as db1_relattn_bwd_ds / _o (drow may be NULL when o is given, and vice versa), writing P and dS in the TILED
 * layout the consumer kernels of csrc/relattn_bwd.cu read: [B][H][nq][nq][2][128][64] fp16 with nq = ceil(L / 128) -
 * tile (I, J) of a head = two contiguous slabs of 128 rows x 64 columns. Only visited (causal, in-window) tiles are
 * written, completely (masked entries and rows beyond L as zeros). No restriction on L. */
extern "C" int db1_relattn_bwd_ds_tiled(const void* qu, const void* qv, const void* k, const void* v, long long ld_qkv,
                                        const void* r, long long ld_r, const void* dout, long long ld_do, const void* o,
                                        long long ld_o, const float* lse2, const float* drow, void* probs, void* ds,
                                        int B, int L, int H, int dh, int window, float scale, void* stream_) {
  DB1_CHECK_ARG(v && dout && (o || drow) && probs && ds, "relattn_bwd_ds_tiled: null pointer");
  DB1_CHECK_ARG(!o || (ld_o % 8 == 0 && (((uintptr_t)o | (uintptr_t)dout) & 15) == 0),
                "relattn_bwd_ds_tiled: o / dout must be 16-byte aligned rows");
  DB1_CHECK_ARG((((uintptr_t)probs | (uintptr_t)ds) & 31) == 0, "relattn_bwd_ds_tiled: probs / ds must be 32-byte aligned");
  return relattn_launch(qu, qv, k, v, ld_qkv, r, ld_r, nullptr, 0, const_cast<float*>(lse2), probs, dout, ld_do, drow, ds,
                        B, L, H, dh, window, scale, 2, 0, (cudaStream_t)stream_, drow ? nullptr : o, ld_o, 1);
}
