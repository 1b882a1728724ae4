// Attention backward, second half: everything that consumes dS = P * (dP - D) * scale (written once, fp16, by the
// recompute kernel) on tcgen05 tensor cores, without ever materialising the relative-position re-layout of dS.
//
//   dq_i  = sum_j dS[i,j] k_j  +  sum_j dS[i,j] r[j + L-1-i]          (adjoint of transformer_xl.py:161-171 w.r.t. q)
//   du_h  = sum_{b,i} (dS K)_i ,   dv_h = sum_{b,i} (skew^-1(dS) R)_i   (shared biases r_w_bias / r_r_bias)
//   dR[c] = sum_{b,(i,j): j+L-1-i = c} dS[i,j] (q_i + v)                (adjoint of _rel_shift :98-110 and of :167)
//
// The per-row shift ("un-shift", the adjoint of _rel_shift) is done on registers by the thread that owns the row: it
// reads its 128 dS values of the tile straight from global memory, shifts them right by 127 - r elements inside a
// 256-wide zero-padded row (8-element chunks by address arithmetic, words by two select stages, the odd half by a
// funnel shift) and stores the result as two K-major 128x128 fp16 tiles ("band" tiles: columns = relative positions
// of the chunk of r this tile pairs with, `new` = positions [cb, cb+128), `prev` = [cb+128, cb+256)) in the 128-byte
// swizzled layout tcgen05.mma reads. No dSr matrix, no rel_unshift pass, no P / dSr re-reads.
//
// Two persistent kernels share the code (template KIND), both with accumulators resident in TMEM:
//   KIND 0 (query-outer, item = (query tile, head, sequence)): dQu += dS . K_J, dQv += band . R_window per key tile;
//           epilogue writes dq = dQu + dQv (fp16) and adds the column sums to du / dv.
//   KIND 1 (diagonal-outer, item = (tile diagonal t, head, pair of sequences)): dR_new += band_new^T . Qv_I,
//           dR_prev += band_prev^T . Qv_I over all tiles of the diagonal; one fp32 reduction per item. (Reducing dR
//           per key step costs a 64 KB fp32 reduction per step: measured 6.3k cycles with red.v4, 3.0k with
//           cp.reduce.async.bulk when all SMs do it - as long as the whole step. tools/ubench_red.cu.)
// Roles (448 threads): warp 0 TMA producer, warp 1 MMA issuer, warps 2-9 un-shift (two threads per tile row: warps 2-5
// produce the row's shifted chunks 0-8, warps 6-9 chunks 9-16 - each reads only the half of the row it needs, and the
// compiler prunes the other half of the shifter), warps 10-13 TMEM drain (epilogue / dR reduction; accumulators are
// double-buffered so it overlaps the next item).
#include <type_traits>

#include "../../include/db1_sm100.h"
#include "common.cuh"
#include "ptx.cuh"

namespace db1 {

constexpr int BW_THREADS = 448;  // TMA warp, MMA warp, 8 un-shift warps (two threads per tile row), 4 drain warps

struct BandParams {
  int L, H, B, dh, window;
  const __half* dS;  // tiled [B,H,nq,nq,2,128,64] fp16 (db1_relattn_bwd_ds_tiled): visited tiles valid, masked entries exactly 0
  __half* dq;        // KIND 0: [B*L, lddq], head h at column h*dh
  long long lddq;
  float* du;         // KIND 0: [H*dh] fp32, accumulated
  float* dv;
  float* dR;         // KIND 1: [L, lddr] fp32, accumulated (row c <-> r row c)
  long long lddr;
};

template <int D, int KIND>
struct BandSmem {
  static constexpr int TILE = 128 * D * 2;  // one [128][D] fp16 operand tile (D/64 slabs of [128][64])
  static constexpr int DS = 0;              // 2 x [128][128] fp16 dS tiles, K-major (2 slabs each), written by TMA
  static constexpr int BAND = 65536;        // [128][256] fp16 K-major (4 slabs): new = slabs 0,1, prev = slabs 2,3
  static constexpr int KT = BAND + 65536;   // KIND 0: K_J tile                     KIND 1: Qv_I tiles (2 buffers)
  static constexpr int RR = KT + TILE;      // KIND 0: 2-slot ring of 128-row chunks of r
  static constexpr int BARS = (KIND == 0) ? (RR + 2 * TILE) : (KT + 2 * TILE);
  static constexpr int TOTAL = BARS + 256;
};

// ---- the un-shift of one tile row on registers ------------------------------------------------------------------
// W[64]: the row's 128 fp16 values (word m = elements 2m, 2m+1). out[c][0..3] (c = 0..16): the 16-byte chunks of the
// row shifted right by rem = (127 - r) & 7 elements; chunk c belongs at chunk position ((127 - r) >> 3) + c of the
// 32-chunk (256-element) band row.
DEVI void unshift_row(const uint32_t (&W)[64], int rem, uint32_t (&out)[17][4]) {
  const uint32_t sb = (uint32_t)(rem & 1) * 16u;
  const bool w2 = rem & 4, w1 = rem & 2;
  // X(m) = elements [2m - (rem&1), +2) of the row, m in [0, 64]; zero outside
  uint32_t X[65];
#pragma unroll
  for (int m = 0; m < 65; ++m) {
    const uint32_t lo = (m >= 1) ? W[m - 1] : 0u;
    const uint32_t hi = (m < 64) ? W[m] : 0u;
    X[m] = __funnelshift_l(lo, hi, sb);
  }
  // T(m) = X(m - 2*w2), m in [-1, 67];  Y(m) = T(m - w1), m in [0, 68)
  uint32_t T[69];  // T[m + 1]
#pragma unroll
  for (int m = -1; m < 68; ++m) {
    const uint32_t x0 = (m >= 0 && m <= 64) ? X[m] : 0u;
    const uint32_t x2 = (m - 2 >= 0 && m - 2 <= 64) ? X[m - 2] : 0u;
    T[m + 1] = w2 ? x2 : x0;
  }
#pragma unroll
  for (int c = 0; c < 17; ++c) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int m = 4 * c + j;
      out[c][j] = w1 ? T[m] : T[m + 1];  // T(m - 1) : T(m)
    }
  }
}

DEVI void sts128(uint32_t saddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
DEVI void red_add_v4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// band chunk positions [16*part, 16*part + 16) of row r: data chunks out[c] at position a + c, zeros elsewhere. The row's
// two threads split the work: thread HS writes the data chunks c in [9*HS, 9 + 8*HS) and the zero chunks q in [8*HS, 8*HS + 8).
template <int HS>
DEVI void store_band_half(uint32_t band_base, int r, int a, int part, const uint32_t (&out)[17][4]) {
  const uint32_t rowb = band_base + (uint32_t)r * 128u;
  const int rx = r & 7;
#pragma unroll
  for (int c = 9 * HS; c < 9 + 8 * HS; ++c) {
    const int pos = a + c;
    if ((pos >> 4) == part)
      sts128(rowb + (uint32_t)(pos >> 3) * 16384u + (uint32_t)(((pos & 7) ^ rx) * 16), out[c][0], out[c][1], out[c][2],
             out[c][3]);
  }
#pragma unroll
  for (int q = 8 * HS; q < 8 * HS + 8; ++q) {
    const int pos = part * 16 + q;
    if (pos < a || pos > a + 16) sts128(rowb + (uint32_t)(pos >> 3) * 16384u + (uint32_t)(((pos & 7) ^ rx) * 16), 0u, 0u, 0u, 0u);
  }
}

// lane l ends with the sum over the warp's 32 lanes of v[l]
DEVI float warp_colsum32(float (&v)[32], int lane) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const bool up = lane & off;
#pragma unroll
    for (int k = 0; k < off; ++k) {
      const float send = up ? v[k] : v[k + off];
      const float keep = up ? v[k + off] : v[k];
      v[k] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  return v[0];
}

template <int D, int KIND>
__global__ void __launch_bounds__(BW_THREADS, 1)
relattn_bwd_band_kernel(const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmR,
                        const __grid_constant__ CUtensorMap tmQv, const __grid_constant__ CUtensorMap tmDS,
                        const BandParams p) {
  using SM = BandSmem<D, KIND>;
  constexpr int NSLAB = D / 64;
  constexpr int TILE = SM::TILE;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SM::BARS);
  // per-buffer barriers (one phase per two steps)
  uint64_t* bar_ds = bars + 0;     // [2] dS tile landed (TMA)
  uint64_t* ds_free = bars + 2;    // [2] un-shift warps have read it (+ KIND 0: the dQu MMAs have)
  uint64_t* bar_b = bars + 4;      // [2] KIND 1: Qv_I landed              KIND 0: [0] K_J landed, [1] r chunk landed (per step)
  uint64_t* b_free = bars + 6;     // [2] KIND 1: the step's MMAs are done  KIND 0: [0] dQu MMAs done (K_J free), per step
  // per-step barriers
  uint64_t* full_prev = bars + 8;  // un-shift warps -> MMA
  uint64_t* full_new = bars + 9;
  uint64_t* free_prev = bars + 10; // MMA commit -> un-shift warps (KIND 0: and the r ring)
  uint64_t* free_new = bars + 11;
  // per-item
  uint64_t* acc_full = bars + 12;  // [2] accumulator set complete (commit)
  uint64_t* acc_free = bars + 14;  // [2] drained (4 warps)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int nq = (p.L + 127) / 128;
  const int Wt = (p.window - 1 + 127) / 128;            // largest tile diagonal that holds an unmasked pair
  const int nT = (nq - 1 < Wt ? nq - 1 : Wt) + 1;       // number of tile diagonals
  const int HB = p.H * p.B;
  const int nbp = (p.B + 1) / 2;
  const int n_items = (KIND == 0) ? nq * HB : nT * p.H * nbp;
  const int G = gridDim.x;
  auto item_of = [&](int pass) -> int {
    const int c = (pass & 1) ? (G - 1 - (int)blockIdx.x) : (int)blockIdx.x;
    const int k = pass * G + c;
    return k < n_items ? k : -1;
  };
  const int n_pass = (n_items + G - 1) / G;
  // KIND 0: (I, h, b), steps st = 0..nsteps-1 over key tiles J = I - st.
  // KIND 1: (t, h, b0, nb), steps over (b, I): tile (I, J = I - t), I = t..nq-1.
  struct Item {
    int I, t, h, b, nb, nsteps;
  };
  auto decode = [&](int k) -> Item {
    Item it;
    if (KIND == 0) {
      it.I = nq - 1 - k / HB;
      const int hb = k % HB;
      it.h = hb % p.H;
      it.b = hb / p.H;
      it.t = 0;
      it.nb = 1;
      it.nsteps = (it.I < Wt ? it.I : Wt) + 1;
    } else {
      it.t = k / (p.H * nbp);
      const int rest = k % (p.H * nbp);
      it.h = rest % p.H;
      it.b = (rest / p.H) * 2;
      it.nb = (p.B - it.b) < 2 ? (p.B - it.b) : 2;
      it.I = it.t;
      it.nsteps = it.nb * (nq - it.t);
    }
    return it;
  };
  // tile of step st of an item: query tile origin, key tile origin, sequence
  auto tile_of = [&](const Item& it, int st, int& I0, int& J0, int& b) {
    if (KIND == 0) {
      I0 = it.I * 128;
      J0 = (it.I - st) * 128;
      b = it.b;
    } else {
      const int per = nq - it.t;
      b = it.b + st / per;
      I0 = (it.t + st % per) * 128;
      J0 = I0 - it.t * 128;
    }
  };

  if ((smem_u32(smem) & 1023u) != 0) __trap();
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmR);
    tma_prefetch_desc(&tmQv);
    tma_prefetch_desc(&tmDS);
  }
  if (warp == 1) {
    if (lane == 0) {
      mbar_init(bar_ds + 0, 1);
      mbar_init(bar_ds + 1, 1);
      mbar_init(ds_free + 0, KIND == 0 ? 9 : 8);  // the 8 un-shift warps (+ KIND 0: the dQu MMAs' commit)
      mbar_init(ds_free + 1, KIND == 0 ? 9 : 8);
      mbar_init(bar_b + 0, 1);
      mbar_init(bar_b + 1, 1);
      mbar_init(b_free + 0, 1);
      mbar_init(b_free + 1, 1);
      mbar_init(full_prev, 8);
      mbar_init(full_new, 8);
      mbar_init(free_prev, 1);
      mbar_init(free_new, 1);
      mbar_init(acc_full + 0, 1);
      mbar_init(acc_full + 1, 1);
      mbar_init(acc_free + 0, 4);
      mbar_init(acc_free + 1, 4);
      mbar_fence_init();
    }
    __syncwarp();
    tmem_alloc<512>(tmem_slot);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();
  pdl_wait();

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int gs = 0;
      for (int pass = 0; pass < n_pass; ++pass) {
        const int k = item_of(pass);
        if (k < 0) continue;
        const Item it = decode(k);
        for (int st = 0; st < it.nsteps; ++st, ++gs) {
          int I0, J0, b;
          tile_of(it, st, I0, J0, b);
          const int buf = gs & 1;
          if (gs >= 2) mbar_wait(ds_free + buf, ((gs >> 1) - 1) & 1);
          mbar_expect_tx(bar_ds + buf, 32768);
          const int tile2 = (((b * p.H + it.h) * nq + (I0 >> 7)) * nq + (J0 >> 7)) * 2;  // slab index in the tiled dS
          tma_load_3d(smem + SM::DS + buf * 32768, &tmDS, bar_ds + buf, 0, 0, tile2);
          tma_load_3d(smem + SM::DS + buf * 32768 + 16384, &tmDS, bar_ds + buf, 0, 0, tile2 + 1);
          if (KIND == 0) {
            const int cb = p.L - 128 - 128 * st;  // first r row of this step's new chunk
            if (gs > 0) mbar_wait(b_free + 0, (gs - 1) & 1);
            mbar_expect_tx(bar_b + 0, TILE);
#pragma unroll
            for (int s = 0; s < NSLAB; ++s) tma_load_4d(smem + SM::KT + s * 16384, &tmK, bar_b + 0, s * 64, J0, it.h, b);
            if (gs > 0) mbar_wait(free_prev, (gs - 1) & 1);  // the slot's previous chunk was last read by step gs-1
            mbar_expect_tx(bar_b + 1, TILE);
#pragma unroll
            for (int s = 0; s < NSLAB; ++s)
              tma_load_4d(smem + SM::RR + (gs & 1) * TILE + s * 16384, &tmR, bar_b + 1, s * 64, cb, it.h, 0);
          } else {
            if (gs >= 2) mbar_wait(b_free + buf, ((gs >> 1) - 1) & 1);
            mbar_expect_tx(bar_b + buf, TILE);
#pragma unroll
            for (int s = 0; s < NSLAB; ++s)
              tma_load_4d(smem + SM::KT + buf * TILE + s * 16384, &tmQv, bar_b + buf, s * 64, I0, it.h, b);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      const uint32_t idesc_q = umma_idesc(128, D, 0, 1, 0);  // A K-major (dS / band), B MN-major (K_J / r chunk)
      const uint32_t idesc_r = umma_idesc(128, D, 1, 1, 0);  // A MN-major (band^T), B MN-major (Qv_I)
      const uint32_t band = smem_u32(smem + SM::BAND);
      int gs = 0, ti = 0;
      for (int pass = 0; pass < n_pass; ++pass) {
        const int k = item_of(pass);
        if (k < 0) continue;
        const Item it = decode(k);
        const int as = ti & 1;
        const uint32_t T0 = tmem_base + as * (2 * D);  // dQu | dR_new
        const uint32_t T1 = T0 + D;                    // dQv | dR_prev
        for (int st = 0; st < it.nsteps; ++st, ++gs) {
          const int buf = gs & 1;
          if (KIND == 0) {
            const uint32_t ds = smem_u32(smem + SM::DS + buf * 32768), kt = smem_u32(smem + SM::KT);
            const uint32_t r_new = smem_u32(smem + SM::RR + (gs & 1) * TILE);
            const uint32_t r_prev = smem_u32(smem + SM::RR + ((gs + 1) & 1) * TILE);
            // The r chunk is waited for before anything of this step is committed: the producer re-arms its barrier for
            // step gs+1 as soon as free_prev(gs) fires (immediately on a step without G2), and an arrive.expect_tx on a
            // barrier whose previous phase still has bytes in flight underflows its pending count (a device fault).
            mbar_wait(bar_b + 1, gs & 1);
            // G1: dQu += dS . K_J   (the dS tile is the TMA-written K-major operand itself)
            mbar_wait(bar_b + 0, gs & 1);
            mbar_wait(bar_ds + buf, (gs >> 1) & 1);
            if (st == 0 && ti >= 2) mbar_wait(acc_free + as, ((ti >> 1) - 1) & 1);
            tc_fence_after();
#pragma unroll
            for (int kk = 0; kk < 8; ++kk)
              umma_ss(T0, umma_smem_desc(ds + (kk >> 2) * 16384 + (kk & 3) * 32, 16, 1024),
                      umma_smem_desc(kt + kk * 2048, 16384, 1024), idesc_q, (st | kk) ? 1u : 0u);
            umma_commit(b_free + 0);
            umma_commit(ds_free + buf);
            // G2: dQv += band_prev . r[cb+128 .. cb+256)   (nothing on the diagonal tile: its upper triangle is masked).
            // The barrier is waited for on every step, also when there is nothing to issue: committing free_prev
            // without having seen this step's full_prev would let it run two phases ahead of the un-shift warps.
            mbar_wait(full_prev, gs & 1);
            if (st > 0) {
              tc_fence_after();
#pragma unroll
              for (int kk = 0; kk < 8; ++kk)
                umma_ss(T1, umma_smem_desc(band + (2 + (kk >> 2)) * 16384 + (kk & 3) * 32, 16, 1024),
                        umma_smem_desc(r_prev + kk * 2048, 16384, 1024), idesc_q, 1u);
            }
            umma_commit(free_prev);
            // G3: dQv += band_new . r[cb .. cb+128)
            mbar_wait(full_new, gs & 1);
            tc_fence_after();
#pragma unroll
            for (int kk = 0; kk < 8; ++kk)
              umma_ss(T1, umma_smem_desc(band + (kk >> 2) * 16384 + (kk & 3) * 32, 16, 1024),
                      umma_smem_desc(r_new + kk * 2048, 16384, 1024), idesc_q, (st | kk) ? 1u : 0u);
            umma_commit(free_new);
          } else {
            const uint32_t qv = smem_u32(smem + SM::KT + buf * TILE);
            mbar_wait(bar_b + buf, (gs >> 1) & 1);
            mbar_wait(full_new, gs & 1);
            if (st == 0 && ti >= 2) mbar_wait(acc_free + as, ((ti >> 1) - 1) & 1);
            tc_fence_after();
#pragma unroll
            for (int kk = 0; kk < 8; ++kk)
              umma_ss(T0, umma_smem_desc(band + kk * 2048, 16384, 1024), umma_smem_desc(qv + kk * 2048, 16384, 1024),
                      idesc_r, (st | kk) ? 1u : 0u);
            umma_commit(free_new);
            mbar_wait(full_prev, gs & 1);
            if (it.t > 0) {
              tc_fence_after();
#pragma unroll
              for (int kk = 0; kk < 8; ++kk)
                umma_ss(T1, umma_smem_desc(band + 2 * 16384 + kk * 2048, 16384, 1024),
                        umma_smem_desc(qv + kk * 2048, 16384, 1024), idesc_r, (st | kk) ? 1u : 0u);
            }
            umma_commit(free_prev);
            umma_commit(b_free + buf);
          }
        }
        umma_commit(acc_full + as);
        ++ti;
      }
    }
  } else if (warp < 10) {
    // ------------------------------------------------------------------ un-shift warps: two threads per tile row
    const int r = ((warp - 2) & 3) * 32 + lane;
    const int sft = 127 - r;
    const int a = sft >> 3, rem = sft & 7;
    const int rx = r & 7;
    const uint32_t band = smem_u32(smem + SM::BAND);
    auto body = [&](auto hs_tag) {
      constexpr int HS = decltype(hs_tag)::value;
      int gs = 0;
      for (int pass = 0; pass < n_pass; ++pass) {
        const int k = item_of(pass);
        if (k < 0) continue;
        const Item it = decode(k);
        for (int st = 0; st < it.nsteps; ++st, ++gs) {
          const int buf = gs & 1;
          const bool with_prev = (KIND == 0) ? (st > 0) : (it.t > 0);
          // this thread's half of the row from the TMA-written tile: conflict-free 128-bit shared loads (128-byte
          // swizzle); words outside [32*HS, 36 + 28*HS) are never used by the chunks this thread produces
          uint32_t W[64];
#pragma unroll
          for (int m = 0; m < 64; ++m) W[m] = 0u;
          mbar_wait(bar_ds + buf, (gs >> 1) & 1);
          {
            const uint32_t dsrow = smem_u32(smem + SM::DS + buf * 32768) + (uint32_t)r * 128u;
#pragma unroll
            for (int c = 8 * HS; c < 9 + 7 * HS; ++c) {
              const Half8 h8 = lds_half8(dsrow + (uint32_t)(c >> 3) * 16384u + (uint32_t)(((c & 7) ^ rx) * 16));
              W[4 * c] = h8.u.x; W[4 * c + 1] = h8.u.y; W[4 * c + 2] = h8.u.z; W[4 * c + 3] = h8.u.w;
            }
          }
          uint32_t out[17][4];
          unshift_row(W, rem, out);
          __syncwarp();
          if (lane == 0) mbar_arrive(ds_free + buf);  // the row is in registers
          if (KIND == 0) {
            if (gs > 0) mbar_wait(free_prev, (gs - 1) & 1);
            if (with_prev) store_band_half<HS>(band, r, a, 1, out);
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(full_prev);
            if (gs > 0) mbar_wait(free_new, (gs - 1) & 1);
            store_band_half<HS>(band, r, a, 0, out);
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(full_new);
          } else {
            if (gs > 0) mbar_wait(free_new, (gs - 1) & 1);
            store_band_half<HS>(band, r, a, 0, out);
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(full_new);
            if (gs > 0) mbar_wait(free_prev, (gs - 1) & 1);
            if (with_prev) store_band_half<HS>(band, r, a, 1, out);
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(full_prev);
          }
        }
      }
    };
    if (warp < 6) body(std::integral_constant<int, 0>{});
    else body(std::integral_constant<int, 1>{});
  } else {
    // ------------------------------------------------------------------ TMEM drain warps (one per lane quadrant)
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    int ti = 0;
    for (int pass = 0; pass < n_pass; ++pass) {
      const int k = item_of(pass);
      if (k < 0) continue;
      const Item it = decode(k);
      const int as = ti & 1;
      const uint32_t T0 = tmem_base + as * (2 * D) + lane_off;
      const uint32_t T1 = T0 + D;
      mbar_wait_backoff(acc_full + as, (ti >> 1) & 1, 64);
      tc_fence_after();
      if (KIND == 0) {
        const int i = it.I * 128 + row;
        __half* qrow = p.dq + ((long long)it.b * p.L + i) * p.lddq + (long long)it.h * p.dh;
        const bool al32 = ((reinterpret_cast<uintptr_t>(qrow)) & 31) == 0;
#pragma unroll 1
        for (int c = 0; c < D / 32; ++c) {
          uint32_t xu[32], xv[32];
          tmem_ld32(T0 + c * 32, xu);
          tmem_ld32(T1 + c * 32, xv);
          tmem_ld_wait();
          int n = p.dh - c * 32;
          n = n < 0 ? 0 : (n > 32 ? 32 : n);
          if (i < p.L && n > 0) {
            uint32_t pk[16];
#pragma unroll
            for (int e = 0; e < 16; ++e)
              pk[e] = pack_half2(__uint_as_float(xu[2 * e]) + __uint_as_float(xv[2 * e]),
                                 __uint_as_float(xu[2 * e + 1]) + __uint_as_float(xv[2 * e + 1]));
            stg_row32(qrow + c * 32, pk, n, al32);
          }
          float fu[32], fv[32];
#pragma unroll
          for (int e = 0; e < 32; ++e) {
            fu[e] = __uint_as_float(xu[e]);
            fv[e] = __uint_as_float(xv[e]);
          }
          const float su = warp_colsum32(fu, lane);
          const float sv = warp_colsum32(fv, lane);
          if (c * 32 + lane < p.dh) {
            atomicAdd(p.du + it.h * p.dh + c * 32 + lane, su);
            atomicAdd(p.dv + it.h * p.dh + c * 32 + lane, sv);
          }
        }
      } else {
        const int cb = p.L - 128 - 128 * it.t;
#pragma unroll 1
        for (int part = 0; part < 2; ++part) {
          if (part == 1 && it.t == 0) break;
          const int crow = cb + part * 128 + row;
          float* drow = p.dR + (long long)crow * p.lddr + (long long)it.h * p.dh;
#pragma unroll 1
          for (int c = 0; c < D / 32; ++c) {
            uint32_t x[32];
            tmem_ld32((part ? T1 : T0) + c * 32, x);
            tmem_ld_wait();
            if (crow >= 0 && crow < p.L) {
#pragma unroll
              for (int g = 0; g < 8; ++g)
                if (c * 32 + g * 4 < p.dh)
                  red_add_v4(drow + c * 32 + g * 4, __uint_as_float(x[4 * g]), __uint_as_float(x[4 * g + 1]),
                             __uint_as_float(x[4 * g + 2]), __uint_as_float(x[4 * g + 3]));
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_free + as);
      ++ti;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// dV = P^T . dO and dK = dS^T . (q+u), key-outer: item = (key tile J, head, sequence), accumulators [128 keys x dh]
// resident in TMEM over all query tiles I >= J of the item; P / dS tiles arrive by TMA and are read as MN-major A
// operands (M = key), dO_I / Qu_I as MN-major B operands - a pure TMA -> tcgen05 pipeline with no CUDA-core work per
// step (3-stage ring of {score tile, row tile} pairs, accumulators double-buffered for the epilogue). Replaces two
// batched causal GEMM launches that each re-read a score-sized operand with 128-byte row segments.
// Roles (192 threads): warp 0 TMA, warp 1 MMA, warps 2-5 epilogue (thread = key row).
// ------------------------------------------------------------------------------------------------------------------
constexpr int KV_THREADS = 192;
constexpr int KV_STAGES = 3;

struct KvParams {
  int L, H, B, dh, window;
  __half* dk;  // [B*L, ld], head h at column h*dh
  __half* dv;
  long long ld;
};

template <int D>
struct KvSmem {
  static constexpr int TILE = 128 * D * 2;
  static constexpr int STAGE = 32768 + TILE;  // [128 i][128 j] score tile (2 slabs) + [128 i][D] row tile
  static constexpr int BARS = KV_STAGES * STAGE;
  static constexpr int TOTAL = BARS + 256;
};

template <int D>
__global__ void __launch_bounds__(KV_THREADS, 1)
relattn_bwd_dkdv_kernel(const __grid_constant__ CUtensorMap tmP, const __grid_constant__ CUtensorMap tmDS,
                        const __grid_constant__ CUtensorMap tmDO, const __grid_constant__ CUtensorMap tmQu,
                        const KvParams p) {
  using SM = KvSmem<D>;
  constexpr int NSLAB = D / 64;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SM::BARS);
  uint64_t* full = bars;                   // [KV_STAGES] TMA landed
  uint64_t* empty = bars + KV_STAGES;      // [KV_STAGES] MMAs done reading
  uint64_t* acc_full = bars + 2 * KV_STAGES;
  uint64_t* acc_free = acc_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_free + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int nq = (p.L + 127) / 128;
  const int Wt = (p.window - 1 + 127) / 128;
  const int HB = p.H * p.B;
  const int n_items = nq * HB;
  const int G = gridDim.x;
  auto item_of = [&](int pass) -> int {
    const int c = (pass & 1) ? (G - 1 - (int)blockIdx.x) : (int)blockIdx.x;
    const int k = pass * G + c;
    return k < n_items ? k : -1;
  };
  const int n_pass = (n_items + G - 1) / G;
  struct Item {
    int J, h, b, nsteps;
  };
  auto decode = [&](int k) -> Item {  // heaviest first: key tile 0 is attended by the most query tiles
    Item it;
    it.J = k / HB;
    const int hb = k % HB;
    it.h = hb % p.H;
    it.b = hb / p.H;
    const int last = (it.J + Wt < nq - 1) ? it.J + Wt : nq - 1;
    it.nsteps = last - it.J + 1;
    return it;
  };

  if ((smem_u32(smem) & 1023u) != 0) __trap();
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmP);
    tma_prefetch_desc(&tmDS);
    tma_prefetch_desc(&tmDO);
    tma_prefetch_desc(&tmQu);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int i = 0; i < KV_STAGES; ++i) {
        mbar_init(full + i, 1);
        mbar_init(empty + i, 1);
      }
      mbar_init(acc_full + 0, 1);
      mbar_init(acc_full + 1, 1);
      mbar_init(acc_free + 0, 4);
      mbar_init(acc_free + 1, 4);
      mbar_fence_init();
    }
    __syncwarp();
    tmem_alloc<512>(tmem_slot);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();
  pdl_wait();

  if (warp == 0) {
    if (lane == 0) {
      int n = 0;  // sub-step counter: stage = n % KV_STAGES, use = n / KV_STAGES
      for (int pass = 0; pass < n_pass; ++pass) {
        const int k = item_of(pass);
        if (k < 0) continue;
        const Item it = decode(k);
        const int bh = it.b * p.H + it.h;
        for (int st = 0; st < it.nsteps; ++st) {
          const int I0 = (it.J + st) * 128;
#pragma unroll
          for (int sub = 0; sub < 2; ++sub, ++n) {
            const int sg = n % KV_STAGES;
            const int use = n / KV_STAGES;
            if (use > 0) mbar_wait(empty + sg, (use - 1) & 1);
            uint8_t* base = smem + sg * SM::STAGE;
            mbar_expect_tx(full + sg, SM::STAGE);
            const CUtensorMap* ms = sub ? &tmDS : &tmP;
            const CUtensorMap* mr = sub ? &tmQu : &tmDO;
            const int tile2 = ((bh * nq + (I0 >> 7)) * nq + it.J) * 2;  // slab index in the tiled P / dS
            tma_load_3d(base, ms, full + sg, 0, 0, tile2);
            tma_load_3d(base + 16384, ms, full + sg, 0, 0, tile2 + 1);
#pragma unroll
            for (int s = 0; s < NSLAB; ++s) tma_load_4d(base + 32768 + s * 16384, mr, full + sg, s * 64, I0, it.h, it.b);
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = umma_idesc(128, D, 1, 1, 0);  // A MN-major (score tile transposed), B MN-major (row tile)
      int n = 0, ti = 0;
      for (int pass = 0; pass < n_pass; ++pass) {
        const int k = item_of(pass);
        if (k < 0) continue;
        const Item it = decode(k);
        const int as = ti & 1;
        if (ti >= 2) {
          mbar_wait(acc_free + as, ((ti >> 1) - 1) & 1);
          tc_fence_after();
        }
        for (int st = 0; st < it.nsteps; ++st) {
#pragma unroll
          for (int sub = 0; sub < 2; ++sub, ++n) {
            const int sg = n % KV_STAGES;
            const int use = n / KV_STAGES;
            mbar_wait(full + sg, use & 1);
            tc_fence_after();
            const uint32_t a = smem_u32(smem + sg * SM::STAGE), b = a + 32768;
            const uint32_t acc = tmem_base + as * (2 * D) + sub * D;  // sub 0: dV, sub 1: dK
#pragma unroll
            for (int kk = 0; kk < 8; ++kk)
              umma_ss(acc, umma_smem_desc(a + kk * 2048, 16384, 1024), umma_smem_desc(b + kk * 2048, 16384, 1024), idesc,
                      (st | kk) ? 1u : 0u);
            umma_commit(empty + sg);
          }
        }
        umma_commit(acc_full + as);
        ++ti;
      }
    }
  } else {
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    int ti = 0;
    for (int pass = 0; pass < n_pass; ++pass) {
      const int k = item_of(pass);
      if (k < 0) continue;
      const Item it = decode(k);
      const int as = ti & 1;
      const int j = it.J * 128 + row;
      mbar_wait_backoff(acc_full + as, (ti >> 1) & 1, 64);
      tc_fence_after();
#pragma unroll 1
      for (int sub = 0; sub < 2; ++sub) {
        __half* orow = (sub ? p.dk : p.dv) + ((long long)it.b * p.L + j) * p.ld + (long long)it.h * p.dh;
        const bool al32 = ((reinterpret_cast<uintptr_t>(orow)) & 31) == 0;
#pragma unroll 1
        for (int c = 0; c < D / 32; ++c) {
          uint32_t x[32];
          tmem_ld32(tmem_base + as * (2 * D) + sub * D + lane_off + c * 32, x);
          tmem_ld_wait();
          int nv = p.dh - c * 32;
          nv = nv < 0 ? 0 : (nv > 32 ? 32 : nv);
          if (j < p.L && nv > 0) {
            uint32_t pk[16];
#pragma unroll
            for (int e = 0; e < 16; ++e) pk[e] = pack_half2(__uint_as_float(x[2 * e]), __uint_as_float(x[2 * e + 1]));
            stg_row32(orow + c * 32, pk, nv, al32);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_free + as);
      ++ti;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

static int make_head_map_bw(CUtensorMap* tm, const void* base, int dh, int L, int H, int B, long long ld) {
  uint64_t dims[4] = {(uint64_t)dh, (uint64_t)L, (uint64_t)H, (uint64_t)(B > 0 ? B : 1)};
  uint64_t str[3] = {(uint64_t)ld * 2, (uint64_t)dh * 2, (uint64_t)L * (uint64_t)ld * 2};
  uint32_t box[4] = {64, 128, 1, 1};
  return make_tmap_f16(tm, base, 4, dims, str, box);
}

// P / dS in the tiled layout written by db1_relattn_bwd_ds_tiled: [B*H][nq][nq][2 slabs][128 rows][64 columns] fp16.
// One box = one slab = 16 KB contiguous (128-byte swizzle on the way into shared memory): index = tile * 2 + slab.
static int make_ds_map(CUtensorMap* tm, const void* ds, int L, int BH) {
  const int nq = (L + 127) / 128;
  uint64_t dims[3] = {64, 128, (uint64_t)BH * nq * nq * 2};
  uint64_t str[2] = {128, 16384};
  uint32_t box[3] = {64, 128, 1};
  return make_tmap_f16(tm, ds, 3, dims, str, box);
}

template <int D, int KIND>
static int launch_band(const CUtensorMap* tm, const BandParams& p, cudaStream_t stream) {
  using SM = BandSmem<D, KIND>;
  static bool configured = false;
  if (!configured) {
    DB1_CUDA(cudaFuncSetAttribute(relattn_bwd_band_kernel<D, KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM::TOTAL));
    configured = true;
  }
  const int nq = (p.L + 127) / 128;
  const int Wt = (p.window - 1 + 127) / 128;
  const int nT = (nq - 1 < Wt ? nq - 1 : Wt) + 1;
  const int n_items = (KIND == 0) ? nq * p.H * p.B : nT * p.H * ((p.B + 1) / 2);
  const int grid = n_items < sm_count() ? n_items : sm_count();
  DB1_CUDA(launch_pdl(relattn_bwd_band_kernel<D, KIND>, dim3(grid), dim3(BW_THREADS), SM::TOTAL, stream, 1, tm[0], tm[1],
                      tm[2], tm[3], p));
  return 0;
}

template <int D>
static int launch_dkdv(const CUtensorMap* tm, const KvParams& p, cudaStream_t stream) {
  using SM = KvSmem<D>;
  static bool configured = false;
  if (!configured) {
    DB1_CUDA(cudaFuncSetAttribute(relattn_bwd_dkdv_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM::TOTAL));
    configured = true;
  }
  const int n_items = ((p.L + 127) / 128) * p.H * p.B;
  const int grid = n_items < sm_count() ? n_items : sm_count();
  DB1_CUDA(launch_pdl(relattn_bwd_dkdv_kernel<D>, dim3(grid), dim3(KV_THREADS), SM::TOTAL, stream, 1, tm[0], tm[1], tm[2],
                      tm[3], p));
  return 0;
}

}  // namespace db1

using namespace db1;

static int band_common_checks(const void* ds, int B, int L, int H, int dh, int window, long long ld_qkv, long long ld_r) {
  DB1_CHECK_ARG(ds != nullptr, "relattn_bwd: null dS");
  DB1_CHECK_ARG(B > 0 && L > 0 && H > 0, "relattn_bwd: bad shape B=%d L=%d H=%d", B, L, H);
  DB1_CHECK_ARG(dh % 8 == 0 && dh >= 8 && dh <= 128, "relattn_bwd: head dim %d unsupported (multiple of 8, <= 128)", dh);
  DB1_CHECK_ARG(window > 0, "relattn_bwd: window must be > 0");
  DB1_CHECK_ARG(ld_qkv % 8 == 0 && ld_r % 8 == 0, "relattn_bwd: row strides must be multiples of 8");
  DB1_CHECK_ARG((reinterpret_cast<uintptr_t>(ds) & 15) == 0, "relattn_bwd: dS must be 16-byte aligned");
  return 0;
}

extern "C" int db1_relattn_bwd_dq(const void* ds, const void* k, long long ld_qkv, const void* r, long long ld_r,
                                  void* dq, long long ld_dq, float* du, float* dv, int B, int L, int H, int dh,
                                  int window, void* stream_) {
  int e = band_common_checks(ds, B, L, H, dh, window, ld_qkv, ld_r);
  if (e) return e;
  DB1_CHECK_ARG(k && r && dq && du && dv, "relattn_bwd_dq: null pointer");
  DB1_CHECK_ARG(ld_dq % 8 == 0 && (reinterpret_cast<uintptr_t>(dq) & 15) == 0, "relattn_bwd_dq: dq rows must be 16-byte aligned");
  BandParams p;
  memset(&p, 0, sizeof(p));
  p.L = L; p.H = H; p.B = B; p.dh = dh; p.window = window;
  p.dS = (const __half*)ds; p.dq = (__half*)dq; p.lddq = ld_dq; p.du = du; p.dv = dv;
  CUtensorMap tm[4];
  if ((e = make_head_map_bw(&tm[0], k, dh, L, H, B, ld_qkv))) return e;
  if ((e = make_head_map_bw(&tm[1], r, dh, L, H, 1, ld_r))) return e;
  tm[2] = tm[0];
  if ((e = make_ds_map(&tm[3], ds, L, B * H))) return e;
  if (dh <= 64) return launch_band<64, 0>(tm, p, (cudaStream_t)stream_);
  return launch_band<128, 0>(tm, p, (cudaStream_t)stream_);
}

extern "C" int db1_relattn_bwd_dr(const void* ds, const void* qv, long long ld_qkv, float* dr, long long ld_dr, int B,
                                  int L, int H, int dh, int window, void* stream_) {
  int e = band_common_checks(ds, B, L, H, dh, window, ld_qkv, 8);
  if (e) return e;
  DB1_CHECK_ARG(qv && dr, "relattn_bwd_dr: null pointer");
  DB1_CHECK_ARG(ld_dr % 4 == 0 && (reinterpret_cast<uintptr_t>(dr) & 15) == 0, "relattn_bwd_dr: dR rows must be 16-byte aligned");
  BandParams p;
  memset(&p, 0, sizeof(p));
  p.L = L; p.H = H; p.B = B; p.dh = dh; p.window = window;
  p.dS = (const __half*)ds; p.dR = dr; p.lddr = ld_dr;
  CUtensorMap tm[4];
  if ((e = make_head_map_bw(&tm[2], qv, dh, L, H, B, ld_qkv))) return e;
  tm[0] = tm[2];
  tm[1] = tm[2];
  if ((e = make_ds_map(&tm[3], ds, L, B * H))) return e;
  if (dh <= 64) return launch_band<64, 1>(tm, p, (cudaStream_t)stream_);
  return launch_band<128, 1>(tm, p, (cudaStream_t)stream_);
}

extern "C" int db1_relattn_bwd_dkdv(const void* probs, const void* ds, const void* dout, long long ld_do, const void* qu,
                                    long long ld_qkv, void* dk, void* dv, long long ld_dkv, int B, int L, int H, int dh,
                                    int window, void* stream_) {
  int e = band_common_checks(ds, B, L, H, dh, window, ld_qkv, 8);
  if (e) return e;
  DB1_CHECK_ARG(probs && dout && qu && dk && dv, "relattn_bwd_dkdv: null pointer");
  DB1_CHECK_ARG((reinterpret_cast<uintptr_t>(probs) & 15) == 0, "relattn_bwd_dkdv: probs must be 16-byte aligned");
  DB1_CHECK_ARG(ld_do % 8 == 0 && ld_dkv % 8 == 0 &&
                    ((reinterpret_cast<uintptr_t>(dk) | reinterpret_cast<uintptr_t>(dv)) & 15) == 0,
                "relattn_bwd_dkdv: dk / dv rows must be 16-byte aligned");
  KvParams p;
  memset(&p, 0, sizeof(p));
  p.L = L; p.H = H; p.B = B; p.dh = dh; p.window = window;
  p.dk = (__half*)dk; p.dv = (__half*)dv; p.ld = ld_dkv;
  CUtensorMap tm[4];
  if ((e = make_ds_map(&tm[0], probs, L, B * H))) return e;
  if ((e = make_ds_map(&tm[1], ds, L, B * H))) return e;
  if ((e = make_head_map_bw(&tm[2], dout, dh, L, H, B, ld_do))) return e;
  if ((e = make_head_map_bw(&tm[3], qu, dh, L, H, B, ld_qkv))) return e;
  if (dh <= 64) return launch_dkdv<64>(tm, p, (cudaStream_t)stream_);
  return launch_dkdv<128>(tm, p, (cudaStream_t)stream_);
}
