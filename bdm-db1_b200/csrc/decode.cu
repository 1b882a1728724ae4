// Memory-augmented decode step: what evaluate_rl.py:157-266 drives through
// TransformerXL.forward(..., mems=...) (transformer_xl.py:124-133, :470-504), with the per-layer keys / values of the
// memory rows CACHED instead of recomputed from the hidden-state memory on every call.
//
//   db1_ring_append      new rows -> ring buffer slots (the `cat(mem, h)[:, -mem_len:]` of _update_mem, in place)
//   db1_relattn_decode   few-query relative-position attention over [ring cache | new rows]; HBM-bound on the cache:
//                        grid = (sequence, head) x key splits, K / V / r tiles staged in shared memory and shared by all
//                        the queries, fp32 online softmax per split, then a merge
//   db1_masked_argmax    masked_logits_for_action + argmax (evaluate_rl.py:96-138): argmax over a token range
//
// The decode attention is a CUDA-core kernel on purpose: with <= a few query rows per (sequence, head) there is no tile
// for a tensor core to fill; the work is one pass over K, V and r (2 * K * dh * 2 B + K * dh * 2 B per head).
#include <stdlib.h>

#include "../../include/db1_sm100.h"
#include "common.cuh"
#include "ptx.cuh"

namespace db1 {

constexpr int DEC_THREADS = 128;  // 4 warps; a warp takes one key at a time, lanes split the head dimension (4 each)

struct DecParams {
  const __half* qu;
  const __half* qv;
  const __half* knew;
  const __half* vnew;
  long long ld_qkv;   // row stride of the fused qkv buffer of the NEW rows ([B*Q, 4d])
  const __half* kc;
  const __half* vc;   // ring caches [B, cap, H*dh]
  int cap, head;      // logical memory row j lives in slot (head + j) % cap
  const int* head_dev;  // if not NULL the head is read from device memory (a captured CUDA graph replays with a moving head)
  const __half* r;    // [cap + Q, H*dh], row c <-> distance cap + Q - 1 - c
  long long ld_r;
  float* ws;          // [B*H*Q, S, dh + 2] partial (max, sum, acc)
  int B, Q, H, dh, S, window;
  float scale_log2;
};

DEVI float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

DEVI void ld4(const __half* p, float (&f)[4]) {
  const uint2 u = *reinterpret_cast<const uint2*>(p);
  const float2 a = unpack_half2(u.x), b = unpack_half2(u.y);
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y;
}

__global__ void __launch_bounds__(DEC_THREADS) relattn_decode_kernel(const DecParams p) {
  __shared__ float sm_m[4], sm_l[4];
  __shared__ float sm_acc[4][128];
  const int bhq = blockIdx.x, s = blockIdx.y;
  const int i = bhq % p.Q, h = (bhq / p.Q) % p.H, b = bhq / (p.Q * p.H);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int d0 = lane * 4;
  const bool act = d0 < p.dh;
  const int M = p.cap, K = p.cap + p.Q;
  const int head = p.head_dev ? *p.head_dev : p.head;
  const long long hoff = (long long)h * p.dh + d0;
  float qu[4] = {0, 0, 0, 0}, qv[4] = {0, 0, 0, 0};
  if (act) {
    ld4(p.qu + (long long)(b * p.Q + i) * p.ld_qkv + hoff, qu);
    ld4(p.qv + (long long)(b * p.Q + i) * p.ld_qkv + hoff, qv);
  }
  // keys query i may see: j <= i + M (causal) and M + i - j < window
  int jlo = M + i - p.window + 1;
  if (jlo < 0) jlo = 0;
  const int jhi = i + M;  // inclusive
  const int chunk = (K + p.S - 1) / p.S;
  int j0 = s * chunk, j1 = j0 + chunk;
  if (j0 < jlo) j0 = jlo;
  if (j1 > jhi + 1) j1 = jhi + 1;
  float m = -INFINITY, l = 0.f, acc[4] = {0, 0, 0, 0};
  for (int j = j0 + warp; j < j1; j += 4) {
    const __half *kp, *vp;
    if (j < M) {
      int slot = head + j;
      if (slot >= p.cap) slot -= p.cap;
      const long long ro = ((long long)b * p.cap + slot) * ((long long)p.H * p.dh) + hoff;
      kp = p.kc + ro;
      vp = p.vc + ro;
    } else {
      const long long ro = (long long)(b * p.Q + (j - M)) * p.ld_qkv + hoff;
      kp = p.knew + ro;
      vp = p.vnew + ro;
    }
    float k4[4] = {0, 0, 0, 0}, r4[4] = {0, 0, 0, 0}, v4[4] = {0, 0, 0, 0};
    if (act) {
      ld4(kp, k4);
      ld4(p.r + (long long)(p.Q - 1 - i + j) * p.ld_r + hoff, r4);
      ld4(vp, v4);
    }
    float part = 0.f;
#pragma unroll
    for (int e = 0; e < 4; ++e) part = fmaf(qu[e], k4[e], fmaf(qv[e], r4[e], part));
    const float sc = warp_sum(part) * p.scale_log2;
    const float mn = fmaxf(m, sc);
    const float f = exp2f(m - mn), pr = exp2f(sc - mn);
    l = l * f + pr;
#pragma unroll
    for (int e = 0; e < 4; ++e) acc[e] = acc[e] * f + pr * v4[e];
    m = mn;
  }
  if (lane == 0) {
    sm_m[warp] = m;
    sm_l[warp] = l;
  }
#pragma unroll
  for (int e = 0; e < 4; ++e) sm_acc[warp][d0 + e] = acc[e];
  __syncthreads();
  if (warp == 0) {
    float mm = fmaxf(fmaxf(sm_m[0], sm_m[1]), fmaxf(sm_m[2], sm_m[3]));
    float ll = 0.f, o[4] = {0, 0, 0, 0};
#pragma unroll
    for (int w = 0; w < 4; ++w) {
      const float f = (sm_m[w] == -INFINITY) ? 0.f : exp2f(sm_m[w] - mm);
      ll += sm_l[w] * f;
#pragma unroll
      for (int e = 0; e < 4; ++e) o[e] += sm_acc[w][d0 + e] * f;
    }
    float* wp = p.ws + ((long long)bhq * p.S + s) * (p.dh + 2);
    if (lane == 0) {
      wp[0] = mm;
      wp[1] = ll;
    }
    if (act) {
#pragma unroll
      for (int e = 0; e < 4; ++e) wp[2 + d0 + e] = o[e];
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// Tiled variant (head dim % 8 == 0): CTA = (sequence, head, key split). The split's keys are walked in tiles of 64; per
// tile the K, V and r rows are staged in shared memory with every thread's 16-byte loads issued back to back (the
// per-key kernel above exposes one memory round trip per key), then for blocks of up to 8 queries: scores (thread =
// key x query subset, 128-bit conflict-free shared loads: row pitch dh + 8 halves), online softmax per query (one warp
// per query), P . V (thread = 8 head dims x key slice, the slices reduced through shared memory at the end). All
// queries of a (sequence, head) share the staged rows, so K / V / r cross HBM once per split whatever Q is.
// ------------------------------------------------------------------------------------------------------------------
constexpr int DT_KEYS = 64;  // keys per tile
constexpr int DT_QB = 8;     // queries per block

template <int DH>
struct DecSmem {
  static constexpr int PITCH = DH + 8;                      // halves per staged row
  static constexpr int K_OFF = 0;                           // [64][PITCH] halves
  static constexpr int V_OFF = K_OFF + DT_KEYS * PITCH * 2;
  static constexpr int R_OFF = V_OFF + DT_KEYS * PITCH * 2; // [64 + 7][PITCH]
  static constexpr int Q_OFF = R_OFF + (DT_KEYS + DT_QB - 1) * PITCH * 2;  // qu[8][DH], qv[8][DH] halves
  static constexpr int S_OFF = Q_OFF + 2 * DT_QB * DH * 2;  // scores / probabilities [8][64] floats
  static constexpr int ST_OFF = S_OFF + DT_QB * DT_KEYS * 4;  // m[8], l[8], f[8] floats
  static constexpr int PART_OFF = ST_OFF + 3 * DT_QB * 4;   // [8 groups][DH] floats
  static constexpr int TOTAL = PART_OFF + 8 * DH * 4;
};

DEVI float dot8h(const uint4& a, const uint4& b, float acc) {
  const float2 a0 = unpack_half2(a.x), a1 = unpack_half2(a.y), a2 = unpack_half2(a.z), a3 = unpack_half2(a.w);
  const float2 b0 = unpack_half2(b.x), b1 = unpack_half2(b.y), b2 = unpack_half2(b.z), b3 = unpack_half2(b.w);
  acc = fmaf(a0.x, b0.x, acc); acc = fmaf(a0.y, b0.y, acc);
  acc = fmaf(a1.x, b1.x, acc); acc = fmaf(a1.y, b1.y, acc);
  acc = fmaf(a2.x, b2.x, acc); acc = fmaf(a2.y, b2.y, acc);
  acc = fmaf(a3.x, b3.x, acc); acc = fmaf(a3.y, b3.y, acc);
  return acc;
}

template <int DH>
__global__ void __launch_bounds__(DEC_THREADS) relattn_decode_tiled_kernel(const DecParams p) {
  using SM = DecSmem<DH>;
  constexpr int PITCH = SM::PITCH;
  constexpr int NC = DH / 8;  // 16-byte chunks per row
  extern __shared__ __align__(16) uint8_t dsm[];
  __half* sK = reinterpret_cast<__half*>(dsm + SM::K_OFF);
  __half* sV = reinterpret_cast<__half*>(dsm + SM::V_OFF);
  __half* sR = reinterpret_cast<__half*>(dsm + SM::R_OFF);
  __half* sQu = reinterpret_cast<__half*>(dsm + SM::Q_OFF);
  __half* sQv = sQu + DT_QB * DH;
  float* sS = reinterpret_cast<float*>(dsm + SM::S_OFF);
  float* sM = reinterpret_cast<float*>(dsm + SM::ST_OFF);
  float* sL = sM + DT_QB;
  float* sF = sL + DT_QB;
  float* sPart = reinterpret_cast<float*>(dsm + SM::PART_OFF);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int h = blockIdx.x % p.H, b = blockIdx.x / p.H, s = blockIdx.y;
  const int M = p.cap, Q = p.Q, K = p.cap + p.Q;
  const int head = p.head_dev ? *p.head_dev : p.head;
  const long long hoff = (long long)h * p.dh;
  const long long crow = (long long)p.H * p.dh;  // cache row stride
  const int chunk = (K + p.S - 1) / p.S;
  const int c0 = s * chunk;
  const int c1 = (c0 + chunk < K) ? c0 + chunk : K;

  for (int qb = 0; qb < Q; qb += DT_QB) {
    const int nq = (Q - qb < DT_QB) ? Q - qb : DT_QB;
    // key range any query of the block may see, cut to this split
    int jlo = M + qb - p.window + 1;
    if (jlo < 0) jlo = 0;
    const int jhi = M + qb + nq - 1;  // inclusive
    const int j0 = c0 > jlo ? c0 : jlo;
    const int j1 = (c1 < jhi + 1) ? c1 : jhi + 1;
    // P.V mapping: 8 groups of 16 threads; query qi gets KS groups, each a key slice jj = ks (mod KS)
    const int nq2 = nq <= 1 ? 1 : (nq <= 2 ? 2 : (nq <= 4 ? 4 : 8));
    const int KS = 8 / nq2;
    const int grp = tid >> 4, dc = tid & 15;
    const int pv_q = grp / KS, pv_ks = grp % KS;
    float acc[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = 0.f;
    __syncthreads();  // the previous block's readers of sQu / sM / sPart are done
    for (int idx = tid; idx < nq * NC; idx += DEC_THREADS) {
      const int qi = idx / NC, c = idx % NC;
      const long long ro = (long long)(b * Q + qb + qi) * p.ld_qkv + hoff + c * 8;
      const bool act = c * 8 < p.dh;
      *reinterpret_cast<uint4*>(sQu + qi * DH + c * 8) = act ? *reinterpret_cast<const uint4*>(p.qu + ro) : make_uint4(0, 0, 0, 0);
      *reinterpret_cast<uint4*>(sQv + qi * DH + c * 8) = act ? *reinterpret_cast<const uint4*>(p.qv + ro) : make_uint4(0, 0, 0, 0);
    }
    if (tid < DT_QB) {
      sM[tid] = -INFINITY;
      sL[tid] = 0.f;
    }
    for (int t0 = j0; t0 < j1; t0 += DT_KEYS) {
      const int len = (j1 - t0 < DT_KEYS) ? j1 - t0 : DT_KEYS;
      __syncthreads();  // previous tile fully consumed (and the block prologue visible)
      // ---- stage K, V rows [t0, t0 + len) and r rows [rbase, rbase + len + nq - 1): fixed trip counts, every load of
      // a thread issued before its first shared store (one memory round trip per tile instead of one per row)
      const int rbase = t0 + Q - 1 - (qb + nq - 1);  // r row of (last query of the block, first key of the tile)
      {
        constexpr int KIT = DT_KEYS * NC / DEC_THREADS;
        constexpr int RIT = ((DT_KEYS + DT_QB - 1) * NC + DEC_THREADS - 1) / DEC_THREADS;
        uint4 kreg[KIT], vreg[KIT], rreg[RIT];
#pragma unroll
        for (int it = 0; it < KIT; ++it) {
          const int idx = tid + it * DEC_THREADS;
          const int jj = idx / NC, c = idx % NC;
          const int j = t0 + jj;
          kreg[it] = vreg[it] = make_uint4(0u, 0u, 0u, 0u);
          if (jj < len && c * 8 < p.dh) {
            if (j < M) {
              int slot = head + j;
              if (slot >= p.cap) slot -= p.cap;
              const long long ro = ((long long)b * p.cap + slot) * crow + hoff + c * 8;
              kreg[it] = *reinterpret_cast<const uint4*>(p.kc + ro);
              vreg[it] = *reinterpret_cast<const uint4*>(p.vc + ro);
            } else {
              const long long ro = (long long)(b * Q + (j - M)) * p.ld_qkv + hoff + c * 8;
              kreg[it] = *reinterpret_cast<const uint4*>(p.knew + ro);
              vreg[it] = *reinterpret_cast<const uint4*>(p.vnew + ro);
            }
          }
        }
#pragma unroll
        for (int it = 0; it < RIT; ++it) {
          const int idx = tid + it * DEC_THREADS;
          const int rr = idx / NC, c = idx % NC;
          const int ri = rbase + rr;
          rreg[it] = make_uint4(0u, 0u, 0u, 0u);
          if (rr < len + nq - 1 && c * 8 < p.dh && ri >= 0 && ri < K)
            rreg[it] = *reinterpret_cast<const uint4*>(p.r + (long long)ri * p.ld_r + hoff + c * 8);
        }
#pragma unroll
        for (int it = 0; it < KIT; ++it) {
          const int idx = tid + it * DEC_THREADS;
          const int jj = idx / NC, c = idx % NC;
          *reinterpret_cast<uint4*>(sK + jj * PITCH + c * 8) = kreg[it];
          *reinterpret_cast<uint4*>(sV + jj * PITCH + c * 8) = vreg[it];
        }
#pragma unroll
        for (int it = 0; it < RIT; ++it) {
          const int idx = tid + it * DEC_THREADS;
          const int rr = idx / NC, c = idx % NC;
          if (rr < DT_KEYS + DT_QB - 1) *reinterpret_cast<uint4*>(sR + rr * PITCH + c * 8) = rreg[it];
        }
      }
      __syncthreads();
      // ---- scores: thread = key jj x queries qs, qs + 2, ...
      {
        const int jj = tid & 63, qs = tid >> 6;
        float sc[4] = {0.f, 0.f, 0.f, 0.f};
        if (jj < len) {
#pragma unroll 4
          for (int c = 0; c < NC; ++c) {
            const uint4 kc = *reinterpret_cast<const uint4*>(sK + jj * PITCH + c * 8);
#pragma unroll
            for (int z = 0; z < 4; ++z) {
              const int qi = qs + 2 * z;
              if (qi < nq) {
                const uint4 qu = *reinterpret_cast<const uint4*>(sQu + qi * DH + c * 8);
                const uint4 qv = *reinterpret_cast<const uint4*>(sQv + qi * DH + c * 8);
                const uint4 rc = *reinterpret_cast<const uint4*>(sR + (jj + nq - 1 - qi) * PITCH + c * 8);
                sc[z] = dot8h(qv, rc, dot8h(qu, kc, sc[z]));
              }
            }
          }
        }
#pragma unroll
        for (int z = 0; z < 4; ++z) {
          const int qi = qs + 2 * z;
          if (qi < nq) {
            const int i = qb + qi, j = t0 + jj;
            const bool ok = jj < len && j <= M + i && M + i - j < p.window;
            sS[qi * DT_KEYS + jj] = ok ? sc[z] * p.scale_log2 : -INFINITY;
          }
        }
      }
      __syncthreads();
      // ---- online softmax: warp w takes queries w, w + 4
      for (int qi = warp; qi < nq; qi += 4) {
        const float s0 = sS[qi * DT_KEYS + lane], s1 = sS[qi * DT_KEYS + 32 + lane];
        float mx = fmaxf(s0, s1);
#pragma unroll
        for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        const float m_old = sM[qi];
        const float m_new = fmaxf(m_old, mx);
        float p0 = 0.f, p1 = 0.f, f = 1.f;
        if (m_new != -INFINITY) {
          p0 = exp2f(s0 - m_new);
          p1 = exp2f(s1 - m_new);
          f = (m_old == -INFINITY) ? 0.f : exp2f(m_old - m_new);
        }
        float sum = p0 + p1;
#pragma unroll
        for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        sS[qi * DT_KEYS + lane] = p0;
        sS[qi * DT_KEYS + 32 + lane] = p1;
        if (lane == 0) {
          sM[qi] = m_new;
          sL[qi] = sL[qi] * f + sum;
          sF[qi] = f;
        }
      }
      __syncthreads();
      // ---- P . V: thread = 8 head dims (dc) of query pv_q over the keys jj = pv_ks (mod KS)
      if (pv_q < nq) {
        const float f = sF[pv_q];
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] *= f;
        if (dc < NC) {
          for (int jj = pv_ks; jj < len; jj += KS) {
            const float pr = sS[pv_q * DT_KEYS + jj];
            const uint4 vv = *reinterpret_cast<const uint4*>(sV + jj * PITCH + dc * 8);
            const float2 v0 = unpack_half2(vv.x), v1 = unpack_half2(vv.y), v2 = unpack_half2(vv.z), v3 = unpack_half2(vv.w);
            acc[0] = fmaf(pr, v0.x, acc[0]); acc[1] = fmaf(pr, v0.y, acc[1]);
            acc[2] = fmaf(pr, v1.x, acc[2]); acc[3] = fmaf(pr, v1.y, acc[3]);
            acc[4] = fmaf(pr, v2.x, acc[4]); acc[5] = fmaf(pr, v2.y, acc[5]);
            acc[6] = fmaf(pr, v3.x, acc[6]); acc[7] = fmaf(pr, v3.y, acc[7]);
          }
        }
      }
    }
    // ---- reduce the key slices of every query and write the split's partial (max, sum, acc)
    __syncthreads();
    if (dc < NC) {
#pragma unroll
      for (int e = 0; e < 8; ++e) sPart[grp * DH + dc * 8 + e] = acc[e];
    }
    __syncthreads();
    for (int idx = tid; idx < nq * DH; idx += DEC_THREADS) {
      const int qi = idx / DH, d = idx % DH;
      if (d < p.dh) {
        float o = 0.f;
        for (int ks = 0; ks < KS; ++ks) o += sPart[(qi * KS + ks) * DH + d];
        const long long bhq = ((long long)b * p.H + h) * Q + qb + qi;
        float* wp = p.ws + (bhq * p.S + s) * (p.dh + 2);
        wp[2 + d] = o;
        if (d == 0) {
          wp[0] = sM[qi];
          wp[1] = sL[qi];
        }
      }
    }
  }
}

__global__ void relattn_decode_merge_kernel(const float* __restrict__ ws, __half* __restrict__ out, long long ld_out, int B,
                                            int Q, int H, int dh, int S) {
  const int bhq = blockIdx.x;
  const int i = bhq % Q, h = (bhq / Q) % H, b = bhq / (Q * H);
  const int d = threadIdx.x;
  if (d >= dh) return;
  const float* wp = ws + (long long)bhq * S * (dh + 2);
  float mm = -INFINITY;
  for (int s = 0; s < S; ++s) mm = fmaxf(mm, wp[s * (dh + 2)]);
  float ll = 0.f, o = 0.f;
  for (int s = 0; s < S; ++s) {
    const float ms = wp[s * (dh + 2)];
    const float f = (ms == -INFINITY) ? 0.f : exp2f(ms - mm);
    ll += wp[s * (dh + 2) + 1] * f;
    o += wp[s * (dh + 2) + 2 + d] * f;
  }
  out[(long long)(b * Q + i) * ld_out + (long long)h * dh + d] = __float2half_rn(o / ll);
}

// ring_z[b][(head + t) % cap][:] = src_z[b*Q + t][:] for up to three (source, ring) pairs in one launch (blockIdx.y):
// the layer input rows and the new k / v rows (n % 8 == 0 halves per row)
struct RingArgs {
  const __half* src[3];
  long long ld[3];
  __half* dst[3];
};
__global__ void ring_append_kernel(const RingArgs a, int cap, int head_val, const int* __restrict__ head_dev, int Q, int n) {
  const int row = blockIdx.x;  // b * Q + t
  const int z = blockIdx.y;
  const int b = row / Q, t = row % Q;
  int slot = (head_dev ? *head_dev : head_val) + t;
  if (slot >= cap) slot -= cap;
  const __half* s = a.src[z] + (long long)row * a.ld[z];
  __half* d = a.dst[z] + ((long long)b * cap + slot) * n;
  for (int c = threadIdx.x * 8; c < n; c += blockDim.x * 8) st_half8(d + c, ld_half8(s + c));
}

__global__ void masked_argmax_kernel(const __half* __restrict__ logits, long long ld, int lo, int hi,
                                     const float* __restrict__ add_mask, long long* __restrict__ out) {
  __shared__ float sv[32];
  __shared__ int si[32];
  const __half* z = logits + (long long)blockIdx.x * ld;
  float best = -INFINITY;
  int bi = hi;
  for (int c = lo + threadIdx.x; c < hi; c += blockDim.x) {
    float v = __half2float(z[c]);
    if (add_mask) v -= add_mask[c - lo];
    if (v > best) {  // ascending c per thread: the first maximum is kept
      best = v;
      bi = c;
    }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov > best || (ov == best && oi < bi)) {
      best = ov;
      bi = oi;
    }
  }
  if (lane == 0) {
    sv[warp] = best;
    si[warp] = bi;
  }
  __syncthreads();
  if (warp == 0) {
    const int nw = blockDim.x >> 5;
    best = lane < nw ? sv[lane] : -INFINITY;
    bi = lane < nw ? si[lane] : hi;
#pragma unroll
    for (int o = 16; o; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > best || (ov == best && oi < bi)) {
        best = ov;
        bi = oi;
      }
    }
    if (lane == 0) out[blockIdx.x] = bi;
  }
}

}  // namespace db1

using namespace db1;

extern "C" int db1_decode_splits(int B, int Q, int H) {
  (void)Q;  // the queries of a (sequence, head) share a CTA: the split count follows B * H alone
  int s = (2 * sm_count_physical()) / (B * H > 0 ? B * H : 1);
  if (s < 1) s = 1;
  if (s > 32) s = 32;
  return s;
}

extern "C" int db1_relattn_decode(const void* qu, const void* qv, const void* knew, const void* vnew, long long ld_qkv,
                                  const void* kcache, const void* vcache, int cap, int head, const int* head_dev,
                                  const void* r, long long ld_r, void* out, long long ld_out, float* ws, long long ws_floats,
                                  int B, int Q, int H, int dh, int window, float scale, void* stream_) {
  DB1_CHECK_ARG(qu && qv && knew && vnew && kcache && vcache && r && out && ws, "relattn_decode: null pointer");
  DB1_CHECK_ARG(B > 0 && Q > 0 && H > 0 && cap > 0 && head >= 0 && head < cap, "relattn_decode: bad shape");
  DB1_CHECK_ARG(dh % 4 == 0 && dh >= 4 && dh <= 128, "relattn_decode: head dim %d unsupported (multiple of 4, <= 128)", dh);
  DB1_CHECK_ARG(ld_qkv % 4 == 0 && ld_r % 4 == 0 && window > 0, "relattn_decode: bad strides / window");
  const int S = db1_decode_splits(B, Q, H);
  DB1_CHECK_ARG(ws_floats >= (long long)B * Q * H * S * (dh + 2), "relattn_decode: workspace too small");
  DecParams p;
  p.qu = (const __half*)qu; p.qv = (const __half*)qv; p.knew = (const __half*)knew; p.vnew = (const __half*)vnew;
  p.ld_qkv = ld_qkv; p.kc = (const __half*)kcache; p.vc = (const __half*)vcache; p.cap = cap; p.head = head; p.head_dev = head_dev;
  p.r = (const __half*)r; p.ld_r = ld_r; p.ws = ws; p.B = B; p.Q = Q; p.H = H; p.dh = dh; p.S = S; p.window = window;
  p.scale_log2 = scale * 1.4426950408889634f;
  cudaStream_t st = (cudaStream_t)stream_;
  const bool al16 = ((((uintptr_t)qu | (uintptr_t)qv | (uintptr_t)knew | (uintptr_t)vnew | (uintptr_t)kcache |
                       (uintptr_t)vcache | (uintptr_t)r) & 15) == 0) && ld_qkv % 8 == 0 && ld_r % 8 == 0 && dh % 8 == 0;
  static int force_old = -1;
  if (force_old < 0) force_old = getenv("DB1_DECODE_PER_KEY") ? 1 : 0;
  if (al16 && !force_old) {
    if (dh <= 64) {
      static bool cfg64 = false;
      if (!cfg64) {
        DB1_CUDA(cudaFuncSetAttribute(relattn_decode_tiled_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, DecSmem<64>::TOTAL));
        cfg64 = true;
      }
      relattn_decode_tiled_kernel<64><<<dim3(B * H, S), DEC_THREADS, DecSmem<64>::TOTAL, st>>>(p);
    } else {
      static bool cfg128 = false;
      if (!cfg128) {
        DB1_CUDA(cudaFuncSetAttribute(relattn_decode_tiled_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, DecSmem<128>::TOTAL));
        cfg128 = true;
      }
      relattn_decode_tiled_kernel<128><<<dim3(B * H, S), DEC_THREADS, DecSmem<128>::TOTAL, st>>>(p);
    }
  } else {
    relattn_decode_kernel<<<dim3(B * Q * H, S), DEC_THREADS, 0, st>>>(p);
  }
  relattn_decode_merge_kernel<<<B * Q * H, 128, 0, st>>>(ws, (__half*)out, ld_out, B, Q, H, dh, S);
  DB1_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int db1_ring_append(const void* const* srcs, const long long* ld_srcs, void* const* rings, int nring, int cap,
                               int head, const int* head_dev, int B, int Q, int n, void* stream_) {
  DB1_CHECK_ARG(srcs && ld_srcs && rings && nring >= 1 && nring <= 3 && cap > 0 && head >= 0 && head < cap && B > 0 && Q > 0 &&
                    Q <= cap && n > 0 && n % 8 == 0,
                "ring_append: bad arguments");
  RingArgs a;
  for (int z = 0; z < 3; ++z) {
    const int y = z < nring ? z : 0;
    DB1_CHECK_ARG(srcs[y] && rings[y] && ld_srcs[y] % 8 == 0, "ring_append: bad source / ring %d", y);
    a.src[z] = (const __half*)srcs[y];
    a.ld[z] = ld_srcs[y];
    a.dst[z] = (__half*)rings[y];
  }
  ring_append_kernel<<<dim3(B * Q, nring), 128, 0, (cudaStream_t)stream_>>>(a, cap, head, head_dev, Q, n);
  DB1_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int db1_masked_argmax(const void* logits, long long ld, int rows, int lo, int hi, const float* add_mask,
                                 long long* out, void* stream_) {
  DB1_CHECK_ARG(logits && out && rows > 0 && lo >= 0 && hi > lo, "masked_argmax: bad arguments");
  masked_argmax_kernel<<<rows, 256, 0, (cudaStream_t)stream_>>>((const __half*)logits, ld, lo, hi, add_mask, out);
  DB1_CUDA(cudaGetLastError());
  return 0;
}
