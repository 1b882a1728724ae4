// Memory-augmented decode step: what evaluate_rl.py:157-266 drives through
// TransformerXL.forward(..., mems=...) (transformer_xl.py:124-133, :470-504), with the per-layer keys / values of the
// memory rows CACHED instead of recomputed from the hidden-state memory on every call.
//
//   db1_ring_append      new rows -> ring buffer slots (the `cat(mem, h)[:, -mem_len:]` of _update_mem, in place)
//   db1_relattn_decode   few-query relative-position attention over [ring cache | new rows]; HBM-bound on the cache:
//                        grid = (sequence, head, query) x key splits, fp32 online softmax per split, then a merge
//   db1_masked_argmax    masked_logits_for_action + argmax (evaluate_rl.py:96-138): argmax over a token range
//
// The decode attention is a CUDA-core kernel on purpose: with <= a few query rows per (sequence, head) there is no tile
// for a tensor core to fill; the work is one pass over K, V and r (2 * K * dh * 2 B + K * dh * 2 B per head).
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
  int s = (2 * sm_count_physical()) / (B * Q * H > 0 ? B * Q * H : 1);
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
  relattn_decode_kernel<<<dim3(B * Q * H, S), DEC_THREADS, 0, st>>>(p);
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
