// HBM-bound kernels of the DB1 path: LayerNorm (+ dropout-mask replay, bias/affine grads), masked cross-entropy,
// embedding assembly, small reductions. All are single-pass over their big operand, 16-byte vectorised, fp32 math.
//
// Reference call sites: nn.LayerNorm at transformer_xl.py:238, :290; embedding assembly :621-703; dropout :545, :575;
// loss :602-613; PositionalEmbedding :34-50, :569-575.
#include "../../include/db1_sm100.h"
#include "common.cuh"
#include "ptx.cuh"

#include <stdlib.h>

namespace db1 {

typedef Half8 H8;  // 8 x fp16 carried in a uint4 (ptx.cuh): one 128-bit access
DEVI void h8_to_f(const H8& v, float (&f)[8]) { half8_to_float(v, f); }
DEVI H8 f_to_h8(const float (&f)[8]) { return float_to_half8(f); }

DEVI float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
DEVI float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// block-wide sum of two values (blockDim.x multiple of 32, <= 1024); result broadcast to all threads
DEVI float2 block_sum2(float a, float b, float2* scratch) {
  a = warp_sum(a);
  b = warp_sum(b);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = blockDim.x >> 5;
  __syncthreads();
  if (l == 0) scratch[w] = make_float2(a, b);
  __syncthreads();
  float2 t = (l < nw) ? scratch[l] : make_float2(0.f, 0.f);
  t.x = warp_sum(t.x);
  t.y = warp_sum(t.y);
  return t;
}
DEVI float block_max(float a, float* scratch) {
  a = warp_max(a);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = blockDim.x >> 5;
  __syncthreads();
  if (l == 0) scratch[w] = a;
  __syncthreads();
  float t = (l < nw) ? scratch[l] : -INFINITY;
  return warp_max(t);
}

// ------------------------------------------------------------------------------------------------ LayerNorm forward
// one CTA (256 threads) per row; d % 8 == 0; the row is kept in registers (up to LN_MAXC 16-byte chunks per thread)
constexpr int LN_THREADS = 256;
constexpr int LN_MAXC = 4;  // d <= 256 * 8 * 4 = 8192

__global__ void __launch_bounds__(LN_THREADS)
ln_fwd_kernel(const __half* __restrict__ y, const __half* __restrict__ gamma, const __half* __restrict__ beta,
              __half* __restrict__ out, float* __restrict__ stats, int rows, int d, float eps) {
  __shared__ float2 scratch[32];
  const int nch = d / 8;
  for (int row = blockIdx.x; row < rows; row += gridDim.x) {
    const __half* yr = y + (size_t)row * d;
    float x[LN_MAXC][8];
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < LN_MAXC; ++c) {
      const int ch = threadIdx.x + c * LN_THREADS;
      if (ch < nch) {
        h8_to_f(*reinterpret_cast<const H8*>(yr + ch * 8), x[c]);
#pragma unroll
        for (int i = 0; i < 8; ++i) s += x[c][i];
      }
    }
    const float mean = block_sum2(s, 0.f, scratch).x / (float)d;
    float v = 0.f;
#pragma unroll
    for (int c = 0; c < LN_MAXC; ++c) {
      const int ch = threadIdx.x + c * LN_THREADS;
      if (ch < nch) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float t = x[c][i] - mean;
          v += t * t;
        }
      }
    }
    const float var = block_sum2(v, 0.f, scratch).x / (float)d;
    const float rstd = rsqrtf(var + eps);
    if (threadIdx.x == 0) {
      stats[2 * row] = mean;
      stats[2 * row + 1] = rstd;
    }
#pragma unroll
    for (int c = 0; c < LN_MAXC; ++c) {
      const int ch = threadIdx.x + c * LN_THREADS;
      if (ch < nch) {
        float g[8], b[8], o[8];
        h8_to_f(*reinterpret_cast<const H8*>(gamma + ch * 8), g);
        h8_to_f(*reinterpret_cast<const H8*>(beta + ch * 8), b);
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = (x[c][i] - mean) * rstd * g[i] + b[i];
        *reinterpret_cast<H8*>(out + (size_t)row * d + ch * 8) = f_to_h8(o);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ LayerNorm backward
// dy = rstd * (dxhat - mean(dxhat) - xhat * mean(dxhat * xhat)),  dxhat = dout * gamma
// dz = dy * dropout_mask * scale  (mask replayed from (seed, row*d + col), same stream as the GEMM epilogue)
// column reductions (fp32 atomics, one per column per CTA): dgamma += dout*xhat, dbeta += dout, dbias += dz
__global__ void __launch_bounds__(LN_THREADS)
ln_bwd_kernel(const __half* __restrict__ dout, const __half* __restrict__ y, const __half* __restrict__ gamma,
              const float* __restrict__ stats, __half* __restrict__ dy, __half* __restrict__ dz,
              float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ dbias, int rows, int d,
              uint32_t drop_thr16, float drop_scale, uint64_t seed) {
  __shared__ float2 scratch[32];
  const int nch = d / 8;
  float ag[LN_MAXC][8], ab[LN_MAXC][8], az[LN_MAXC][8];
#pragma unroll
  for (int c = 0; c < LN_MAXC; ++c)
#pragma unroll
    for (int i = 0; i < 8; ++i) ag[c][i] = ab[c][i] = az[c][i] = 0.f;

  for (int row = blockIdx.x; row < rows; row += gridDim.x) {
    const float mean = stats[2 * row], rstd = stats[2 * row + 1];
    float xh[LN_MAXC][8], dx[LN_MAXC][8];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int c = 0; c < LN_MAXC; ++c) {
      const int ch = threadIdx.x + c * LN_THREADS;
      if (ch < nch) {
        float g[8], go[8], yv[8];
        h8_to_f(*reinterpret_cast<const H8*>(gamma + ch * 8), g);
        h8_to_f(*reinterpret_cast<const H8*>(dout + (size_t)row * d + ch * 8), go);
        h8_to_f(*reinterpret_cast<const H8*>(y + (size_t)row * d + ch * 8), yv);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          xh[c][i] = (yv[i] - mean) * rstd;
          dx[c][i] = go[i] * g[i];
          s1 += dx[c][i];
          s2 += dx[c][i] * xh[c][i];
          ag[c][i] += go[i] * xh[c][i];
          ab[c][i] += go[i];
        }
      }
    }
    const float2 t = block_sum2(s1, s2, scratch);
    const float m1 = t.x / (float)d, m2 = t.y / (float)d;
#pragma unroll
    for (int c = 0; c < LN_MAXC; ++c) {
      const int ch = threadIdx.x + c * LN_THREADS;
      if (ch < nch) {
        float o[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = rstd * (dx[c][i] - m1 - xh[c][i] * m2);
        const H8 hv = f_to_h8(o);
        *reinterpret_cast<H8*>(dy + (size_t)row * d + ch * 8) = hv;
        if (dz != nullptr || dbias != nullptr) {
          float z[8];
          h8_to_f(hv, z);  // the GEMM branch sees the fp16-rounded dy
          if (drop_thr16 && dz != nullptr) {
            const uint64_t e = (uint64_t)row * (uint64_t)d + (uint64_t)ch * 8;
            const uint64_t b0 = rng64(seed, e >> 2), b1 = rng64(seed, (e >> 2) + 1);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              z[i] = dropout_keep(b0, i, drop_thr16) ? z[i] * drop_scale : 0.f;
              z[4 + i] = dropout_keep(b1, i, drop_thr16) ? z[4 + i] * drop_scale : 0.f;
            }
            *reinterpret_cast<H8*>(dz + (size_t)row * d + ch * 8) = f_to_h8(z);
          }
          if (dbias != nullptr) {
#pragma unroll
            for (int i = 0; i < 8; ++i) az[c][i] += z[i];
          }
        }
      }
    }
  }
#pragma unroll
  for (int c = 0; c < LN_MAXC; ++c) {
    const int ch = threadIdx.x + c * LN_THREADS;
    if (ch < nch) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        atomicAdd(dgamma + ch * 8 + i, ag[c][i]);
        atomicAdd(dbeta + ch * 8 + i, ab[c][i]);
        if (dbias != nullptr) atomicAdd(dbias + ch * 8 + i, az[c][i]);
      }
    }
  }
}


// ---- warp-per-row variants (d == 256 * CH, CH <= 8): no block barriers, 8 rows per CTA in flight
template <int CH>
__global__ void __launch_bounds__(256)
ln_fwd_warp_kernel(const __half* __restrict__ y, const __half* __restrict__ gamma, const __half* __restrict__ beta,
                   __half* __restrict__ out, float* __restrict__ stats, int rows, float eps) {
  pdl_launch_dependents();
  pdl_wait();
  constexpr int d = 256 * CH;
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const __half* yr = y + (size_t)row * d;
  float x[CH][8];
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < CH; ++c) {
    h8_to_f(*reinterpret_cast<const H8*>(yr + (c * 32 + lane) * 8), x[c]);
#pragma unroll
    for (int i = 0; i < 8; ++i) s += x[c][i];
  }
  const float mean = warp_sum(s) / (float)d;
  float v = 0.f;
#pragma unroll
  for (int c = 0; c < CH; ++c)
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float t = x[c][i] - mean;
      v += t * t;
    }
  const float rstd = rsqrtf(warp_sum(v) / (float)d + eps);
  if (lane == 0) {
    stats[2 * row] = mean;
    stats[2 * row + 1] = rstd;
  }
#pragma unroll
  for (int c = 0; c < CH; ++c) {
    float g[8], b[8], o[8];
    h8_to_f(*reinterpret_cast<const H8*>(gamma + (c * 32 + lane) * 8), g);
    h8_to_f(*reinterpret_cast<const H8*>(beta + (c * 32 + lane) * 8), b);
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i] = (x[c][i] - mean) * rstd * g[i] + b[i];
    *reinterpret_cast<H8*>(out + (size_t)row * d + (c * 32 + lane) * 8) = f_to_h8(o);
  }
}

// Fused backward for d == THREADS * 8: every thread owns one 16-byte column chunk, a CTA walks groups of ROWS rows.
// Per group: one pass over dout / y (kept in registers), ONE block reduction for the 2 * ROWS row sums, dy (and
// dz = dy * dropout mask) written, and the column sums dgamma / dbeta / dbias accumulated in registers across all the
// CTA's rows; one vector reduction per column chunk per CTA at the end. Single pass: dout, y read once; dy, dz written once.
DEVI void red_add_f32x4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// ROWS rows per group; the NEXT group's dout / y (packed fp16, 8 registers per row) are already in flight while the
// current group is reduced and written, so a CTA keeps 2 * ROWS * 2 16-byte loads per thread outstanding all the time.
template <int THREADS, int ROWS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB)
ln_bwd_fused_kernel(const __half* __restrict__ dout, const __half* __restrict__ y, const __half* __restrict__ gamma,
                    const float* __restrict__ stats, __half* __restrict__ dy, __half* __restrict__ dz,
                    float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ dbias, int rows,
                    uint32_t drop_thr16, float drop_scale, uint64_t seed) {
  pdl_launch_dependents();
  pdl_wait();
  constexpr int d = THREADS * 8;
  constexpr int NW = THREADS / 32;
  __shared__ float red[2][NW][2 * ROWS];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int off = tid * 8;
  float2 g2[4], ag2[4], ab2[4], az2[4];
  {
    const H8 hg = *reinterpret_cast<const H8*>(gamma + off);
    const uint32_t* w = reinterpret_cast<const uint32_t*>(&hg);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      g2[i] = unpack_half2(w[i]);
      ag2[i] = ab2[i] = az2[i] = make_float2(0.f, 0.f);
    }
  }
  const int ngroups = (rows + ROWS - 1) / ROWS;
  H8 ngo[ROWS], ny[ROWS];
  auto fetch = [&](int grp) {
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
      const int row = grp * ROWS + r;
      if (grp < ngroups && row < rows) {
        ngo[r] = *reinterpret_cast<const H8*>(dout + (size_t)row * d + off);
        ny[r] = *reinterpret_cast<const H8*>(y + (size_t)row * d + off);
      } else {
        ngo[r] = ny[r] = half8_zero();
      }
    }
  };
  fetch(blockIdx.x);
  int par = 0;
  for (int grp = blockIdx.x; grp < ngroups; grp += gridDim.x, par ^= 1) {
    const int r0 = grp * ROWS;
    H8 hgo[ROWS], hy[ROWS];
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
      hgo[r] = ngo[r];
      hy[r] = ny[r];
    }
    fetch(grp + gridDim.x);  // next group's loads fly under this group's reduction and stores
    // the packed fp16 inputs stay in registers (4 regs per row each); xhat / dxhat are recomputed in the second phase.
    // All per-element arithmetic runs on Blackwell's packed fp32 pairs (FFMA2 / FADD2 / FMUL2: one issue slot per two
    // elements): xhat = fma(y, rstd, -mean * rstd), and dy = fma(-c1, xhat, fma(dout * gamma, rstd, -c0)) with
    // c0 = rstd * mean(dxhat), c1 = rstd * mean(dxhat * xhat).
    float mean[ROWS], rstd[ROWS], part[2 * ROWS];
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
      const int rr = (r0 + r < rows) ? r0 + r : rows - 1;
      mean[r] = stats[2 * rr];
      rstd[r] = stats[2 * rr + 1];
      const float2 rs2 = make_float2(rstd[r], rstd[r]);
      const float nm = -mean[r] * rstd[r];
      const float2 nm2 = make_float2(nm, nm);
      float2 s1 = make_float2(0.f, 0.f), s2 = make_float2(0.f, 0.f);
      const uint32_t* wg = reinterpret_cast<const uint32_t*>(&hgo[r]);
      const uint32_t* wy = reinterpret_cast<const uint32_t*>(&hy[r]);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 go = unpack_half2(wg[i]);
        const float2 xh = __ffma2_rn(unpack_half2(wy[i]), rs2, nm2);
        const float2 dx = __fmul2_rn(go, g2[i]);
        s1 = __fadd2_rn(s1, dx);
        s2 = __ffma2_rn(dx, xh, s2);
        ag2[i] = __ffma2_rn(go, xh, ag2[i]);  // out-of-range rows contribute go == 0
        ab2[i] = __fadd2_rn(ab2[i], go);
      }
      part[2 * r] = s1.x + s1.y;
      part[2 * r + 1] = s2.x + s2.y;
    }
#pragma unroll
    for (int k = 0; k < 2 * ROWS; ++k) part[k] = warp_sum(part[k]);
    // double-buffered scratch: one barrier per group (a warp can be at most one group ahead of the slowest reader)
    if (lane == 0) {
#pragma unroll
      for (int k = 0; k < 2 * ROWS; ++k) red[par][warp][k] = part[k];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 2 * ROWS; ++k) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < NW; ++w) t += red[par][w][k];
      part[k] = t * (1.0f / (float)d);
    }
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
      if (r0 + r >= rows) break;
      const float2 rs2 = make_float2(rstd[r], rstd[r]);
      const float nm = -mean[r] * rstd[r];
      const float2 nm2 = make_float2(nm, nm);
      const float c0 = -rstd[r] * part[2 * r], c1 = -rstd[r] * part[2 * r + 1];
      const float2 c02 = make_float2(c0, c0), c12 = make_float2(c1, c1);
      const uint32_t* wg = reinterpret_cast<const uint32_t*>(&hgo[r]);
      const uint32_t* wy = reinterpret_cast<const uint32_t*>(&hy[r]);
      H8 hv;
      uint32_t* wv = reinterpret_cast<uint32_t*>(&hv);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 xh = __ffma2_rn(unpack_half2(wy[i]), rs2, nm2);
        const float2 dx = __fmul2_rn(unpack_half2(wg[i]), g2[i]);
        const float2 o = __ffma2_rn(c12, xh, __ffma2_rn(dx, rs2, c02));
        wv[i] = pack_half2(o.x, o.y);
      }
      *reinterpret_cast<H8*>(dy + (size_t)(r0 + r) * d + off) = hv;
      if (dz != nullptr || dbias != nullptr) {
        float z[8];
        h8_to_f(hv, z);  // the GEMM branch sees the fp16-rounded dy
        if (drop_thr16 && dz != nullptr) {
          const uint64_t e = (uint64_t)(r0 + r) * (uint64_t)d + (uint64_t)off;
          const uint64_t b0 = rng64(seed, e >> 2), b1 = rng64(seed, (e >> 2) + 1);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            z[i] = dropout_keep(b0, i, drop_thr16) ? z[i] * drop_scale : 0.f;
            z[4 + i] = dropout_keep(b1, i, drop_thr16) ? z[4 + i] * drop_scale : 0.f;
          }
          const H8 hz = f_to_h8(z);
          *reinterpret_cast<H8*>(dz + (size_t)(r0 + r) * d + off) = hz;
          h8_to_f(hz, z);
        }
        if (dbias != nullptr) {
#pragma unroll
          for (int i = 0; i < 4; ++i) az2[i] = __fadd2_rn(az2[i], make_float2(z[2 * i], z[2 * i + 1]));
        }
      }
    }
  }
  float ag[8], ab[8], az[8];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    ag[2 * i] = ag2[i].x; ag[2 * i + 1] = ag2[i].y;
    ab[2 * i] = ab2[i].x; ab[2 * i + 1] = ab2[i].y;
    az[2 * i] = az2[i].x; az[2 * i + 1] = az2[i].y;
  }
  red_add_f32x4(dgamma + off, ag[0], ag[1], ag[2], ag[3]);
  red_add_f32x4(dgamma + off + 4, ag[4], ag[5], ag[6], ag[7]);
  red_add_f32x4(dbeta + off, ab[0], ab[1], ab[2], ab[3]);
  red_add_f32x4(dbeta + off + 4, ab[4], ab[5], ab[6], ab[7]);
  if (dbias != nullptr) {
    red_add_f32x4(dbias + off, az[0], az[1], az[2], az[3]);
    red_add_f32x4(dbias + off + 4, az[4], az[5], az[6], az[7]);
  }
}

// dSr[z][i][c] = dS[z][i][c - (L-1-i)] for c >= L-1-i, 0 below: the adjoint of _rel_shift (transformer_xl.py:98-110)
// as a pure re-layout; reads of dS and writes of dSr are 16-byte aligned and fully coalesced.
// One WARP per row, rows dealt round-robin over a persistent grid: no shared memory and no block barriers. Output chunk
// k (8 fp16 = 16 bytes, aligned) of row i is the 16-byte window of the source row that starts e = (i + 1) mod 8 elements
// into source chunk k - 1 (k for e == 0): a lane loads one aligned source chunk, takes its right neighbour's from the
// next lane by shuffle, and funnel-shifts the pair. Elements right of the diagonal are exact zeros in dS (masked
// probabilities), chunks left of column L-1-i are never written (the workspace is zero-filled once).
DEVI uint32_t sel4(uint32_t a, uint32_t b, uint32_t c, uint32_t d, int s) {
  return s == 0 ? a : (s == 1 ? b : (s == 2 ? c : d));
}

__global__ void __launch_bounds__(256)
rel_unshift_kernel(const __half* __restrict__ ds, __half* __restrict__ dsr, int L, long long nrows) {
  pdl_launch_dependents();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const long long nwarps = (long long)gridDim.x * 8;
  const int nchunk = L >> 3;
  for (long long r = (long long)blockIdx.x * 8 + (threadIdx.x >> 5); r < nrows; r += nwarps) {
    const int i = (int)(r % L);
    const size_t zrow = (size_t)r * (size_t)L;
    const int c_lo = L - 1 - i;
    const int e = (8 - (c_lo & 7)) & 7;
    const int cb = c_lo >> 3;        // first output chunk
    const int nout = nchunk - cb;    // output chunks cb .. L/8 - 1
    const int nsrc = (i + 8) >> 3;   // source chunks that can hold non-zero data
    const uint4* src = reinterpret_cast<const uint4*>(ds + zrow);
    uint4* dst = reinterpret_cast<uint4*>(dsr + zrow) + cb;
    for (int k0 = 0; k0 < nout; k0 += 32) {
      const int k = k0 + lane;
      const int q = e ? k - 1 : k;
      uint4 a = make_uint4(0u, 0u, 0u, 0u);
      if (q >= 0 && q < nsrc) a = src[q];
      uint4 b;
      b.x = __shfl_down_sync(0xffffffffu, a.x, 1);
      b.y = __shfl_down_sync(0xffffffffu, a.y, 1);
      b.z = __shfl_down_sync(0xffffffffu, a.z, 1);
      b.w = __shfl_down_sync(0xffffffffu, a.w, 1);
      if (lane == 31) b = (q + 1 >= 0 && q + 1 < nsrc) ? src[q + 1] : make_uint4(0u, 0u, 0u, 0u);
      if (k < nout) {
        // halves [e + 2m, e + 2m + 1], m = 0..3, of the 16 halves in (a, b); e is warp-uniform
        const int sh = e >> 1;
        const uint32_t l0 = sel4(a.x, a.y, a.z, a.w, sh), l1 = sel4(a.y, a.z, a.w, b.x, sh), l2 = sel4(a.z, a.w, b.x, b.y, sh),
                       l3 = sel4(a.w, b.x, b.y, b.z, sh), l4 = sel4(b.x, b.y, b.z, b.w, sh);
        dst[k] = (e & 1) ? make_uint4(__funnelshift_r(l0, l1, 16), __funnelshift_r(l1, l2, 16), __funnelshift_r(l2, l3, 16),
                                      __funnelshift_r(l3, l4, 16))
                         : make_uint4(l0, l1, l2, l3);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ masked CE
// one CTA per row: online (max, sum-exp) over V fp16 logits, fp32 math. loss_row = (lse - z[label]) * mask.
constexpr int CE_THREADS = 256;

__global__ void __launch_bounds__(CE_THREADS)
ce_fwd_kernel(const __half* __restrict__ logits, long long ld, const long long* __restrict__ labels,
              const float* __restrict__ mask, float* __restrict__ row_loss, float* __restrict__ row_lse, int rows,
              int V) {
  __shared__ float2 scratch[32];
  __shared__ float fscr[32];
  const int row = blockIdx.x;
  const __half* z = logits + (size_t)row * ld;
  const int nch = (V + 7) / 8;
  float mx = -INFINITY, sm = 0.f;
  for (int ch = threadIdx.x; ch < nch; ch += CE_THREADS) {
    float f[8];
    h8_to_f(*reinterpret_cast<const H8*>(z + ch * 8), f);
    float cm = -INFINITY;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (ch * 8 + i >= V) f[i] = -INFINITY;
      cm = fmaxf(cm, f[i]);
    }
    if (cm > mx) {
      sm *= __expf(mx - cm);
      mx = cm;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) sm += __expf(f[i] - mx);
  }
  const float bm = block_max(mx, fscr);
  sm = (mx == -INFINITY) ? 0.f : sm * __expf(mx - bm);
  const float tot = block_sum2(sm, 0.f, scratch).x;
  if (threadIdx.x == 0) {
    const float lse = bm + logf(tot);
    // A label outside [0, V) is never dereferenced: the row contributes zero loss (and zero gradient in ce_bwd), which is
    // what nn.CrossEntropyLoss does for its ignore_index (-100); sum(mask) in the denominator is unchanged, as in
    // transformer_xl.py:606-609. (Other out-of-range labels make the reference raise a device assert.)
    const long long lab = labels[row];
    const bool ok = lab >= 0 && lab < (long long)V;
    const float zl = ok ? __half2float(z[lab]) : 0.f;
    row_lse[row] = lse;
    row_loss[row] = ok ? (lse - zl) * mask[row] : 0.f;
  }
}

// loss = sum(row_loss) / sum(mask); also keeps sum(mask) for the backward. Single CTA (rows is a few thousand).
__global__ void ce_finalize_kernel(const float* __restrict__ row_loss, const float* __restrict__ mask, int rows,
                                   float* __restrict__ out2) {
  __shared__ float2 scratch[32];
  float a = 0.f, b = 0.f;
  for (int i = threadIdx.x; i < rows; i += blockDim.x) {
    a += row_loss[i];
    b += mask[i];
  }
  const float2 t = block_sum2(a, b, scratch);
  if (threadIdx.x == 0) {
    out2[0] = t.x / t.y;
    out2[1] = t.y;
  }
}

// dlogits[row, c] = (softmax - onehot) * mask[row] / sum(mask) * gscale   (gscale = upstream dLoss, a device scalar)
__global__ void __launch_bounds__(CE_THREADS)
ce_bwd_kernel(const __half* __restrict__ logits, long long ld, const long long* __restrict__ labels,
              const float* __restrict__ mask, const float* __restrict__ row_lse, const float* __restrict__ loss2,
              const float* __restrict__ gscale, __half* __restrict__ dlogits, long long ldd, int rows, int V) {
  const int row = blockIdx.x;
  const __half* z = logits + (size_t)row * ld;
  __half* dzp = dlogits + (size_t)row * ldd;
  const long long lab_ll = labels[row];
  const bool lab_ok = lab_ll >= 0 && lab_ll < (long long)V;
  const float w = lab_ok ? mask[row] / loss2[1] * gscale[0] : 0.f;  // ignored row: no gradient (see ce_fwd_kernel)
  const float lse = row_lse[row];
  const int lab = lab_ok ? (int)lab_ll : -1;
  const int nch = (V + 7) / 8;
  for (int ch = threadIdx.x; ch < nch; ch += CE_THREADS) {
    float f[8], o[8];
    h8_to_f(*reinterpret_cast<const H8*>(z + ch * 8), f);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int c = ch * 8 + i;
      float pr = (w == 0.f) ? 0.f : __expf(f[i] - lse);
      if (c == lab) pr -= 1.f;
      o[i] = (c < V) ? pr * w : 0.f;
    }
    *reinterpret_cast<H8*>(dzp + ch * 8) = f_to_h8(o);
  }
}

// ------------------------------------------------------------------------------------------------ embedding assembly
// slot[b,l] = number of image-patch slots (token == -1) strictly before l in row b  (transformer_xl.py:639-642)
__global__ void embed_slots_kernel(const long long* __restrict__ tok, int* __restrict__ slot, int B, int L) {
  const int b = blockIdx.x;
  const int lane = threadIdx.x;  // 32 threads
  int base = 0;
  for (int l0 = 0; l0 < L; l0 += 32) {
    const int l = l0 + lane;
    const bool is = (l < L) && (tok[(size_t)b * L + l] == -1);
    const unsigned m = __ballot_sync(0xffffffffu, is);
    if (l < L) slot[(size_t)b * L + l] = base + __popc(m & ((1u << lane) - 1u));
    base += __popc(m);
  }
}

// out[b, l, :] = dropout( (tok >= 0 ? W[tok] : vis[b, slot]) + (pos ? T[pos] : 0) )
// one warp per token; out row = out + b*out_bs + l*d
__global__ void __launch_bounds__(256)
embed_fwd_kernel(const long long* __restrict__ tok, const long long* __restrict__ pos, const int* __restrict__ slot,
                 const __half* __restrict__ W, const __half* __restrict__ T, const __half* __restrict__ vis,
                 long long vis_bs, int nvis, __half* __restrict__ out, long long out_bs, int B, int L, int d, int V,
                 uint32_t drop_thr16, float drop_scale, uint64_t seed, long long seed_row0) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= B * L) return;
  const int b = warp / L, l = warp % L;
  const long long t = tok[warp];
  const __half* src = nullptr;
  if (t >= 0 && t < V) src = W + (size_t)t * d;
  else if (t == -1 && vis != nullptr) {
    const int s = slot[warp];
    if (s < nvis) src = vis + (size_t)b * vis_bs + (size_t)s * d;
  }
  const __half* tp = nullptr;
  if (pos != nullptr) tp = T + (size_t)pos[warp] * d;
  __half* dst = out + (size_t)b * out_bs + (size_t)l * d;
  for (int ch = lane; ch < d / 8; ch += 32) {
    float f[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) f[i] = 0.f;
    if (src) h8_to_f(*reinterpret_cast<const H8*>(src + ch * 8), f);
    if (tp) {
      float g[8];
      h8_to_f(*reinterpret_cast<const H8*>(tp + ch * 8), g);
      // the reference adds two fp16 tensors: round the sum once
#pragma unroll
      for (int i = 0; i < 8; ++i) f[i] += g[i];
    }
    if (drop_thr16) {
      const uint64_t e = (uint64_t)(seed_row0 + warp) * (uint64_t)d + (uint64_t)ch * 8;
      const uint64_t b0 = rng64(seed, e >> 2), b1 = rng64(seed, (e >> 2) + 1);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        f[i] = dropout_keep(b0, i, drop_thr16) ? f[i] * drop_scale : 0.f;
        f[4 + i] = dropout_keep(b1, i, drop_thr16) ? f[4 + i] * drop_scale : 0.f;
      }
    }
    *reinterpret_cast<H8*>(dst + ch * 8) = f_to_h8(f);
  }
}

// scatter-add of the embedding gradient: dW[tok] += g, dT[pos] += g, dvis[b, slot] = g   (g = dout * dropout mask)
__global__ void __launch_bounds__(256)
embed_bwd_kernel(const long long* __restrict__ tok, const long long* __restrict__ pos, const int* __restrict__ slot,
                 const __half* __restrict__ dout, long long dout_bs, __half* __restrict__ dW, __half* __restrict__ dT,
                 __half* __restrict__ dvis, long long vis_bs, int nvis, int B, int L, int d, int V,
                 uint32_t drop_thr16, float drop_scale, uint64_t seed, long long seed_row0) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= B * L) return;
  const int b = warp / L, l = warp % L;
  const long long t = tok[warp];
  const __half* src = dout + (size_t)b * dout_bs + (size_t)l * d;
  __half* wdst = (t >= 0 && t < V && dW != nullptr) ? dW + (size_t)t * d : nullptr;
  __half* tdst = (pos != nullptr && dT != nullptr) ? dT + (size_t)pos[warp] * d : nullptr;
  __half* vdst = nullptr;
  if (t == -1 && dvis != nullptr) {
    const int s = slot[warp];
    if (s < nvis) vdst = dvis + (size_t)b * vis_bs + (size_t)s * d;
  }
  for (int ch = lane; ch < d / 8; ch += 32) {
    H8 hv = *reinterpret_cast<const H8*>(src + ch * 8);
    if (drop_thr16) {
      float f[8];
      h8_to_f(hv, f);
      const uint64_t e = (uint64_t)(seed_row0 + warp) * (uint64_t)d + (uint64_t)ch * 8;
      const uint64_t b0 = rng64(seed, e >> 2), b1 = rng64(seed, (e >> 2) + 1);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        f[i] = dropout_keep(b0, i, drop_thr16) ? f[i] * drop_scale : 0.f;
        f[4 + i] = dropout_keep(b1, i, drop_thr16) ? f[4 + i] * drop_scale : 0.f;
      }
      hv = f_to_h8(f);
    }
    if (wdst) {
#pragma unroll
      for (int i = 0; i < 4; ++i)
        atomicAdd(reinterpret_cast<__half2*>(wdst + ch * 8) + i, reinterpret_cast<const __half2*>(&hv.u)[i]);
    }
    if (tdst) {
#pragma unroll
      for (int i = 0; i < 4; ++i)
        atomicAdd(reinterpret_cast<__half2*>(tdst + ch * 8) + i, reinterpret_cast<const __half2*>(&hv.u)[i]);
    }
    if (vdst) *reinterpret_cast<H8*>(vdst + ch * 8) = hv;
  }
}

// ------------------------------------------------------------------------------------------------ small reductions
// out[c] += sum_r in[r, c]   (fp16 in, fp32 accumulate). A CTA owns a 64-column strip and a block of rows: thread
// (cx = tid % 8, ry = tid / 8) streams rows ry, ry + 32, ... of its 16-byte column chunk (a warp reads 4 rows x 128
// contiguous bytes per step, 8 loads in flight per thread), the 32 row-lanes are folded through shared memory and
// each column issues ONE atomic per CTA.
constexpr int CS_ROWS_PER_CTA = 512;
__global__ void __launch_bounds__(256)
colsum_kernel(const __half* __restrict__ in, long long ld, float* __restrict__ out, int rows, int n) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float red[32][65];
  const int cx = threadIdx.x & 7, ry = threadIdx.x >> 3;
  const int col = blockIdx.x * 64 + cx * 8;
  const int r_begin = blockIdx.y * CS_ROWS_PER_CTA;
  const int r_end = min(rows, r_begin + CS_ROWS_PER_CTA);
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  if (col < n) {
    int r = r_begin + ry;
    for (; r + 7 * 32 < r_end; r += 8 * 32) {
      H8 v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = *reinterpret_cast<const H8*>(in + (size_t)(r + u * 32) * ld + col);
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        float f[8];
        h8_to_f(v[u], f);
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] += f[i];
      }
    }
    for (; r < r_end; r += 32) {
      float f[8];
      h8_to_f(*reinterpret_cast<const H8*>(in + (size_t)r * ld + col), f);
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] += f[i];
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) red[ry][cx * 8 + i] = acc[i];
  __syncthreads();
  if (threadIdx.x < 64) {
    const int c = blockIdx.x * 64 + threadIdx.x;
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 32; ++k) t += red[k][threadIdx.x];
    if (c < n) atomicAdd(out + c, t);
  }
}

// dq = dqu + dqv (fp16 out, written into the fused dQKV buffer) and du += colsum(dqu), dv += colsum(dqv).
// Same strip decomposition as colsum_kernel: a CTA owns 64 columns x CS_ROWS_PER_CTA rows, a warp touches 4 rows x 128
// contiguous bytes per access, 4 row-iterations in flight; one atomic per column per CTA.
__global__ void __launch_bounds__(256)
dq_finalize_kernel(const __half* __restrict__ dqu, const __half* __restrict__ dqv, long long ld_in,
                   __half* __restrict__ dq, long long ld_out, float* __restrict__ du, float* __restrict__ dv, int rows,
                   int n) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float red[2][32][65];
  const int cx = threadIdx.x & 7, ry = threadIdx.x >> 3;
  const int col = blockIdx.x * 64 + cx * 8;
  const int r_begin = blockIdx.y * CS_ROWS_PER_CTA;
  const int r_end = min(rows, r_begin + CS_ROWS_PER_CTA);
  float au[8], av[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) au[i] = av[i] = 0.f;
  if (col < n) {
    int r = r_begin + ry;
    for (; r + 3 * 32 < r_end; r += 4 * 32) {
      H8 a[4], b[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        a[u] = *reinterpret_cast<const H8*>(dqu + (size_t)(r + u * 32) * ld_in + col);
        b[u] = *reinterpret_cast<const H8*>(dqv + (size_t)(r + u * 32) * ld_in + col);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        float fa[8], fb[8], o[8];
        h8_to_f(a[u], fa);
        h8_to_f(b[u], fb);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          o[i] = fa[i] + fb[i];
          au[i] += fa[i];
          av[i] += fb[i];
        }
        *reinterpret_cast<H8*>(dq + (size_t)(r + u * 32) * ld_out + col) = f_to_h8(o);
      }
    }
    for (; r < r_end; r += 32) {
      float fa[8], fb[8], o[8];
      h8_to_f(*reinterpret_cast<const H8*>(dqu + (size_t)r * ld_in + col), fa);
      h8_to_f(*reinterpret_cast<const H8*>(dqv + (size_t)r * ld_in + col), fb);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        o[i] = fa[i] + fb[i];
        au[i] += fa[i];
        av[i] += fb[i];
      }
      *reinterpret_cast<H8*>(dq + (size_t)r * ld_out + col) = f_to_h8(o);
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    red[0][ry][cx * 8 + i] = au[i];
    red[1][ry][cx * 8 + i] = av[i];
  }
  __syncthreads();
  if (threadIdx.x < 128) {
    const int which = threadIdx.x >> 6, c = threadIdx.x & 63;
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 32; ++k) t += red[which][k][c];
    const int gc = blockIdx.x * 64 + c;
    if (gc < n) atomicAdd((which ? dv : du) + gc, t);
  }
}

// D[b,h,i] = sum_d dO[b,i,h,d] * O[b,i,h,d]   (one warp per (b,i,h))
__global__ void __launch_bounds__(256)
rowdot_kernel(const __half* __restrict__ a, const __half* __restrict__ b, long long ld, float* __restrict__ out, int B,
              int L, int H, int dh) {
  pdl_launch_dependents();
  pdl_wait();
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w >= B * L * H) return;
  const int h = w % H;
  const int bi = w / H;  // b*L + i
  const __half* pa = a + (size_t)bi * ld + h * dh;
  const __half* pb = b + (size_t)bi * ld + h * dh;
  float s = 0.f;
  for (int ch = lane; ch < dh / 8; ch += 32) {
    float x[8], y[8];
    h8_to_f(*reinterpret_cast<const H8*>(pa + ch * 8), x);
    h8_to_f(*reinterpret_cast<const H8*>(pb + ch * 8), y);
#pragma unroll
    for (int i = 0; i < 8; ++i) s += x[i] * y[i];
  }
  s = warp_sum(s);
  if (lane == 0) {
    const int bb = bi / L, i = bi % L;
    out[((size_t)bb * H + h) * L + i] = s;
  }
}

// Same, one warp per (b, i) row covering all heads: lanes read consecutive 16-byte chunks (fully coalesced), a head's
// dh / 8 chunks sit in CPH consecutive lanes and are folded with xor-shuffles. d % 256 == 0, CPH = dh / 8 in {1..32} pow2.
template <int CPH>
__global__ void __launch_bounds__(256)
rowdot_rows_kernel(const __half* __restrict__ a, const __half* __restrict__ b, long long ld, float* __restrict__ out,
                   int B, int L, int H, int d) {
  pdl_launch_dependents();
  pdl_wait();
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= B * L) return;
  const __half* pa = a + (size_t)row * ld;
  const __half* pb = b + (size_t)row * ld;
  const int bb = row / L, i = row % L;
  for (int c0 = 0; c0 < d / 8; c0 += 32) {
    const int c = c0 + lane;
    float x[8], y[8];
    h8_to_f(*reinterpret_cast<const H8*>(pa + c * 8), x);
    h8_to_f(*reinterpret_cast<const H8*>(pb + c * 8), y);
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) s = fmaf(x[k], y[k], s);
#pragma unroll
    for (int o = 1; o < CPH; o <<= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((lane & (CPH - 1)) == 0) out[((size_t)bb * H + c / CPH) * L + i] = s;
  }
}

// sinusoid rows in the reference order: row c <-> distance min(klen-1-c, clamp); [sin | cos]; fp32 math -> fp16, dropout
__global__ void posemb_kernel(__half* __restrict__ out, const float* __restrict__ inv_freq, int klen, int d,
                              int clamp_len, uint32_t drop_thr16, float drop_scale, uint64_t seed, int half_phase) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int half_d = d / 2;
  if (idx >= klen * half_d) return;
  const int c = idx / half_d, k = idx % half_d;
  float pos = (float)(klen - 1 - c);
  if (clamp_len > 0 && pos > (float)clamp_len) pos = (float)clamp_len;
  // inv_freq is the module's registered buffer (transformer_xl.py:40-41), so both sides use identical frequencies
  float ang = pos * inv_freq[k];
  if (half_phase) {
    // what the reference computes after module.half() (transformer_xl.py:569-571 with fp16 hidden states, :44 with the
    // fp16-cast inv_freq buffer): position, frequency and their product are each rounded to fp16 before sin / cos
    ang = __half2float(__float2half_rn(__half2float(__float2half_rn(pos)) * __half2float(__float2half_rn(inv_freq[k]))));
  }
  float sv = sinf(ang), cv = cosf(ang);
  if (drop_thr16) {
    const uint64_t e1 = (uint64_t)c * d + k, e2 = (uint64_t)c * d + half_d + k;
    const uint64_t b1 = rng64(seed, e1 >> 2), b2 = rng64(seed, e2 >> 2);
    sv = dropout_keep(b1, (int)(e1 & 3), drop_thr16) ? sv * drop_scale : 0.f;
    cv = dropout_keep(b2, (int)(e2 & 3), drop_thr16) ? cv * drop_scale : 0.f;
  }
  out[(size_t)c * d + k] = __float2half_rn(sv);
  out[(size_t)c * d + half_d + k] = __float2half_rn(cv);
}

// dst(fp16) = src(fp32)   and   dst(fp16) += src(fp32)
__global__ void f32_to_f16_kernel(const float* __restrict__ src, __half* __restrict__ dst, long long n, int accumulate) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float v = src[i];
  if (accumulate) v += __half2float(dst[i]);
  dst[i] = __float2half_rn(v);
}

// up to 8 independent (fp32 slice -> fp16 destination) conversions in one launch: blockIdx.y = segment
struct CvtSegs {
  __half* dst[8];
  long long off[8];
  long long n[8];
  int acc[8];
};
__global__ void f32_to_f16_multi_kernel(const float* __restrict__ src, const CvtSegs sg) {
  pdl_launch_dependents();
  pdl_wait();
  const int k = blockIdx.y;
  const long long n = sg.n[k];
  const float* sp = src + sg.off[k];
  __half* dp = sg.dst[k];
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float v = sp[i];
    if (sg.acc[k]) v += __half2float(dp[i]);
    dp[i] = __float2half_rn(v);
  }
}

static inline uint32_t thr16(float p) { return (uint32_t)(p * 65536.0f + 0.5f); }
static inline float dscale(uint32_t t) { return t ? 65536.0f / (65536.0f - (float)t) : 1.0f; }

}  // namespace db1

using namespace db1;

extern "C" int db1_layernorm_fwd(const void* y, const void* gamma, const void* beta, void* out, float* stats, int rows,
                                 int d, float eps, void* stream) {
  DB1_CHECK_ARG(y && gamma && beta && out && stats, "layernorm_fwd: null pointer");
  DB1_CHECK_ARG(rows > 0 && d > 0 && d % 8 == 0 && d <= LN_THREADS * 8 * LN_MAXC, "layernorm_fwd: bad shape %d x %d", rows, d);
  cudaStream_t st = (cudaStream_t)stream;
  const __half *yy = (const __half*)y, *gg = (const __half*)gamma, *bb = (const __half*)beta;
  const int g8 = (rows + 7) / 8;
  if (d == 2048) DB1_CUDA(launch_pdl(ln_fwd_warp_kernel<8>, dim3(g8), dim3(256), 0, st, 1, yy, gg, bb, (__half*)out, stats, rows, eps));
  else if (d == 1024) DB1_CUDA(launch_pdl(ln_fwd_warp_kernel<4>, dim3(g8), dim3(256), 0, st, 1, yy, gg, bb, (__half*)out, stats, rows, eps));
  else if (d == 512) DB1_CUDA(launch_pdl(ln_fwd_warp_kernel<2>, dim3(g8), dim3(256), 0, st, 1, yy, gg, bb, (__half*)out, stats, rows, eps));
  else if (d == 256) DB1_CUDA(launch_pdl(ln_fwd_warp_kernel<1>, dim3(g8), dim3(256), 0, st, 1, yy, gg, bb, (__half*)out, stats, rows, eps));
  else {
    const int grid = rows < 148 * 8 ? rows : 148 * 8;
    ln_fwd_kernel<<<grid, LN_THREADS, 0, st>>>(yy, gg, bb, (__half*)out, stats, rows, d, eps);
  }
  DB1_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int db1_layernorm_bwd(const void* dout, const void* y, const void* gamma, const float* stats, void* dy,
                                 void* dz, float* dgamma, float* dbeta, float* dbias, int rows, int d, float drop_p,
                                 uint64_t seed, void* stream) {
  DB1_CHECK_ARG(dout && y && gamma && stats && dy && dgamma && dbeta, "layernorm_bwd: null pointer");
  DB1_CHECK_ARG(rows > 0 && d > 0 && d % 8 == 0 && d <= LN_THREADS * 8 * LN_MAXC, "layernorm_bwd: bad shape %d x %d", rows, d);
  DB1_CHECK_ARG(drop_p >= 0.f && drop_p < 1.f, "layernorm_bwd: bad dropout p");
  DB1_CHECK_ARG(drop_p == 0.f || dz != nullptr, "layernorm_bwd: dropout needs a dz buffer");
  const uint32_t t = thr16(drop_p);
  cudaStream_t st = (cudaStream_t)stream;
  const __half *go = (const __half*)dout, *yy = (const __half*)y, *gg = (const __half*)gamma;
  if (d == 4096 || d == 2048 || d == 1024 || d == 512 || d == 256) {
    DB1_CHECK_ARG(((uintptr_t)dgamma & 15) == 0 && ((uintptr_t)dbeta & 15) == 0 && ((uintptr_t)dbias & 15) == 0,
                  "layernorm_bwd: dgamma / dbeta / dbias must be 16-byte aligned");
    __half *o1 = (__half*)dy, *o2 = (__half*)dz;
    // two rows per group (more rows per group spill once the next group's loads are kept in flight: measured 26 us vs
    // 21 us at 4096 x 2048); whole waves of equally loaded CTAs: ceil(ngroups / k) CTAs with k groups each
    constexpr int ROWS = 2;
    const int ngroups = (rows + ROWS - 1) / ROWS;
    const int slots = sm_count() * 2;
    const int k = (ngroups + slots - 1) / slots;
    const int grid = (ngroups + k - 1) / k;
#define DB1_LNB(T, B) DB1_CUDA(launch_pdl(ln_bwd_fused_kernel<T, ROWS, B>, dim3(grid), dim3(T), 0, st, 1, go, yy, gg, stats, o1, o2, dgamma, dbeta, dbias, rows, t, dscale(t), seed))
    if (d == 4096) DB1_LNB(512, 1);
    else if (d == 2048) DB1_LNB(256, 2);
    else if (d == 1024) DB1_LNB(128, 2);
    else if (d == 512) DB1_LNB(64, 2);
    else DB1_LNB(32, 2);
#undef DB1_LNB
  } else {
    const int grid = rows < 148 * 2 ? rows : 148 * 2;
    ln_bwd_kernel<<<grid, LN_THREADS, 0, st>>>(go, yy, gg, stats, (__half*)dy, (__half*)dz, dgamma, dbeta, dbias, rows,
                                               d, t, dscale(t), seed);
  }
  DB1_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int db1_ce_fwd(const void* logits, long long ld, const long long* labels, const float* mask, float* row_loss,
                          float* row_lse, float* loss2, int rows, int V, void* stream) {
  DB1_CHECK_ARG(logits && labels && mask && row_loss && row_lse && loss2, "ce_fwd: null pointer");
  DB1_CHECK_ARG(rows > 0 && V > 0 && ld % 8 == 0 && ld >= V, "ce_fwd: bad shape rows=%d V=%d ld=%lld", rows, V, ld);
  ce_fwd_kernel<<<rows, CE_THREADS, 0, (cudaStream_t)stream>>>((const __half*)logits, ld, labels, mask, row_loss,
                                                             row_lse, rows, V);
  ce_finalize_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(row_loss, mask, rows, loss2);
  DB1_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int db1_ce_bwd(const void* logits, long long ld, const long long* labels, const float* mask,
                          const float* row_lse, const float* loss2, const float* gscale, void* dlogits, long long ldd,
                          int rows, int V, void* stream) {
  DB1_CHECK_ARG(logits && labels && mask && row_lse && loss2 && gscale && dlogits, "ce_bwd: null pointer");
  DB1_CHECK_ARG(rows > 0 && V > 0 && ld % 8 == 0 && ldd % 8 == 0 && ld >= V && ldd >= ((V + 7) / 8) * 8,
                "ce_bwd: bad shape rows=%d V=%d", rows, V);
  ce_bwd_kernel<<<rows, CE_THREADS, 0, (cudaStream_t)stream>>>((const __half*)logits, ld, labels, mask, row_lse, loss2,
                                                             gscale, (__half*)dlogits, ldd, rows, V);
  DB1_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int db1_embed_fwd(const long long* tok, const long long* pos, int* slot, const void* W, const void* T,
                             const void* vis, long long vis_bs, int nvis, void* out, long long out_bs, int B, int L,
                             int d, int V, float drop_p, uint64_t seed, long long seed_row0, void* stream) {
  DB1_CHECK_ARG(tok && W && out && slot, "embed_fwd: null pointer");
  DB1_CHECK_ARG(B > 0 && L > 0 && d % 8 == 0, "embed_fwd: bad shape");
  DB1_CHECK_ARG(pos == nullptr || T != nullptr, "embed_fwd: position ids need the timestep table");
  embed_slots_kernel<<<B, 32, 0, (cudaStream_t)stream>>>(tok, slot, B, L);
  const uint32_t t = thr16(drop_p);
  const long long nthreads = (long long)B * L * 32;
  embed_fwd_kernel<<<(unsigned)((nthreads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      tok, pos, slot, (const __half*)W, (const __half*)T, (const __half*)vis, vis_bs, nvis, (__half*)out, out_bs, B, L,
      d, V, t, dscale(t), seed, seed_row0);
  DB1_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int db1_embed_bwd(const long long* tok, const long long* pos, const int* slot, const void* dout,
                             long long dout_bs, void* dW, void* dT, void* dvis, long long vis_bs, int nvis, int B, int L,
                             int d, int V, float drop_p, uint64_t seed, long long seed_row0, void* stream) {
  DB1_CHECK_ARG(tok && dout && slot, "embed_bwd: null pointer");
  DB1_CHECK_ARG(B > 0 && L > 0 && d % 8 == 0, "embed_bwd: bad shape");
  const uint32_t t = thr16(drop_p);
  const long long nthreads = (long long)B * L * 32;
  embed_bwd_kernel<<<(unsigned)((nthreads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      tok, pos, slot, (const __half*)dout, dout_bs, (__half*)dW, (__half*)dT, (__half*)dvis, vis_bs, nvis, B, L, d, V, t,
      dscale(t), seed, seed_row0);
  DB1_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int db1_colsum(const void* in, long long ld, float* out, int rows, int n, void* stream) {
  DB1_CHECK_ARG(in && out && rows > 0 && n > 0 && n % 8 == 0 && ld % 8 == 0, "colsum: bad arguments");
  dim3 grid((n + 63) / 64, (rows + CS_ROWS_PER_CTA - 1) / CS_ROWS_PER_CTA);
  DB1_CUDA(launch_pdl(colsum_kernel, grid, dim3(256), 0, (cudaStream_t)stream, 1, (const __half*)in, ld, out, rows, n));
  DB1_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int db1_dq_finalize(const void* dqu, const void* dqv, long long ld_in, void* dq, long long ld_out, float* du,
                               float* dv, int rows, int n, void* stream) {
  DB1_CHECK_ARG(dqu && dqv && dq && du && dv && rows > 0 && n % 8 == 0 && ld_in % 8 == 0 && ld_out % 8 == 0,
                "dq_finalize: bad arguments");
  dim3 grid((n + 63) / 64, (rows + CS_ROWS_PER_CTA - 1) / CS_ROWS_PER_CTA);
  DB1_CUDA(launch_pdl(dq_finalize_kernel, grid, dim3(256), 0, (cudaStream_t)stream, 1, (const __half*)dqu,
                      (const __half*)dqv, ld_in, (__half*)dq, ld_out, du, dv, rows, n));
  DB1_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int db1_rowdot(const void* a, const void* b, long long ld, float* out, int B, int L, int H, int dh,
                          void* stream) {
  DB1_CHECK_ARG(a && b && out && B > 0 && L > 0 && H > 0 && dh % 8 == 0 && ld % 8 == 0, "rowdot: bad arguments");
  const int cph = dh / 8, dd = H * dh;
  if (dd % 256 == 0 && cph <= 32 && (cph & (cph - 1)) == 0) {
    const dim3 grid((unsigned)(((long long)B * L * 32 + 255) / 256));
    const __half *pa = (const __half*)a, *pb = (const __half*)b;
    cudaStream_t st = (cudaStream_t)stream;
#define DB1_RD(C) DB1_CUDA(launch_pdl(rowdot_rows_kernel<C>, grid, dim3(256), 0, st, 1, pa, pb, ld, out, B, L, H, dd))
    switch (cph) {
      case 1: DB1_RD(1); break;
      case 2: DB1_RD(2); break;
      case 4: DB1_RD(4); break;
      case 8: DB1_RD(8); break;
      case 16: DB1_RD(16); break;
      default: DB1_RD(32); break;
    }
#undef DB1_RD
    return 0;
  }
  const long long nthreads = (long long)B * L * H * 32;
  DB1_CUDA(launch_pdl(rowdot_kernel, dim3((unsigned)((nthreads + 255) / 256)), dim3(256), 0, (cudaStream_t)stream, 1,
                      (const __half*)a, (const __half*)b, ld, out, B, L, H, dh));
  DB1_CUDA(cudaGetLastError());
  return 0;
}

static int posemb_launch(void* out, const float* inv_freq, int klen, int d, int clamp_len, float drop_p, uint64_t seed,
                         int half_phase, void* stream) {
  DB1_CHECK_ARG(out && inv_freq && klen > 0 && d > 0 && d % 2 == 0, "posemb: bad arguments");
  const uint32_t t = thr16(drop_p);
  const int n = klen * (d / 2);
  posemb_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>((__half*)out, inv_freq, klen, d, clamp_len, t, dscale(t),
                                                                 seed, half_phase);
  DB1_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int db1_posemb(void* out, const float* inv_freq, int klen, int d, int clamp_len, float drop_p, uint64_t seed,
                          void* stream) {
  return posemb_launch(out, inv_freq, klen, d, clamp_len, drop_p, seed, 0, stream);
}

extern "C" int db1_posemb_half_phase(void* out, const float* inv_freq, int klen, int d, int clamp_len, float drop_p,
                                     uint64_t seed, void* stream) {
  return posemb_launch(out, inv_freq, klen, d, clamp_len, drop_p, seed, 1, stream);
}

extern "C" int db1_rel_unshift(const void* ds, void* dsr, int Z, int L, void* stream) {
  DB1_CHECK_ARG(ds && dsr && Z > 0 && L > 0 && L % 8 == 0 && L <= 16384, "rel_unshift: bad arguments");
  const long long nrows = (long long)Z * L;
  long long blocks = (nrows + 7) / 8;
  const long long cap = (long long)sm_count() * 8;  // persistent: 8 CTAs of 8 warps per SM, rows round-robin
  if (blocks > cap) blocks = cap;
  DB1_CUDA(launch_pdl(rel_unshift_kernel, dim3((unsigned)blocks), dim3(256), 0, (cudaStream_t)stream, 1, (const __half*)ds,
                      (__half*)dsr, L, nrows));
  DB1_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int db1_f32_to_f16(const float* src, void* dst, long long n, int accumulate, void* stream) {
  DB1_CHECK_ARG(src && dst && n > 0, "f32_to_f16: bad arguments");
  f32_to_f16_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(src, (__half*)dst, n, accumulate);
  DB1_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int db1_f32_to_f16_multi(const float* src, int nseg, void* const* dst, const long long* off, const long long* n,
                                    const int* accumulate, void* stream) {
  DB1_CHECK_ARG(src && dst && off && n && accumulate && nseg > 0 && nseg <= 8, "f32_to_f16_multi: bad arguments");
  CvtSegs sg;
  long long mx = 0;
  for (int k = 0; k < nseg; ++k) {
    DB1_CHECK_ARG(dst[k] != nullptr && n[k] > 0 && off[k] >= 0, "f32_to_f16_multi: bad segment %d", k);
    sg.dst[k] = (__half*)dst[k];
    sg.off[k] = off[k];
    sg.n[k] = n[k];
    sg.acc[k] = accumulate[k];
    if (n[k] > mx) mx = n[k];
  }
  long long bx = (mx + 255) / 256;
  if (bx > 1024) bx = 1024;
  DB1_CUDA(launch_pdl(f32_to_f16_multi_kernel, dim3((unsigned)bx, nseg), dim3(256), 0, (cudaStream_t)stream, 1, src, sg));
  DB1_CUDA(cudaGetLastError());
  return 0;
}
