// ResNet-v2-block image patch embedder (src/tokenizer/vision_embedding.py:36-86), the pieces around the tensor-core
// GEMMs: per-patch standardisation + conv3x3(3->64), GroupNorm(32 groups) + exact GELU + im2col for the two
// conv3x3(64->64) layers (which then run as db1_gemm_f16 [P*256, 576] x [64, 576]^T), their adjoints, and the small
// re-layout kernels. Patches are the batch dimension (zero padding is per patch), activations are stored
// pixel-major / channels-last: [P, 256 pixels, 64 channels] fp16, so a conv output row is one 128-byte line.
//
// Thread mapping of the GroupNorm kernels: tid -> (c8 = tid & 7: channels 8*c8..+7 = groups 4*c8..+3,
// pl = tid >> 3: pixels pl + 32*k, k = 0..7). A warp then touches 4 pixel rows x 128 contiguous bytes per access and a
// group's 512 elements live in the 32 threads that share c8 (4 lanes of each of the 8 warps).
#include "../../include/db1_sm100.h"
#include "common.cuh"
#include "ptx.cuh"

namespace db1 {

constexpr int PX = 256;   // pixels per 16x16 patch
constexpr int CH = 64;    // channels of the ResNet block
constexpr int PADW = 18;  // padded patch edge

DEVI float warp_sum_f(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

DEVI float gelu_exact(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }
DEVI float gelu_exact_grad(float x) {
  return 0.5f * (1.0f + erff(x * 0.70710678118654752f)) + x * 0.3989422804014327f * __expf(-0.5f * x * x);
}

// ------------------------------------------------------------------------------------------------ batched transpose
// out[b][c][r] = in[b][r][c]  (weight re-layouts: conv [co][ci][9] <-> [co][9][ci], projection [d][64][256] <-> [d][256][64])
__global__ void __launch_bounds__(256)
transpose_kernel(const __half* __restrict__ in, __half* __restrict__ out, int rows, int cols) {
  __shared__ __half tile[32][33];
  const size_t boff = (size_t)blockIdx.z * rows * cols;
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int k = ty; k < 32; k += 8) {
    const int r = r0 + k, c = c0 + tx;
    if (r < rows && c < cols) tile[k][tx] = in[boff + (size_t)r * cols + c];
  }
  __syncthreads();
  for (int k = ty; k < 32; k += 8) {
    const int c = c0 + k, r = r0 + tx;
    if (r < rows && c < cols) out[boff + (size_t)c * rows + r] = tile[tx][k];
  }
}

// ------------------------------------------------------------------------------------------------ standardise + conv1
// One CTA (256 threads = pixels) per patch. xs [P, C, 256] keeps the standardised pixels (fp16) for the weight
// gradient; y1 [P, 256, 64] = conv3x3(xs) + bias.
DEVI float px_to_float(__half v) { return __half2float(v); }
DEVI float px_to_float(float v) { return v; }

// PT = pixel storage type: fp16, or fp32 (the reference standardises in the INPUT dtype, :73-77, and only then casts to
// the module dtype, :79 - with fp32 frames the mean / std must see the unrounded pixels)
template <int C, typename PT>
__global__ void __launch_bounds__(256)
patch_conv1_fwd_kernel(const PT* __restrict__ pixels, const __half* __restrict__ W1, const __half* __restrict__ b1,
                       __half* __restrict__ xs, __half* __restrict__ y1, int Himg, int Wimg, int h0, int w0) {
  __shared__ float xpad[C][PADW * PADW];
  __shared__ float wsm[C * 9][CH];  // [ci*9 + tap][co]
  __shared__ float red[8][2 * C];
  const int p = blockIdx.x;
  const int n = p / (h0 * w0), pr = (p / w0) % h0, pc = p % w0;
  const int t = threadIdx.x, y = t >> 4, x = t & 15;
  const int lane = t & 31, warp = t >> 5;
  for (int i = t; i < C * PADW * PADW; i += 256) (&xpad[0][0])[i] = 0.f;
  for (int i = t; i < CH * C * 9; i += 256) {
    const int co = i / (C * 9), rem = i % (C * 9);  // W1 [co][ci][ky][kx]
    wsm[rem][co] = __half2float(W1[i]);
  }
  float v[C];
#pragma unroll
  for (int c = 0; c < C; ++c)
    v[c] = px_to_float(pixels[(((size_t)n * C + c) * Himg + (pr * 16 + y)) * Wimg + pc * 16 + x]);
  // mean, then centred unbiased variance (torch.std default), per channel over the 256 pixels
  float part[2 * C];
#pragma unroll
  for (int c = 0; c < C; ++c) part[c] = warp_sum_f(v[c]);
  if (lane == 0)
#pragma unroll
    for (int c = 0; c < C; ++c) red[warp][c] = part[c];
  __syncthreads();
  float mean[C];
#pragma unroll
  for (int c = 0; c < C; ++c) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += red[w][c];
    mean[c] = s * (1.0f / PX);
  }
#pragma unroll
  for (int c = 0; c < C; ++c) part[c] = warp_sum_f((v[c] - mean[c]) * (v[c] - mean[c]));
  if (lane == 0)
#pragma unroll
    for (int c = 0; c < C; ++c) red[warp][C + c] = part[c];
  __syncthreads();
#pragma unroll
  for (int c = 0; c < C; ++c) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += red[w][C + c];
    const float sd = sqrtf(s * (1.0f / (PX - 1)));
    // (x - mean) / (1e-6 + std) / sqrt(16), rounded to fp16 as the reference's `.to(data_type)` does
    const __half hv = __float2half_rn((v[c] - mean[c]) / (1e-6f + sd) * 0.25f);
    xs[((size_t)p * C + c) * PX + t] = hv;
    xpad[c][(y + 1) * PADW + x + 1] = __half2float(hv);
  }
  __syncthreads();
  float acc[CH];
#pragma unroll
  for (int co = 0; co < CH; ++co) acc[co] = 0.f;
#pragma unroll
  for (int c = 0; c < C; ++c)
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
      const float xv = xpad[c][(y + tap / 3) * PADW + x + tap % 3];
      const float* wr = wsm[c * 9 + tap];
#pragma unroll
      for (int co = 0; co < CH; ++co) acc[co] = fmaf(xv, wr[co], acc[co]);
    }
  __half* dst = y1 + ((size_t)p * PX + t) * CH;
#pragma unroll
  for (int g = 0; g < 8; ++g) {
    float f[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) f[i] = acc[g * 8 + i] + __half2float(b1[g * 8 + i]);
    st_half8(dst + g * 8, float_to_half8(f));
  }
}

// dW1[co][ci][tap] += sum_px dy1[px][co] * xs_pad[ci][px + tap offset]  (fp32 atomics, one set per CTA; CTAs loop over patches)
template <int C>
__global__ void __launch_bounds__(256)
patch_conv1_bwd_kernel(const __half* __restrict__ xs, const __half* __restrict__ dy1, float* __restrict__ dW1, int P) {
  __shared__ float xpad[C][PADW * PADW];
  __shared__ __half dsm[PX][CH + 2];
  const int t = threadIdx.x;
  const int co = t & 63, tg = t >> 6;  // this thread: output channel co, taps tg, tg+4, ... of the C*9
  constexpr int NT = (C * 9 + 3) / 4;
  float acc[NT];
#pragma unroll
  for (int k = 0; k < NT; ++k) acc[k] = 0.f;
  for (int i = t; i < C * PADW * PADW; i += 256) (&xpad[0][0])[i] = 0.f;
  for (int p = blockIdx.x; p < P; p += gridDim.x) {
    __syncthreads();
    for (int c = 0; c < C; ++c) xpad[c][((t >> 4) + 1) * PADW + (t & 15) + 1] = __half2float(xs[((size_t)p * C + c) * PX + t]);
    for (int i = t; i < PX * CH / 8; i += 256) {
      const int px = i >> 3, c8 = i & 7;
      const Half8 h = ld_half8(dy1 + ((size_t)p * PX + px) * CH + c8 * 8);
      const __half* hs = reinterpret_cast<const __half*>(&h);
#pragma unroll
      for (int j = 0; j < 8; ++j) dsm[px][c8 * 8 + j] = hs[j];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < NT; ++k) {
      const int ct = tg + 4 * k;  // ci*9 + tap
      if (ct < C * 9) {
        const int ci = ct / 9, tap = ct % 9;
        float a = 0.f;
        for (int px = 0; px < PX; ++px)
          a = fmaf(__half2float(dsm[px][co]), xpad[ci][((px >> 4) + tap / 3) * PADW + (px & 15) + tap % 3], a);
        acc[k] += a;
      }
    }
  }
#pragma unroll
  for (int k = 0; k < NT; ++k) {
    const int ct = tg + 4 * k;
    if (ct < C * 9) atomicAdd(dW1 + (size_t)co * C * 9 + ct, acc[k]);
  }
}

// ------------------------------------------------------------------------------------------------ GN + GELU + im2col
// Reduce 8 per-thread values (4 groups x {sum, sumsq} or 8 channels) over the 32 threads sharing c8: lanes c8 + 8*j of
// every warp -> xor-shuffles 8, 16, then across the 8 warps through shared memory. Result broadcast to those threads.
DEVI void reduce_c8(float (&v)[8], float (*scr)[8][8], int warp, int lane) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    v[i] += __shfl_xor_sync(0xffffffffu, v[i], 8);
    v[i] += __shfl_xor_sync(0xffffffffu, v[i], 16);
  }
  __syncthreads();
  if (lane < 8)
#pragma unroll
    for (int i = 0; i < 8; ++i) scr[warp][lane][i] = v[i];
  __syncthreads();
  const int c8 = lane & 7;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += scr[w][c8][i];
    v[i] = s;
  }
}

// x [P,256,64] -> a = gelu(GN(x)) (kept only in shared memory) -> col [P*256, 576], col[q][tap*64 + ci] = a[q + tap offset][ci]
// stats [P, 32, 2] = (mean, rstd) per group, kept for the backward.
__global__ void __launch_bounds__(256)
gn_gelu_im2col_kernel(const __half* __restrict__ x, const __half* __restrict__ gamma, const __half* __restrict__ beta,
                      __half* __restrict__ col, float* __restrict__ stats, float eps) {
  extern __shared__ __align__(16) uint8_t psm[];
  __half* apad = reinterpret_cast<__half*>(psm);                         // [324][64] fp16, zero border
  float(*scr)[8][8] = reinterpret_cast<float(*)[8][8]>(psm + PADW * PADW * CH * 2);
  const int p = blockIdx.x, t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int c8 = t & 7, pl = t >> 3;
  for (int i = t; i < PADW * PADW * CH / 8; i += 256) reinterpret_cast<uint4*>(apad)[i] = make_uint4(0, 0, 0, 0);
  float xv[8][8];
  float gs[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) gs[i] = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    half8_to_float(ld_half8(x + ((size_t)p * PX + pl + 32 * k) * CH + c8 * 8), xv[k]);
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      gs[2 * g] += xv[k][2 * g] + xv[k][2 * g + 1];
      gs[2 * g + 1] += xv[k][2 * g] * xv[k][2 * g] + xv[k][2 * g + 1] * xv[k][2 * g + 1];
    }
  }
  reduce_c8(gs, scr, warp, lane);
  float mean[4], rstd[4];
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    mean[g] = gs[2 * g] * (1.0f / 512.0f);
    const float var = fmaxf(gs[2 * g + 1] * (1.0f / 512.0f) - mean[g] * mean[g], 0.f);
    rstd[g] = rsqrtf(var + eps);
    if (pl == 0) {
      stats[((size_t)p * 32 + c8 * 4 + g) * 2] = mean[g];
      stats[((size_t)p * 32 + c8 * 4 + g) * 2 + 1] = rstd[g];
    }
  }
  float ga[8], be[8];
  half8_to_float(ld_half8(gamma + c8 * 8), ga);
  half8_to_float(ld_half8(beta + c8 * 8), be);
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int px = pl + 32 * k;
    float a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      // GroupNorm output and GELU output are both rounded to fp16 in the reference's half-precision module
      const float z = __half2float(__float2half_rn((xv[k][i] - mean[i >> 1]) * rstd[i >> 1] * ga[i] + be[i]));
      a[i] = gelu_exact(z);
    }
    st_half8(apad + (((px >> 4) + 1) * PADW + (px & 15) + 1) * CH + c8 * 8, float_to_half8(a));
  }
  __syncthreads();
  // im2col: 256 rows x 72 16-byte pieces, consecutive threads -> consecutive pieces (contiguous global writes)
  __half* crow = col + (size_t)p * PX * 576;
  for (int item = t; item < PX * 72; item += 256) {
    const int q = item / 72, piece = item % 72;
    const int tap = piece >> 3, cc = piece & 7;
    const int sy = (q >> 4) + tap / 3, sx = (q & 15) + tap % 3;
    st_half8(crow + (size_t)q * 576 + piece * 8, ld_half8(apad + (sy * PADW + sx) * CH + cc * 8));
  }
}

// Adjoint: da[q'][ci] = sum_tap dcol[q' - tap offset][tap*64 + ci] (col2im), then GELU' and GroupNorm backward.
// dx [P,256,64] (+= dres when given: the residual branch's gradient), dgamma/dbeta fp32 atomics.
__global__ void __launch_bounds__(256)
col2im_gn_gelu_bwd_kernel(const __half* __restrict__ dcol, const __half* __restrict__ x, const float* __restrict__ stats,
                          const __half* __restrict__ gamma, const __half* __restrict__ beta,
                          const __half* __restrict__ dres, __half* __restrict__ dx, float* __restrict__ dgamma,
                          float* __restrict__ dbeta) {
  extern __shared__ __align__(16) uint8_t psm[];
  __half* tapbuf = reinterpret_cast<__half*>(psm);  // [256][64] one tap slice of dcol
  float(*scr)[8][8] = reinterpret_cast<float(*)[8][8]>(psm + PX * CH * 2);
  const int p = blockIdx.x, t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int c8 = t & 7, pl = t >> 3;
  float da[8][8];
#pragma unroll
  for (int k = 0; k < 8; ++k)
#pragma unroll
    for (int i = 0; i < 8; ++i) da[k][i] = 0.f;
  const __half* dc = dcol + (size_t)p * PX * 576;
  for (int tap = 0; tap < 9; ++tap) {
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int q = pl + 32 * k;
      st_half8(tapbuf + q * CH + c8 * 8, ld_half8(dc + (size_t)q * 576 + tap * 64 + c8 * 8));
    }
    __syncthreads();
    const int oy = tap / 3 - 1, ox = tap % 3 - 1;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int px = pl + 32 * k;
      const int qy = (px >> 4) - oy, qx = (px & 15) - ox;  // the output pixel that read this input pixel through `tap`
      if (qy >= 0 && qy < 16 && qx >= 0 && qx < 16) {
        float f[8];
        half8_to_float(ld_half8(tapbuf + (qy * 16 + qx) * CH + c8 * 8), f);
#pragma unroll
        for (int i = 0; i < 8; ++i) da[k][i] += f[i];
      }
    }
  }
  float mean[4], rstd[4], ga[8], be[8];
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    mean[g] = stats[((size_t)p * 32 + c8 * 4 + g) * 2];
    rstd[g] = stats[((size_t)p * 32 + c8 * 4 + g) * 2 + 1];
  }
  half8_to_float(ld_half8(gamma + c8 * 8), ga);
  half8_to_float(ld_half8(beta + c8 * 8), be);
  float xh[8][8];
  float cs[8], cb[8], gsum[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) cs[i] = cb[i] = gsum[i] = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    float xv[8];
    half8_to_float(ld_half8(x + ((size_t)p * PX + pl + 32 * k) * CH + c8 * 8), xv);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      xh[k][i] = (xv[i] - mean[i >> 1]) * rstd[i >> 1];
      const float z = __half2float(__float2half_rn(xh[k][i] * ga[i] + be[i]));
      const float dz = da[k][i] * gelu_exact_grad(z);
      cs[i] += dz * xh[k][i];  // dgamma
      cb[i] += dz;             // dbeta
      da[k][i] = dz * ga[i];   // dxhat
      gsum[(i >> 1) * 2] += da[k][i];
      gsum[(i >> 1) * 2 + 1] += da[k][i] * xh[k][i];
    }
  }
  reduce_c8(gsum, scr, warp, lane);
  reduce_c8(cs, scr, warp, lane);
  reduce_c8(cb, scr, warp, lane);
  if (pl == 0) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      atomicAdd(dgamma + c8 * 8 + i, cs[i]);
      atomicAdd(dbeta + c8 * 8 + i, cb[i]);
    }
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const size_t off = ((size_t)p * PX + pl + 32 * k) * CH + c8 * 8;
    float o[8], r[8];
    if (dres != nullptr) half8_to_float(ld_half8(dres + off), r);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int g = i >> 1;
      o[i] = rstd[g] * (da[k][i] - gsum[2 * g] * (1.0f / 512.0f) - xh[k][i] * gsum[2 * g + 1] * (1.0f / 512.0f));
      if (dres != nullptr) o[i] += r[i];
    }
    st_half8(dx + off, float_to_half8(o));
  }
}

// out(fp16) = x * keep(seed, flat index) / (1 - p)   (embedding dropout on image-patch rows, transformer_xl.py:545)
__global__ void dropout_kernel(const __half* __restrict__ x, __half* __restrict__ out, long long n8, uint32_t thr16,
                               float scale, uint64_t seed) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n8) return;
  float f[8];
  half8_to_float(ld_half8(x + i * 8), f);
  const uint64_t b0 = rng64(seed, (uint64_t)i * 2), b1 = rng64(seed, (uint64_t)i * 2 + 1);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    f[k] = dropout_keep(b0, k, thr16) ? f[k] * scale : 0.f;
    f[4 + k] = dropout_keep(b1, k, thr16) ? f[4 + k] * scale : 0.f;
  }
  st_half8(out + i * 8, float_to_half8(f));
}

}  // namespace db1

using namespace db1;

extern "C" int db1_transpose_f16(const void* in, void* out, int batch, int rows, int cols, void* stream) {
  DB1_CHECK_ARG(in && out && batch > 0 && rows > 0 && cols > 0 && batch <= 65535, "transpose: bad arguments");
  dim3 grid((cols + 31) / 32, (rows + 31) / 32, batch);
  transpose_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const __half*)in, (__half*)out, rows, cols);
  DB1_CUDA(cudaGetLastError());
  return 0;
}

template <typename PT>
static int patch_conv1_fwd_launch(const void* pixels, const void* W1, const void* b1, void* xs, void* y1, int N, int C,
                                  int Himg, int Wimg, void* stream) {
  DB1_CHECK_ARG(pixels && W1 && b1 && xs && y1, "patch_conv1_fwd: null pointer");
  DB1_CHECK_ARG(N > 0 && Himg > 0 && Wimg > 0 && Himg % 16 == 0 && Wimg % 16 == 0,
                "patch_conv1_fwd: image %dx%d must be a multiple of the 16x16 patch (as the reference's rearrange requires)",
                Himg, Wimg);
  DB1_CHECK_ARG(C == 3 || C == 1, "patch_conv1_fwd: %d input channels unsupported (1 or 3)", C);
  const int h0 = Himg / 16, w0 = Wimg / 16;
  const int P = N * h0 * w0;
  if (C == 3)
    patch_conv1_fwd_kernel<3, PT><<<P, 256, 0, (cudaStream_t)stream>>>((const PT*)pixels, (const __half*)W1,
                                                                      (const __half*)b1, (__half*)xs, (__half*)y1, Himg,
                                                                      Wimg, h0, w0);
  else
    patch_conv1_fwd_kernel<1, PT><<<P, 256, 0, (cudaStream_t)stream>>>((const PT*)pixels, (const __half*)W1,
                                                                      (const __half*)b1, (__half*)xs, (__half*)y1, Himg,
                                                                      Wimg, h0, w0);
  DB1_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int db1_patch_conv1_fwd(const void* pixels, const void* W1, const void* b1, void* xs, void* y1, int N, int C,
                                   int Himg, int Wimg, void* stream) {
  return patch_conv1_fwd_launch<__half>(pixels, W1, b1, xs, y1, N, C, Himg, Wimg, stream);
}

extern "C" int db1_patch_conv1_fwd_f32(const float* pixels, const void* W1, const void* b1, void* xs, void* y1, int N,
                                       int C, int Himg, int Wimg, void* stream) {
  return patch_conv1_fwd_launch<float>(pixels, W1, b1, xs, y1, N, C, Himg, Wimg, stream);
}

extern "C" int db1_patch_conv1_bwd(const void* xs, const void* dy1, float* dW1, int P, int C, void* stream) {
  DB1_CHECK_ARG(xs && dy1 && dW1 && P > 0 && (C == 3 || C == 1), "patch_conv1_bwd: bad arguments");
  const int grid = P < 2 * sm_count() ? P : 2 * sm_count();
  if (C == 3)
    patch_conv1_bwd_kernel<3><<<grid, 256, 0, (cudaStream_t)stream>>>((const __half*)xs, (const __half*)dy1, dW1, P);
  else
    patch_conv1_bwd_kernel<1><<<grid, 256, 0, (cudaStream_t)stream>>>((const __half*)xs, (const __half*)dy1, dW1, P);
  DB1_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int db1_gn_gelu_im2col(const void* x, const void* gamma, const void* beta, void* col, float* stats, int P,
                                  float eps, void* stream) {
  DB1_CHECK_ARG(x && gamma && beta && col && stats && P > 0, "gn_gelu_im2col: bad arguments");
  constexpr int SMEM = PADW * PADW * CH * 2 + 8 * 8 * 8 * 4;
  static bool configured = false;
  if (!configured) {
    DB1_CUDA(cudaFuncSetAttribute(gn_gelu_im2col_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    configured = true;
  }
  gn_gelu_im2col_kernel<<<P, 256, SMEM, (cudaStream_t)stream>>>((const __half*)x, (const __half*)gamma,
                                                               (const __half*)beta, (__half*)col, stats, eps);
  DB1_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int db1_col2im_gn_gelu_bwd(const void* dcol, const void* x, const float* stats, const void* gamma,
                                      const void* beta, const void* dres, void* dx, float* dgamma, float* dbeta, int P,
                                      void* stream) {
  DB1_CHECK_ARG(dcol && x && stats && gamma && beta && dx && dgamma && dbeta && P > 0, "col2im_gn_gelu_bwd: bad arguments");
  constexpr int SMEM = PX * CH * 2 + 8 * 8 * 8 * 4;
  col2im_gn_gelu_bwd_kernel<<<P, 256, SMEM, (cudaStream_t)stream>>>((const __half*)dcol, (const __half*)x, stats,
                                                                   (const __half*)gamma, (const __half*)beta,
                                                                   (const __half*)dres, (__half*)dx, dgamma, dbeta);
  DB1_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int db1_dropout_f16(const void* x, void* out, long long n, float drop_p, uint64_t seed, void* stream) {
  DB1_CHECK_ARG(x && out && n > 0 && n % 8 == 0 && drop_p >= 0.f && drop_p < 1.f, "dropout: bad arguments");
  const uint32_t t = (uint32_t)(drop_p * 65536.0f + 0.5f);
  const float sc = t ? 65536.0f / (65536.0f - (float)t) : 1.0f;
  const long long n8 = n / 8;
  dropout_kernel<<<(unsigned)((n8 + 255) / 256), 256, 0, (cudaStream_t)stream>>>((const __half*)x, (__half*)out, n8, t,
                                                                               sc, seed);
  DB1_CUDA(cudaGetLastError());
  return 0;
}
