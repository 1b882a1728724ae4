// Fused optimizer step over the flat gradient buckets: what DeepSpeed's fp16 optimizer wrapper +
// FusedAdam do after backward in the reference's loop (src/train_utils/train.py:231-232, flags train_config.py:212-235):
// unscale by the loss scale, global-norm clip, overflow check, Adam(W) on fp32 master weights, fp16 parameters
// refreshed. Two HBM-bound passes:
//   db1_grad_sumsq : sum of squares of an fp16 bucket (fp32 accumulation; inf/nan propagate -> overflow detection)
//   db1_adam_step  : reads g16, master, m, v (14 B / parameter), writes master, m, v, p16 (14 B / parameter)
#include "../../include/db1_sm100.h"
#include "common.cuh"
#include "ptx.cuh"

namespace db1 {

__global__ void __launch_bounds__(256)
grad_sumsq_kernel(const __half* __restrict__ g, long long n8, float* __restrict__ out) {
  __shared__ float red[8];
  float acc = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
    float f[8];
    half8_to_float(ld_half8(g + i * 8), f);
#pragma unroll
    for (int k = 0; k < 8; ++k) acc = fmaf(f[k], f[k], acc);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += red[w];
    atomicAdd(out, t);
  }
}

struct AdamParams {
  float lr, beta1, beta2, eps, weight_decay;
  float bc1, bc2;       // bias corrections 1 - beta^t
  int adamw;            // 1: decoupled weight decay (AdamW), 0: L2 added to the gradient
};

// gcoef[0] (device scalar) = 1 / loss_scale * clip coefficient, computed on the device from the global norm
__global__ void __launch_bounds__(256)
adam_step_kernel(const __half* __restrict__ g, __half* __restrict__ p16, float* __restrict__ master,
                 float* __restrict__ m, float* __restrict__ v, long long n8, const float* __restrict__ gcoef,
                 const AdamParams a) {
  const float coef = gcoef[0];
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
    float gf[8];
    half8_to_float(ld_half8(g + i * 8), gf);
    float4 w0 = reinterpret_cast<const float4*>(master)[2 * i], w1 = reinterpret_cast<const float4*>(master)[2 * i + 1];
    float4 m0 = reinterpret_cast<const float4*>(m)[2 * i], m1 = reinterpret_cast<const float4*>(m)[2 * i + 1];
    float4 v0 = reinterpret_cast<const float4*>(v)[2 * i], v1 = reinterpret_cast<const float4*>(v)[2 * i + 1];
    float w[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
    float mm[8] = {m0.x, m0.y, m0.z, m0.w, m1.x, m1.y, m1.z, m1.w};
    float vv[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      float gr = gf[k] * coef;
      if (!a.adamw) gr = fmaf(a.weight_decay, w[k], gr);
      mm[k] = fmaf(a.beta1, mm[k], (1.0f - a.beta1) * gr);
      vv[k] = fmaf(a.beta2, vv[k], (1.0f - a.beta2) * gr * gr);
      float upd = (mm[k] / a.bc1) / (sqrtf(vv[k] / a.bc2) + a.eps);
      if (a.adamw) upd = fmaf(a.weight_decay, w[k], upd);
      w[k] = fmaf(-a.lr, upd, w[k]);
    }
    reinterpret_cast<float4*>(master)[2 * i] = make_float4(w[0], w[1], w[2], w[3]);
    reinterpret_cast<float4*>(master)[2 * i + 1] = make_float4(w[4], w[5], w[6], w[7]);
    reinterpret_cast<float4*>(m)[2 * i] = make_float4(mm[0], mm[1], mm[2], mm[3]);
    reinterpret_cast<float4*>(m)[2 * i + 1] = make_float4(mm[4], mm[5], mm[6], mm[7]);
    reinterpret_cast<float4*>(v)[2 * i] = make_float4(vv[0], vv[1], vv[2], vv[3]);
    reinterpret_cast<float4*>(v)[2 * i + 1] = make_float4(vv[4], vv[5], vv[6], vv[7]);
    st_half8(p16 + i * 8, float_to_half8(w));
  }
}

// gcoef = inv_scale * min(1, clip / (sqrt(sumsq) * inv_scale + 1e-6)); flag[0] = 1 if sumsq is not finite (overflow)
__global__ void clip_coef_kernel(const float* __restrict__ sumsq, float inv_scale, float clip, float* __restrict__ gcoef,
                                 int* __restrict__ flag) {
  const float s = sumsq[0];
  const bool bad = !(s == s) || s > 3.0e38f;
  flag[0] = bad ? 1 : 0;
  const float norm = sqrtf(s) * inv_scale;
  float c = inv_scale;
  if (clip > 0.f && norm > clip) c = inv_scale * clip / (norm + 1e-6f);
  gcoef[0] = bad ? 0.f : c;
  gcoef[1] = norm;
}

}  // namespace db1

using namespace db1;

extern "C" int db1_grad_sumsq(const void* g16, long long n, float* out, void* stream) {
  DB1_CHECK_ARG(g16 && out && n > 0 && n % 8 == 0, "grad_sumsq: bad arguments (n must be a multiple of 8)");
  const long long n8 = n / 8;
  long long blocks = (n8 + 255) / 256;
  const long long cap = (long long)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  grad_sumsq_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((const __half*)g16, n8, out);
  DB1_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int db1_clip_coef(const float* sumsq, float inv_scale, float clip, float* gcoef2, int* overflow_flag,
                             void* stream) {
  DB1_CHECK_ARG(sumsq && gcoef2 && overflow_flag, "clip_coef: null pointer");
  clip_coef_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(sumsq, inv_scale, clip, gcoef2, overflow_flag);
  DB1_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int db1_adam_step(const void* g16, void* p16, float* master, float* m, float* v, long long n,
                             const float* gcoef, float lr, float beta1, float beta2, float eps, float weight_decay,
                             int step, int adamw, void* stream) {
  DB1_CHECK_ARG(g16 && p16 && master && m && v && gcoef && n > 0 && n % 8 == 0 && step >= 1,
                "adam_step: bad arguments (n must be a multiple of 8, step >= 1)");
  AdamParams a;
  a.lr = lr; a.beta1 = beta1; a.beta2 = beta2; a.eps = eps; a.weight_decay = weight_decay;
  a.bc1 = 1.0f - powf(beta1, (float)step);
  a.bc2 = 1.0f - powf(beta2, (float)step);
  a.adamw = adamw ? 1 : 0;
  const long long n8 = n / 8;
  long long blocks = (n8 + 255) / 256;
  const long long cap = (long long)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  adam_step_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((const __half*)g16, (__half*)p16, master, m, v,
                                                                     n8, gcoef, a);
  DB1_CUDA(cudaGetLastError());
  return 0;
}
