// Micro-benchmark (development tool, not product): how fast can 148 CTAs reduce fp32 tiles into a small global buffer?
// Pattern = the dR flush of the attention-backward dQ kernel: a CTA adds a [128 rows x 128 floats] chunk (64 KB) to
// rows [c0, c0+128) x columns [h*128, h*128+128) of an [L=1024, d=2048] fp32 matrix, thread = row.
//   mode 0: red.global.add.v4.f32 straight from registers (32 per thread and chunk)
//   mode 1: cp.reduce.async.bulk (shared -> global, add.f32), one 512-byte row per thread from a 64 KB staging tile
//   mode 2: red.global.add.f32 scalar (baseline)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/ubench_red tools/ubench_red.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int MODE>
__global__ void __launch_bounds__(128, 1) red_kernel(float* dst, int L, int d, int iters, long long* cycles) {
  extern __shared__ __align__(128) float stage[];  // [128][128] fp32 (mode 1)
  const int r = threadIdx.x;
  float v[4] = {1.f, 2.f, 3.f, 4.f};
  if (MODE == 1) {
    for (int k = 0; k < 128; ++k) stage[r * 128 + k] = 1.0f;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
  }
  uint32_t s = blockIdx.x * 2654435761u + 12345u;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    s = s * 1664525u + 1013904223u;
    const int c0 = ((s >> 8) % (L / 128)) * 128;
    const int h = (s >> 20) % (d / 128);
    float* row = dst + (size_t)(c0 + r) * d + h * 128;
    if (MODE == 0) {
#pragma unroll
      for (int k = 0; k < 32; ++k)
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(row + 4 * k), "f"(v[0]), "f"(v[1]), "f"(v[2]),
                     "f"(v[3])
                     : "memory");
    } else if (MODE == 2) {
#pragma unroll 8
      for (int k = 0; k < 128; ++k) asm volatile("red.global.add.f32 [%0], %1;" ::"l"(row + k), "f"(v[k & 3]) : "memory");
    } else {
      asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], 512;" ::"l"(row),
                   "r"(smem_u32(stage + r * 128))
                   : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
  }
  if (MODE == 1) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  __threadfence();
  const long long t1 = clock64();
  if (r == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int MODE>
static void run(const char* name, int L, int d, int iters) {
  float* dst;
  long long* cyc;
  cudaMalloc(&dst, (size_t)L * d * 4);
  cudaMemset(dst, 0, (size_t)L * d * 4);
  cudaMalloc(&cyc, 148 * 8);
  cudaFuncSetAttribute(red_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  red_kernel<MODE><<<148, 128, 65536>>>(dst, L, d, 4, cyc);
  cudaEventRecord(e0);
  red_kernel<MODE><<<148, 128, 65536>>>(dst, L, d, iters, cyc);
  cudaEventRecord(e1);
  cudaError_t e = cudaDeviceSynchronize();
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  long long h[148];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  long long mx = 0;
  for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
  const double bytes = 148.0 * iters * 65536.0;
  printf("%-28s L=%d d=%d: %.3f ms, %.1f GB/s reduced, %.1f B/cycle/SM (max %lld cycles / %d chunks = %.0f cycles per 64 KB chunk) [%s]\n",
         name, L, d, ms, bytes / ms / 1e6, 65536.0 * iters / (double)mx, mx, iters, (double)mx / iters,
         cudaGetErrorString(e));
  cudaFree(dst);
  cudaFree(cyc);
}

int main() {
  run<0>("red.global.add.v4.f32", 1024, 2048, 200);
  run<0>("red.global.add.v4.f32", 4096, 2048, 200);
  run<1>("cp.reduce.async.bulk 512B", 1024, 2048, 200);
  run<1>("cp.reduce.async.bulk 512B", 4096, 2048, 200);
  run<2>("red.global.add.f32", 1024, 2048, 50);
  return 0;
}
