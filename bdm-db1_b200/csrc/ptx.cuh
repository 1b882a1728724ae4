// Thin inline-PTX wrappers for sm_100a: mbarrier, TMA, tcgen05 (UMMA / TMEM).
// Everything here is hand-written against the PTX ISA; bit layouts of the UMMA
// shared-memory and instruction descriptors are documented next to the builders.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace db1 {

#define DEVI __device__ __forceinline__

DEVI uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
DEVI uint32_t lane_id() { return threadIdx.x & 31; }

DEVI bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.b32 %0, 1, 0, P;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// ---------------------------------------------------------------- mbarrier
DEVI void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
DEVI void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
DEVI void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
DEVI void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
DEVI bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, P;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
DEVI void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}

// Long waits (an epilogue warp waiting out a whole main loop): poll with a back-off instead of spinning - eight warps
// spinning on try_wait take issue slots and power from the MMA / TMA threads of a power-capped chip.
DEVI void mbar_wait_backoff(uint64_t* bar, uint32_t parity, uint32_t ns) {
  while (!mbar_try_wait(bar, parity)) __nanosleep(ns);
}

// ---------------------------------------------------------------- fences
DEVI void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
DEVI void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
DEVI void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// ---------------------------------------------------------------- TMA
DEVI void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(m) : "memory");
}
DEVI void tma_load_2d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(m), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
DEVI void tma_load_4d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(m), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
// multicast variant: the box lands at the same smem offset (and signals the mbarrier at the same offset) in every CTA
// of the cluster whose bit is set in cta_mask
DEVI void tma_load_4d_mc(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2, int c3,
                         uint16_t cta_mask) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
      " [%0], [%1, {%3, %4, %5, %6}], [%2], %7;"
      ::"r"(smem_u32(smem_dst)), "l"(m), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "h"(cta_mask)
      : "memory");
}
// L2 prefetch of a box (no shared-memory destination, no barrier): hides the DRAM leg of a later tma_load of the same box
DEVI void tma_prefetch_4d(const CUtensorMap* m, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global [%0, {%1, %2, %3, %4}];"
               ::"l"(m), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
DEVI void tma_load_3d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(m), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

// shared -> global tile store (bulk async-group completion); out-of-bounds parts of the box are clipped
DEVI void tma_store_2d(const CUtensorMap* m, const void* smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(m), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
               : "memory");
}
DEVI void tma_store_3d(const CUtensorMap* m, const void* smem_src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(m), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
DEVI void bulk_commit_group() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// wait until at most N of this thread's bulk groups still have to READ their shared-memory source
template <int N>
DEVI void bulk_wait_group_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}

// ---------------------------------------------------------------- TMEM
template <int NCOLS>
DEVI void tmem_alloc(uint32_t* smem_slot) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)),
               "r"(NCOLS)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int NCOLS>
DEVI void tmem_dealloc(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(NCOLS) : "memory");
}

// D[tmem] (+)= A[smem desc] * B[smem desc]; kind::f16 covers fp16 and bf16 inputs, fp32 accumulate.
DEVI void umma_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// A operand from TMEM (K-major only), B from smem.
DEVI void umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrive on an mbarrier when all previously issued tcgen05.mma of this thread have completed.
DEVI void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// same, arriving on the barrier at this offset in every CTA of the cluster selected by cta_mask
DEVI void umma_commit_mc(uint64_t* bar, uint16_t cta_mask) {
  asm volatile(
      "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
      ::"r"(smem_u32(bar)), "h"(cta_mask)
      : "memory");
}
// ---------------------------------------------------------------- CTA-pair (cta_group::2) variants
// The two CTAs of a cluster (ranks 0/1, one TPC) run ONE M=256 MMA: each supplies 128 rows of A and N/2 rows of B from
// the same shared-memory offsets and owns the accumulator rows of its own 128-lane TMEM. Per flop every SM then reads
// and writes half as many B bytes of shared memory as with cta_group::1 (which is shared-memory-bandwidth bound:
// 12 KB read + 12 KB TMA-written per 128-cycle 128x256x16 MMA against a 128 B/clk port).
constexpr uint32_t PEER_BIT_MASK = 0xFEFFFFFFu;  // clears the bit that selects the odd CTA of the pair in a shared::cluster address

template <int NCOLS>
DEVI void tmem_alloc2(uint32_t* smem_slot) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)),
               "r"(NCOLS)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <int NCOLS>
DEVI void tmem_dealloc2(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(NCOLS) : "memory");
}
DEVI void umma_ss2(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive (once all prior cta_group::2 MMAs of this thread completed) on the barrier at this offset in both CTAs
DEVI void umma_commit2_mc(uint64_t* bar, uint16_t cta_mask) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
      ::"r"(smem_u32(bar)), "h"(cta_mask)
      : "memory");
}
// TMA load into this CTA's shared memory whose transaction bytes are counted on the LEADER (even) CTA's mbarrier
DEVI void tma_load_4d_2sm(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(m), "r"(smem_u32(bar) & PEER_BIT_MASK), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
// same, delivered to the same shared-memory offset of every CTA in cta_mask; each destination's bytes are counted on
// the barrier at `bar`'s offset in the leader (even) CTA of THAT destination's pair
DEVI void tma_load_4d_2sm_mc(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2, int c3,
                             uint16_t cta_mask) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
      " [%0], [%1, {%3, %4, %5, %6}], [%2], %7;"
      ::"r"(smem_u32(smem_dst)), "l"(m), "r"(smem_u32(bar) & PEER_BIT_MASK), "r"(c0), "r"(c1), "r"(c2), "r"(c3),
        "h"(cta_mask)
      : "memory");
}
// arrive on the barrier at this offset in the leader CTA's shared memory (works from either CTA of the pair)
DEVI void mbar_arrive_leader(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(smem_u32(bar) & PEER_BIT_MASK) : "memory");
}

DEVI uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
DEVI void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

DEVI void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
DEVI void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// 32 lanes x 32-bit, 32 consecutive columns: thread t of the warp gets lane (quadrant*32 + t), columns [c, c+32).
DEVI void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
DEVI void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
DEVI void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%32], "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31};"
      ::"r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
        "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),
        "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]),
        "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31]),
        "r"(taddr)
      : "memory");
}
DEVI void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%16], "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15};"
      ::"r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
        "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),
        "r"(taddr)
      : "memory");
}

// ---------------------------------------------------------------- UMMA descriptors
// Shared-memory matrix descriptor (64 bit), SWIZZLE_128B flavour only:
//   [ 0,14) start address >> 4      [16,30) leading-dim byte offset >> 4
//   [32,46) stride byte offset >> 4 [46,48) version = 1 (Blackwell)
//   [49,52) base offset = 0         [61,64) layout type: 2 = SWIZZLE_128B
// K-major operand tile  [rows][64 fp16] (row = 128 B, 8-row swizzle atom = 1024 B):
//   SBO = 1024 (next 8-row group), LBO unused (1). Advancing K by 16 elements = +32 B on the start address.
// MN-major operand tile [k rows][64 fp16 along MN] per 64-wide MN block:
//   SBO = 1024 (next 8 k-rows), LBO = byte distance between consecutive 64-wide MN blocks.
//   Advancing K by 16 = +2048 B on the start address.
DEVI uint64_t umma_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// Instruction descriptor (32 bit) for kind::f16:
//   [4,6) D format (1 = f32)  [7,10) A format (0 = f16, 1 = bf16)  [10,13) B format
//   [15] A major (0 = K, 1 = MN)  [16] B major  [17,23) N >> 3  [24,29) M >> 4
__host__ __device__ constexpr uint32_t umma_idesc(int M, int N, int a_mn_major, int b_mn_major, int bf16) {
  return (1u << 4) | ((uint32_t)bf16 << 7) | ((uint32_t)bf16 << 10) | ((uint32_t)a_mn_major << 15) |
         ((uint32_t)b_mn_major << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ---------------------------------------------------------------- programmatic dependent launch
// Kernels are launched with cudaLaunchAttributeProgrammaticStreamSerialization (common.cuh: launch_pdl): the grid may
// become resident while its predecessor in the stream is still draining, so its prologue (barrier init, TMEM
// allocation, descriptor prefetch) and the launch latency overlap the predecessor's tail. pdl_wait() must precede the
// first access to global memory (it returns once every prerequisite grid has completed and its writes are visible);
// pdl_launch_dependents() lets the NEXT kernel in the stream start being scheduled.
DEVI void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
DEVI void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// ---------------------------------------------------------------- misc
DEVI uint32_t pack_half2(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
DEVI float2 unpack_half2(uint32_t v) {
  __half2 h = *reinterpret_cast<__half2*>(&v);
  return __half22float2(h);
}
// Eight fp16 values moved as ONE 128-bit access. (A struct of four __half2 is copied member-wise by nvcc - four 32-bit
// LDG/STG per "vector" access, i.e. 4x the memory instructions and quarter-sector writes - so the carrier is a uint4.)
struct alignas(16) Half8 {
  uint4 u;
};
DEVI Half8 ld_half8(const __half* p) {
  Half8 v;
  v.u = *reinterpret_cast<const uint4*>(p);
  return v;
}
DEVI void st_half8(__half* p, const Half8& v) { *reinterpret_cast<uint4*>(p) = v.u; }
DEVI Half8 lds_half8(uint32_t saddr) {
  Half8 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.u.x), "=r"(v.u.y), "=r"(v.u.z), "=r"(v.u.w) : "r"(saddr));
  return v;
}
DEVI void sts_half8(uint32_t saddr, const Half8& v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "r"(v.u.x), "r"(v.u.y), "r"(v.u.z), "r"(v.u.w) : "memory");
}
DEVI void half8_to_float(const Half8& v, float (&f)[8]) {
  float2 t;
  t = unpack_half2(v.u.x); f[0] = t.x; f[1] = t.y;
  t = unpack_half2(v.u.y); f[2] = t.x; f[3] = t.y;
  t = unpack_half2(v.u.z); f[4] = t.x; f[5] = t.y;
  t = unpack_half2(v.u.w); f[6] = t.x; f[7] = t.y;
}
DEVI Half8 float_to_half8(const float (&f)[8]) {
  Half8 v;
  v.u.x = pack_half2(f[0], f[1]);
  v.u.y = pack_half2(f[2], f[3]);
  v.u.z = pack_half2(f[4], f[5]);
  v.u.w = pack_half2(f[6], f[7]);
  return v;
}
// 32 consecutive fp16 of one row (16 packed words) straight between registers and global memory: two 256-bit accesses
// (sm_100: LDG / STG.256 = one full 32-byte sector per thread and instruction), or four 128-bit ones when the row is
// only 16-byte aligned; `n` = valid elements (multiple of 8, <= 32), the rest is skipped / zero.
DEVI void ldg_row32(const __half* p, uint32_t (&v)[16], int n, bool al32) {
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = 0u;
  if (al32) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      if (n >= 16 * (h + 1)) {
        asm volatile("ld.global.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(v[8 * h]), "=r"(v[8 * h + 1]), "=r"(v[8 * h + 2]), "=r"(v[8 * h + 3]), "=r"(v[8 * h + 4]),
                       "=r"(v[8 * h + 5]), "=r"(v[8 * h + 6]), "=r"(v[8 * h + 7])
                     : "l"(p + 16 * h));
      } else if (n > 16 * h) {
        const uint4 t = *reinterpret_cast<const uint4*>(p + 16 * h);
        v[8 * h] = t.x; v[8 * h + 1] = t.y; v[8 * h + 2] = t.z; v[8 * h + 3] = t.w;
      }
    }
  } else {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      if (n >= 8 * (q + 1)) {
        const uint4 t = *reinterpret_cast<const uint4*>(p + 8 * q);
        v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
      }
    }
  }
}
DEVI void stg_row32(__half* p, const uint32_t (&v)[16], int n, bool al32) {
  if (al32) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      if (n >= 16 * (h + 1)) {
        asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                     ::"l"(p + 16 * h), "r"(v[8 * h]), "r"(v[8 * h + 1]), "r"(v[8 * h + 2]), "r"(v[8 * h + 3]),
                       "r"(v[8 * h + 4]), "r"(v[8 * h + 5]), "r"(v[8 * h + 6]), "r"(v[8 * h + 7])
                     : "memory");
      } else if (n > 16 * h) {
        *reinterpret_cast<uint4*>(p + 16 * h) = make_uint4(v[8 * h], v[8 * h + 1], v[8 * h + 2], v[8 * h + 3]);
      }
    }
  } else {
#pragma unroll
    for (int q = 0; q < 4; ++q)
      if (n >= 8 * (q + 1))
        *reinterpret_cast<uint4*>(p + 8 * q) = make_uint4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
  }
}

DEVI Half8 half8_zero() {
  Half8 v;
  v.u = make_uint4(0u, 0u, 0u, 0u);
  return v;
}
// 2^x on the MUFU alone (no denormal rescue around it: results below 2^-126 flush to zero, which is what a softmax wants)
DEVI float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
DEVI void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// Exact-erf GELU (F.gelu default, activations.py:29) and its derivative from ONE exp and ONE reciprocal:
// erf(t) = 1 - (a1 s + ... + a5 s^5) exp(-t^2), s = 1 / (1 + p t), t >= 0   (Abramowitz & Stegun 7.1.26, |err| <= 1.5e-7,
// far below the fp16 resolution of the stored activations); exp(-t^2) with t = |x| / sqrt(2) is also the Gaussian
// factor of gelu'(x) = Phi(x) + x * phi(x). erff() + expf() cost ~4x more issue slots and made the GeGLU-backward
// epilogue, not the MMA, the pacing stage.
DEVI void gelu_erf_both(float x, float& gl, float& dgl) {
  const float t = fabsf(x) * 0.70710678118654752f;
  const float e = __expf(-t * t);
  const float s = __fdividef(1.0f, fmaf(0.3275911f, t, 1.0f));
  float poly = fmaf(s, 1.061405429f, -1.453152027f);
  poly = fmaf(poly, s, 1.421413741f);
  poly = fmaf(poly, s, -0.284496736f);
  poly = fmaf(poly, s, 0.254829592f);
  const float erf_abs = fmaf(-poly * s, e, 1.0f);
  const float cdf = 0.5f + 0.5f * copysignf(erf_abs, x);
  gl = x * cdf;
  dgl = fmaf(x * 0.3989422804014327f, e, cdf);
}
DEVI float gelu_erf(float x) {
  float a, b;
  gelu_erf_both(x, a, b);
  return a;
}

// Stateless counter-based RNG for dropout (splitmix64 of seed + group index). One 64-bit draw covers four
// consecutive elements (16 bits each), so the mask is a pure function of (seed, flat element index) and
// forward and backward regenerate it without storing it. keep iff draw16 >= thr16 (thr16 = p * 65536).
DEVI uint64_t rng64(uint64_t seed, uint64_t group) {
  uint64_t z = seed + (group + 1) * 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
DEVI bool dropout_keep(uint64_t bits, int lane4, uint32_t thr16) {
  return (uint32_t)((bits >> (16 * lane4)) & 0xFFFFull) >= thr16;
}

}  // namespace db1
