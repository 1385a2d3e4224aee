// Inline-PTX wrappers shared by the tcgen05 kernels (sm_100a): mbarrier, TMA (cp.async.bulk.tensor), UMMA
// shared-memory descriptors, tcgen05.mma / commit / ld / st, proxy fences.  Bit layouts follow
// cute/arch/mma_sm100_desc.hpp (read-only reference in the vendored CUTLASS tree).
#pragma once

#include <cuda.h>

#include "common.cuh"

namespace pd {

// ---- PTX wrappers ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t s_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(s_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
// try_wait with a suspend-time hint (ns): the waiting thread is parked by the hardware until the phase completes or the
// hint expires instead of returning to its polling loop every few dozen cycles — for role warps (TMA, MMA issue) that
// share a sub-partition with warps whose issue slots matter.
__device__ __forceinline__ bool mbar_try_wait_hint(uint64_t* bar, uint32_t parity, uint32_t ns) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(s_u32(bar)), "r"(parity), "r"(ns) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait_parked(uint64_t* bar, uint32_t parity, int tag) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait_hint(bar, parity, 4000u)) {
    if (clock64() - t0 > 4000000000LL) {
      printf("pd_b200 tcgen05 kernel: mbarrier timeout tag=%d block=%d thread=%d parity=%u\n", tag, blockIdx.x,
             threadIdx.x, parity);
      __trap();
    }
  }
}
// Bounded wait: a protocol bug traps (-> launch error) instead of hanging the GPU.
#ifndef PD_PARK
#define PD_PARK 0          // 1: every bounded wait parks its thread with a suspend-time hint instead of re-polling
#endif
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, int tag) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!(PD_PARK ? mbar_try_wait_hint(bar, parity, 4000u) : mbar_try_wait(bar, parity))) {
    if (clock64() - t0 > 4000000000LL) {
      printf("pd_b200 tcgen05 kernel: mbarrier timeout tag=%d block=%d thread=%d parity=%u\n", tag, blockIdx.x,
             threadIdx.x, parity);
      __trap();
    }
  }
}
// One lane of a CONVERGED warp (elect.sync).  Role loops run warp-uniformly and guard only the single-thread
// instructions (TMA, tcgen05.mma / commit) with this: ptxas then emits a plain predicated UTMALDG / UTCHMMA, whereas
// the same instructions inside an `if (lane == 0)` region are wrapped in an ELECT ... BRA.U.ANY retry loop each —
// several hundred cycles per k-block on the issuing thread, enough to starve the tensor pipe.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\t"
      "elect.sync rx|px, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, px;\n\t}"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1,
                                            int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(s_u32(dst)), "l"(map), "r"(s_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(s_u32(dst)), "l"(map), "r"(s_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(s_u32(dst)), "l"(map), "r"(s_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, const void* src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
               ::"l"(map), "r"(s_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void tma_store_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void epi_bar_sync(int id) { asm volatile("bar.sync %0, 128;" ::"r"(id) : "memory"); }
// both epilogue groups (8 warps)
__device__ __forceinline__ void epi_bar_sync_all(int id) { asm volatile("bar.sync %0, 256;" ::"r"(id) : "memory"); }
// TMA prefetch of a 4-D box into L2 (no shared-memory destination, no barrier)
__device__ __forceinline__ void tma_prefetch_4d(const CUtensorMap* map, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global.tile [%0, {%1, %2, %3, %4}];"
               ::"l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ unsigned long long gtimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor layout):
// start>>4 [0,14), LBO>>4 [16,30) (=1, unused for swizzled K-major), SBO>>4 [32,46) (= 1024 B:
// eight 128-byte rows per swizzle atom), version=1 [46,48), layout SWIZZLE_128B=2 [61,64).
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,"
      "%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }


// MN-major (the "transposed" B operand: N contiguous), SWIZZLE_128B: atoms of 64 N-elements x 8 K-rows.
// LBO = byte stride between 64-element N chunks, SBO = byte stride between groups of 8 K rows.
__device__ __forceinline__ uint64_t make_smem_desc_mn(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
        "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
               ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}


// D[tmem] (+)= A[tmem] * B[smem]: the A operand (M = 128 rows = TMEM lanes, K packed two bf16 per 32-bit column)
// is read from tensor memory — used for P.V with P written by the softmax threads straight into TMEM.
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// pointer forms (caller guarantees full unrolling so that r[] stays in registers)
__device__ __forceinline__ void tmem_ld32p(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,"
      "%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_st16p(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
        "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}
// 16-byte shared-memory accesses by 32-bit shared address: a bf16x8 store through a GENERIC pointer derived from the
// aligned dynamic-smem base compiles to four 4-byte generic ST.E (ptxas loses both the address space and the vector
// width), which made the epilogue of the short-K layers instruction- and LSU-bound
__device__ __forceinline__ void sts_bf16x8(uint32_t saddr, const bf16x8& v) {
  const uint32_t* w = reinterpret_cast<const uint32_t*>(&v);
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]) : "memory");
}
__device__ __forceinline__ bf16x8 lds_bf16x8(uint32_t saddr) {
  bf16x8 v;
  uint32_t* w = reinterpret_cast<uint32_t*>(&v);
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]) : "r"(saddr) : "memory");
  return v;
}
// packed fp32x2 arithmetic (sm_100: FFMA2 / FADD2 halve the issue slots of the softmax inner loop)
__device__ __forceinline__ void ffma2(float& d0, float& d1, float a0, float a1, float b0, float b1, float c0, float c1) {
  asm("{\n\t.reg .b64 ra, rb, rc, rd;\n\t"
      "mov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmov.b64 rc, {%6, %7};\n\t"
      "fma.rn.f32x2 rd, ra, rb, rc;\n\t"
      "mov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(d0), "=f"(d1) : "f"(a0), "f"(a1), "f"(b0), "f"(b1), "f"(c0), "f"(c1));
}
__device__ __forceinline__ void fadd2(float& d0, float& d1, float a0, float a1, float b0, float b1) {
  asm("{\n\t.reg .b64 ra, rb, rd;\n\t"
      "mov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\t"
      "add.rn.f32x2 rd, ra, rb;\n\t"
      "mov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(d0), "=f"(d1) : "f"(a0), "f"(a1), "f"(b0), "f"(b1));
}
__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}

// register-typed (b32) variants for values that live in tcgen05.ld result registers; the _v forms are volatile so
// that the order written in the source (software pipelining of MUFU vs. its consumers) survives the compiler
__device__ __forceinline__ void ffma2_b32(uint32_t& x0, uint32_t& x1, float a, float c) {
  asm("{\n\t.reg .b64 rx, ra, rc;\n\t"
      "mov.b64 rx, {%0, %1};\n\tmov.b64 ra, {%2, %2};\n\tmov.b64 rc, {%3, %3};\n\t"
      "fma.rn.f32x2 rx, rx, ra, rc;\n\t"
      "mov.b64 {%0, %1}, rx;\n\t}"
      : "+r"(x0), "+r"(x1) : "f"(a), "f"(c));
}
__device__ __forceinline__ void ffma2_b32_v(uint32_t& x0, uint32_t& x1, float a, float c) {
  asm volatile("{\n\t.reg .b64 rx, ra, rc;\n\t"
               "mov.b64 rx, {%0, %1};\n\tmov.b64 ra, {%2, %2};\n\tmov.b64 rc, {%3, %3};\n\t"
               "fma.rn.f32x2 rx, rx, ra, rc;\n\t"
               "mov.b64 {%0, %1}, rx;\n\t}"
               : "+r"(x0), "+r"(x1) : "f"(a), "f"(c));
}
__device__ __forceinline__ void ffma2_b32_v2(uint32_t& x0, uint32_t& x1, float a, float c0, float c1) {   // per-lane addend
  asm volatile("{\n\t.reg .b64 rx, ra, rc;\n\t"
               "mov.b64 rx, {%0, %1};\n\tmov.b64 ra, {%2, %2};\n\tmov.b64 rc, {%3, %4};\n\t"
               "fma.rn.f32x2 rx, rx, ra, rc;\n\t"
               "mov.b64 {%0, %1}, rx;\n\t}"
               : "+r"(x0), "+r"(x1) : "f"(a), "f"(c0), "f"(c1));
}
// {c0, c1} = {l0, l1} * 0 + c: the value of c with a data dependency on l (a scheduling throttle, see attention_tc.cu)
__device__ __forceinline__ void dep_on(float& c0, float& c1, float l0, float l1, float c) {
  asm volatile("{\n\t.reg .b64 rl, rz, rc;\n\t"
               "mov.b64 rl, {%2, %3};\n\tmov.b64 rz, {0f00000000, 0f00000000};\n\tmov.b64 rc, {%4, %4};\n\t"
               "fma.rn.f32x2 rc, rl, rz, rc;\n\t"
               "mov.b64 {%0, %1}, rc;\n\t}"
               : "=f"(c0), "=f"(c1) : "f"(l0), "f"(l1), "f"(c));
}
__device__ __forceinline__ void ex2_b32_v(uint32_t& x) { asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+r"(x)); }
__device__ __forceinline__ void fadd2_b32_v(float& l0, float& l1, uint32_t x0, uint32_t x1) {
  asm volatile("{\n\t.reg .b64 rl, rx;\n\t"
               "mov.b64 rl, {%0, %1};\n\tmov.b64 rx, {%2, %3};\n\t"
               "add.rn.f32x2 rl, rl, rx;\n\t"
               "mov.b64 {%0, %1}, rl;\n\t}"
               : "+f"(l0), "+f"(l1) : "r"(x0), "r"(x1));
}
__device__ __forceinline__ uint32_t pack_bf16x2_b32_v(uint32_t lo, uint32_t hi) {
  uint32_t r;
  asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "r"(hi), "r"(lo));
  return r;
}


// ---- CTA-pair (cta_group::2) variants ---------------------------------------------------------------------
// Two CTAs of a cluster (same TPC) execute ONE tcgen05.mma of M = 256: each CTA supplies its 128 rows of A and
// half of the B tile from its own shared memory and receives its 128 accumulator rows in its own TMEM.  Only the
// even CTA issues MMAs; TMA loads of both CTAs report their bytes to the even CTA's mbarrier.
constexpr uint32_t PEER_BIT_MASK = 0xFEFFFFFFu;   // clears the CTA-rank bit of a shared::cluster address -> even CTA
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_4d_2sm(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1,
                                                int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(s_u32(dst)), "l"(map), "r"(s_u32(bar) & PEER_BIT_MASK), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_load_3d_2sm(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(s_u32(dst)), "l"(map), "r"(s_u32(bar) & PEER_BIT_MASK), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(s_u32(dst)), "l"(map), "r"(s_u32(bar) & PEER_BIT_MASK), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void umma_bf16_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                              uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// arrives (once all prior MMAs of this thread retire) on the barrier at the same offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma_commit_2sm(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(s_u32(bar)), "h"((uint16_t)3) : "memory");
}
// arrive on the barrier at this offset in the cluster's CTA `rank` (remote shared memory)
__device__ __forceinline__ void mbar_arrive_cluster(uint64_t* bar, uint32_t rank) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t}"
      ::"r"(s_u32(bar)), "r"(rank) : "memory");
}


// exp2 of two values on the FMA / ALU pipes (no SFU): Cody-Waite split x = n + f with the 1.5 * 2^23 magic-number
// add, degree-3 minimax polynomial for 2^f on [-0.5, 0.5] (max rel. error 1.9e-4: far below the bf16 rounding the
// result gets as an MMA operand), exponent insertion by integer add.  Used for a fraction of the softmax
// exponentials so that MUFU.EX2 (16 lanes / clk / SM) stops being the only pipe the attention kernel waits on.
__device__ __forceinline__ void exp2_poly2(float& x0, float& x1) {
  const float MAGIC = 12582912.0f;                       // 1.5 * 2^23: round-to-nearest-integer in the low mantissa bits
  x0 = fmaxf(x0, -125.0f);                               // keeps the biased exponent positive (also maps -inf)
  x1 = fmaxf(x1, -125.0f);
  float t0, t1, n0, n1, f0, f1, p0, p1;
  fadd2(t0, t1, x0, x1, MAGIC, MAGIC);
  fadd2(n0, n1, t0, t1, -MAGIC, -MAGIC);
  ffma2(f0, f1, n0, n1, -1.0f, -1.0f, x0, x1);           // f = x - round(x)
  ffma2(p0, p1, f0, f1, 0.0558755025f, 0.0558755025f, 0.2422944456f, 0.2422944456f);
  ffma2(p0, p1, p0, p1, f0, f1, 0.6931272745f, 0.6931272745f);
  ffma2(p0, p1, p0, p1, f0, f1, 0.9999482632f, 0.9999482632f);
  x0 = __int_as_float(__float_as_int(p0) + (__float_as_int(t0) << 23));
  x1 = __int_as_float(__float_as_int(p1) + (__float_as_int(t1) << 23));
}

// The polynomial exp2 of exp2_poly2 in six separately placeable stages on b32 registers (volatile: the hand-scheduled
// exponential section of attention_tc.cu spreads them between its MUFU.EX2 issues).  x: the pair, t: scratch pair.
__device__ __forceinline__ void poly_s1(uint32_t& x0, uint32_t& x1, uint32_t& t0, uint32_t& t1) {   // clamp, t = x + 1.5 * 2^23
  asm volatile("{\n\t.reg .b64 rx, rt, rm;\n\t"
               "max.f32 %0, %0, 0fC2FA0000;\n\tmax.f32 %1, %1, 0fC2FA0000;\n\t"
               "mov.b64 rx, {%0, %1};\n\tmov.b64 rm, {0f4B400000, 0f4B400000};\n\t"
               "add.rn.f32x2 rt, rx, rm;\n\t"
               "mov.b64 {%2, %3}, rt;\n\t}"
               : "+r"(x0), "+r"(x1), "=r"(t0), "=r"(t1));
}
__device__ __forceinline__ void poly_s2a(uint32_t& n0, uint32_t& n1, uint32_t t0, uint32_t t1) {   // n = t - 1.5 * 2^23 = round(x)
  asm volatile("{\n\t.reg .b64 rn, rt, rm;\n\t"
               "mov.b64 rt, {%2, %3};\n\tmov.b64 rm, {0fCB400000, 0fCB400000};\n\t"
               "add.rn.f32x2 rn, rt, rm;\n\t"
               "mov.b64 {%0, %1}, rn;\n\t}"
               : "=r"(n0), "=r"(n1) : "r"(t0), "r"(t1));
}
__device__ __forceinline__ void poly_s2b(uint32_t& x0, uint32_t& x1, uint32_t n0, uint32_t n1) {   // f = x - n  (in x)
  asm volatile("{\n\t.reg .b64 rx, rn, rm;\n\t"
               "mov.b64 rx, {%0, %1};\n\tmov.b64 rn, {%2, %3};\n\tmov.b64 rm, {0fBF800000, 0fBF800000};\n\t"
               "fma.rn.f32x2 rx, rn, rm, rx;\n\t"
               "mov.b64 {%0, %1}, rx;\n\t}"
               : "+r"(x0), "+r"(x1) : "r"(n0), "r"(n1));
}
__device__ __forceinline__ void poly_s3(uint32_t& p0, uint32_t& p1, uint32_t f0, uint32_t f1) {   // p = c3 f + c2
  asm volatile("{\n\t.reg .b64 rp, rf, ra, rb;\n\t"
               "mov.b64 rf, {%2, %3};\n\tmov.b64 ra, {0f3D64DDB6, 0f3D64DDB6};\n\tmov.b64 rb, {0f3E781C09, 0f3E781C09};\n\t"
               "fma.rn.f32x2 rp, rf, ra, rb;\n\t"
               "mov.b64 {%0, %1}, rp;\n\t}"
               : "=r"(p0), "=r"(p1) : "r"(f0), "r"(f1));
}
template <int STAGE>                                                                               // p = p f + c1 | c0
__device__ __forceinline__ void poly_s45(uint32_t& p0, uint32_t& p1, uint32_t f0, uint32_t f1) {
  if constexpr (STAGE == 4)
    asm volatile("{\n\t.reg .b64 rp, rf, rc;\n\t"
                 "mov.b64 rp, {%0, %1};\n\tmov.b64 rf, {%2, %3};\n\tmov.b64 rc, {0f3F3170CA, 0f3F3170CA};\n\t"
                 "fma.rn.f32x2 rp, rp, rf, rc;\n\t"
                 "mov.b64 {%0, %1}, rp;\n\t}"
                 : "+r"(p0), "+r"(p1) : "r"(f0), "r"(f1));
  else
    asm volatile("{\n\t.reg .b64 rp, rf, rc;\n\t"
                 "mov.b64 rp, {%0, %1};\n\tmov.b64 rf, {%2, %3};\n\tmov.b64 rc, {0f3F7FFC9C, 0f3F7FFC9C};\n\t"
                 "fma.rn.f32x2 rp, rp, rf, rc;\n\t"
                 "mov.b64 {%0, %1}, rp;\n\t}"
                 : "+r"(p0), "+r"(p1) : "r"(f0), "r"(f1));
}
__device__ __forceinline__ void poly_s6(uint32_t& x0, uint32_t& x1, uint32_t p0, uint32_t p1, uint32_t t0, uint32_t t1) {  // 2^n * p
  asm volatile("mad.lo.s32 %0, %4, 0x800000, %2;\n\tmad.lo.s32 %1, %5, 0x800000, %3;"
               : "=r"(x0), "=r"(x1) : "r"(p0), "r"(p1), "r"(t0), "r"(t1));
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode_fn();
int encode_map(CUtensorMap* map, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
               const uint32_t* box, const uint32_t* estrides, const char* what,
               CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B);

}  // namespace pd
