// sm_100a building blocks shared by the tensor-core kernels: mbarrier, bulk async copy (TMA engine, 1-D),
// tcgen05 (UMMA) descriptors / issue / commit, TMEM allocation and access.  Inline PTX only.
#pragma once
#include <cuda_bf16.h>
#include <stdint.h>

namespace gldm {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// one lane of a fully converged warp (the caller's control flow must be warp-uniform)
__device__ __forceinline__ bool elect_one_sync() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}

// ------------------------------------------------------------------------------------------ mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// non-blocking probe of a phase
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// blocking wait: the suspend-time hint lets the hardware park the thread until the phase completes instead of
// re-issuing try_wait (a spinning warp would steal issue slots from the warp that drives the tensor core)
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
      "@p bra WAIT_DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t}" ::"r"(smem_u32(bar)),
      "r"(parity), "r"(0x989680u)
      : "memory");
}

// ------------------------------------------------------------------------------------------ bulk copy
// 1-D bulk async copy global -> shared (TMA engine, SASS UBLKCP); completion counted on `bar` in bytes.
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// generic-proxy writes (st.shared) must be fenced before the async proxy (UMMA / TMA) reads them
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ------------------------------------------------------------------------------------------ tcgen05
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// whole warp; writes the TMEM base address to *slot (shared memory)
template <int COLS>
__device__ __forceinline__ void tmem_alloc(uint32_t* slot) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "n"(COLS)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t addr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "n"(COLS) : "memory");
}

enum : uint32_t { SW_NONE = 0, SW_128 = 2, SW_64 = 4, SW_32 = 6 };

// shared-memory matrix descriptor, K-major operand: 8-row groups `sbo_bytes` apart, rows `swizzle span` apart
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t sbo_bytes, uint32_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;                               // leading byte offset: unused for swizzled K-major
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;                               // descriptor version (Blackwell)
  d |= (uint64_t)layout << 61;
  return d;
}

// instruction descriptor, kind::f16: BF16 x BF16 -> F32, both operands K-major
__host__ __device__ constexpr uint32_t idesc_bf16(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// D[tmem] (+)= A[smem] * B[smem]; issued by ONE thread
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Warp-collective forms (the whole warp executes them with identical operands; one elected lane issues): no
// C++-level divergence around the tensor-core instructions, so the operands stay in uniform registers.
// KSTEPS consecutive K=16 steps of one swizzled K-block: descriptors advance by 32 bytes (2 x 16-byte units).
template <int KSTEPS>
__device__ __forceinline__ void umma_bf16_block_elect(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                                      uint32_t accumulate_first) {
  static_assert(KSTEPS == 1 || KSTEPS == 2 || KSTEPS == 4, "KSTEPS");
  if (KSTEPS == 1) {
    asm volatile(
        "{\n\t.reg .pred q, p;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate_first)
        : "memory");
  } else if (KSTEPS == 2) {
    asm volatile(
        "{\n\t.reg .pred q, p, t;\n\t.reg .b64 a1, b1;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "setp.eq.b32 t, 0, 0;\n\t"
        "add.u64 a1, %1, 2;\n\tadd.u64 b1, %2, 2;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], a1, b1, %3, t;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate_first)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred q, p, t;\n\t.reg .b64 a1, b1, a2, b2, a3, b3;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "setp.eq.b32 t, 0, 0;\n\t"
        "add.u64 a1, %1, 2;\n\tadd.u64 b1, %2, 2;\n\t"
        "add.u64 a2, %1, 4;\n\tadd.u64 b2, %2, 4;\n\t"
        "add.u64 a3, %1, 6;\n\tadd.u64 b3, %2, 6;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], a1, b1, %3, t;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], a2, b2, %3, t;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], a3, b3, %3, t;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate_first)
        : "memory");
  }
}
// Three UMMAs behind one elect: accumulators d, d + 64, d + 128 columns; A descriptors 8 units (one 128-byte row) apart,
// B descriptors b_step units apart.  The issuing warp is instruction bound when every UMMA carries its own address
// arithmetic and elect (measured: ~86 cycles per 128 x 48 x 16 UMMA in the Conv3d issue loop), so the per-UMMA work
// is kept to the adds below.
__device__ __forceinline__ void umma_bf16_x3_elect(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint64_t b_step,
                                                   uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred q, p;\n\t.reg .b64 a1, a2, b1, b2;\n\t.reg .b32 d1, d2;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "setp.ne.b32 p, %5, 0;\n\t"
      "add.u64 a1, %1, 8;\n\tadd.u64 a2, %1, 16;\n\t"
      "add.u64 b1, %2, %4;\n\tadd.u64 b2, b1, %4;\n\t"
      "add.u32 d1, %0, 64;\n\tadd.u32 d2, %0, 128;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [d1], a1, b1, %3, p;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [d2], a2, b2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "l"(b_step), "r"(accumulate)
      : "memory");
}
// the same for operand rows of ASTEP 16-byte units (32-byte rows of the narrow first layer: ASTEP = 2)
template <int ASTEP>
__device__ __forceinline__ void umma_bf16_x3_elect_a(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint64_t b_step,
                                                     uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred q, p;\n\t.reg .b64 a1, a2, b1, b2;\n\t.reg .b32 d1, d2;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "setp.ne.b32 p, %5, 0;\n\t"
      "add.u64 a1, %1, %6;\n\tadd.u64 a2, %1, %7;\n\t"
      "add.u64 b1, %2, %4;\n\tadd.u64 b2, b1, %4;\n\t"
      "add.u32 d1, %0, 64;\n\tadd.u32 d2, %0, 128;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [d1], a1, b1, %3, p;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [d2], a2, b2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "l"(b_step), "r"(accumulate), "n"(ASTEP), "n"(2 * ASTEP)
      : "memory");
}
// three taps into ONE accumulator (the first with the caller's accumulate flag, the others accumulating)
template <int ASTEP>
__device__ __forceinline__ void umma_bf16_x3_same_elect_a(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint64_t b_step,
                                                          uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred q, p, t;\n\t.reg .b64 a1, a2, b1, b2;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "setp.ne.b32 p, %5, 0;\n\t"
      "setp.eq.b32 t, 0, 0;\n\t"
      "add.u64 a1, %1, %6;\n\tadd.u64 a2, %1, %7;\n\t"
      "add.u64 b1, %2, %4;\n\tadd.u64 b2, b1, %4;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], a1, b1, %3, t;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], a2, b2, %3, t;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "l"(b_step), "r"(accumulate), "n"(ASTEP), "n"(2 * ASTEP)
      : "memory");
}
__device__ __forceinline__ void umma_commit_elect(uint64_t* bar) {
  asm volatile(
      "{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\t"
      "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(smem_u32(bar))
      : "memory");
}
// expect_tx + bulk copy by one elected lane
__device__ __forceinline__ void bulk_g2s_elect(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "{\n\t.reg .pred q;\n\t.reg .b64 st;\n\telect.sync _|q, 0xffffffff;\n\t"
      "@q mbarrier.arrive.expect_tx.shared::cta.b64 st, [%3], %2;\n\t"
      "@q cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n\t}" ::"r"(
          smem_u32(smem_dst)),
      "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

// arrive on `bar` when every previously issued UMMA of this thread has completed (implies fence::before_thread_sync)
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// TMEM -> registers: the warp's 32 lanes x N consecutive 32-bit columns (thread i <-> lane base+i)
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&v)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, uint32_t (&v)[4]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3])
               : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,"
      "%30,%31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// registers -> TMEM
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(v[0]),
               "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st4(uint32_t taddr, const uint32_t (&v)[4]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]),
               "r"(v[3])
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// byte offset of element (row r, 16-byte chunk j) inside a K-major swizzled operand slab whose rows are SWB bytes
// (SWB = 32, 64 or 128; the slab base must be 1024-byte aligned): Swizzle<log2(SWB/16), 4, 3>
template <int SWB>
__host__ __device__ __forceinline__ uint32_t swz_off(uint32_t r, uint32_t j) {
  const uint32_t x = (SWB == 128) ? (r & 7) : (SWB == 64) ? ((r >> 1) & 3) : ((r >> 2) & 1);
  return r * SWB + ((j ^ x) << 4);
}

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}

// reduce-scatter step: NOUT pairs (i, i + NOUT); lanes with bit `OFF` set keep the upper element
template <int OFF, int NOUT, int NV>
__device__ __forceinline__ void rs_step(float (&a)[NV], int lane) {
  const bool hi = (lane & OFF) != 0;
#pragma unroll
  for (int i = 0; i < NOUT; ++i) {
    const float send = hi ? a[i] : a[i + NOUT];
    const float keep = hi ? a[i + NOUT] : a[i];
    a[i] = keep + __shfl_xor_sync(0xffffffffu, send, OFF);
  }
}

}  // namespace tc
}  // namespace gldm
