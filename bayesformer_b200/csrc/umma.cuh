// sm_100a tensor-core plumbing: tcgen05.mma (UMMA) with operands in shared memory and the fp32
// accumulator in tensor memory (TMEM), mbarrier completion, TMEM load/store, 1-D bulk copies (TMA).
// Inline PTX only; the bit layouts follow the PTX ISA "tcgen05" matrix/instruction descriptors.
#pragma once

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace bfb {
namespace umma {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ------------------------------------------------------------------------------------------
// mbarrier
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t arrivals) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(arrivals) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
  } while (!done);
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

// 1-D bulk copy global -> shared (TMA engine, SASS UBLKCP), completion counted on `bar` in bytes.
// dst/src 16-byte aligned, bytes a multiple of 16.
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// generic-proxy writes to shared memory -> visible to the async proxy (tcgen05.mma / TMA reads)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ------------------------------------------------------------------------------------------
// TMEM allocation (one full warp executes these)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_result, uint32_t n_cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)),
               "r"(n_cols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t n_cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(n_cols) : "memory");
}
__device__ __forceinline__ void tc_fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// ------------------------------------------------------------------------------------------
// descriptors
// ------------------------------------------------------------------------------------------
// Shared-memory matrix descriptor of a K-major operand tile stored as rows of 128 bytes (64 bf16) with the
// 128-byte swizzle: byte (row r, column byte c) lives at r*128 + (((c >> 4) ^ (r & 7)) << 4) + (c & 15),
// tile base 1024-byte aligned; 8-row groups are 1024 bytes apart (stride byte offset).
// Advancing along K by one UMMA step (16 bf16 = 32 bytes) adds 32 bytes to the start address.
__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);   // [0,14)  start address >> 4
  d |= (uint64_t)0 << 16;                         // [16,30) leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)(1024u >> 4) << 32;              // [32,46) stride byte offset >> 4
  d |= (uint64_t)1 << 46;                         // [46,48) descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                         // [61,64) layout: SWIZZLE_128B
  return d;
}

// Instruction descriptor of kind::f16: 16-bit operands (format 0 = F16, 1 = BF16) -> fp32, both operands
// K-major, shape M x N x 16.
constexpr uint32_t kFmtF16 = 0, kFmtBF16 = 1;
__host__ __device__ constexpr uint32_t instr_desc_16bit(int M, int N, uint32_t fmt) {
  return (1u << 4)                      // [4,6)   accumulator format: F32
         | (fmt << 7)                   // [7,10)  A format
         | (fmt << 10)                  // [10,13) B format
         | (0u << 15) | (0u << 16)      // A, B K-major
         | ((uint32_t)(N >> 3) << 17)   // [17,23) N >> 3
         | ((uint32_t)(M >> 4) << 24);  // [24,29) M >> 4
}
__host__ __device__ constexpr uint32_t instr_desc_bf16(int M, int N) { return instr_desc_16bit(M, N, kFmtBF16); }
__host__ __device__ constexpr uint32_t instr_desc_f16(int M, int N) { return instr_desc_16bit(M, N, kFmtF16); }

// D[tmem] (+)= A[smem] * B[smem]^T ; issued by ONE thread.
__device__ __forceinline__ void mma_bf16_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

// D[tmem] (+)= A[tmem] * B[smem]^T : the "TS" form — the A operand (16-bit, two K elements per 32-bit column, lane =
// row) is read from tensor memory, where the threads that produced it left it with tcgen05.st.
__device__ __forceinline__ void mma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                           uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

// All previously issued MMAs of this thread arrive on `bar` when complete (implies fence::before_thread_sync).
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// ------------------------------------------------------------------------------------------
// TMEM <-> registers.  32x32b: the warp's lane i addresses TMEM lane (warp % 4) * 32 + i; the N registers are
// N consecutive 32-bit columns.  `taddr` = (lane_base << 16) | column.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,"
      "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
        "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
        "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st4(uint32_t taddr, const uint32_t (&r)[4]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]),
               "r"(r[3])
               : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// named barrier for a sub-group of the CTA (ids 1..15; 0 is __syncthreads)
__device__ __forceinline__ void group_sync(int id, int n_threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n_threads) : "memory");
}

// ------------------------------------------------------------------------------------------
// operand staging helpers (swizzled 128-byte rows)
// ------------------------------------------------------------------------------------------
// byte offset of 16-byte chunk `chunk` (0..7) of row `row` inside a swizzled tile
__device__ __forceinline__ uint32_t sw128_offset(uint32_t row, uint32_t chunk) {
  return row * 128u + (((chunk ^ (row & 7u)) & 7u) << 4);
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  const __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&v);
}
// fp16 with saturation to +-65504 (an overflowing activation must not become inf inside the tensor core)
__device__ __forceinline__ uint32_t pack_f16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ float2 unpack_f16x2(uint32_t packed) {
  return __half22float2(*reinterpret_cast<const __half2*>(&packed));
}
__device__ __forceinline__ float bf16_lo(uint32_t packed) { return __uint_as_float(packed << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t packed) { return __uint_as_float(packed & 0xFFFF0000u); }

// ------------------------------------------------------------------------------------------
// Mixed-precision FMA, d = a * b + c with fp16 a, b taken straight from a half of a packed word and fp32 c, d
// (PTX fma.rn.f32.f16, SASS FHFMA with .H0/.H1 operand selectors): fp16 tiles are multiplied (or, against a packed 1.0, added)
// without unpacking them.  The product of two fp16 values is exact in fp32, so this is the arithmetic of
// "unpack, then FFMA" at half the instructions.
#define BFB_FH(NAME, ASEL, BSEL)                                                                              \
  __device__ __forceinline__ float NAME(uint32_t a, uint32_t b, float c) {                                    \
    float d;                                                                                                  \
    asm("{\n\t.reg .b16 al, ah, bl, bh;\n\tmov.b32 {al, ah}, %1;\n\tmov.b32 {bl, bh}, %2;\n\t"                  \
        "fma.rn.f32.f16 %0, " ASEL ", " BSEL ", %3;\n\t}"                                                     \
        : "=f"(d)                                                                                             \
        : "r"(a), "r"(b), "f"(c));                                                                            \
    return d;                                                                                                 \
  }
BFB_FH(fh_ll, "al", "bl")
BFB_FH(fh_hh, "ah", "bh")
BFB_FH(fh_lh, "al", "bh")
BFB_FH(fh_hl, "ah", "bl")
#undef BFB_FH

// fp32 pair -> fp16 hi + fp16 lo words (a in the low halves): hi = fp16(v) (saturating), lo = fp16(v - hi); v = hi + lo
// to 2^-22 relative.  The residuals come from one mixed-precision FMA each (hi * -1 + v, exact in fp32).
__device__ __forceinline__ void split_f16x2(float a, float b, uint32_t& hi, uint32_t& lo) {
  hi = pack_f16x2(a, b);
  constexpr uint32_t kNegOne = 0xBC00BC00u;
  lo = pack_f16x2(fh_ll(hi, kNegOne, a), fh_hl(hi, kNegOne, b));
}
// acc += float(half `HI` of hi) + float(same half of lo)
__device__ __forceinline__ float add_hilo_l(uint32_t hi, uint32_t lo, float acc) {
  constexpr uint32_t kOne = 0x3C003C00u;
  return fh_ll(lo, kOne, fh_ll(hi, kOne, acc));
}
__device__ __forceinline__ float add_hilo_h(uint32_t hi, uint32_t lo, float acc) {
  constexpr uint32_t kOne = 0x3C003C00u;
  return fh_hl(lo, kOne, fh_hl(hi, kOne, acc));
}


}  // namespace umma
}  // namespace bfb
