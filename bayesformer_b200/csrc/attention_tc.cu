// Fused causal attention for heads of width 64 on tcgen05 tensor cores (the scaled configuration: d_model 512,
// 8 heads, sequences of up to 256 tokens).  Reference arithmetic: src/model/transformer.py:58-65 (per head:
// softmax(Q K^T / sqrt(dh) + causal mask) V) and :116-124 (optional head-output dropout, concatenation).
//
// One CTA (256 threads: two per query row, alternating 32-key chunks) per (sequence, head, tile of 128 queries);
// two to four CTAs share an SM (48 / 96 KB of shared memory, 128 / 256 TMEM columns each), so one CTA's softmax
// overlaps the others' copies and tensor-core work:
//   1. one thread issues 2-D TMA loads (128-byte swizzle) of the Q tile and of the K and V rows this tile can attend
//      to, straight out of the packed (R, 3d) QKV activation;
//   2. S = Q K^T: tcgen05.mma M = 128, N = keys (<= 256), K = 64, fp32 accumulator in 256 TMEM columns;
//   3. softmax, thread = query row: two passes over the row with tcgen05.ld (max, then exp / sum with the causal
//      mask), the un-normalised probabilities go to shared memory as the 16-bit A operand (swizzled K-major slabs,
//      written over the dead Q / K tiles);
//   4. O = P V: tcgen05.mma M = 128, N = 64, K = keys.  V is consumed exactly as TMA delivered it — keys x dims,
//      i.e. an MN-major B operand (instruction-descriptor bit 16) — and O accumulates into the first 64 of the S
//      columns, dead once every probability is in shared memory;
//   5. epilogue: O row / row sum (+ head-output dropout), converted, one 128-byte line per query row to HBM.
// Q, K, V, P are 16-bit (fp16 or bf16); scores, softmax statistics and both accumulators are fp32.
#include <cuda.h>

#include "common.cuh"
#include "kernels.cuh"
#include "umma.cuh"

namespace bfb {

namespace {

constexpr int AT_DH = 64;
constexpr int AT_M = 128;        // queries per CTA
constexpr int AT_MAXKV = 256;    // most keys a CTA attends to
constexpr uint32_t AT_Q_BYTES = AT_M * 128;
// shared memory of the variant that attends to at most MAXKV keys (128: the first query tile of a sequence, four
// CTAs per SM; 256: the second tile, two CTAs per SM):  [ P slabs over Q | K ][ V ]
template <int MAXKV> struct AtSmem {
  // The 256-key variant keeps the softmax weights P in TENSOR memory (A operand of the PV product read from TMEM,
  // "TS" form of tcgen05.mma) and needs no P region; the 128-key variant writes P over the dead Q / K tiles.
  static constexpr bool TS = MAXKV == 256;
  static constexpr uint32_t K_BYTES = MAXKV * 128;
  static constexpr uint32_t V_BYTES = MAXKV * 128;                   // keys x 64 dims as delivered by TMA
  static constexpr uint32_t P_BYTES = TS ? AT_Q_BYTES + K_BYTES : (MAXKV / 64) * AT_M * 128;   // slabs of 128 rows x 64 keys
  static constexpr uint32_t TOTAL = 1024 + P_BYTES + V_BYTES;
  static_assert(P_BYTES >= AT_Q_BYTES + K_BYTES, "P must cover the Q and K tiles it overwrites");
};

// D[tmem] (+)= A[tmem] * B[smem]: the A operand (16-bit, two elements per 32-bit column, lane = row) comes from tensor memory
__device__ __forceinline__ void at_mma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void at_tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}

__device__ __forceinline__ void at_tma_load_2d(void* smem_dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(umma::smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(umma::smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}

struct AttnArgs {
  const uint16_t* qkv;   // (R, 3d)
  uint16_t* out;         // (R, d)
  int64_t n_rows;        // R = n_seq * len
  int64_t n_items;       // n_seq * n_heads (sequence, head) work items per query tile
  int len, d, n_heads, m_tile;   // this launch handles query tile m_tile of every (sequence, head)
  uint32_t fmt;
  float inv_sqrt_dh;
  int site;              // head-output dropout site or -1
  MaskSource ms;
};

__device__ __forceinline__ uint32_t at_pack2(float a, float b, uint32_t fmt) {
  return fmt == umma::kFmtF16 ? umma::pack_f16x2(a, b) : umma::pack_bf16x2(a, b);
}

#ifdef BFB_ATTN_TIMING
// debug build only (make EXTRA=-DBFB_ATTN_TIMING, tools/attn_zoo.py prints it): cycles thread 0 of every CTA of the
// 256-key variant spends up to each milestone
__device__ unsigned long long bfb_attn_dbg[8];
#define AT_MARK(slot) do { if (MAXKV == 256 && tid == 0) atomicAdd(&bfb_attn_dbg[slot], (unsigned long long)(clock64() - at_t0)); } while (0)
#else
#define AT_MARK(slot) do { } while (0)
#endif

template <int MAXKV, bool F16>
__global__ void __launch_bounds__(2 * AT_M, MAXKV == 128 ? 3 : 2) attention_tc_kernel(const __grid_constant__ CUtensorMap map_qkv,
                                                                                  const AttnArgs a) {
  constexpr uint32_t FMT = F16 ? umma::kFmtF16 : umma::kFmtBF16;   // compile-time: no per-element format branches
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (umma::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + AT_Q_BYTES;
  uint8_t* sP = smem;                 // written only after S = Q K^T has completed
  uint8_t* sV = smem + AtSmem<MAXKV>::P_BYTES;
  __shared__ uint64_t bar_qk, bar_v, bar_s, bar_o;
  __shared__ uint32_t tmem_slot;
  __shared__ float s_max[2][AT_M], s_sum[2][AT_M];

  const int tid = threadIdx.x, warp = tid >> 5;
#ifdef BFB_ATTN_TIMING
  long long at_t0 = clock64();
#endif
  const int row = tid & (AT_M - 1), half = tid >> 7;   // warps w and w + 4 share TMEM lane quarter w & 3
  const int m0 = a.m_tile * AT_M;
  const int kv_len = min(a.len, m0 + AT_M);
  const int kvp = (kv_len + 15) & ~15;

  if (warp == 0) {
    umma::tmem_alloc(&tmem_slot, MAXKV);
    umma::tmem_relinquish();
  }
  if (tid == 0) {
    umma::mbar_init(&bar_qk, 1);
    umma::mbar_init(&bar_v, 1);
    umma::mbar_init(&bar_s, 1);
    umma::mbar_init(&bar_o, 1);
    umma::fence_mbar_init();
  }
  umma::tc_fence_before_sync();
  __syncthreads();
  umma::tc_fence_after_sync();
  const uint32_t tmem_base = tmem_slot;
  AT_MARK(0);   // TMEM allocated, barriers initialised

  // Persistent over (sequence, head) items: TMEM and the barriers are set up once per CTA, and an item's tiles are
  // requested as soon as the previous item's PV product has released the shared memory — while that item's output is
  // still being normalised and stored.
  constexpr bool TS = AtSmem<MAXKV>::TS;
  constexpr uint32_t o_col = TS ? MAXKV / 4 : 0;   // first accumulator column of O
  constexpr int HK = MAXKV / 2;   // TS: thread `half` of a row owns the CONTIGUOUS keys [half HK, half HK + HK)
  auto issue_qk = [&](int64_t item) {
    const int hh = (int)(item % a.n_heads);
    const int64_t rb = (item / a.n_heads) * a.len;
    const int k_boxes = (kvp + AT_M - 1) / AT_M;
    umma::mbar_expect_tx(&bar_qk, AT_Q_BYTES + (uint32_t)k_boxes * AT_M * 128u);
    at_tma_load_2d(sQ, &map_qkv, hh * AT_DH, (int)(rb + m0), &bar_qk);
    for (int b = 0; b < k_boxes; ++b)
      at_tma_load_2d(sK + b * AT_M * 128, &map_qkv, a.d + hh * AT_DH, (int)(rb + b * AT_M), &bar_qk);
  };
  auto issue_v = [&](int64_t item) {
    const int hh = (int)(item % a.n_heads);
    const int64_t rb = (item / a.n_heads) * a.len;
    const int k_boxes = (kvp + AT_M - 1) / AT_M;
    umma::mbar_expect_tx(&bar_v, (uint32_t)k_boxes * AT_M * 128u);
    for (int b = 0; b < k_boxes; ++b)
      at_tma_load_2d(sV + b * AT_M * 128, &map_qkv, 2 * a.d + hh * AT_DH, (int)(rb + b * AT_M), &bar_v);
  };
  auto issue_loads = [&](int64_t item) { issue_qk(item); issue_v(item); };
  if (tid == 0 && (int64_t)blockIdx.x < a.n_items) issue_loads(blockIdx.x);

  uint32_t phase = 0;
  for (int64_t item = blockIdx.x; item < a.n_items; item += gridDim.x, phase ^= 1u) {
  const int h = (int)(item % a.n_heads);
  const int64_t s = item / a.n_heads;
  const int64_t row_base = s * a.len;
#ifdef BFB_ATTN_TIMING
  if (item != (int64_t)blockIdx.x) at_t0 = clock64();   // milestones are per item (the first one includes the set-up)
#endif

  // ---- S = Q K^T ----
  if (tid == 0) {
    umma::mbar_wait(&bar_qk, phase);
    umma::tc_fence_after_sync();
    const uint64_t dq = umma::smem_desc_sw128(umma::smem_u32(sQ)), dk = umma::smem_desc_sw128(umma::smem_u32(sK));
    const uint32_t idesc = umma::instr_desc_16bit(AT_M, kvp, FMT);
#pragma unroll
    for (int k = 0; k < AT_DH / 16; ++k) umma::mma_bf16_ss(tmem_base, dq + 2 * k, dk + 2 * k, idesc, k > 0);
    umma::mma_commit(&bar_s);
  }
  umma::mbar_wait(&bar_s, phase);
  umma::tc_fence_after_sync();
  AT_MARK(1);   // Q / K landed and S = Q K^T complete
  if (TS && tid == 0 && item + gridDim.x < a.n_items) issue_qk(item + gridDim.x);   // Q / K tiles are dead already

  // ---- softmax over the keys j <= query position; two threads per query row, alternating 32-key chunks ----
  // A warp's 32 rows are queries m0 + 32 w + lane and a chunk is 32 keys, so every chunk is, for the WHOLE warp,
  // either entirely visible (c0 + 31 <= first query of the warp), the diagonal one (c0 = first query: key i of the
  // chunk is visible to lane l iff i <= l), or entirely masked — no per-element causal test outside the diagonal
  // (generate_subsequent_mask, transformer.py:275-287).
  const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
  const int lane = tid & 31;
  const int tq = m0 + row;                       // query position inside the sequence
  const bool q_ok = tq < a.len;
  const int warp_first = m0 + (warp & 3) * 32;   // first query row of this warp = its diagonal chunk
  const float cs = a.inv_sqrt_dh * 1.4426950408889634f;   // scores / sqrt(dh), as a base-2 exponent
  float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
  const int c_begin = TS ? half * HK : 32 * half;            // first chunk of this thread
  const int c_end = TS ? min(kvp, half * HK + HK) : kvp;     // one past its last key
  const int c_step = TS ? 32 : 64;
  for (int c0 = c_begin; c0 <= warp_first && c0 < c_end; c0 += c_step) {
    uint32_t r[32];
    umma::tmem_ld32(tmem_base + lane_base + c0, r);
    umma::tmem_ld_wait();
    if (c0 < warp_first) {
#pragma unroll
      for (int i = 0; i < 32; i += 4) {
        mx0 = fmaxf(mx0, __uint_as_float(r[i]));     mx1 = fmaxf(mx1, __uint_as_float(r[i + 1]));
        mx2 = fmaxf(mx2, __uint_as_float(r[i + 2])); mx3 = fmaxf(mx3, __uint_as_float(r[i + 3]));
      }
    } else {
#pragma unroll
      for (int i = 0; i < 32; i += 4) {
        mx0 = fmaxf(mx0, i <= lane ? __uint_as_float(r[i]) : -INFINITY);
        mx1 = fmaxf(mx1, i + 1 <= lane ? __uint_as_float(r[i + 1]) : -INFINITY);
        mx2 = fmaxf(mx2, i + 2 <= lane ? __uint_as_float(r[i + 2]) : -INFINITY);
        mx3 = fmaxf(mx3, i + 3 <= lane ? __uint_as_float(r[i + 3]) : -INFINITY);
      }
    }
  }
  float mx = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
  s_max[half][row] = mx;
  __syncthreads();
  AT_MARK(2);   // row maxima
  mx = fmaxf(mx, s_max[half ^ 1][row]);   // finite: key 0 is visible to every query
  const float nmx = -mx * cs;
  float sm0 = 0.f, sm1 = 0.f, sm2 = 0.f, sm3 = 0.f;
  for (int c0 = c_begin; c0 < c_end; c0 += c_step) {
    float e[32];
    if (c0 <= warp_first) {
      uint32_t r[32];
      umma::tmem_ld32(tmem_base + lane_base + c0, r);
      umma::tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        float ex;
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ex) : "f"(fmaf(__uint_as_float(r[i]), cs, nmx)));
        e[i] = ex;
      }
      if (c0 == warp_first) {
#pragma unroll
        for (int i = 0; i < 32; ++i) e[i] = i <= lane ? e[i] : 0.f;
      }
#pragma unroll
      for (int i = 0; i < 32; i += 4) { sm0 += e[i]; sm1 += e[i + 1]; sm2 += e[i + 2]; sm3 += e[i + 3]; }
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i) e[i] = 0.f;
    }
    if constexpr (TS) {
      // P stays in tensor memory: the 32 weights of this chunk, packed two per column, go to the 16 columns
      // half HK + (c0 - half HK) / 2 — inside the S columns this very thread has already consumed (its own chunks
      // up to the current one), so neither thread of a row can overwrite a score the other still has to read
      uint32_t pk[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) pk[i] = at_pack2(e[2 * i], e[2 * i + 1], FMT);
      at_tmem_st16(tmem_base + lane_base + (uint32_t)(half * HK + ((c0 - half * HK) >> 1)), pk);
    } else {
      uint8_t* slab = sP + (c0 >> 6) * (AT_M * 128);
      const uint32_t ch0 = (uint32_t)(c0 & 63) >> 3;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint4 w;
        w.x = at_pack2(e[8 * c], e[8 * c + 1], FMT); w.y = at_pack2(e[8 * c + 2], e[8 * c + 3], FMT);
        w.z = at_pack2(e[8 * c + 4], e[8 * c + 5], FMT); w.w = at_pack2(e[8 * c + 6], e[8 * c + 7], FMT);
        *reinterpret_cast<uint4*>(slab + umma::sw128_offset(row, ch0 + c)) = w;
      }
    }
  }
  float sum = (sm0 + sm1) + (sm2 + sm3);
  s_sum[half][row] = sum;
  if constexpr (TS) umma::tmem_st_wait();
  else umma::fence_async_smem();
  umma::tc_fence_before_sync();
  __syncthreads();
  AT_MARK(3);   // weights written
  sum += s_sum[half ^ 1][row];

  // ---- O = P V ----
  if (tid == 0) {
    umma::mbar_wait(&bar_v, phase);
    umma::tc_fence_after_sync();
    // B = V as loaded (keys x dims): MN-major, 128-byte rows of 64 dims, 8-key groups 1024 bytes apart;
    // one K = 16 step consumes two such groups -> 2048 bytes per step
    const uint32_t idesc = umma::instr_desc_16bit(AT_M, AT_DH, FMT) | (1u << 16);
    const uint32_t pP = umma::smem_u32(sP), pV = umma::smem_u32(sV);
    for (int ks = 0; ks < kvp / 16; ++ks) {
      const uint64_t dv = umma::smem_desc_sw128(pV + ks * 2048);
      if constexpr (TS) {
        // 16 keys = 8 packed columns of P; keys [HK, 2 HK) start at column HK.  O accumulates into the 64 columns
        // [HK / 2, HK / 2 + 64): scores of the first thread's later chunks, all consumed
        const int k0 = 16 * ks;
        const uint32_t pa = tmem_base + (uint32_t)(k0 < HK ? k0 >> 1 : HK + ((k0 - HK) >> 1));
        at_mma_ts(tmem_base + o_col, pa, dv, idesc, ks > 0);
      } else {
        const uint64_t dp = umma::smem_desc_sw128(pP + (ks >> 2) * (AT_M * 128) + (ks & 3) * 32);
        umma::mma_bf16_ss(tmem_base, dp, dv, idesc, ks > 0);
      }
    }
    umma::mma_commit(&bar_o);
  }
  umma::mbar_wait(&bar_o, phase);
  umma::tc_fence_after_sync();
  AT_MARK(4);   // O = P V complete
  if (tid == 0 && item + gridDim.x < a.n_items) {   // V (and, in the 128-key variant, P over Q / K) is dead: next item's tiles
    if (TS) issue_v(item + gridDim.x);
    else issue_loads(item + gridDim.x);
  }

  // ---- epilogue: thread (row, half) takes output dims [32 half, 32 half + 32) ----
  {
    uint32_t r0[32];
    umma::tmem_ld32(tmem_base + lane_base + o_col + 32 * half, r0);
    umma::tmem_ld_wait();
    if (q_ok) {
      const float inv = 1.0f / sum;
      float o[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) o[i] = __uint_as_float(r0[i]) * inv;
      if (a.site >= 0) {   // transformer.py:116-122
#pragma unroll
        for (int f4 = 0; f4 < 8; ++f4) {
          const uint32_t keep = keep4(a.ms, s, (uint32_t)tq, (uint32_t)a.site, (uint32_t)(h * (AT_DH / 4) + 8 * half + f4));
#pragma unroll
          for (int q = 0; q < 4; ++q) o[4 * f4 + q] = ((keep >> q) & 1u) ? o[4 * f4 + q] * a.ms.scale : 0.f;
        }
      }
      uint4* dst = reinterpret_cast<uint4*>(a.out + (row_base + tq) * (int64_t)a.d + h * AT_DH + 32 * half);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint4 w;
        w.x = at_pack2(o[8 * c], o[8 * c + 1], FMT); w.y = at_pack2(o[8 * c + 2], o[8 * c + 3], FMT);
        w.z = at_pack2(o[8 * c + 4], o[8 * c + 5], FMT); w.w = at_pack2(o[8 * c + 6], o[8 * c + 7], FMT);
        dst[c] = w;
      }
    }
  }
  umma::tc_fence_before_sync();
  __syncthreads();   // every O row has been read: the next item's S may overwrite the accumulator columns
  umma::tc_fence_after_sync();
  AT_MARK(5);   // output stored
#ifdef BFB_ATTN_TIMING
  if (MAXKV == 256 && tid == 0) atomicAdd(&bfb_attn_dbg[6], 1ull);
#endif
  }
  if (warp == 0) umma::tmem_dealloc(tmem_base, MAXKV);
}

typedef CUresult (*AtEncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

}  // namespace

int launch_attention_tc(const void* qkv16, void* out16, int64_t n_seq, int len, int d, int n_heads, uint32_t fmt,
                        const MaskSource& ms, int site, cudaStream_t st) {
  if (n_seq == 0) return BFB200_OK;
  BFB_REQUIRE(d == n_heads * AT_DH, BFB200_ERR_UNSUPPORTED, "tensor-core attention needs heads of width 64 (d=%d, H=%d)", d,
              n_heads);
  BFB_REQUIRE(len >= 1 && len <= AT_MAXKV, BFB200_ERR_UNSUPPORTED, "tensor-core attention takes sequences of 1..%d tokens (got %d)",
              AT_MAXKV, len);
  static AtEncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<AtEncodeTiledFn>(p);
  }
  BFB_REQUIRE(fn != nullptr, BFB200_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
  const int64_t R = n_seq * len;
  CUtensorMap map;
  const cuuint64_t dims[2] = {(cuuint64_t)(3 * d), (cuuint64_t)R};
  const cuuint64_t strides[1] = {(cuuint64_t)(3 * d) * 2};
  const cuuint32_t box[2] = {(cuuint32_t)AT_DH, (cuuint32_t)AT_M};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = fn(&map, fmt == umma::kFmtF16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                        const_cast<void*>(qkv16), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  BFB_REQUIRE(r == CUDA_SUCCESS, BFB200_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) for the QKV activation", (int)r);
  AttnArgs a;
  memset(&a, 0, sizeof(a));
  a.qkv = (const uint16_t*)qkv16; a.out = (uint16_t*)out16; a.n_rows = R;
  a.len = len; a.d = d; a.n_heads = n_heads;
  a.fmt = fmt; a.inv_sqrt_dh = 1.0f / sqrtf((float)AT_DH); a.site = site; a.ms = ms;
  a.n_items = n_seq * n_heads;
  int dev_id = 0, sms = 148;
  cudaGetDevice(&dev_id);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev_id);
  if (sms <= 0) sms = 148;
  // one launch per query tile: tile 0 attends to <= 128 keys (48 KB of shared memory, 128 TMEM columns: four CTAs
  // per SM), tile 1 to <= 256 (96 KB, 256 columns: two CTAs per SM)
  const bool f16 = fmt == umma::kFmtF16;
#define AT_LAUNCH(KV)                                                                                                   \
  do {                                                                                                                  \
    const int64_t resident = (int64_t)sms * (KV == 128 ? 3 : 2);   /* persistent CTAs: what fits on the chip at once */ \
    const int64_t grid = a.n_items < resident ? a.n_items : resident;                                                   \
    if (f16) {                                                                                                          \
      cudaFuncSetAttribute(attention_tc_kernel<KV, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)AtSmem<KV>::TOTAL); \
      attention_tc_kernel<KV, true><<<(unsigned)grid, 2 * AT_M, AtSmem<KV>::TOTAL, st>>>(map, a);                       \
    } else {                                                                                                            \
      cudaFuncSetAttribute(attention_tc_kernel<KV, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)AtSmem<KV>::TOTAL); \
      attention_tc_kernel<KV, false><<<(unsigned)grid, 2 * AT_M, AtSmem<KV>::TOTAL, st>>>(map, a);                      \
    }                                                                                                                   \
  } while (0)
  a.m_tile = 0;
  AT_LAUNCH(128);
  BFB_LAUNCH_CHECK("attention_tc_kernel<128>", st);
  if (len > AT_M) {
    a.m_tile = 1;
    AT_LAUNCH(256);
    BFB_LAUNCH_CHECK("attention_tc_kernel<256>", st);
  }
#undef AT_LAUNCH
  return BFB200_OK;
}

}  // namespace bfb

#ifdef BFB_ATTN_TIMING
extern "C" int bfb200_debug_attn_timing(unsigned long long* out8, int reset) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out8, bfb::bfb_attn_dbg, sizeof(unsigned long long) * 8);
  if (reset) {
    unsigned long long z[8] = {0};
    cudaMemcpyToSymbol(bfb::bfb_attn_dbg, z, sizeof(z));
  }
  return 0;
}
#endif

// causal multi-head attention on a packed 16-bit QKV activation: qkv (n_seq * len, 3 d) -> out (n_seq * len, d);
// fmt 0 = fp16, 1 = bf16; heads of width 64, len <= 256.
extern "C" int bfb200_attention_tc(const void* qkv, void* out, int64_t n_seq, int32_t len, int32_t d_model, int32_t n_heads,
                                   int32_t fmt, void* stream) {
  using namespace bfb;
  BFB_REQUIRE(n_seq >= 0 && len >= 1 && d_model >= 1 && n_heads >= 1, BFB200_ERR_INVALID_ARG, "bad attention shape");
  BFB_REQUIRE(fmt == 0 || fmt == 1, BFB200_ERR_INVALID_ARG, "fmt must be 0 (fp16) or 1 (bf16)");
  if (n_seq == 0) return BFB200_OK;
  BFB_DEV(qkv); BFB_DEV(out);
  MaskSource ms;
  memset(&ms, 0, sizeof(ms));
  ms.n_draws = 1;
  return launch_attention_tc(qkv, out, n_seq, len, d_model, n_heads, (uint32_t)fmt, ms, -1, (cudaStream_t)stream);
}
