// MC-batched tensor-core linear layer: C[M, N] = A[M, K] * W[N, K]^T (+ bias) (ReLU)
// nn.Linear of src/model/transformer.py (:33-35 Q/K/V, :100 W_out, :152-154 FFN, :336 vocabulary head) over the
// R = B * T_mc * len token rows of a chunk, for models outside the fused kernel's class (d_model 512, F 2048 ...).
//
// sm_100a design (the canonical Blackwell GEMM): persistent CTAs, 320 threads in three roles
//   warp 0      one lane issues 2-D TMA loads (cp.async.bulk.tensor, 128-byte swizzle) of the 128 x 64 A tile and
//               the BN x 64 W tile into a 4-stage shared-memory ring, completion on `full` mbarriers;
//   warp 1      one lane issues tcgen05.mma.cta_group::1.kind::f16 (M = 128, N = BN, K = 16 x 4 per stage) into one of
//               two TMEM accumulator stages; tcgen05.commit releases the ring slot (`empty`) and, after the last
//               K block, hands the accumulator to the epilogue (`acc_full`);
//   warps 2-9   epilogue: thread = accumulator row (two warps per TMEM lane quarter, half of the columns each); tcgen05.ld 32 columns at a time, + bias,
//               ReLU, convert, write a 32-row x 128-byte slab into swizzled shared memory and hand it to a 2-D TMA
//               store (cp.async.bulk.tensor ... global.shared::cta), double-buffered per warp, so every global
//               write is a full 128-byte line and M / N tails are clipped by the copy engine; then `acc_empty`.
// Operands are 16-bit (fp16 or bf16, instruction-descriptor format bit), accumulation fp32.  M / N / K tails are
// zero-filled by TMA and masked in the epilogue.  Tiles are walked N-fastest so an A row block stays in L2 while
// all of its N tiles are produced.
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"
#include "kernels.cuh"
#include "umma.cuh"

namespace bfb {

namespace {

constexpr int GM = 128;       // tile rows (UMMA M)
constexpr int GK = 64;        // K per stage: one 128-byte swizzle row of 16-bit elements
constexpr int GACC = 2;       // TMEM accumulator stages
constexpr int GEPI_WARPS = 8;   // two epilogue warps per TMEM lane quarter, half of the tile columns each
constexpr int GTHREADS = 64 + 32 * GEPI_WARPS;

__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(umma::smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(umma::smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(map)),
               "r"(umma::smem_u32(smem_src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tma_store_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(umma::smem_u32(bar)) : "memory");
}

struct GemmArgs {
  int64_t M;
  int N, K;
  const float* bias;       // (N) or nullptr
  void* out16;             // (M, N) 16-bit or nullptr
  float* out32;            // (M, N) fp32 or nullptr
  int relu;
  uint32_t fmt;            // umma::kFmtF16 / kFmtBF16
  int64_t ldc;             // output row pitch in elements (= N)
  // fused epilogues of the tensor-core transformer path (launch_gemm_tc_ex):
  int resid;               // two-SM kernel: out16 = acc (+ bias) + R16, a 16-bit matrix shaped like the output whose tiles the
                           // TMA producer streams through the operand ring (map_r)
  RowStats* stats_out;     // with resid: per-row (sum, sum of squares) of the fp32 result, accumulated atomically
  const RowStats* stats_in;  // row-affine mode: out = rstd_r * acc + col_c[n] (weight rows centred, so acc carries no mean term), rstd from the
  const float* col_s;      //   (sum, sum of squares) of a row of `stats_width` values — LayerNorm folded into this product
  const float* col_c;
  float stats_inv_width, stats_eps;
  int vec_ok;              // bias / col_s / col_c are 16-byte aligned: full chunks read them with 128-bit loads
};

// Compile-time epilogue specialisations of the two-SM kernel: the launcher instantiates the combinations the
// transformer path uses (so their epilogues carry no dead branches: the K = 512 products are epilogue-paced) and
// falls back to EPI_DYNAMIC, which consults the runtime fields of GemmArgs for everything.
constexpr uint32_t EPI_BIAS = 1u, EPI_RELU = 2u, EPI_AFFINE = 4u, EPI_RESID = 8u, EPI_STATS = 16u, EPI_F16 = 32u,
                   EPI_OUT16_ONLY = 64u, EPI_DYNAMIC = 0x80000000u;
template <uint32_t F> struct Epi {
  static constexpr bool dyn = (F & EPI_DYNAMIC) != 0;
  __device__ static __forceinline__ bool bias(const GemmArgs& g) { return dyn ? g.bias != nullptr : (F & EPI_BIAS) != 0; }
  __device__ static __forceinline__ bool relu(const GemmArgs& g) { return dyn ? g.relu != 0 : (F & EPI_RELU) != 0; }
  __device__ static __forceinline__ bool affine(const GemmArgs& g) { return dyn ? g.stats_in != nullptr : (F & EPI_AFFINE) != 0; }
  __device__ static __forceinline__ bool resid(const GemmArgs& g) { return dyn ? g.resid != 0 : (F & EPI_RESID) != 0; }
  __device__ static __forceinline__ bool stats(const GemmArgs& g) { return dyn ? g.stats_out != nullptr : (F & EPI_STATS) != 0; }
  __device__ static __forceinline__ bool f16(const GemmArgs& g) { return dyn ? g.fmt == umma::kFmtF16 : (F & EPI_F16) != 0; }
  __device__ static __forceinline__ bool out32(const GemmArgs& g) { return (dyn || !(F & EPI_OUT16_ONLY)) ? g.out32 != nullptr : false; }
  __device__ static __forceinline__ bool out16(const GemmArgs& g) { return (dyn || !(F & EPI_OUT16_ONLY)) ? g.out16 != nullptr : true; }
};

// The arithmetic of one 32-column chunk of one accumulator row.  FULL: all 32 columns are inside the matrix and the
// per-column vectors are 16-byte aligned (128-bit broadcast loads, no guards).
//   plain     : v = acc (+ bias) (ReLU)
//   affine    : v = rstd acc + col_c (ReLU)                           LayerNorm folded into this product (centred weight rows)
//   residual  : v += R16 row (read from the ring slot TMA filled), and the row's running sum / sum of squares
template <bool FULL, uint32_t F>
__device__ __forceinline__ void chunk_math(const GemmArgs& g, const uint32_t (&acc)[32], float (&v)[32], int n0, float r_mean,
                                           float r_rstd, const uint8_t* box, int resid_row, int half, float& st_sum,
                                           float& st_sq) {
  auto load32 = [&](const float* src, float (&dst)[32]) {
    if (FULL) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(src + n0) + i);
        dst[4 * i] = t.x; dst[4 * i + 1] = t.y; dst[4 * i + 2] = t.z; dst[4 * i + 3] = t.w;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i) dst[i] = n0 + i < g.N ? __ldg(src + n0 + i) : 0.f;
    }
  };
  if (Epi<F>::affine(g)) {
    load32(g.col_c, v);   // the weight rows are centred (fold_ln_pack_kernel): acc is already (v - mean) (W g)^T
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = fmaf(r_rstd, __uint_as_float(acc[i]), v[i]);
  } else if (Epi<F>::bias(g)) {
    load32(g.bias, v);
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] += __uint_as_float(acc[i]);
  } else {
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(acc[i]);
  }
  if (Epi<F>::relu(g)) {
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
  }
  if (Epi<F>::resid(g)) {
    // chunk = bytes [64 half, +64) of row `resid_row` of a swizzled box of 128 rows x 64 columns
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const uint4 w = *reinterpret_cast<const uint4*>(box + umma::sw128_offset((uint32_t)resid_row, 4 * half + c));
      const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (Epi<F>::f16(g)) {   // v += x: one mixed-precision FMA against a packed 1.0 per value
          v[8 * c + 2 * k] = umma::fh_ll(ww[k], 0x3C003C00u, v[8 * c + 2 * k]);
          v[8 * c + 2 * k + 1] = umma::fh_hl(ww[k], 0x3C003C00u, v[8 * c + 2 * k + 1]);
        } else {
          v[8 * c + 2 * k] += umma::bf16_lo(ww[k]);
          v[8 * c + 2 * k + 1] += umma::bf16_hi(ww[k]);
        }
      }
    }
    if (Epi<F>::stats(g)) {
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const float x = (FULL || n0 + i < g.N) ? v[i] : 0.f;
        st_sum += x;
        st_sq = fmaf(x, x, st_sq);
      }
    }
  }
}

// LayerNorm statistics of the output row a thread owns (row-affine epilogue): fetched BEFORE the thread waits for the
// accumulator, so the global-memory latency hides behind the tile's MMAs.
template <uint32_t F>
__device__ __forceinline__ void load_row_stats(const GemmArgs& g, int row, float& r_mean, float& r_rstd) {
  r_mean = 0.f;
  r_rstd = 1.f;
  if (Epi<F>::affine(g) && row < g.M) {
    const float2 st = row_stats_load(g.stats_in + row);
    r_mean = st.x * g.stats_inv_width;
    r_rstd = rsqrtf(fmaxf(st.y * g.stats_inv_width - r_mean * r_mean, 0.f) + g.stats_eps);
  }
}

// Epilogue of one accumulator stage for one warp: columns [c_begin, c_begin + 32 * NCH) of its 32 TMEM lanes ->
// (+ bias)(ReLU) -> swizzled 32-row slabs -> 2-D TMA stores.  The tcgen05.ld of chunk i + 1 is in flight while chunk
// i is converted and staged.
template <int NCH, uint32_t F = EPI_DYNAMIC, int NSLAB = 2>
__device__ __forceinline__ void epilogue_columns(const GemmArgs& g, const CUtensorMap* map_c16, const CUtensorMap* map_c32,
                                                 uint32_t tacc, int n_base, int row0,
                                                 int c_begin, uint8_t* my_slabs, uint32_t& slab_i, int lane,
                                                 const uint8_t* resid_rows, int resid_row, float r_mean, float r_rstd) {
  constexpr uint32_t SLAB_BYTES = 32 * 128;
  float st_sum = 0.f, st_sq = 0.f;
  uint32_t r[2][32];
  if (n_base + c_begin < g.N) umma::tmem_ld32(tacc + c_begin, r[0]);
#pragma unroll
  for (int ch = 0; ch < NCH; ++ch) {
    const int c0 = c_begin + 32 * ch;
    const int n0 = n_base + c0;
    if (n0 < g.N) {   // warp-uniform: columns past N are outside the matrix
      umma::tmem_ld_wait();
      if (ch + 1 < NCH && n0 + 32 < g.N) umma::tmem_ld32(tacc + c0 + 32, r[(ch + 1) & 1]);
      float v[32];
      const uint8_t* box = resid_rows + (uint32_t)(ch >> 1) * (GM * 128u);
      if (n0 + 32 <= g.N && g.vec_ok)   // warp-uniform: the whole chunk is inside the matrix -> no per-column guards
        chunk_math<true, F>(g, r[ch & 1], v, n0, r_mean, r_rstd, box, resid_row, ch & 1, st_sum, st_sq);
      else
        chunk_math<false, F>(g, r[ch & 1], v, n0, r_mean, r_rstd, box, resid_row, ch & 1, st_sum, st_sq);
      if (Epi<F>::out32(g)) {
        uint8_t* slab = my_slabs + (slab_i % NSLAB) * SLAB_BYTES;
        tma_store_wait_read<NSLAB - 1>();   // the slab used NSLAB stores ago has been read by the copy engine
        __syncwarp();
#pragma unroll
        for (int c = 0; c < 8; ++c)
          *reinterpret_cast<float4*>(slab + umma::sw128_offset(lane, c)) =
              make_float4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
        umma::fence_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_store_2d(map_c32, slab, n0, row0);
          tma_store_commit();
        }
        ++slab_i;
      }
      if (Epi<F>::out16(g)) {
        // two 32-column chunks share one 128-byte slab row: even chunk -> bytes [0, 64), odd chunk -> [64, 128)
        uint8_t* slab = my_slabs + (slab_i % NSLAB) * SLAB_BYTES;
        const int half = ch & 1;   // c_begin is a multiple of 64
        if (half == 0) {
          tma_store_wait_read<NSLAB - 1>();
          __syncwarp();
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint4 w;
          const int i = 8 * c;
          if (g.fmt == umma::kFmtF16) {
            w.x = umma::pack_f16x2(v[i], v[i + 1]); w.y = umma::pack_f16x2(v[i + 2], v[i + 3]);
            w.z = umma::pack_f16x2(v[i + 4], v[i + 5]); w.w = umma::pack_f16x2(v[i + 6], v[i + 7]);
          } else {
            w.x = umma::pack_bf16x2(v[i], v[i + 1]); w.y = umma::pack_bf16x2(v[i + 2], v[i + 3]);
            w.z = umma::pack_bf16x2(v[i + 4], v[i + 5]); w.w = umma::pack_bf16x2(v[i + 6], v[i + 7]);
          }
          *reinterpret_cast<uint4*>(slab + umma::sw128_offset(lane, 4 * half + c)) = w;
        }
        const bool last_chunk = half == 1 || n0 + 32 >= g.N || ch + 1 >= NCH;
        if (last_chunk) {
          umma::fence_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(map_c16, slab, n0 - 32 * half, row0);
            tma_store_commit();
          }
          ++slab_i;
        }
      }
    }
  }
  if (Epi<F>::resid(g) && Epi<F>::stats(g) && row0 + lane < g.M) {
    row_stats_add(&g.stats_out[row0 + lane], st_sum, st_sq);
  }
}

template <int BN, int GSTAGES>
__global__ void __launch_bounds__(GTHREADS, 1) gemm_tc_kernel(const __grid_constant__ CUtensorMap map_a,
                                                              const __grid_constant__ CUtensorMap map_w,
                                                              const __grid_constant__ CUtensorMap map_c16,
                                                              const __grid_constant__ CUtensorMap map_c32,
                                                              const GemmArgs g) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (umma::smem_u32(smem_raw) & 1023u)) & 1023u);
  constexpr uint32_t A_BYTES = GM * GK * 2, W_BYTES = BN * GK * 2, STAGE_BYTES = A_BYTES + W_BYTES;
  constexpr uint32_t SLAB_BYTES = 32 * 128;   // 32 rows x 128 bytes, one TMA-store box
  uint8_t* slabs = smem + GSTAGES * STAGE_BYTES;            // GEPI_WARPS x 2 slabs
  __shared__ uint64_t full_bar[GSTAGES], empty_bar[GSTAGES], acc_full[GACC], acc_empty[GACC];
  __shared__ uint32_t tmem_slot;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t tiles_m = (g.M + GM - 1) / GM;
  const int tiles_n = (g.N + BN - 1) / BN;
  const int64_t n_tiles = tiles_m * tiles_n;
  const int k_blocks = (g.K + GK - 1) / GK;

  if (warp == 0) {
    umma::tmem_alloc(&tmem_slot, GACC * BN);   // 256 or 512 columns: a power of two
    umma::tmem_relinquish();
  }
  if (threadIdx.x == 32) {
    for (int i = 0; i < GSTAGES; ++i) {
      umma::mbar_init(&full_bar[i], 1);
      umma::mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < GACC; ++i) {
      umma::mbar_init(&acc_full[i], 1);
      umma::mbar_init(&acc_empty[i], GEPI_WARPS);   // one arrival per epilogue warp
    }
    umma::fence_mbar_init();
  }
  umma::tc_fence_before_sync();
  __syncthreads();
  umma::tc_fence_after_sync();
  const uint32_t tmem_base = tmem_slot;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      uint32_t it = 0;
      for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t mb = tile / tiles_n;
        const int nb = (int)(tile - mb * tiles_n);
        for (int kb = 0; kb < k_blocks; ++kb, ++it) {
          const uint32_t s = it % GSTAGES, ph = (it / GSTAGES) & 1u;
          umma::mbar_wait(&empty_bar[s], ph ^ 1u);
          uint8_t* sa = smem + s * STAGE_BYTES;
          umma::mbar_expect_tx(&full_bar[s], STAGE_BYTES);
          tma_load_2d(sa, &map_a, kb * GK, (int)(mb * GM), &full_bar[s]);
          tma_load_2d(sa + A_BYTES, &map_w, kb * GK, nb * BN, &full_bar[s]);
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      const uint32_t idesc = umma::instr_desc_16bit(GM, BN, g.fmt);
      uint32_t it = 0, tcount = 0;
      for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tcount) {
        const uint32_t as = tcount % GACC, aph = (tcount / GACC) & 1u;
        umma::mbar_wait(&acc_empty[as], aph ^ 1u);   // epilogue has drained this accumulator stage
        umma::tc_fence_after_sync();
        const uint32_t d_tmem = tmem_base + as * BN;
        for (int kb = 0; kb < k_blocks; ++kb, ++it) {
          const uint32_t s = it % GSTAGES, ph = (it / GSTAGES) & 1u;
          umma::mbar_wait(&full_bar[s], ph);
          umma::tc_fence_after_sync();
          const uint32_t sa = umma::smem_u32(smem + s * STAGE_BYTES);
          const uint64_t da = umma::smem_desc_sw128(sa), dw = umma::smem_desc_sw128(sa + A_BYTES);
#pragma unroll
          for (int k = 0; k < GK / 16; ++k) umma::mma_bf16_ss(d_tmem, da + 2 * k, dw + 2 * k, idesc, (kb | k) != 0);
          umma::mma_commit(&empty_bar[s]);          // ring slot free once these MMAs have read it
        }
        umma::mma_commit(&acc_full[as]);            // accumulator complete
      }
    }
  } else {
    // ===== epilogue: warps 2..9; warp w drains TMEM lane quarter w & 3, column half (w - 2) >> 2 =====
    const int q = warp & 3, chalf = (warp - 2) >> 2;
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    uint8_t* my_slabs = slabs + (uint32_t)(warp - 2) * 2u * SLAB_BYTES;
    uint32_t tcount = 0, slab_i = 0;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tcount) {
      const int64_t mb = tile / tiles_n;
      const int nb = (int)(tile - mb * tiles_n);
      const uint32_t as = tcount % GACC, aph = (tcount / GACC) & 1u;
      const int row0 = (int)(mb * GM) + q * 32;
      float r_mean, r_rstd;
      load_row_stats<EPI_DYNAMIC>(g, row0 + lane, r_mean, r_rstd);
      umma::mbar_wait(&acc_full[as], aph);
      umma::tc_fence_after_sync();
      epilogue_columns<BN / 64>(g, &map_c16, &map_c32, tmem_base + lane_base + as * BN, nb * BN, row0, chalf * (BN / 2),
                                my_slabs, slab_i, lane, nullptr, 0, r_mean, r_rstd);
      umma::tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[as]);
    }
    tma_store_wait_read<0>();
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // stores complete before the CTA retires
  }

  umma::tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tmem_base, GACC * BN);
}

// ------------------------------------------------------------------------------------------
// Two-SM variant (N > 128): a cluster of two CTAs owns a 256 x 256 output tile and the even CTA issues
// tcgen05.mma.cta_group::2 (M = 256) for the pair.  Each CTA stages only ITS 128 rows of A and ITS 128 of the 256
// W rows per K block — 32 KB instead of 48 KB per 128 x 256 x 64 of work — which is what lifts the kernel off the
// L2 -> SM bandwidth ceiling the one-SM tiling sits on (128 x 256 tiles with K = 512 need ~11.5 TB/s of L2 reads at
// 1 PFLOP/s; the L2 delivers ~12 TB/s).  Protocol (all mbarriers count 1 unless noted):
//   full[s]      in the LEADER's shared memory; the leader's producer expects the bytes of BOTH CTAs' copies, the
//                follower's TMA loads complete on it (cp.async.bulk.tensor ... .cta_group::2, remote mbarrier);
//   empty[s]     one per CTA; the leader's MMA warp releases both with one multicast tcgen05.commit;
//   acc_full[a]  one per CTA, same multicast commit after the last K block;
//   acc_empty[a] in the leader (count 16): every epilogue warp of both CTAs arrives (the follower's remotely).
// Each CTA's epilogue drains its own 128 TMEM lanes exactly like the one-SM kernel.
// ------------------------------------------------------------------------------------------
constexpr int G2_BN = 256;          // tile columns (UMMA N); each CTA stages G2_BN / 2 rows of W
constexpr int G2_STAGES = 6;
constexpr int G2_SLABS = 1;        // store slabs per epilogue warp

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t mapa_rank(uint32_t smem_addr, uint32_t rank) {   // shared::cta -> shared::cluster of `rank`
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// 2-D TMA load into this CTA's shared memory, completion bytes on an mbarrier that may live in the peer CTA
__device__ __forceinline__ void tma_load_2d_2sm(void* smem_dst, const CUtensorMap* map, int c0, int c1, uint32_t bar_cluster_addr) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(umma::smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void mma_f16_ss_2sm(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// all MMAs issued so far by this thread -> arrive on the barrier at this shared-memory offset in BOTH CTAs
__device__ __forceinline__ void mma_commit_2sm(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   umma::smem_u32(bar)),
               "h"((uint16_t)3)
               : "memory");
}

#ifdef BFB_GEMM_TIMING
// debug build only (make EXTRA=-DBFB_GEMM_TIMING, tools/gemm_timing.py): where the two-SM kernel's roles wait
__device__ unsigned long long bfb_gemm_dbg[16];
#define GT_BEGIN() const long long _gt0 = clock64()
#define GT_END(slot) gdbg[slot] += clock64() - _gt0
#else
#define GT_BEGIN() do { } while (0)
#define GT_END(slot) do { } while (0)
#endif

template <uint32_t F>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GTHREADS, 1)
    gemm2_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w,
                    const __grid_constant__ CUtensorMap map_c16, const __grid_constant__ CUtensorMap map_c32,
                    const __grid_constant__ CUtensorMap map_r, const GemmArgs g) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (umma::smem_u32(smem_raw) & 1023u)) & 1023u);
  constexpr int BN = G2_BN, WROWS = G2_BN / 2, GSTAGES = G2_STAGES;
  constexpr uint32_t A_BYTES = GM * GK * 2, W_BYTES = WROWS * GK * 2, STAGE_BYTES = A_BYTES + W_BYTES;
  constexpr uint32_t SLAB_BYTES = 32 * 128;
  uint8_t* slabs = smem + GSTAGES * STAGE_BYTES;
  __shared__ uint64_t full_bar[GSTAGES], empty_bar[GSTAGES], rfull_bar[GSTAGES], acc_full[GACC], acc_empty[GACC];
  __shared__ uint32_t tmem_slot;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int64_t cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;
  const int64_t tiles_m = (g.M + 2 * GM - 1) / (2 * GM);
  const int tiles_n = (g.N + BN - 1) / BN;
  const int64_t n_tiles = tiles_m * tiles_n;
  const int k_blocks = (g.K + GK - 1) / GK;

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(umma::smem_u32(&tmem_slot)),
                 "r"((uint32_t)(GACC * BN))
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  if (threadIdx.x == 32) {
    for (int i = 0; i < GSTAGES; ++i) {
      umma::mbar_init(&full_bar[i], 1);
      umma::mbar_init(&empty_bar[i], 1);
      umma::mbar_init(&rfull_bar[i], 1);
    }
    for (int i = 0; i < GACC; ++i) {
      umma::mbar_init(&acc_full[i], 1);
      umma::mbar_init(&acc_empty[i], 2 * GEPI_WARPS);   // every epilogue warp of both CTAs
    }
    umma::fence_mbar_init();
  }
  umma::tc_fence_before_sync();
  __syncthreads();
  cluster_sync_all();   // the peer's barriers exist before anything signals them
  umma::tc_fence_after_sync();
  const uint32_t tmem_base = tmem_slot;
#ifdef BFB_GEMM_TIMING
  long long gdbg[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  const long long g_begin = clock64();
#endif

  if (warp == 0) {
    // ===== TMA producer (both CTAs): own 128 rows of A, own 128 rows of the W tile =====
    if (lane == 0) {
      uint32_t it = 0;
      for (int64_t tile = cluster_id; tile < n_tiles; tile += n_clusters) {
        const int64_t mb = tile / tiles_n;
        const int nb = (int)(tile - mb * tiles_n);
        for (int kb = 0; kb < k_blocks; ++kb, ++it) {
          const uint32_t s = it % GSTAGES, ph = (it / GSTAGES) & 1u;
          { GT_BEGIN(); umma::mbar_wait(&empty_bar[s], ph ^ 1u); GT_END(0); }
          uint8_t* sa = smem + s * STAGE_BYTES;
          if (rank == 0) umma::mbar_expect_tx(&full_bar[s], 2 * STAGE_BYTES);
          const uint32_t bar = mapa_rank(umma::smem_u32(&full_bar[s]), 0);
          tma_load_2d_2sm(sa, &map_a, kb * GK, (int)(mb * 2 * GM + rank * GM), bar);
          tma_load_2d_2sm(sa + A_BYTES, &map_w, kb * GK, nb * BN + (int)rank * WROWS, bar);
        }
        if (Epi<F>::resid(g)) {
          // the tile's residual rows of this CTA ride the same ring: one slot per 128 output columns (two swizzled boxes of
          // 128 rows x 64 columns), completion on this CTA's own rfull barrier; the epilogue warps hand the slot back
          const int n_rs = nb * BN + 128 < g.N ? 2 : 1;
          for (int rs = 0; rs < n_rs; ++rs, ++it) {
            const uint32_t s = it % GSTAGES, ph = (it / GSTAGES) & 1u;
            umma::mbar_wait(&empty_bar[s], ph ^ 1u);
            uint8_t* sr = smem + s * STAGE_BYTES;
            umma::mbar_expect_tx(&rfull_bar[s], 2u * GM * 128u);
            tma_load_2d(sr, &map_r, nb * BN + rs * 128, (int)(mb * 2 * GM + rank * GM), &rfull_bar[s]);
            tma_load_2d(sr + GM * 128, &map_r, nb * BN + rs * 128 + 64, (int)(mb * 2 * GM + rank * GM), &rfull_bar[s]);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: the leader CTA drives both tensor cores =====
    if (lane == 0 && rank == 0) {
      const uint32_t idesc = umma::instr_desc_16bit(2 * GM, BN, g.fmt);
      // full_bar[s] completes a phase only when slot s carries operands (residual slots complete rfull_bar instead),
      // so its parity is tracked per slot rather than derived from the ring position
      uint32_t it = 0, tcount = 0, fphase = 0;
      for (int64_t tile = cluster_id; tile < n_tiles; tile += n_clusters, ++tcount) {
        const uint32_t as = tcount % GACC, aph = (tcount / GACC) & 1u;
        { GT_BEGIN(); umma::mbar_wait(&acc_empty[as], aph ^ 1u); GT_END(1); }   // both epilogues have drained this stage
        umma::tc_fence_after_sync();
        const uint32_t d_tmem = tmem_base + as * BN;
        for (int kb = 0; kb < k_blocks; ++kb, ++it) {
          const uint32_t s = it % GSTAGES;
          { GT_BEGIN(); umma::mbar_wait(&full_bar[s], (fphase >> s) & 1u); GT_END(2); }
          fphase ^= 1u << s;
          umma::tc_fence_after_sync();
          const uint32_t sa = umma::smem_u32(smem + s * STAGE_BYTES);
          const uint64_t da = umma::smem_desc_sw128(sa), dw = umma::smem_desc_sw128(sa + A_BYTES);
#pragma unroll
          for (int k = 0; k < GK / 16; ++k) mma_f16_ss_2sm(d_tmem, da + 2 * k, dw + 2 * k, idesc, (kb | k) != 0);
          mma_commit_2sm(&empty_bar[s]);
        }
        mma_commit_2sm(&acc_full[as]);
        if (Epi<F>::resid(g)) {   // the residual slots of this tile are not MMA operands
          const int64_t mb_ = tile / tiles_n;
          const int nb_ = (int)(tile - mb_ * tiles_n);
          it += nb_ * BN + 128 < g.N ? 2u : 1u;
        }
      }
    }
  } else {
    // ===== epilogue (both CTAs): warps 2..9 over this CTA's 128 rows; lane quarter w & 3, column half (w - 2) >> 2 =====
    const int q = warp & 3, chalf = (warp - 2) >> 2;
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    uint8_t* my_slabs = slabs + (uint32_t)(warp - 2) * (uint32_t)G2_SLABS * SLAB_BYTES;
    const uint32_t acc_empty_leader0 = mapa_rank(umma::smem_u32(&acc_empty[0]), 0);
    uint32_t tcount = 0, slab_i = 0, ring_it = 0, rphase = 0;
    for (int64_t tile = cluster_id; tile < n_tiles; tile += n_clusters, ++tcount) {
      const int64_t mb = tile / tiles_n;
      const int nb = (int)(tile - mb * tiles_n);
      const uint32_t as = tcount % GACC, aph = (tcount / GACC) & 1u;
      const int row0 = (int)(mb * 2 * GM) + (int)rank * GM + q * 32;
      float r_mean, r_rstd;
      load_row_stats<F>(g, row0 + lane, r_mean, r_rstd);
      { GT_BEGIN(); umma::mbar_wait(&acc_full[as], aph); GT_END(3); }
      umma::tc_fence_after_sync();
#ifdef BFB_GEMM_TIMING
      const long long _ge0 = clock64();
#endif
      const uint8_t* resid_rows = nullptr;
      uint32_t rslot = 0;
      const int n_rs = nb * BN + 128 < g.N ? 2 : 1;
      if (Epi<F>::resid(g)) {
        ring_it += (uint32_t)k_blocks;                 // this tile's operand slots
        for (int rs = 0; rs < n_rs; ++rs) {            // every warp tracks the phase of every residual slot
          const uint32_t sl = (ring_it + (uint32_t)rs) % GSTAGES;
          if (rs == chalf) {
            rslot = sl;
            umma::mbar_wait(&rfull_bar[sl], (rphase >> sl) & 1u);
            resid_rows = smem + sl * STAGE_BYTES;
          }
          rphase ^= 1u << sl;
        }
      }
      epilogue_columns<BN / 64, F, G2_SLABS>(g, &map_c16, &map_c32, tmem_base + lane_base + as * BN, nb * BN, row0, chalf * (BN / 2),
                                   my_slabs, slab_i, lane, resid_rows, q * 32 + lane, r_mean, r_rstd);
      if (Epi<F>::resid(g)) {
        if (chalf < n_rs) {   // the four warps that read this slot hand it back to the producer
          asm volatile("bar.sync %0, 128;" ::"r"(2 + chalf) : "memory");
          if (q == 0 && lane == 0) mbar_arrive(&empty_bar[rslot]);
        }
        ring_it += (uint32_t)n_rs;
      }
#ifdef BFB_GEMM_TIMING
      gdbg[4] += clock64() - _ge0;
      gdbg[5] += 1;
#endif
      umma::tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(acc_empty_leader0 + as * (uint32_t)sizeof(uint64_t));
    }
    tma_store_wait_read<0>();
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }

#ifdef BFB_GEMM_TIMING
  if (lane == 0) {
    const unsigned long long total = (unsigned long long)(clock64() - g_begin);
    if (warp == 0) { atomicAdd(&bfb_gemm_dbg[0], (unsigned long long)gdbg[0]); atomicAdd(&bfb_gemm_dbg[1], total); }
    if (warp == 1 && rank == 0) {
      atomicAdd(&bfb_gemm_dbg[2], (unsigned long long)gdbg[1]); atomicAdd(&bfb_gemm_dbg[3], (unsigned long long)gdbg[2]);
      atomicAdd(&bfb_gemm_dbg[4], total);
    }
    if (warp == 2) {
      atomicAdd(&bfb_gemm_dbg[5], (unsigned long long)gdbg[3]); atomicAdd(&bfb_gemm_dbg[6], (unsigned long long)gdbg[4]);
      atomicAdd(&bfb_gemm_dbg[7], (unsigned long long)gdbg[5]); atomicAdd(&bfb_gemm_dbg[8], total);
    }
  }
#endif
  umma::tc_fence_before_sync();
  __syncthreads();
  cluster_sync_all();   // nobody retires while its peer may still signal its barriers or read its shared memory
  if (warp == 0)
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)(GACC * BN)) : "memory");
}

// ------------------------------------------------------------------------------------------
// host: tensor maps through the driver entry point (no link-time dependency on libcuda)
// ------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// 2-D map over a (rows, cols) matrix of 16-bit (elem_bytes 2) or fp32 (4) elements with row pitch `pitch_elems`;
// box = (128 bytes of columns) x box_rows, 128-byte swizzle
int make_map(CUtensorMap* map, const void* base, int64_t rows, int K, int64_t pitch_elems, int box_rows, uint32_t fmt,
             int elem_bytes = 2) {
  EncodeTiledFn fn = encode_tiled_fn();
  BFB_REQUIRE(fn != nullptr, BFB200_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
  const cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)pitch_elems * (cuuint64_t)elem_bytes};
  const cuuint32_t box[2] = {(cuuint32_t)(128 / elem_bytes), (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUtensorMapDataType dt = elem_bytes == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32
                                 : fmt == umma::kFmtF16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  const CUresult r = fn(map, dt, 2,
                        const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  BFB_REQUIRE(r == CUDA_SUCCESS, BFB200_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) for rows=%lld K=%d pitch=%lld",
              (int)r, (long long)rows, K, (long long)pitch_elems);
  return BFB200_OK;
}

}  // namespace

int launch_gemm_tc(const void* A, int64_t a_pitch, const void* W, const float* bias, void* out16, float* out32,
                   int64_t ldc, int64_t M, int N, int K, int relu, uint32_t fmt, cudaStream_t st) {
  return launch_gemm_tc_ex(A, a_pitch, W, bias, out16, out32, ldc, M, N, K, relu, fmt, GemmFusion(), st);
}

int launch_gemm_tc_ex(const void* A, int64_t a_pitch, const void* W, const float* bias, void* out16, float* out32,
                      int64_t ldc, int64_t M, int N, int K, int relu, uint32_t fmt, const GemmFusion& fu, cudaStream_t st) {
  if (M == 0 || N == 0) return BFB200_OK;
  BFB_REQUIRE(fu.resid16 == nullptr || (fu.resid16 != out16 && (reinterpret_cast<uintptr_t>(fu.resid16) & 15u) == 0 &&
                                        (fu.resid_pitch & 7) == 0 && (N & 7) == 0),
              BFB200_ERR_INVALID_ARG, "a residual operand needs its own 16-byte aligned buffer, a row pitch and an N that are "
              "multiples of 8");
  BFB_REQUIRE(fu.stats_in == nullptr || (fu.col_s != nullptr && fu.col_c != nullptr && fu.stats_width > 0),
              BFB200_ERR_INVALID_ARG, "row-affine epilogue needs col_s, col_c and stats_width");
  BFB_REQUIRE((K & 7) == 0 && (a_pitch & 7) == 0, BFB200_ERR_UNSUPPORTED,
              "tensor-core GEMM needs K (%d) and the A row pitch (%lld) to be multiples of 8", K, (long long)a_pitch);
  BFB_REQUIRE((reinterpret_cast<uintptr_t>(A) & 15u) == 0 && (reinterpret_cast<uintptr_t>(W) & 15u) == 0,
              BFB200_ERR_INVALID_ARG, "GEMM operands must be 16-byte aligned");
  BFB_REQUIRE(ldc >= N, BFB200_ERR_INVALID_ARG, "output pitch %lld < N = %d", (long long)ldc, N);
  BFB_REQUIRE(out16 == nullptr || ((ldc & 7) == 0 && (reinterpret_cast<uintptr_t>(out16) & 15u) == 0),
              BFB200_ERR_UNSUPPORTED, "16-bit GEMM output needs a row pitch (%lld) that is a multiple of 8 elements and a "
              "16-byte aligned buffer", (long long)ldc);
  BFB_REQUIRE(out32 == nullptr || ((ldc & 3) == 0 && (reinterpret_cast<uintptr_t>(out32) & 15u) == 0),
              BFB200_ERR_UNSUPPORTED, "fp32 GEMM output needs a row pitch (%lld) that is a multiple of 4 elements and a "
              "16-byte aligned buffer", (long long)ldc);
  int sms = 0, cur_dev = 0;   // per call: the attribute belongs to the CURRENT device (a process may drive several)
  cudaGetDevice(&cur_dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, cur_dev);
  if (sms <= 0) sms = 148;
  static const bool one_sm_only = getenv("BFB200_GEMM_1SM") != nullptr;   // A/B switch for measurements
  const bool two_sm = N > 128 && !one_sm_only;
  const int BN = N > 128 ? 256 : 128;
  CUtensorMap map_a, map_w, map_c16, map_c32;
  BFB_CALL(make_map(&map_a, A, M, K, a_pitch, GM, fmt));
  BFB_CALL(make_map(&map_w, W, N, K, K, two_sm ? G2_BN / 2 : BN, fmt));
  // outputs: 32-row x 128-byte store boxes; an absent output reuses the other's map (never dereferenced)
  if (out16 != nullptr) BFB_CALL(make_map(&map_c16, out16, M, N, ldc, 32, fmt, 2));
  if (out32 != nullptr) BFB_CALL(make_map(&map_c32, out32, M, N, ldc, 32, fmt, 4));
  if (out16 == nullptr) map_c16 = map_c32;
  if (out32 == nullptr) map_c32 = map_c16;
  GemmArgs g;
  g.M = M; g.N = N; g.K = K; g.bias = bias; g.out16 = out16; g.out32 = out32; g.relu = relu; g.fmt = fmt; g.ldc = ldc;
  g.resid = fu.resid16 != nullptr ? 1 : 0;
  BFB_REQUIRE(!g.resid || two_sm, BFB200_ERR_UNSUPPORTED, "the residual epilogue is served by the two-SM kernel (N > 128)");
  CUtensorMap map_r = map_c16;
  if (g.resid) BFB_CALL(make_map(&map_r, fu.resid16, M, N, fu.resid_pitch, GM, fmt, 2));
  g.stats_out = fu.stats_out; g.stats_in = fu.stats_in; g.col_s = fu.col_s; g.col_c = fu.col_c;
  g.stats_inv_width = fu.stats_width > 0 ? 1.0f / (float)fu.stats_width : 0.f;
  g.stats_eps = fu.stats_eps;
  g.vec_ok = ((reinterpret_cast<uintptr_t>(bias) | reinterpret_cast<uintptr_t>(fu.col_s) | reinterpret_cast<uintptr_t>(fu.col_c)) & 15u) == 0;
  if (two_sm) {
    const int64_t n_tiles = ((M + 2 * GM - 1) / (2 * GM)) * ((N + G2_BN - 1) / G2_BN);
    const int64_t max_clusters = sms / 2;
    const int grid = 2 * (int)(n_tiles < max_clusters ? n_tiles : max_clusters);
    const size_t smem = 1024 + (size_t)G2_STAGES * (GM * GK * 2 + (G2_BN / 2) * GK * 2) + G2_SLABS * GEPI_WARPS * 32 * 128;
    // the specialisations the transformer path launches; anything else takes the dynamic kernel
    uint32_t want = (bias != nullptr ? EPI_BIAS : 0u) | (relu ? EPI_RELU : 0u) | (g.stats_in != nullptr ? EPI_AFFINE : 0u) |
                    (g.resid ? EPI_RESID : 0u) | (g.stats_out != nullptr ? EPI_STATS : 0u) | (fmt == umma::kFmtF16 ? EPI_F16 : 0u) |
                    ((out16 != nullptr && out32 == nullptr) ? EPI_OUT16_ONLY : 0u);
    if (g.stats_in != nullptr) want &= ~EPI_BIAS;   // the affine epilogue carries its own constant term
#define G2_LAUNCH(FLAGS)                                                                                               \
  do {                                                                                                                 \
    cudaFuncSetAttribute(gemm2_tc_kernel<FLAGS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);             \
    gemm2_tc_kernel<FLAGS><<<grid, GTHREADS, smem, st>>>(map_a, map_w, map_c16, map_c32, map_r, g);                   \
  } while (0)
    constexpr uint32_t kPlain = EPI_OUT16_ONLY, kF = EPI_F16;
    const uint32_t mode = want & ~kF;
#define G2_MODE(M)                                  \
  if (mode == (M)) {                                \
    if (want & kF) G2_LAUNCH((M) | kF);             \
    else G2_LAUNCH(M);                              \
  }
    G2_MODE(kPlain)                                              // QKV
    else G2_MODE(kPlain | EPI_RESID | EPI_STATS)                 // W_out
    else G2_MODE(kPlain | EPI_AFFINE | EPI_RELU)                 // FFN1, LayerNorm folded
    else G2_MODE(kPlain | EPI_BIAS | EPI_RESID | EPI_STATS)      // FFN2 (fp16 residual stream)
    else G2_MODE(kPlain | EPI_BIAS | EPI_RELU)                   // FFN1, unfused
    else G2_MODE(kPlain | EPI_BIAS)                              // FFN2, unfused
    else G2_LAUNCH(EPI_DYNAMIC);
#undef G2_MODE
#undef G2_LAUNCH
    BFB_LAUNCH_CHECK(g.resid ? "gemm_tc_kernel[+residual,+row stats]" : g.stats_in != nullptr ? "gemm_tc_kernel[LayerNorm folded]" : "gemm_tc_kernel", st);
    return BFB200_OK;
  }
  const int64_t n_tiles = ((M + GM - 1) / GM) * ((N + BN - 1) / BN);
  const int grid = (int)(n_tiles < sms ? n_tiles : sms);
  const int stages = BN == 256 ? 3 : 4;
  const size_t smem = 1024 + (size_t)stages * (GM * GK * 2 + BN * GK * 2) + 2 * GEPI_WARPS * 32 * 128;
  if (BN == 256) {
    cudaFuncSetAttribute(gemm_tc_kernel<256, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    gemm_tc_kernel<256, 3><<<grid, GTHREADS, smem, st>>>(map_a, map_w, map_c16, map_c32, g);
  } else {
    cudaFuncSetAttribute(gemm_tc_kernel<128, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    gemm_tc_kernel<128, 4><<<grid, GTHREADS, smem, st>>>(map_a, map_w, map_c16, map_c32, g);
  }
  BFB_LAUNCH_CHECK("gemm_tc_kernel", st);
  return BFB200_OK;
}

}  // namespace bfb

#ifdef BFB_GEMM_TIMING
extern "C" int bfb200_debug_gemm_timing(unsigned long long* out16, int reset) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out16, bfb::bfb_gemm_dbg, sizeof(unsigned long long) * 16);
  if (reset) {
    unsigned long long z[16] = {0};
    cudaMemcpyToSymbol(bfb::bfb_gemm_dbg, z, sizeof(z));
  }
  return 0;
}
#endif

// C[M, N] = A[M, K] W[N, K]^T (+ bias)(ReLU); A, W 16-bit (fmt 0 = fp16, 1 = bf16) device pointers, A rows `a_pitch`
// elements apart; out16 (same format) and/or out32 (fp32) with row pitch `ldc` elements, either may be NULL.
extern "C" int bfb200_linear_tc(const void* A, int64_t a_pitch, const void* W, const float* bias, void* out16,
                                float* out32, int64_t ldc, int64_t M, int32_t N, int32_t K, int32_t relu, int32_t fmt,
                                void* stream) {
  using namespace bfb;
  BFB_REQUIRE(M >= 0 && N >= 1 && K >= 8, BFB200_ERR_INVALID_ARG, "bad GEMM shape (%lld, %d, %d)", (long long)M, N, K);
  BFB_REQUIRE(fmt == 0 || fmt == 1, BFB200_ERR_INVALID_ARG, "fmt must be 0 (fp16) or 1 (bf16)");
  BFB_REQUIRE(out16 != nullptr || out32 != nullptr, BFB200_ERR_INVALID_ARG, "no output requested");
  if (M == 0) return BFB200_OK;
  BFB_DEV(A); BFB_DEV(W);
  BFB_DEV_OPT(bias); BFB_DEV_OPT(out16); BFB_DEV_OPT(out32);
  return launch_gemm_tc(A, a_pitch, W, bias, out16, out32, ldc, M, N, K, relu, (uint32_t)fmt, (cudaStream_t)stream);
}
