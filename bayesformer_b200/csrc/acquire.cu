// Acquisition kernels: compute_uncertainty, the T x C -> (mean p, max, margin, entropy)
// reduction, and the two toy classifiers' MC loops.
//
// All of these are HBM-streaming kernels (SURVEY §8d "Acquisition kernel standalone"):
// one warp owns one pool item, lanes stride over the class dimension so every global
// load is a coalesced 128-byte line, reductions stay in registers / warp shuffles.
#include "common.cuh"
#include "umma.cuh"

namespace bfb {

// ---------------------------------------------------------------------------
// compute_uncertainty (src/libs/active_learning.py:10-35): probs (B, C) -> (B)
// ---------------------------------------------------------------------------
template <int NPL>
__global__ void __launch_bounds__(256) compute_uncertainty_kernel(const float* __restrict__ probs,
                                                                  int64_t row_begin, int64_t n_rows,
                                                                  int n_classes, int mode,
                                                                  float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t warps_total = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t row = row_begin + (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); row < n_rows;
       row += warps_total) {
    const float* src = probs + row * n_classes;
    float p[NPL];
#pragma unroll
    for (int j = 0; j < NPL; ++j) {
      const int c = lane + 32 * j;
      p[j] = c < n_classes ? __ldg(src + c) : 0.f;
    }
    float s0, s1, s2;
    warp_scores<NPL>(p, n_classes, lane, s0, s1, s2);
    if (lane == 0) out[row] = mode == BFB200_MODE_MAX ? s0 : (mode == BFB200_MODE_MARGIN ? s1 : s2);
  }
}

// ---------------------------------------------------------------------------
// acquire: x (T, B, C) -> mean_p (B, C), score_of_mean (3, B), mean_of_scores (3, B)
// draws are accumulated in index order (the reference's `uncertainties += ...` loop,
// active_learning.py:141-155), then divided by T.
// ---------------------------------------------------------------------------
template <int NPL>
__global__ void __launch_bounds__(256) acquire_kernel(const float* __restrict__ x, int is_logp,
                                                      int n_draws, int64_t item_begin, int64_t n_items,
                                                      int n_classes, float* __restrict__ mean_p,
                                                      float* __restrict__ score_of_mean,
                                                      float* __restrict__ mean_of_scores) {
  const int lane = threadIdx.x & 31;
  const int64_t warps_total = (int64_t)gridDim.x * (blockDim.x >> 5);
  const int64_t draw_stride = n_items * (int64_t)n_classes;
  for (int64_t item = item_begin + (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); item < n_items;
       item += warps_total) {
    const float* src = x + item * n_classes;
    float acc[NPL];
#pragma unroll
    for (int j = 0; j < NPL; ++j) acc[j] = 0.f;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f;
#pragma unroll 2
    for (int t = 0; t < n_draws; ++t) {
      float p[NPL];
#pragma unroll
      for (int j = 0; j < NPL; ++j) {
        const int c = lane + 32 * j;
        p[j] = c < n_classes ? __ldg(src + (int64_t)t * draw_stride + c) : (is_logp ? -INFINITY : 0.f);
      }
      if (is_logp) {
        // probs = softmax(log-probs), as dropout_uncertainty does (active_learning.py:147)
        float m = -INFINITY;
#pragma unroll
        for (int j = 0; j < NPL; ++j) m = fmaxf(m, p[j]);
        m = warp_max(m);
        float z = 0.f;
#pragma unroll
        for (int j = 0; j < NPL; ++j) {
          p[j] = (lane + 32 * j) < n_classes ? expf(p[j] - m) : 0.f;
          z += p[j];
        }
        z = warp_sum(z);
#pragma unroll
        for (int j = 0; j < NPL; ++j) p[j] = p[j] / z;
      }
      float s0, s1, s2;
      warp_scores<NPL>(p, n_classes, lane, s0, s1, s2);
      a0 += s0; a1 += s1; a2 += s2;
#pragma unroll
      for (int j = 0; j < NPL; ++j) acc[j] += p[j];
    }
    const float inv_t = (float)n_draws;
#pragma unroll
    for (int j = 0; j < NPL; ++j) acc[j] = acc[j] / inv_t;
    if (mean_p != nullptr) {
#pragma unroll
      for (int j = 0; j < NPL; ++j) {
        const int c = lane + 32 * j;
        if (c < n_classes) mean_p[item * n_classes + c] = acc[j];
      }
    }
    if (score_of_mean != nullptr) {
      float s0, s1, s2;
      warp_scores<NPL>(acc, n_classes, lane, s0, s1, s2);
      if (lane == 0) {
        score_of_mean[item] = s0;
        score_of_mean[n_items + item] = s1;
        score_of_mean[2 * n_items + item] = s2;
      }
    }
    if (mean_of_scores != nullptr && lane == 0) {
      mean_of_scores[item] = a0 / inv_t;
      mean_of_scores[n_items + item] = a1 / inv_t;
      mean_of_scores[2 * n_items + item] = a2 / inv_t;
    }
  }
}


// ---------------------------------------------------------------------------
// HBM-streaming acquisition kernel for n_classes <= 128 (SURVEY 8d "Acquisition kernel standalone").
// A tile of 128 consecutive items of one draw is a contiguous byte range of x, so it is fetched by ONE 1-D bulk
// TMA copy (cp.async.bulk -> shared memory, completion on an mbarrier) into a ring of stages.  Sixteen warps drain
// the ring; each releases a stage (per-stage "empty" barrier) as soon as its rows are in registers, and one lane
// refills released stages opportunistically (it only blocks when the load it is about to consume is missing).  An item is
// owned by FOUR adjacent lanes (a warp = 8 items): each lane pulls its quarter of the row from shared memory into
// registers with 64-bit loads (32-bit when n_classes is odd), releases the stage at once, and reduces in
// registers; the four partial results meet through two xor-shuffles.  The predictive mean accumulates in
// registers across the draws, so a row is read from shared memory exactly once per draw.
// Bank conflicts: item i starts its walk at a rotation s_i chosen so that the 4 (8 for 32-bit loads) items served
// by one shared-memory wavefront start 4 load-units apart: conflict-free for every n_classes.
// Draws are visited in index order and divided by T at the end, as the reference's `uncertainties += ...` loop
// (active_learning.py:141-155).
// ---------------------------------------------------------------------------
constexpr int AQ_ROWS = 128;                      // items per tile
constexpr int AQ_SPLIT = 4;                       // lanes per item
constexpr int AQ_CONSUMERS = AQ_ROWS * AQ_SPLIT;  // 16 consumer warps
constexpr int AQ_THREADS = AQ_CONSUMERS;          // lane 0 of warp 0 doubles as the producer (a 17th warp would
                                                  // cost every thread a quarter of its registers)
constexpr int AQ_MAX_STAGES = 8;

__device__ __forceinline__ void aq_mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(umma::smem_u32(bar)) : "memory");
}
__device__ __forceinline__ float aq_ex2(float v) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
  return r;
}
__device__ __forceinline__ float aq_lg2(float v) {   // v >= 1e-10 here: flush-to-zero never triggers
  float r;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
  return r;
}
__device__ __forceinline__ void aq_top2(float& t1, float& t2, float v) {
  t2 = fmaxf(t2, fminf(t1, v));
  t1 = fmaxf(t1, v);
}
// merge the (largest, second largest) pairs of the four lanes of an item
__device__ __forceinline__ void aq_top2_quad(float& b1, float& b2) {
#pragma unroll
  for (int o = 1; o <= 2; o <<= 1) {
    const float o1 = __shfl_xor_sync(0xffffffffu, b1, o);
    const float o2 = __shfl_xor_sync(0xffffffffu, b2, o);
    b2 = fmaxf(fminf(b1, o1), fmaxf(b2, o2));
    b1 = fmaxf(b1, o1);
  }
}
__device__ __forceinline__ float aq_sum_quad(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  return v;
}

// VEC = floats per shared-memory load (2 when n_classes is even); NSTEPS = loads per lane and draw
// (ceil(ceil(C / VEC) / 4) <= NSTEPS).
template <int VEC, int NSTEPS, bool IS_LOGP, bool WANT_MEAN_P>
__global__ void __launch_bounds__(AQ_THREADS, 1) acquire_tma_kernel(const float* __restrict__ x, int n_draws,
                                                                     int64_t n_items, int64_t n_tiles, int n_classes,
                                                                     int n_stages, uint32_t need,  // bit0 max, 1 margin, 2 entropy
                                                                     float* __restrict__ mean_p,
                                                                     float* __restrict__ score_of_mean,
                                                                     float* __restrict__ mean_of_scores,
                                                                     float* __restrict__ single_out, int single_mode) {
  extern __shared__ __align__(128) uint8_t aq_smem[];
  __shared__ uint64_t full_bar[AQ_MAX_STAGES], empty_bar[AQ_MAX_STAGES];
  constexpr int NV = NSTEPS * VEC;
  constexpr float LOG2E = 1.4426950408889634f, LN2 = 0.6931471805599453f;
  const int C = n_classes;
  const uint32_t tile_bytes = (uint32_t)AQ_ROWS * C * 4u;
  const int tid = threadIdx.x;
  const int64_t draw_stride = n_items * (int64_t)C;

  if (tid == 0) {
    for (int i = 0; i < n_stages; ++i) {
      umma::mbar_init(&full_bar[i], 1);
      umma::mbar_init(&empty_bar[i], AQ_CONSUMERS / 32);
    }
    umma::fence_mbar_init();
  }
  __syncthreads();

  // flattened sequence of (tile, draw) loads of this CTA; load k goes to stage k % n_stages
  const int64_t my_tiles = (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x;

  // ---- consumers ----
  const int lane = tid & 31;
  const int q = lane & (AQ_SPLIT - 1);
  const int i_tile = tid >> 2;                     // item of the tile
  const int nvec = (C + VEC - 1) / VEC;            // load units per row (VEC == 2 only for even C)
  constexpr int NB = 32 / VEC, G = NB / AQ_SPLIT;  // load units per wavefront; items per wavefront
  int rot = (AQ_SPLIT * (i_tile % G) - (int)(((int64_t)i_tile * (C / VEC)) % NB)) % NB;
  if (rot < 0) rot += NB;
  rot %= nvec;
  uint32_t off[NSTEPS];      // byte offset inside a stage of this lane's j-th load unit (valid iff 4j + q < nvec)
#pragma unroll
  for (int j = 0; j < NSTEPS; ++j) {
    int v = AQ_SPLIT * j + q + rot;
    if (v >= nvec) v -= nvec;
    off[j] = AQ_SPLIT * j + q < nvec ? (uint32_t)(i_tile * C + v * VEC) * 4u : 0u;
  }
  const float pad = IS_LOGP ? -INFINITY : 0.f;
  const float inv_t = 1.0f / (float)n_draws;
  int st = 0;
  uint32_t phase = 0;

  // producer state (thread 0 only): loads issued so far, and where the next one goes
  const int64_t n_loads = my_tiles * n_draws;
  int64_t issued = 0, consumed = 0, p_it = 0;
  int p_t = 0, p_st = 0;
  uint32_t p_phase = 0;
  auto produce = [&](bool block) {
    while (issued < n_loads) {
      if (issued >= n_stages) {   // the stage must have been released by all sixteen warps
        if (block) {
          umma::mbar_wait(&empty_bar[p_st], p_phase ^ 1u);
        } else {
          uint32_t done;
          asm volatile(
              "{\n\t.reg .pred p;\n\t"
              "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
              "selp.u32 %0, 1, 0, p;\n\t}"
              : "=r"(done)
              : "r"(umma::smem_u32(&empty_bar[p_st])), "r"(p_phase ^ 1u)
              : "memory");
          if (!done) break;
        }
      }
      const float* src = x + (blockIdx.x + p_it * gridDim.x) * AQ_ROWS * (int64_t)C + p_t * draw_stride;
      umma::mbar_expect_tx(&full_bar[p_st], tile_bytes);
      umma::bulk_g2s(aq_smem + (size_t)p_st * tile_bytes, src, tile_bytes, &full_bar[p_st]);
      ++issued;
      if (++p_t == n_draws) { p_t = 0; ++p_it; }
      if (++p_st == n_stages) { p_st = 0; p_phase ^= 1u; }
      block = false;
    }
  };
  if (tid == 0) produce(false);
  for (int64_t it = 0; it < my_tiles; ++it) {
    const int64_t tile = blockIdx.x + it * gridDim.x;
    const int64_t item = tile * AQ_ROWS + i_tile;
    float a_max = 0.f, a_margin = 0.f, a_ent = 0.f;
    float acc[WANT_MEAN_P ? NV : 1];
    if (WANT_MEAN_P) {
#pragma unroll
      for (int j = 0; j < NV; ++j) acc[j] = 0.f;
    }
    for (int t = 0; t < n_draws; ++t) {
      if (tid == 0 && issued <= consumed) produce(true);
      umma::mbar_wait(&full_bar[st], phase);
      const uint8_t* base = aq_smem + (size_t)st * tile_bytes;
      float val[NV];
#pragma unroll
      for (int j = 0; j < NSTEPS; ++j) {
        const bool ok = AQ_SPLIT * j + q < nvec;
        if (VEC == 2) {
          const float2 v2 = ok ? *reinterpret_cast<const float2*>(base + off[j]) : make_float2(pad, pad);
          val[2 * j] = v2.x;
          val[2 * j + 1] = v2.y;
        } else {
          val[j] = ok ? *reinterpret_cast<const float*>(base + off[j]) : pad;
        }
      }
      // the row is in registers: hand the stage back to the producer
      __syncwarp();
      if (lane == 0) aq_mbar_arrive(&empty_bar[st]);
      if (++st == n_stages) { st = 0; phase ^= 1u; }
      if (tid == 0) { ++consumed; produce(false); }

      float b1a = -INFINITY, b2a = -INFINITY, b1b = -INFINITY, b2b = -INFINITY, enta = 0.f, entb = 0.f;
      float b1, b2, ent;
      if (IS_LOGP) {
        // probs = softmax(log-probs), as dropout_uncertainty does (active_learning.py:147):
        // e = exp(l - max), z = sum e, p = e / z.  The scores are taken on e and rescaled, and the entropy uses
        // log p = (l - max) - log z exactly (log(p + 1e-10) differs from it only where p < 1e-8).
        float mxa = -INFINITY, mxb = -INFINITY;
#pragma unroll
        for (int j = 0; j < NV; j += 2) {
          mxa = fmaxf(mxa, val[j]);
          if (j + 1 < NV) mxb = fmaxf(mxb, val[j + 1]);
        }
        float mx = fmaxf(mxa, mxb);
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
        const float mxs = mx * LOG2E;
        float za = 0.f, zb = 0.f;
#pragma unroll
        for (int j = 0; j < NV; ++j) {
          const float d2 = fmaf(val[j], LOG2E, -mxs);   // (l - max) / ln 2
          const float e = aq_ex2(d2);
          if (j & 1) {
            zb += e;
            if (need & 4u) entb = fmaf(e, fmaxf(d2, -1e30f), entb);
            if (need & 3u) aq_top2(b1b, b2b, e);
          } else {
            za += e;
            if (need & 4u) enta = fmaf(e, fmaxf(d2, -1e30f), enta);
            if (need & 3u) aq_top2(b1a, b2a, e);
          }
          if (WANT_MEAN_P) val[j] = e;
        }
        const float z = aq_sum_quad(za + zb);
        const float inv_z = 1.0f / z;
        b1 = b1a; b2 = b2a;
        if (need & 3u) {
          aq_top2(b1, b2, b2b);
          aq_top2(b1, b2, b1b);
          aq_top2_quad(b1, b2);
        }
        b1 *= inv_z;
        b2 *= inv_z;
        ent = 0.f;
        if (need & 4u) ent = aq_sum_quad(enta + entb) * LN2 * inv_z - __logf(z);   // sum p log p
        if (WANT_MEAN_P) {
#pragma unroll
          for (int j = 0; j < NV; ++j) acc[j] = fmaf(val[j], inv_z, acc[j]);
        }
      } else {
#pragma unroll
        for (int j = 0; j < NV; ++j) {
          const float p = val[j];
          if (j & 1) {
            if (need & 3u) aq_top2(b1b, b2b, p);
            if (need & 4u) entb = fmaf(p, aq_lg2(p + 1e-10f), entb);
          } else {
            if (need & 3u) aq_top2(b1a, b2a, p);
            if (need & 4u) enta = fmaf(p, aq_lg2(p + 1e-10f), enta);
          }
          if (WANT_MEAN_P) acc[j] += p;
        }
        b1 = b1a; b2 = b2a;
        if (need & 3u) {
          aq_top2(b1, b2, b2b);
          aq_top2(b1, b2, b1b);
          aq_top2_quad(b1, b2);
        }
        ent = 0.f;
        if (need & 4u) ent = aq_sum_quad(enta + entb) * LN2;   // sum p log(p + 1e-10), logs taken in base 2
      }
      a_max += b1;
      a_margin += b1 - b2;
      a_ent -= ent;
    }
    if (q == 0) {
      if (mean_of_scores != nullptr) {
        if (need & 1u) mean_of_scores[item] = a_max * inv_t;
        if (need & 2u) mean_of_scores[n_items + item] = a_margin * inv_t;
        if (need & 4u) mean_of_scores[2 * n_items + item] = a_ent * inv_t;
      }
      if (single_out != nullptr)
        single_out[item] = (single_mode == BFB200_MODE_MAX ? a_max : single_mode == BFB200_MODE_MARGIN ? a_margin : a_ent) * inv_t;
    }
    if (WANT_MEAN_P) {
      float b1 = -INFINITY, b2 = -INFINITY, ent = 0.f;
      float* dst = mean_p != nullptr ? mean_p + tile * AQ_ROWS * (int64_t)C : nullptr;
#pragma unroll
      for (int j = 0; j < NSTEPS; ++j) {
        if (AQ_SPLIT * j + q < nvec) {
#pragma unroll
          for (int w = 0; w < VEC; ++w) {
            const float p = acc[VEC * j + w] * inv_t;
            acc[VEC * j + w] = p;
            aq_top2(b1, b2, p);
            ent = fmaf(p, aq_lg2(p + 1e-10f), ent);
          }
          if (dst != nullptr) {   // the lanes of an item write adjacent 8-byte pieces of its row
            if (VEC == 2)
              __stcs(reinterpret_cast<float2*>(reinterpret_cast<uint8_t*>(dst) + off[j]),
                     make_float2(acc[2 * j], acc[2 * j + 1]));
            else
              __stcs(reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(dst) + off[j]), acc[j]);
          }
        }
      }
      if (score_of_mean != nullptr) {
        aq_top2_quad(b1, b2);
        ent = aq_sum_quad(ent) * LN2;
        if (q == 0) {
          score_of_mean[item] = b1;
          score_of_mean[n_items + item] = b1 - b2;
          score_of_mean[2 * n_items + item] = -ent;
        }
      }
    }
  }
}

// scores of the two-class distribution [1 - p, p] (compute_uncertainty on a (B, 2) tensor)
__device__ __forceinline__ void binary_scores(float p, float& s_max, float& s_margin, float& s_ent) {
  const float q = 1.f - p;
  const float hi = fmaxf(p, q), lo = fminf(p, q);
  s_max = hi;
  s_margin = hi - lo;
  s_ent = -(q * logf(q + 1e-10f) + p * logf(p + 1e-10f));
}

// ---------------------------------------------------------------------------
// MC-dropout MLP (src/model/bayesian_classification.py:184-198, loop of
// src/libs/visualization.py:105-113).  One thread per pool item: the hidden layer
// + ReLU is computed ONCE and kept in registers, only the masked 'hidden'-wide dot
// product is repeated per draw.  ALU/RNG bound by construction (12 B of HBM per item).
// ---------------------------------------------------------------------------
template <int HCAP>
__global__ void __launch_bounds__(128) mlp_mc_kernel(
    const float* __restrict__ x, int64_t n_items, int in_dim, int hidden,
    const float* __restrict__ w1, const float* __restrict__ b1, const float* __restrict__ w2,
    const float* __restrict__ b2, float scale, uint32_t threshold, int n_draws, uint2 key,
    int64_t index_base, const uint32_t* __restrict__ mask_bits, float* __restrict__ probs_out,
    float* __restrict__ mean_p, float* __restrict__ score_of_mean,
    float* __restrict__ mean_of_scores) {
  extern __shared__ float sm[];
  float* s_w1 = sm;                       // hidden * in_dim
  float* s_b1 = s_w1 + hidden * in_dim;   // hidden
  float* s_w2 = s_b1 + hidden;            // hidden
  for (int i = threadIdx.x; i < hidden * in_dim; i += blockDim.x) s_w1[i] = w1[i];
  for (int i = threadIdx.x; i < hidden; i += blockDim.x) { s_b1[i] = b1[i]; s_w2[i] = w2[i]; }
  __syncthreads();
  const float bias2 = b2[0];
  const int words = (hidden + 31) >> 5;

  for (int64_t item = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; item < n_items;
       item += (int64_t)gridDim.x * blockDim.x) {
    float h[HCAP];
#pragma unroll
    for (int j = 0; j < HCAP; ++j) {
      float a = 0.f;
      if (j < hidden) {
        for (int i = 0; i < in_dim; ++i) a = fmaf(__ldg(x + item * in_dim + i), s_w1[j * in_dim + i], a);
        a = fmaxf(a + s_b1[j], 0.f);
      }
      h[j] = a;
    }
    float sum_p = 0.f, a0 = 0.f, a1 = 0.f, a2 = 0.f;
    for (int t = 0; t < n_draws; ++t) {
      float logit = 0.f;
#pragma unroll
      for (int j8 = 0; j8 < HCAP / 8; ++j8) {
        if (j8 * 8 < hidden) {
          uint32_t keep;   // 8 hidden units per Philox call (16-bit decisions)
          if (mask_bits != nullptr) {
            const uint32_t w = mask_bits[((int64_t)t * n_items + item) * words + (j8 >> 2)];
            keep = (w >> ((j8 & 3) * 8)) & 0xFFu;
          } else {
            const uint4 r = philox4x32_10(
                make_uint4((uint32_t)(index_base + item), (uint32_t)t, 0u,
                           (BFB200_SITE_MLP_HIDDEN << 16) | (uint32_t)j8), key);
            keep = keep8_of(r, threshold);
          }
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const int j = j8 * 8 + q;
            if (j < hidden && ((keep >> q) & 1u)) logit = fmaf(h[j] * scale, s_w2[j], logit);
          }
        }
      }
      logit += bias2;
      const float p = 1.f / (1.f + expf(-logit));
      if (probs_out != nullptr) probs_out[(int64_t)t * n_items + item] = p;
      sum_p += p;
      float s0, s1, s2;
      binary_scores(p, s0, s1, s2);
      a0 += s0; a1 += s1; a2 += s2;
    }
    const float nt = (float)n_draws;
    const float pm = sum_p / nt;
    if (mean_p != nullptr) mean_p[item] = pm;
    if (score_of_mean != nullptr) {
      float s0, s1, s2;
      binary_scores(pm, s0, s1, s2);
      score_of_mean[item] = s0;
      score_of_mean[n_items + item] = s1;
      score_of_mean[2 * n_items + item] = s2;
    }
    if (mean_of_scores != nullptr) {
      mean_of_scores[item] = a0 / nt;
      mean_of_scores[n_items + item] = a1 / nt;
      mean_of_scores[2 * n_items + item] = a2 / nt;
    }
  }
}

// ---------------------------------------------------------------------------
// VariationalMLP (src/model/bayesian_classification.py:118-130, sampling :31-44).
// Every block first materialises the T sampled weight sets in shared memory
// (W = mu + eps * log1p(exp(rho)); one draw shared by the whole pool, :76), then one
// thread per item walks the draws with broadcast shared-memory reads.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(128) vmlp_mc_kernel(
    const float* __restrict__ x, int64_t n_items, int in_dim, int hidden,
    const float* __restrict__ w1_mu, const float* __restrict__ w1_rho, const float* __restrict__ b1,
    const float* __restrict__ w2_mu, const float* __restrict__ w2_rho, const float* __restrict__ b2,
    const float* __restrict__ eps, int n_draws, float* __restrict__ probs_out,
    float* __restrict__ mean_p, float* __restrict__ score_of_mean,
    float* __restrict__ mean_of_scores) {
  extern __shared__ float sm[];
  const int n1 = in_dim * hidden, n2 = hidden, per_draw = n1 + n2;
  float* s_w = sm;                         // n_draws * per_draw
  float* s_b1 = s_w + n_draws * per_draw;  // hidden
  for (int i = threadIdx.x; i < n_draws * per_draw; i += blockDim.x) {
    const int t = i / per_draw, r = i - t * per_draw;
    const float mu = r < n1 ? w1_mu[r] : w2_mu[r - n1];
    const float rho = r < n1 ? w1_rho[r] : w2_rho[r - n1];
    s_w[i] = mu + eps[i] * log1pf(expf(rho));
  }
  for (int i = threadIdx.x; i < hidden; i += blockDim.x) s_b1[i] = b1[i];
  __syncthreads();
  const float bias2 = b2[0];
  for (int64_t item = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; item < n_items;
       item += (int64_t)gridDim.x * blockDim.x) {
    float sum_p = 0.f, a0 = 0.f, a1 = 0.f, a2 = 0.f;
    for (int t = 0; t < n_draws; ++t) {
      const float* W1 = s_w + t * per_draw;  // (in_dim, hidden)
      const float* W2 = W1 + n1;             // (hidden, 1)
      float logit = 0.f;
      for (int j = 0; j < hidden; ++j) {
        float a = 0.f;
        for (int i = 0; i < in_dim; ++i) a = fmaf(__ldg(x + item * in_dim + i), W1[i * hidden + j], a);
        a = fmaxf(a + s_b1[j], 0.f);
        logit = fmaf(a, W2[j], logit);
      }
      logit += bias2;
      const float p = 1.f / (1.f + expf(-logit));
      if (probs_out != nullptr) probs_out[(int64_t)t * n_items + item] = p;
      sum_p += p;
      float s0, s1, s2;
      binary_scores(p, s0, s1, s2);
      a0 += s0; a1 += s1; a2 += s2;
    }
    const float nt = (float)n_draws;
    const float pm = sum_p / nt;
    if (mean_p != nullptr) mean_p[item] = pm;
    if (score_of_mean != nullptr) {
      float s0, s1, s2;
      binary_scores(pm, s0, s1, s2);
      score_of_mean[item] = s0;
      score_of_mean[n_items + item] = s1;
      score_of_mean[2 * n_items + item] = s2;
    }
    if (mean_of_scores != nullptr) {
      mean_of_scores[item] = a0 / nt;
      mean_of_scores[n_items + item] = a1 / nt;
      mean_of_scores[2 * n_items + item] = a2 / nt;
    }
  }
}

static int grid_for(int64_t work_items, int per_block, int blocks_per_sm = 8) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int64_t need = (work_items + per_block - 1) / per_block;
  int64_t cap = (int64_t)sms * blocks_per_sm;  // grid-stride: whole waves of the SM count
  if (need < 1) need = 1;
  return (int)(need < cap ? need : cap);
}


// number of SMs of the current device
static int aq_sm_count() {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return sms > 0 ? sms : 148;
}

// Launch the TMA-staged kernel on the full 128-item tiles of x; returns the number of items it covered (0 when
// the shape or alignment is outside its class, in which case the warp-per-item kernels take everything).
template <int VEC, int NSTEPS>
static void aq_launch(bool is_logp, bool want_mean, int grid, size_t smem, cudaStream_t st, const float* x, int n_draws,
                      int64_t n_items, int64_t n_tiles, int n_classes, int n_stages, uint32_t need, float* mean_p,
                      float* score_of_mean, float* mean_of_scores, float* single_out, int single_mode) {
#define AQ_LAUNCH(LOGP, MEAN)                                                                                    \
  do {                                                                                                           \
    cudaFuncSetAttribute(acquire_tma_kernel<VEC, NSTEPS, LOGP, MEAN>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                         (int)smem);                                                                             \
    acquire_tma_kernel<VEC, NSTEPS, LOGP, MEAN><<<grid, AQ_THREADS, smem, st>>>(                                  \
        x, n_draws, n_items, n_tiles, n_classes, n_stages, need, mean_p, score_of_mean, mean_of_scores,          \
        single_out, single_mode);                                                                                \
  } while (0)
  if (is_logp) { if (want_mean) AQ_LAUNCH(true, true); else AQ_LAUNCH(true, false); }
  else         { if (want_mean) AQ_LAUNCH(false, true); else AQ_LAUNCH(false, false); }
#undef AQ_LAUNCH
}

static int64_t launch_acquire_tma(const float* x, int is_logp, int n_draws, int64_t n_items, int n_classes,
                                  uint32_t need, float* mean_p, float* score_of_mean, float* mean_of_scores,
                                  float* single_out, int single_mode, cudaStream_t st, int* rc) {
  *rc = BFB200_OK;
  const int64_t n_tiles = n_items / AQ_ROWS;
  if (n_tiles == 0 || n_classes > 128 || n_classes < 2) return 0;
  if ((reinterpret_cast<uintptr_t>(x) & 15u) != 0) return 0;
  if (n_draws > 1 && ((n_items * (int64_t)n_classes) & 3) != 0) return 0;   // every draw's tile 16-byte aligned
  const bool want_mean = mean_p != nullptr || score_of_mean != nullptr;
  if (mean_p != nullptr && (reinterpret_cast<uintptr_t>(mean_p) & 7u) != 0) return 0;
  const size_t tile_bytes = (size_t)AQ_ROWS * n_classes * sizeof(float);
  int n_stages = (int)((200 * 1024) / tile_bytes);
  if (n_stages > AQ_MAX_STAGES) n_stages = AQ_MAX_STAGES;
  if (n_stages < 2) return 0;
  const size_t smem = tile_bytes * n_stages;
  const int grid = (int)(n_tiles < aq_sm_count() ? n_tiles : aq_sm_count());
  const int vec = (n_classes & 1) ? 1 : 2;
  const int steps = ((n_classes + vec - 1) / vec + AQ_SPLIT - 1) / AQ_SPLIT;
#define AQ_GO(VEC, NSTEPS)                                                                                        \
  aq_launch<VEC, NSTEPS>(is_logp != 0, want_mean, grid, smem, st, x, n_draws, n_items, n_tiles, n_classes, n_stages, \
                         need, mean_p, score_of_mean, mean_of_scores, single_out, single_mode)
  if (vec == 2) {
    if (steps <= 4) AQ_GO(2, 4);
    else if (steps <= 8) AQ_GO(2, 8);
    else if (steps <= 15) AQ_GO(2, 15);
    else AQ_GO(2, 16);
  } else {
    if (steps <= 8) AQ_GO(1, 8);
    else if (steps <= 17) AQ_GO(1, 17);
    else AQ_GO(1, 32);
  }
#undef AQ_GO
  count_launch("acquire_tma_kernel", st);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("launch of acquire_tma_kernel failed: %s", cudaGetErrorString(e));
    *rc = BFB200_ERR_CUDA;
    return 0;
  }
  return n_tiles * AQ_ROWS;
}

}  // namespace bfb

using namespace bfb;

extern "C" int bfb200_compute_uncertainty(const float* probs, int64_t n_rows, int32_t n_classes,
                                          int32_t mode, float* out, void* stream) {
  BFB_REQUIRE(mode >= BFB200_MODE_MAX && mode <= BFB200_MODE_ENTROPY, BFB200_ERR_INVALID_ARG,
              "Unknown mode %d. Available modes are 'max', 'margin' and 'entropy'.", mode);
  BFB_REQUIRE(n_rows >= 0 && n_classes >= 1, BFB200_ERR_INVALID_ARG, "bad shape (%lld, %d)",
              (long long)n_rows, n_classes);
  BFB_REQUIRE(n_classes <= 1024, BFB200_ERR_UNSUPPORTED, "n_classes %d > 1024", n_classes);
  if (n_rows == 0) return BFB200_OK;
  BFB_DEV(probs);
  BFB_DEV(out);
  cudaStream_t st = (cudaStream_t)stream;
  int rc = BFB200_OK;
  const uint32_t need = mode == BFB200_MODE_ENTROPY ? 4u : 3u;
  const int64_t row_begin = launch_acquire_tma(probs, 0, 1, n_rows, n_classes, need, nullptr, nullptr, nullptr, out, mode,
                                               st, &rc);
  if (rc != BFB200_OK) return rc;
  if (row_begin == n_rows) return BFB200_OK;
  const int grid = grid_for(n_rows - row_begin, 8);
#define LAUNCH_CU(NPL) compute_uncertainty_kernel<NPL><<<grid, 256, 0, st>>>(probs, row_begin, n_rows, n_classes, mode, out)
  if (n_classes <= 32) LAUNCH_CU(1);
  else if (n_classes <= 64) LAUNCH_CU(2);
  else if (n_classes <= 128) LAUNCH_CU(4);
  else if (n_classes <= 256) LAUNCH_CU(8);
  else LAUNCH_CU(32);
#undef LAUNCH_CU
  BFB_LAUNCH_CHECK("compute_uncertainty_kernel", st);
  return BFB200_OK;
}

extern "C" int bfb200_acquire(const float* x, int32_t is_logp, int32_t n_draws, int64_t n_items,
                              int32_t n_classes, float* mean_p, float* score_of_mean,
                              float* mean_of_scores, void* stream) {
  BFB_REQUIRE(n_draws >= 1 && n_items >= 0 && n_classes >= 1, BFB200_ERR_INVALID_ARG,
              "bad shape (T=%d, B=%lld, C=%d)", n_draws, (long long)n_items, n_classes);
  BFB_REQUIRE(n_classes <= 1024, BFB200_ERR_UNSUPPORTED, "n_classes %d > 1024", n_classes);
  if (n_items == 0) return BFB200_OK;
  BFB_DEV(x);
  BFB_DEV_OPT(mean_p);
  BFB_DEV_OPT(score_of_mean);
  BFB_DEV_OPT(mean_of_scores);
  cudaStream_t st = (cudaStream_t)stream;
  int rc = BFB200_OK;
  const int64_t item_begin = launch_acquire_tma(x, is_logp, n_draws, n_items, n_classes, mean_of_scores != nullptr ? 7u : 0u,
                                                mean_p, score_of_mean, mean_of_scores, nullptr, 0, st, &rc);
  if (rc != BFB200_OK) return rc;
  if (item_begin == n_items) return BFB200_OK;
  const int grid = grid_for(n_items - item_begin, 8);
#define LAUNCH_AQ(NPL) acquire_kernel<NPL><<<grid, 256, 0, st>>>(x, is_logp, n_draws, item_begin, n_items, n_classes, mean_p, score_of_mean, mean_of_scores)
  if (n_classes <= 32) LAUNCH_AQ(1);
  else if (n_classes <= 64) LAUNCH_AQ(2);
  else if (n_classes <= 128) LAUNCH_AQ(4);
  else if (n_classes <= 256) LAUNCH_AQ(8);
  else LAUNCH_AQ(32);
#undef LAUNCH_AQ
  BFB_LAUNCH_CHECK("acquire_kernel", st);
  return BFB200_OK;
}

extern "C" int bfb200_mlp_mc_score(const float* x, int64_t n_items, int32_t in_dim, int32_t hidden,
                                   const float* w1, const float* b1, const float* w2, const float* b2,
                                   float p_drop, int32_t n_draws, uint64_t seed, int64_t index_base,
                                   const uint32_t* mask_bits, float* probs_out, float* mean_p,
                                   float* score_of_mean, float* mean_of_scores, void* stream) {
  BFB_REQUIRE(n_items >= 0 && in_dim >= 1 && hidden >= 1 && n_draws >= 1, BFB200_ERR_INVALID_ARG,
              "bad shape (B=%lld, in=%d, hidden=%d, T=%d)", (long long)n_items, in_dim, hidden, n_draws);
  BFB_REQUIRE(p_drop >= 0.f && p_drop < 1.f, BFB200_ERR_INVALID_ARG, "p_drop %f outside [0, 1)", p_drop);
  BFB_REQUIRE(hidden <= 128, BFB200_ERR_UNSUPPORTED, "hidden %d > 128", hidden);
  BFB_REQUIRE((size_t)hidden * (in_dim + 2) * 4 <= 160 * 1024, BFB200_ERR_UNSUPPORTED, "weights do not fit shared memory");
  if (n_items == 0) return BFB200_OK;
  BFB_DEV(x); BFB_DEV(w1); BFB_DEV(b1); BFB_DEV(w2); BFB_DEV(b2);
  BFB_DEV_OPT(mask_bits); BFB_DEV_OPT(probs_out); BFB_DEV_OPT(mean_p);
  BFB_DEV_OPT(score_of_mean); BFB_DEV_OPT(mean_of_scores);
  cudaStream_t st = (cudaStream_t)stream;
  const uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
  const float scale = 1.0f / (1.0f - p_drop);
  const uint32_t thr = keep_threshold(p_drop);
  const size_t smem = (size_t)hidden * (in_dim + 2) * sizeof(float);
  const int grid = grid_for(n_items, 128, 16);
#define LAUNCH_MLP(HCAP)                                                                              \
  do {                                                                                                \
    if (smem > 48 * 1024)                                                                             \
      cudaFuncSetAttribute(mlp_mc_kernel<HCAP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    mlp_mc_kernel<HCAP><<<grid, 128, smem, st>>>(x, n_items, in_dim, hidden, w1, b1, w2, b2, scale, thr, \
                                                 n_draws, key, index_base, mask_bits, probs_out, mean_p, \
                                                 score_of_mean, mean_of_scores);                      \
  } while (0)
  if (hidden <= 64) LAUNCH_MLP(64);
  else LAUNCH_MLP(128);
#undef LAUNCH_MLP
  BFB_LAUNCH_CHECK("mlp_mc_kernel", st);
  return BFB200_OK;
}

extern "C" int bfb200_vmlp_mc_score(const float* x, int64_t n_items, int32_t in_dim, int32_t hidden,
                                    const float* w1_mu, const float* w1_rho, const float* b1,
                                    const float* w2_mu, const float* w2_rho, const float* b2,
                                    const float* eps, int32_t n_draws, float* probs_out, float* mean_p,
                                    float* score_of_mean, float* mean_of_scores, void* stream) {
  BFB_REQUIRE(n_items >= 0 && in_dim >= 1 && hidden >= 1 && n_draws >= 1, BFB200_ERR_INVALID_ARG,
              "bad shape (B=%lld, in=%d, hidden=%d, T=%d)", (long long)n_items, in_dim, hidden, n_draws);
  const size_t smem = ((size_t)n_draws * ((size_t)in_dim * hidden + hidden) + hidden) * sizeof(float);
  BFB_REQUIRE(smem <= 200 * 1024, BFB200_ERR_UNSUPPORTED,
              "sampled weights (%zu B) do not fit shared memory", smem);
  if (n_items == 0) return BFB200_OK;
  BFB_DEV(x); BFB_DEV(w1_mu); BFB_DEV(w1_rho); BFB_DEV(b1); BFB_DEV(w2_mu); BFB_DEV(w2_rho);
  BFB_DEV(b2); BFB_DEV(eps);
  BFB_DEV_OPT(probs_out); BFB_DEV_OPT(mean_p); BFB_DEV_OPT(score_of_mean); BFB_DEV_OPT(mean_of_scores);
  cudaStream_t st = (cudaStream_t)stream;
  if (smem > 48 * 1024)
    cudaFuncSetAttribute(vmlp_mc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int grid = grid_for(n_items, 128, 8);
  vmlp_mc_kernel<<<grid, 128, smem, st>>>(x, n_items, in_dim, hidden, w1_mu, w1_rho, b1, w2_mu, w2_rho, b2,
                                          eps, n_draws, probs_out, mean_p, score_of_mean, mean_of_scores);
  BFB_LAUNCH_CHECK("vmlp_mc_kernel", st);
  return BFB200_OK;
}
