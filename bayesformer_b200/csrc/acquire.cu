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
// TMA copy (cp.async.bulk -> shared memory, completion on an mbarrier) into a ring of stages; the copy engine
// keeps HBM busy while the 128 threads each reduce their own item from shared memory (thread = item: no
// shuffles, 32 items per warp instruction).  A thread walks its row starting at a thread-dependent rotation when
// n_classes is even, which makes the stride-C shared-memory reads conflict-free for every C.
// Draws are visited in index order and divided by T at the end, as the reference's `uncertainties += ...` loop.
// ---------------------------------------------------------------------------
constexpr int AQ_ROWS = 128;

template <bool IS_LOGP, bool WANT_MEAN_P>
__global__ void __launch_bounds__(AQ_ROWS) acquire_tma_kernel(const float* __restrict__ x, int n_draws,
                                                               int64_t n_items, int64_t n_tiles, int n_classes,
                                                               int n_stages, uint32_t need,  // bit0 max, 1 margin, 2 entropy
                                                               float* __restrict__ mean_p,
                                                               float* __restrict__ score_of_mean,
                                                               float* __restrict__ mean_of_scores,
                                                               float* __restrict__ single_out, int single_mode) {
  extern __shared__ __align__(128) uint8_t aq_smem[];
  __shared__ uint64_t full_bar[4];
  const int C = n_classes;
  const uint32_t tile_bytes = (uint32_t)AQ_ROWS * C * 4u;
  float* stage_base = reinterpret_cast<float*>(aq_smem);
  float* acc_tile = WANT_MEAN_P ? stage_base + (size_t)n_stages * AQ_ROWS * C : nullptr;
  const int tid = threadIdx.x;
  const int64_t draw_stride = n_items * (int64_t)C;

  if (tid == 0) {
    for (int i = 0; i < n_stages; ++i) umma::mbar_init(&full_bar[i], 1);
    umma::fence_mbar_init();
  }
  __syncthreads();

  // flattened sequence of (tile, draw) loads of this CTA; load k goes to stage k % n_stages
  const int64_t my_tiles = (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x;
  const int64_t n_loads = my_tiles * n_draws;
  auto issue = [&](int64_t k) {
    const int64_t tile = blockIdx.x + (k / n_draws) * gridDim.x;
    const int t = (int)(k % n_draws);
    const int st = (int)(k % n_stages);
    umma::mbar_expect_tx(&full_bar[st], tile_bytes);
    umma::bulk_g2s(stage_base + (size_t)st * AQ_ROWS * C, x + t * draw_stride + tile * AQ_ROWS * (int64_t)C, tile_bytes,
                   &full_bar[st]);
  };
  if (tid == 0)
    for (int64_t k = 0; k < n_stages && k < n_loads; ++k) issue(k);

  const int rot = (C & 1) ? 0 : tid % C;   // even C: start column differs per thread -> no bank conflicts
  auto col = [&](int j) { const int c = rot + j; return c >= C ? c - C : c; };
  const float inv_t = 1.0f / (float)n_draws;
  int64_t k = 0;
  for (int64_t it = 0; it < my_tiles; ++it) {
    const int64_t tile = blockIdx.x + it * gridDim.x;
    const int64_t item = tile * AQ_ROWS + tid;
    float a_max = 0.f, a_margin = 0.f, a_ent = 0.f;
    float* acc = WANT_MEAN_P ? acc_tile + tid * C : nullptr;
    for (int t = 0; t < n_draws; ++t, ++k) {
      const int st = (int)(k % n_stages);
      umma::mbar_wait(&full_bar[st], (uint32_t)((k / n_stages) & 1));
      float* row = stage_base + (size_t)st * AQ_ROWS * C + tid * C;
      // The row is walked in rotated order with a warp-uniform trip count, four columns per trip into four
      // independent accumulator sets (a single warp per scheduler has nothing else to hide the latency of the
      // max / sum / top-2 dependency chains with).
      const int C4 = C & ~3;
      float b1w[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY}, b2w[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
      float entw[4] = {0.f, 0.f, 0.f, 0.f};
      float b1, b2, ent;
      auto top2 = [](float& t1, float& t2, float v) {
        t2 = fmaxf(t2, fminf(t1, v));
        t1 = fmaxf(t1, v);
      };
      if (IS_LOGP) {
        // probs = softmax(log-probs), as dropout_uncertainty does (active_learning.py:147):
        // e = exp(l - max), z = sum e, p = e / z.  The scores are taken on e and rescaled, and the entropy uses
        // log p = (l - max) - log z exactly (log(p + 1e-10) differs from it only where p < 1e-8).
        float mxw[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
        for (int j = 0; j < C4; j += 4) {
#pragma unroll
          for (int w = 0; w < 4; ++w) mxw[w] = fmaxf(mxw[w], row[col(j + w)]);
        }
        for (int j = C4; j < C; ++j) mxw[0] = fmaxf(mxw[0], row[col(j)]);
        const float mx = fmaxf(fmaxf(mxw[0], mxw[1]), fmaxf(mxw[2], mxw[3]));
        float zw[4] = {0.f, 0.f, 0.f, 0.f};
        auto body = [&](int c, int w) {
          const float d = row[c] - mx;
          const float e = __expf(d);
          zw[w] += e;
          if (need & 4u) entw[w] = fmaf(e, fmaxf(d, -1e30f), entw[w]);
          if (need & 3u) top2(b1w[w], b2w[w], e);
          if (WANT_MEAN_P) row[c] = e;
        };
        for (int j = 0; j < C4; j += 4) {
#pragma unroll
          for (int w = 0; w < 4; ++w) body(col(j + w), w);
        }
        for (int j = C4; j < C; ++j) body(col(j), 0);
        const float z = (zw[0] + zw[1]) + (zw[2] + zw[3]);
        const float inv_z = 1.0f / z;
        b1 = b1w[0]; b2 = b2w[0];
#pragma unroll
        for (int w = 1; w < 4; ++w) { top2(b1, b2, b2w[w]); top2(b1, b2, b1w[w]); }
        b1 *= inv_z;
        b2 *= inv_z;
        ent = ((entw[0] + entw[1]) + (entw[2] + entw[3])) * inv_z - __logf(z);   // sum p log p
        if (WANT_MEAN_P) {
          auto accum = [&](int c) { acc[c] = t == 0 ? row[c] * inv_z : fmaf(row[c], inv_z, acc[c]); };
#pragma unroll 4
          for (int j = 0; j < C; ++j) accum(col(j));
        }
      } else {
        auto body = [&](int c, int w) {
          const float p = row[c];
          if (need & 3u) top2(b1w[w], b2w[w], p);
          if (need & 4u) entw[w] = fmaf(p, __logf(p + 1e-10f), entw[w]);
          if (WANT_MEAN_P) acc[c] = t == 0 ? p : acc[c] + p;
        };
        for (int j = 0; j < C4; j += 4) {
#pragma unroll
          for (int w = 0; w < 4; ++w) body(col(j + w), w);
        }
        for (int j = C4; j < C; ++j) body(col(j), 0);
        b1 = b1w[0]; b2 = b2w[0];
#pragma unroll
        for (int w = 1; w < 4; ++w) { top2(b1, b2, b2w[w]); top2(b1, b2, b1w[w]); }
        ent = (entw[0] + entw[1]) + (entw[2] + entw[3]);
      }
      a_max += b1;
      a_margin += b1 - b2;
      a_ent -= ent;
      // this stage is consumed: refill it with the load n_stages ahead (generic-proxy accesses of the stage are
      // ordered before the copy engine's overwrite)
      umma::fence_async_smem();
      __syncthreads();
      if (tid == 0 && k + n_stages < n_loads) issue(k + n_stages);
    }
    if (mean_of_scores != nullptr) {
      if (need & 1u) mean_of_scores[item] = a_max * inv_t;
      if (need & 2u) mean_of_scores[n_items + item] = a_margin * inv_t;
      if (need & 4u) mean_of_scores[2 * n_items + item] = a_ent * inv_t;
    }
    if (single_out != nullptr)
      single_out[item] = (single_mode == BFB200_MODE_MAX ? a_max : single_mode == BFB200_MODE_MARGIN ? a_margin : a_ent) * inv_t;
    if (WANT_MEAN_P) {
      float b1 = -INFINITY, b2 = -INFINITY, ent = 0.f;
      auto fin = [&](int c) {
        const float p = acc[c] * inv_t;
        acc[c] = p;
        b2 = fmaxf(b2, fminf(b1, p));
        b1 = fmaxf(b1, p);
        ent = fmaf(p, __logf(p + 1e-10f), ent);
      };
#pragma unroll 4
      for (int j = 0; j < C; ++j) fin(col(j));
      if (score_of_mean != nullptr) {
        score_of_mean[item] = b1;
        score_of_mean[n_items + item] = b1 - b2;
        score_of_mean[2 * n_items + item] = -ent;
      }
      __syncthreads();
      if (mean_p != nullptr) {   // the tile of predictive means is contiguous in mean_p: coalesced 128-bit stores
        float4* dst = reinterpret_cast<float4*>(mean_p + tile * AQ_ROWS * (int64_t)C);
        const float4* src = reinterpret_cast<const float4*>(acc_tile);
        for (int i = tid; i < AQ_ROWS * C / 4; i += AQ_ROWS) dst[i] = src[i];
      }
      __syncthreads();
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
static int64_t launch_acquire_tma(const float* x, int is_logp, int n_draws, int64_t n_items, int n_classes,
                                  uint32_t need, float* mean_p, float* score_of_mean, float* mean_of_scores,
                                  float* single_out, int single_mode, cudaStream_t st, int* rc) {
  *rc = BFB200_OK;
  const int64_t n_tiles = n_items / AQ_ROWS;
  if (n_tiles == 0 || n_classes > 128) return 0;
  if ((reinterpret_cast<uintptr_t>(x) & 15u) != 0) return 0;
  if (n_draws > 1 && ((n_items * (int64_t)n_classes) & 3) != 0) return 0;   // every draw's tile 16-byte aligned
  const bool want_mean = mean_p != nullptr || score_of_mean != nullptr;
  if (want_mean && mean_p != nullptr && (reinterpret_cast<uintptr_t>(mean_p) & 15u) != 0) return 0;
  const size_t tile_bytes = (size_t)AQ_ROWS * n_classes * sizeof(float);
  int n_stages = (int)((200 * 1024 - (want_mean ? tile_bytes : 0)) / tile_bytes);
  if (n_stages > 4) n_stages = 4;
  if (n_stages < 2) return 0;
  const size_t smem = tile_bytes * (n_stages + (want_mean ? 1 : 0));
  int64_t grid = n_tiles < aq_sm_count() ? n_tiles : aq_sm_count();
#define AQ_LAUNCH(LOGP, MEAN)                                                                              \
  do {                                                                                                     \
    cudaFuncSetAttribute(acquire_tma_kernel<LOGP, MEAN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    acquire_tma_kernel<LOGP, MEAN><<<(int)grid, AQ_ROWS, smem, st>>>(x, n_draws, n_items, n_tiles, n_classes,   \
                                                                     n_stages, need, mean_p, score_of_mean,  \
                                                                     mean_of_scores, single_out, single_mode); \
  } while (0)
  if (is_logp) { if (want_mean) AQ_LAUNCH(true, true); else AQ_LAUNCH(true, false); }
  else         { if (want_mean) AQ_LAUNCH(false, true); else AQ_LAUNCH(false, false); }
#undef AQ_LAUNCH
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
