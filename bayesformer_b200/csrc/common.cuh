// Shared device/host helpers of libbayesformer_b200 (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <math.h>

#include "../../include/bayesformer_b200.h"

namespace bfb {

// ----------------------------------------------------------------------------
// error plumbing
// ----------------------------------------------------------------------------
void set_error(const char* fmt, ...);
int check_device_ptr(const void* p, const char* name);  // BFB200_OK or error (message set)
void count_launch(const char* name, cudaStream_t st);  // launch accounting + optional per-kernel event timing

#define BFB_REQUIRE(cond, code, ...)      \
  do {                                    \
    if (!(cond)) {                        \
      ::bfb::set_error(__VA_ARGS__);      \
      return (code);                      \
    }                                     \
  } while (0)

#define BFB_DEV(ptr)                                              \
  do {                                                            \
    int _rc = ::bfb::check_device_ptr((ptr), #ptr);               \
    if (_rc != BFB200_OK) return _rc;                             \
  } while (0)

#define BFB_DEV_OPT(ptr)                                          \
  do {                                                            \
    if ((ptr) != nullptr) {                                       \
      int _rc = ::bfb::check_device_ptr((ptr), #ptr);             \
      if (_rc != BFB200_OK) return _rc;                           \
    }                                                             \
  } while (0)

#define BFB_LAUNCH_CHECK(name, stream_)                                            \
  do {                                                                             \
    ::bfb::count_launch((name), (stream_));                                        \
    cudaError_t _e = cudaGetLastError();                                           \
    if (_e != cudaSuccess) {                                                       \
      ::bfb::set_error("launch of %s failed: %s", (name), cudaGetErrorString(_e)); \
      return BFB200_ERR_CUDA;                                                      \
    }                                                                              \
  } while (0)

#define BFB_CALL(expr)                 \
  do {                                 \
    int _rc = (expr);                  \
    if (_rc != BFB200_OK) return _rc;  \
  } while (0)

// ----------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al., SC'11) — counter-based, no state.
// ----------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint32_t mulhi32(uint32_t a, uint32_t b) {
#ifdef __CUDA_ARCH__
  return __umulhi(a, b);
#else
  return (uint32_t)(((uint64_t)a * (uint64_t)b) >> 32);
#endif
}

__host__ __device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
  constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
  constexpr uint32_t W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = mulhi32(M0, c.x), lo0 = M0 * c.x;
    const uint32_t hi1 = mulhi32(M1, c.z), lo1 = M1 * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += W0;
    k.y += W1;
  }
  return c;
}

// Bernoulli decisions are taken on the 16-bit halves of the Philox words (two decisions per word, eight per
// call): a feature is KEPT iff its half >= threshold16(p) = min(65535, floor(p * 65536)).  Keep probability
// = 1 - threshold16 / 65536, i.e. p is resolved to 1.5e-5.
__host__ inline uint32_t keep_threshold(float p) {
  double t = (double)p * 65536.0;
  if (t <= 0.0) return 0u;
  if (t >= 65535.0) return 65535u;
  return (uint32_t)t;
}

// 8 keep bits from one Philox call: bit j <- half (j & 1) of word (j >> 1)
__host__ __device__ __forceinline__ uint32_t keep8_of(const uint4& r, uint32_t thr) {
  return ((r.x & 0xFFFFu) >= thr ? 1u : 0u) | ((r.x >> 16) >= thr ? 2u : 0u) |
         ((r.y & 0xFFFFu) >= thr ? 4u : 0u) | ((r.y >> 16) >= thr ? 8u : 0u) |
         ((r.z & 0xFFFFu) >= thr ? 16u : 0u) | ((r.z >> 16) >= thr ? 32u : 0u) |
         ((r.w & 0xFFFFu) >= thr ? 64u : 0u) | ((r.w >> 16) >= thr ? 128u : 0u);
}

// Where the dropout masks of one launch come from (see bfb200_masks).
struct MaskSource {
  const uint32_t* bits;  // nullptr -> Philox
  uint2 key;
  uint32_t threshold;
  float scale;           // 1 / (1 - p)
  // packed-bit addressing
  int32_t n_sites, max_len, words;  // words = ceil(d / 32)
  int64_t n_seq;                    // sequences covered by `bits` per (step, site)
  int64_t seq_base;                 // first sequence of this launch inside `bits`
  // Philox addressing
  uint32_t step;
  uint32_t n_draws;
  int64_t item_base;     // global item index of local sequence 0's item
  uint32_t draw_base;    // used when n_draws == 1 (plain forward: draw id of the call)
};

// 4 keep bits (bit j = feature 4*f4 + j) for (local sequence s, position pos, site).  Philox counter =
// (item, draw, step << 16 | pos, site << 16 | feature >> 3): one call covers 8 features.
__device__ __forceinline__ uint32_t keep4(const MaskSource& m, int64_t s, uint32_t pos,
                                          uint32_t site, uint32_t f4) {
  if (m.bits != nullptr) {
    const int64_t row = (((int64_t)m.step * m.n_sites + site) * m.n_seq + (m.seq_base + s)) *
                            m.max_len + pos;
    const uint32_t w = m.bits[row * m.words + (f4 >> 3)];
    return (w >> ((f4 & 7u) * 4u)) & 0xFu;
  }
  const uint64_t item = (uint64_t)(m.item_base + s / m.n_draws);
  const uint32_t draw = m.draw_base + (uint32_t)(s % m.n_draws);
  const uint4 ctr = make_uint4((uint32_t)item, draw, (m.step << 16) | pos, (site << 16) | (f4 >> 1));
  const uint4 r = philox4x32_10(ctr, m.key);
  return (keep8_of(r, m.threshold) >> ((f4 & 1u) * 4u)) & 0xFu;
}

// ----------------------------------------------------------------------------
// warp reductions
// ----------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ----------------------------------------------------------------------------
// sortable 64-bit keys: larger key = larger score, ties -> lower index first
// ----------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint32_t orderable_f32(float f) {
#ifdef __CUDA_ARCH__
  uint32_t u = __float_as_uint(f);
#else
  uint32_t u;
  memcpy(&u, &f, 4);
#endif
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__host__ __device__ __forceinline__ float from_orderable_f32(uint32_t o) {
  uint32_t u = (o & 0x80000000u) ? (o & 0x7FFFFFFFu) : ~o;
#ifdef __CUDA_ARCH__
  return __uint_as_float(u);
#else
  float f;
  memcpy(&f, &u, 4);
  return f;
#endif
}
__host__ __device__ __forceinline__ uint64_t pack_key(float score, uint64_t global_index) {
  return ((uint64_t)orderable_f32(score) << 32) | (uint64_t)(0xFFFFFFFFu - (uint32_t)global_index);
}

// ----------------------------------------------------------------------------
// the three scores of compute_uncertainty on one probability row, warp-cooperative.
// Each lane holds p[j] for classes lane, lane+32, ... (register array, NPL entries,
// classes >= n_classes are ignored).  Reference:
// src/libs/active_learning.py:25-31.
// ----------------------------------------------------------------------------
template <int NPL>
__device__ __forceinline__ void warp_scores(const float (&p)[NPL], int n_classes, int lane,
                                            float& s_max, float& s_margin, float& s_entropy) {
  float best = -INFINITY;
  int best_i = 0x7fffffff;
  float ent = 0.f;
#pragma unroll
  for (int j = 0; j < NPL; ++j) {
    const int c = lane + 32 * j;
    if (c < n_classes) {
      if (p[j] > best) { best = p[j]; best_i = c; }
      ent += p[j] * logf(p[j] + 1e-10f);
    }
  }
  // arg-max with lowest-index tie break
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, best_i, o);
    if (ob > best || (ob == best && oi < best_i)) { best = ob; best_i = oi; }
  }
  float second = -INFINITY;
#pragma unroll
  for (int j = 0; j < NPL; ++j) {
    const int c = lane + 32 * j;
    if (c < n_classes && c != best_i) second = fmaxf(second, p[j]);
  }
  second = warp_max(second);
  s_max = best;
  s_margin = best - second;
  s_entropy = -warp_sum(ent);
}

}  // namespace bfb
