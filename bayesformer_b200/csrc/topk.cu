// top-k selection (torch.topk of select_samples, src/libs/active_learning.py:210) on packed 64-bit keys
// (orderable(score) << 32 | ~index: all distinct, larger = better, ties towards the lowest index).
//
// A block takes a chunk of 16,384 keys into shared memory and finds its k-th largest key by RADIX SELECT — six
// histogram rounds over 12-bit digits, most significant first, each narrowing the candidates to the bin that holds
// the k-th key — then compacts the k keys >= that threshold.  Levels repeat until one block remains; the last one
// bitonic-sorts its k survivors (k <= 4096) for the output.  This block-level path serves inputs of up to one chunk and
// bfb200_merge_topk (candidate keys of several ranks).
//
// Scores that do not fit one chunk take the ONE-LAUNCH path (topk_global_kernel): a cooperative grid of one CTA per SM
// runs the same radix select on GLOBAL histograms — per round every CTA histograms the 12-bit digit of its slice in
// shared memory, adds its non-empty bins to the round's global histogram, a grid barrier, then every CTA reads the
// 4,096 bins and derives the same digit — and stops as soon as the keys at or above the chosen prefix fit one
// block's sort buffer (typically after two rounds: sign / exponent / 3 mantissa bits, then 12 more mantissa bits).
// Those candidates are appended to a global list, and CTA 0 bitonic-sorts them and writes the k best.
// 1 M scores, k = 200: one launch, three grid barriers.  Keys are recomputed from the scores in every round
// (4 B per score from L2), so the slice size is unbounded.
#include "common.cuh"

namespace bfb {

constexpr int TOPK_CHUNK = 16384;   // keys per block (128 KB of shared memory)
constexpr int TOPK_THREADS = 1024;
constexpr int TOPK_MAX_K = TOPK_CHUNK / 4;
constexpr int TOPK_BINS = 4096;     // 12-bit digits
constexpr size_t TOPK_SMEM = (size_t)TOPK_CHUNK * 8 + (size_t)TOPK_BINS * 4 + (size_t)TOPK_MAX_K * 8 + 256;

template <bool FROM_SCORES, bool FINAL>
__global__ void __launch_bounds__(TOPK_THREADS) topk_select_kernel(
    const float* __restrict__ scores, const uint64_t* __restrict__ keys_in, int64_t n, int k,
    int64_t index_base, uint64_t* __restrict__ keys_out) {
  extern __shared__ __align__(16) unsigned char topk_smem[];
  uint64_t* skeys = reinterpret_cast<uint64_t*>(topk_smem);
  uint32_t* hist = reinterpret_cast<uint32_t*>(skeys + TOPK_CHUNK);
  uint64_t* sout = reinterpret_cast<uint64_t*>(hist + TOPK_BINS);
  uint32_t* misc = reinterpret_cast<uint32_t*>(sout + TOPK_MAX_K);   // [0] chosen digit, [1] keys above it, [2] output counter, [8..40) warp sums
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t base = (int64_t)blockIdx.x * TOPK_CHUNK;
  const int m = (int)((n - base) < TOPK_CHUNK ? (n - base) : TOPK_CHUNK);   // valid keys of this chunk
  for (int i = tid; i < TOPK_CHUNK; i += TOPK_THREADS) {
    const int64_t g = base + i;
    uint64_t key = 0ull;  // padding sorts last
    if (i < m) key = FROM_SCORES ? pack_key(__ldg(scores + g), (uint64_t)(index_base + g)) : keys_in[g];
    skeys[i] = key;
  }
  if (tid == 0) misc[2] = 0u;
  __syncthreads();

  uint64_t threshold = 0ull;   // k-th largest key of the chunk (0: everything survives)
  if (m > k) {
    uint64_t prefix = 0ull, mask = 0ull;
    uint32_t kk = (uint32_t)k;   // rank of the wanted key among the keys that share `prefix`
#pragma unroll 1
    for (int r = 0; r < 6; ++r) {
      const int shift = r < 5 ? 52 - 12 * r : 0;
      const uint32_t nb = r < 5 ? 4096u : 16u;
      for (int i = tid; i < TOPK_BINS; i += TOPK_THREADS) hist[i] = 0u;
      __syncthreads();
      for (int i = tid; i < TOPK_CHUNK; i += TOPK_THREADS) {   // (uniform trip count: the ballots below are warp-wide)
        const uint64_t key = skeys[i];
        const bool in = i < m && (key & mask) == prefix;
        const uint32_t digit = (uint32_t)(key >> shift) & (nb - 1u);
        const uint32_t act = __ballot_sync(0xffffffffu, in);
        if (in) {   // one atomic per distinct digit of the warp (scores share their leading bits: few distinct digits)
          const uint32_t peers = __match_any_sync(act, digit);
          if (lane == __ffs(peers) - 1) atomicAdd(&hist[digit], (uint32_t)__popc(peers));
        }
      }
      __syncthreads();
      // the digit d with  #(digits > d) < kk <= #(digits >= d): suffix sums over the bins, four bins per thread
      const uint32_t b0 = (uint32_t)tid * 4u;
      const uint32_t h0 = hist[b0], h1 = hist[b0 + 1], h2 = hist[b0 + 2], h3 = hist[b0 + 3];
      uint32_t incl = h0 + h1 + h2 + h3;   // -> sum over this thread's bins and all higher threads'
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t up = __shfl_down_sync(0xffffffffu, incl, o);
        if (lane + o < 32) incl += up;
      }
      if (lane == 0) misc[8 + warp] = incl;
      __syncthreads();
      uint32_t higher_warps = 0u;
      for (int w = warp + 1; w < TOPK_THREADS / 32; ++w) higher_warps += misc[8 + w];
      const uint32_t above_thread = incl + higher_warps - (h0 + h1 + h2 + h3);   // keys in bins above this thread's four
      const uint32_t a3 = above_thread, a2 = a3 + h3, a1 = a2 + h2, a0 = a1 + h1;
      if (a3 < kk && kk <= a3 + h3) { misc[0] = b0 + 3; misc[1] = a3; }
      else if (a2 < kk && kk <= a2 + h2) { misc[0] = b0 + 2; misc[1] = a2; }
      else if (a1 < kk && kk <= a1 + h1) { misc[0] = b0 + 1; misc[1] = a1; }
      else if (a0 < kk && kk <= a0 + h0) { misc[0] = b0; misc[1] = a0; }
      __syncthreads();
      prefix |= (uint64_t)misc[0] << shift;
      mask |= (uint64_t)(nb - 1u) << shift;
      kk -= misc[1];
      __syncthreads();   // misc / hist are rewritten by the next round
    }
    threshold = prefix;   // all 64 bits fixed: the k-th largest key itself
  }

  // compaction of the survivors (exactly min(k, m) keys), one atomic per warp and trip
  for (int i = tid; i < TOPK_CHUNK; i += TOPK_THREADS) {
    const uint64_t key = skeys[i];
    const bool keep = i < m && key >= threshold;
    const uint32_t act = __ballot_sync(0xffffffffu, keep);
    uint32_t pos = 0u;
    if (lane == 0 && act != 0u) pos = atomicAdd(&misc[2], (uint32_t)__popc(act));
    pos = __shfl_sync(0xffffffffu, pos, 0);
    if (keep) {
      const uint32_t at = pos + (uint32_t)__popc(act & ((1u << lane) - 1u));
      if (at < (uint32_t)k) sout[at] = key;   // (keys are distinct, so exactly k survive; duplicates fed to merge_topk must not overrun)
    }
  }
  __syncthreads();
  const int kept = m > k ? k : m;
  if (!FINAL) {   // survivors in any order; short chunks pad with keys that sort last
    for (int i = tid; i < k; i += TOPK_THREADS) keys_out[(int64_t)blockIdx.x * k + i] = i < kept ? sout[i] : 0ull;
    return;
  }
  int p2 = 1;
  while (p2 < k) p2 <<= 1;
  for (int i = kept + tid; i < p2; i += TOPK_THREADS) sout[i] = 0ull;
  __syncthreads();
  for (int size = 2; size <= p2; size <<= 1) {   // bitonic sort, descending
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int t = tid; t < p2 / 2; t += TOPK_THREADS) {
        const int lo = ((t / stride) * (stride << 1)) + (t % stride);
        const int hi = lo + stride;
        const bool desc = ((lo & size) == 0);
        const uint64_t a = sout[lo], b = sout[hi];
        if ((a < b) == desc) { sout[lo] = b; sout[hi] = a; }
      }
      __syncthreads();
    }
  }
  for (int i = tid; i < k; i += TOPK_THREADS) keys_out[i] = sout[i];
}


// ---- one-launch path ------------------------------------------------------------------------------------------
constexpr int TOPK_CAND = 4096;   // most candidates the final sort takes
struct TopkGlobalWs {             // zeroed before the launch
  uint32_t hist[6][TOPK_BINS];
  uint32_t barrier;               // arrivals, monotone over the barriers of one launch
  uint32_t n_cand;
  uint32_t pad[2];
  uint64_t cand[TOPK_CAND];
};

__device__ __forceinline__ void topk_grid_barrier(uint32_t* counter, uint32_t target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(counter, 1u);
    uint32_t seen;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(counter) : "memory");
    } while (seen < target);
  }
  __syncthreads();
}

__global__ void __launch_bounds__(TOPK_THREADS) topk_global_kernel(const float* __restrict__ scores, int64_t n, int k,
                                                                   int64_t index_base, uint64_t* __restrict__ keys_out,
                                                                   TopkGlobalWs* __restrict__ ws) {
  __shared__ uint64_t sortbuf[TOPK_CAND];
  __shared__ uint32_t misc[40];   // [0] chosen digit, [1] keys above it, [2] keys in it, [8..40) warp sums
  uint32_t* hist = reinterpret_cast<uint32_t*>(sortbuf);   // the select rounds' histogram: dead before the final sort
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t G = gridDim.x;
  // slice of this CTA, a multiple of the block size so that every warp-wide vote below sees full warps
  const int64_t per = ((n + G - 1) / G + TOPK_THREADS - 1) / TOPK_THREADS * TOPK_THREADS;
  const int64_t lo = (int64_t)blockIdx.x * per, hi = lo + per;
  uint64_t prefix = 0ull, mask = 0ull;
  uint32_t kk = (uint32_t)k;
  uint32_t barriers = 0u;
#pragma unroll 1
  for (int r = 0; r < 6; ++r) {
    const int shift = r < 5 ? 52 - 12 * r : 0;
    const uint32_t nb = r < 5 ? 4096u : 16u;
    for (int i = tid; i < TOPK_BINS; i += TOPK_THREADS) hist[i] = 0u;
    __syncthreads();
    for (int64_t g = lo + tid; g < hi; g += TOPK_THREADS) {
      const bool valid = g < n;
      const uint64_t key = valid ? pack_key(__ldg(scores + g), (uint64_t)(index_base + g)) : 0ull;
      const bool in = valid && (key & mask) == prefix;
      const uint32_t digit = (uint32_t)(key >> shift) & (nb - 1u);
      const uint32_t act = __ballot_sync(0xffffffffu, in);
      if (in) {
        const uint32_t peers = __match_any_sync(act, digit);
        if (lane == __ffs(peers) - 1) atomicAdd(&hist[digit], (uint32_t)__popc(peers));
      }
    }
    __syncthreads();
    for (int i = tid; i < TOPK_BINS; i += TOPK_THREADS)
      if (hist[i] != 0u) atomicAdd(&ws->hist[r][i], hist[i]);
    topk_grid_barrier(&ws->barrier, ++barriers * G);
    // every CTA derives the same digit from the global histogram: suffix sums, four bins per thread
    const uint32_t b0 = (uint32_t)tid * 4u;
    const uint4 h = __ldcg(reinterpret_cast<const uint4*>(&ws->hist[r][b0]));
    const uint32_t h0 = h.x, h1 = h.y, h2 = h.z, h3 = h.w;
    uint32_t incl = h0 + h1 + h2 + h3;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t up = __shfl_down_sync(0xffffffffu, incl, o);
      if (lane + o < 32) incl += up;
    }
    if (lane == 0) misc[8 + warp] = incl;
    __syncthreads();
    uint32_t higher_warps = 0u;
    for (int w = warp + 1; w < TOPK_THREADS / 32; ++w) higher_warps += misc[8 + w];
    const uint32_t a3 = incl + higher_warps - (h0 + h1 + h2 + h3), a2 = a3 + h3, a1 = a2 + h2, a0 = a1 + h1;
    if (a3 < kk && kk <= a3 + h3) { misc[0] = b0 + 3; misc[1] = a3; misc[2] = h3; }
    else if (a2 < kk && kk <= a2 + h2) { misc[0] = b0 + 2; misc[1] = a2; misc[2] = h2; }
    else if (a1 < kk && kk <= a1 + h1) { misc[0] = b0 + 1; misc[1] = a1; misc[2] = h1; }
    else if (a0 < kk && kk <= a0 + h0) { misc[0] = b0; misc[1] = a0; misc[2] = h0; }
    __syncthreads();
    prefix |= (uint64_t)misc[0] << shift;
    mask |= (uint64_t)(nb - 1u) << shift;
    kk -= misc[1];
    // keys >= prefix (on the masked bits): the (k - kk) already known to be above, plus this bin
    const uint32_t at_or_above = ((uint32_t)k - kk) + misc[2];
    __syncthreads();
    if (at_or_above <= (uint32_t)TOPK_CAND) break;   // (the last round leaves exactly k)
  }
  // candidates: every key whose fixed bits are >= the prefix
  for (int64_t g = lo + tid; g < hi; g += TOPK_THREADS) {
    const bool valid = g < n;
    const uint64_t key = valid ? pack_key(__ldg(scores + g), (uint64_t)(index_base + g)) : 0ull;
    const bool keep = valid && (key & mask) >= prefix;
    const uint32_t act = __ballot_sync(0xffffffffu, keep);
    uint32_t pos = 0u;
    if (lane == 0 && act != 0u) pos = atomicAdd(&ws->n_cand, (uint32_t)__popc(act));
    pos = __shfl_sync(0xffffffffu, pos, 0);
    if (keep) {
      const uint32_t at = pos + (uint32_t)__popc(act & ((1u << lane) - 1u));
      if (at < (uint32_t)TOPK_CAND) ws->cand[at] = key;
    }
  }
  topk_grid_barrier(&ws->barrier, ++barriers * G);
  if (blockIdx.x != 0) return;
  uint32_t nc;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(nc) : "l"(&ws->n_cand) : "memory");
  nc = nc < (uint32_t)TOPK_CAND ? nc : (uint32_t)TOPK_CAND;
  if (nc <= (uint32_t)TOPK_THREADS) {   // few candidates (the usual case): rank sort, one pass and no barriers in it
    const uint64_t mine = tid < (int)nc ? __ldcg(&ws->cand[tid]) : 0ull;
    sortbuf[tid] = mine;
    __syncthreads();
    if (tid < (int)nc) {
      int rank = 0;   // keys are distinct: the number of larger keys is the output position
      for (int j = 0; j < (int)nc; ++j) rank += sortbuf[j] > mine ? 1 : 0;
      if (rank < k) keys_out[rank] = mine;
    }
    return;
  }
  int p2 = 1;
  while (p2 < (int)nc) p2 <<= 1;
  for (int i = tid; i < p2; i += TOPK_THREADS) sortbuf[i] = i < (int)nc ? __ldcg(&ws->cand[i]) : 0ull;
  __syncthreads();
  for (int size = 2; size <= p2; size <<= 1) {   // bitonic sort, descending
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int t = tid; t < p2 / 2; t += TOPK_THREADS) {
        const int l0 = ((t / stride) * (stride << 1)) + (t % stride);
        const int h0 = l0 + stride;
        const bool desc = ((l0 & size) == 0);
        const uint64_t x = sortbuf[l0], y = sortbuf[h0];
        if ((x < y) == desc) { sortbuf[l0] = y; sortbuf[h0] = x; }
      }
      __syncthreads();
    }
  }
  for (int i = tid; i < k; i += TOPK_THREADS) keys_out[i] = sortbuf[i];
}

__global__ void unpack_keys_kernel(const uint64_t* __restrict__ keys, int k, float* __restrict__ scores,
                                   int64_t* __restrict__ indices) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= k) return;
  const uint64_t key = keys[i];
  if (scores != nullptr) scores[i] = from_orderable_f32((uint32_t)(key >> 32));
  if (indices != nullptr) indices[i] = (int64_t)(0xFFFFFFFFu - (uint32_t)key);
}

static int64_t blocks_for(int64_t n) { return (n + TOPK_CHUNK - 1) / TOPK_CHUNK; }

static size_t topk_ws_bytes(int64_t n, int k) {
  // two ping-pong candidate buffers; the first level writes blocks_for(n) * k keys
  const int64_t lvl1 = blocks_for(n) * k;
  const int64_t lvl2 = blocks_for(lvl1) * k;
  const size_t levels = (size_t)(lvl1 + lvl2) * sizeof(uint64_t) + 256;
  return levels > sizeof(TopkGlobalWs) + 256 ? levels : sizeof(TopkGlobalWs) + 256;
}

static int run_topk(const float* scores, const uint64_t* keys, int64_t n, int k, int64_t index_base,
                    uint64_t* keys_out, void* workspace, size_t workspace_bytes, cudaStream_t st) {
  BFB_REQUIRE(k >= 1, BFB200_ERR_INVALID_ARG, "k must be >= 1 (got %d)", k);
  BFB_REQUIRE(n >= k, BFB200_ERR_INVALID_ARG, "selected index k out of range (k=%d > n=%lld)", k, (long long)n);
  BFB_REQUIRE(k <= TOPK_MAX_K, BFB200_ERR_UNSUPPORTED, "k=%d exceeds the supported maximum %d", k, TOPK_MAX_K);
  BFB_REQUIRE(index_base >= 0 && index_base + n <= 0xFFFFFFFFll, BFB200_ERR_UNSUPPORTED,
              "global indices must fit 32 bits");
  BFB_DEV(keys_out);
  int64_t nb = blocks_for(n);
  if (nb > 1) {
    BFB_REQUIRE(workspace_bytes >= topk_ws_bytes(n, k), BFB200_ERR_WORKSPACE,
                "top-k workspace too small: %zu < %zu", workspace_bytes, topk_ws_bytes(n, k));
    BFB_DEV(workspace);
  }
  uint64_t* buf_a = (uint64_t*)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
  if (scores != nullptr && nb > 1) {   // one cooperative launch: global-histogram radix select
    int dev = 0, sms = 0, coop = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
    if (coop && sms > 0) {
      TopkGlobalWs* gws = reinterpret_cast<TopkGlobalWs*>(buf_a);
      cudaError_t e = cudaMemsetAsync(gws, 0, offsetof(TopkGlobalWs, cand), st);
      BFB_REQUIRE(e == cudaSuccess, BFB200_ERR_CUDA, "memset failed: %s", cudaGetErrorString(e));
      int64_t grid = (n + TOPK_THREADS - 1) / TOPK_THREADS;
      if (grid > sms) grid = sms;
      void* args[] = {(void*)&scores, (void*)&n, (void*)&k, (void*)&index_base, (void*)&keys_out, (void*)&gws};
      e = cudaLaunchCooperativeKernel((const void*)topk_global_kernel, dim3((unsigned)grid), dim3(TOPK_THREADS), args, 0, st);
      BFB_REQUIRE(e == cudaSuccess, BFB200_ERR_CUDA, "cooperative launch of topk_global_kernel failed: %s", cudaGetErrorString(e));
      BFB_LAUNCH_CHECK("topk_global_kernel", st);
      return BFB200_OK;
    }
  }
  uint64_t* buf_b = buf_a + blocks_for(n) * k;
  const uint64_t* cur_keys = keys;
  bool from_scores = scores != nullptr;
  int64_t cur_n = n;
  uint64_t* dst = buf_a;
  // the opt-in for > 48 KB of dynamic shared memory is per device: set on every launch path (cheap), never cached
#define TOPK_LAUNCH(FS, FIN, SC, KI, IB, DST)                                                                      \
  do {                                                                                                             \
    cudaFuncSetAttribute(topk_select_kernel<FS, FIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TOPK_SMEM); \
    topk_select_kernel<FS, FIN><<<(int)nb, TOPK_THREADS, TOPK_SMEM, st>>>(SC, KI, cur_n, k, IB, DST);              \
  } while (0)
  while (true) {
    nb = blocks_for(cur_n);
    const bool final_level = nb == 1;
    if (final_level) dst = keys_out;
    if (from_scores) {
      if (final_level) TOPK_LAUNCH(true, true, scores, nullptr, index_base, dst);
      else             TOPK_LAUNCH(true, false, scores, nullptr, index_base, dst);
    } else {
      if (final_level) TOPK_LAUNCH(false, true, nullptr, cur_keys, 0, dst);
      else             TOPK_LAUNCH(false, false, nullptr, cur_keys, 0, dst);
    }
    BFB_LAUNCH_CHECK("topk_select_kernel", st);
    if (final_level) break;
    from_scores = false;
    cur_keys = dst;
    cur_n = nb * k;
    dst = (dst == buf_a) ? buf_b : buf_a;
  }
#undef TOPK_LAUNCH
  return BFB200_OK;
}

}  // namespace bfb

using namespace bfb;

extern "C" size_t bfb200_topk_workspace_bytes(int64_t n, int32_t k) {
  if (n <= 0 || k <= 0) return 256;
  return topk_ws_bytes(n, k);
}

extern "C" int bfb200_topk(const float* scores, int64_t n, int32_t k, int64_t index_base,
                           uint64_t* keys_out, void* workspace, size_t workspace_bytes, void* stream) {
  BFB_REQUIRE(n >= 0, BFB200_ERR_INVALID_ARG, "n < 0");
  if (n > 0) BFB_DEV(scores);
  return run_topk(scores, nullptr, n, k, index_base, keys_out, workspace, workspace_bytes, (cudaStream_t)stream);
}

extern "C" int bfb200_merge_topk(const uint64_t* keys, int64_t n, int32_t k, uint64_t* keys_out,
                                 void* workspace, size_t workspace_bytes, void* stream) {
  BFB_REQUIRE(n >= 0, BFB200_ERR_INVALID_ARG, "n < 0");
  if (n > 0) BFB_DEV(keys);
  return run_topk(nullptr, keys, n, k, 0, keys_out, workspace, workspace_bytes, (cudaStream_t)stream);
}

extern "C" int bfb200_unpack_keys(const uint64_t* keys, int32_t k, float* scores, int64_t* indices,
                                  void* stream) {
  BFB_REQUIRE(k >= 0, BFB200_ERR_INVALID_ARG, "k < 0");
  if (k == 0) return BFB200_OK;
  BFB_DEV(keys);
  BFB_DEV_OPT(scores);
  BFB_DEV_OPT(indices);
  cudaStream_t st = (cudaStream_t)stream;
  unpack_keys_kernel<<<(k + 255) / 256, 256, 0, st>>>(keys, k, scores, indices);
  BFB_LAUNCH_CHECK("unpack_keys_kernel", st);
  return BFB200_OK;
}
