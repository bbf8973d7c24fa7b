// top-k selection (torch.topk of select_samples, src/libs/active_learning.py:210) on packed 64-bit keys
// (orderable(score) << 32 | ~index: all distinct, larger = better, ties towards the lowest index).
//
// A block takes a chunk of 16,384 keys into shared memory and finds its k-th largest key by RADIX SELECT — six
// histogram rounds over 12-bit digits, most significant first, each narrowing the candidates to the bin that holds
// the k-th key — then compacts the k keys >= that threshold.  Levels repeat until one block remains; the last one
// bitonic-sorts its k survivors (k <= 4096) for the output.  1 M scores, k = 200: 62 blocks + 1 block, two launches
// (the previous version bitonic-sorted 8,192 keys per block in three launches of ~90 us each).
// HBM traffic: 4 B per score read once + 8 k B of candidates per block.
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
  return (size_t)(lvl1 + lvl2) * sizeof(uint64_t) + 256;
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
