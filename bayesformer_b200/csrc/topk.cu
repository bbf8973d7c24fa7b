// top-k selection (torch.topk of select_samples, src/libs/active_learning.py:210) on packed
// 64-bit keys.  Each block bitonic-sorts a chunk of CHUNK keys in shared memory and keeps the
// k largest; levels are repeated until one block remains.  HBM traffic: 4 B per score read
// once + 8 k B of candidates per block.
#include "common.cuh"

namespace bfb {

constexpr int TOPK_CHUNK = 8192;    // keys per block (64 KB of shared memory)
constexpr int TOPK_THREADS = 1024;
constexpr int TOPK_MAX_K = TOPK_CHUNK / 2;

template <bool FROM_SCORES>
__global__ void __launch_bounds__(TOPK_THREADS) topk_chunk_kernel(
    const float* __restrict__ scores, const uint64_t* __restrict__ keys_in, int64_t n, int k,
    int64_t index_base, uint64_t* __restrict__ keys_out) {
  extern __shared__ uint64_t skeys[];
  const int64_t base = (int64_t)blockIdx.x * TOPK_CHUNK;
  for (int i = threadIdx.x; i < TOPK_CHUNK; i += TOPK_THREADS) {
    const int64_t g = base + i;
    uint64_t key = 0ull;  // padding sorts last
    if (g < n) key = FROM_SCORES ? pack_key(__ldg(scores + g), (uint64_t)(index_base + g)) : keys_in[g];
    skeys[i] = key;
  }
  __syncthreads();
  // bitonic sort, descending
  for (int size = 2; size <= TOPK_CHUNK; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int t = threadIdx.x; t < TOPK_CHUNK / 2; t += TOPK_THREADS) {
        const int lo = ((t / stride) * (stride << 1)) + (t % stride);
        const int hi = lo + stride;
        const bool desc = ((lo & size) == 0);
        const uint64_t a = skeys[lo], b = skeys[hi];
        if ((a < b) == desc) { skeys[lo] = b; skeys[hi] = a; }
      }
      __syncthreads();
    }
  }
  for (int i = threadIdx.x; i < k; i += TOPK_THREADS) keys_out[(int64_t)blockIdx.x * k + i] = skeys[i];
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
  static bool attr_done = false;
  if (!attr_done) {
    cudaFuncSetAttribute(topk_chunk_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TOPK_CHUNK * 8);
    cudaFuncSetAttribute(topk_chunk_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TOPK_CHUNK * 8);
    attr_done = true;
  }
  const size_t smem = TOPK_CHUNK * sizeof(uint64_t);
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
  uint64_t* dst = nb == 1 ? keys_out : buf_a;
  while (true) {
    nb = blocks_for(cur_n);
    if (nb == 1) dst = keys_out;
    if (from_scores)
      topk_chunk_kernel<true><<<(int)nb, TOPK_THREADS, smem, st>>>(scores, nullptr, cur_n, k, index_base, dst);
    else
      topk_chunk_kernel<false><<<(int)nb, TOPK_THREADS, smem, st>>>(nullptr, cur_keys, cur_n, k, 0, dst);
    BFB_LAUNCH_CHECK("topk_chunk_kernel", st);
    if (nb == 1) break;
    from_scores = false;
    cur_keys = dst;
    cur_n = nb * k;
    dst = (dst == buf_a) ? buf_b : buf_a;
  }
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
