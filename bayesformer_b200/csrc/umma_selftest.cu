// Bring-up check of the tcgen05 plumbing in umma.cuh: one 128 x N x 64 bf16 tile product,
// C = A * W^T, operands staged by the threads into swizzled shared memory exactly as the fused
// MC-forward kernel stages them.  Exported for tests only (not part of the public header).
#include "common.cuh"
#include "umma.cuh"

namespace bfb {

__global__ void __launch_bounds__(128) umma_tile_kernel(const float* __restrict__ A, const float* __restrict__ W,
                                                        int N, float* __restrict__ out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;                 // 128 rows x 128 B
  uint8_t* sW = smem + 128 * 128;     // N rows x 128 B (N <= 256)
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  uint32_t n_cols = 32;
  while ((int)n_cols < N) n_cols <<= 1;

  if (warp == 0) {
    umma::tmem_alloc(&tmem_slot, n_cols);
    umma::tmem_relinquish();
  }
  if (tid == 0) {
    umma::mbar_init(&bar, 1);
    umma::fence_mbar_init();
  }
  // stage A (thread = row) and W rows as bf16, swizzled
  for (int c = 0; c < 8; ++c) {
    const float4 lo = *reinterpret_cast<const float4*>(A + tid * 64 + c * 8);
    const float4 hi = *reinterpret_cast<const float4*>(A + tid * 64 + c * 8 + 4);
    uint4 v;
    v.x = umma::pack_bf16x2(lo.x, lo.y); v.y = umma::pack_bf16x2(lo.z, lo.w);
    v.z = umma::pack_bf16x2(hi.x, hi.y); v.w = umma::pack_bf16x2(hi.z, hi.w);
    *reinterpret_cast<uint4*>(sA + umma::sw128_offset(tid, c)) = v;
  }
  for (int n = tid; n < N; n += 128) {
    for (int c = 0; c < 8; ++c) {
      const float4 lo = *reinterpret_cast<const float4*>(W + n * 64 + c * 8);
      const float4 hi = *reinterpret_cast<const float4*>(W + n * 64 + c * 8 + 4);
      uint4 v;
      v.x = umma::pack_bf16x2(lo.x, lo.y); v.y = umma::pack_bf16x2(lo.z, lo.w);
      v.z = umma::pack_bf16x2(hi.x, hi.y); v.w = umma::pack_bf16x2(hi.z, hi.w);
      *reinterpret_cast<uint4*>(sW + umma::sw128_offset(n, c)) = v;
    }
  }
  umma::fence_async_smem();
  umma::tc_fence_before_sync();
  __syncthreads();
  umma::tc_fence_after_sync();
  const uint32_t tmem_base = tmem_slot;
  if (tid == 0) {
    const uint64_t da = umma::smem_desc_sw128(umma::smem_u32(sA));
    const uint64_t db = umma::smem_desc_sw128(umma::smem_u32(sW));
    const uint32_t idesc = umma::instr_desc_bf16(128, N);
#pragma unroll
    for (int k = 0; k < 4; ++k) umma::mma_bf16_ss(tmem_base, da + 2 * k, db + 2 * k, idesc, k > 0);
    umma::mma_commit(&bar);
  }
  umma::mbar_wait(&bar, 0);
  umma::tc_fence_after_sync();
  const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
  for (int c = 0; c < N; c += 8) {
    uint32_t r[8];
    umma::tmem_ld8(tmem_base + lane_base + c, r);
    umma::tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 8; ++j) out[tid * N + c + j] = __uint_as_float(r[j]);
  }
  umma::tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tmem_base, n_cols);
}

}  // namespace bfb

extern "C" int bfb200_debug_umma_tile(const float* A, const float* W, int32_t N, float* out, void* stream) {
  using namespace bfb;
  BFB_REQUIRE(N >= 16 && N <= 256 && N % 16 == 0, BFB200_ERR_INVALID_ARG, "N must be a multiple of 16 in [16, 256]");
  BFB_DEV(A); BFB_DEV(W); BFB_DEV(out);
  cudaStream_t st = (cudaStream_t)stream;
  const size_t smem = 1024 + 128 * 128 + (size_t)N * 128;
  cudaFuncSetAttribute(umma_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  umma_tile_kernel<<<1, 128, smem, st>>>(A, W, N, out);
  BFB_LAUNCH_CHECK("umma_tile_kernel", st);
  return BFB200_OK;
}
