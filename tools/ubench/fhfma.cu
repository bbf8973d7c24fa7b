// Issue-rate microbenchmark: FFMA vs FHFMA (fma.rn.f32.f16) vs HADD2.F32 unpack + FFMA, sm_100a.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fhfma fhfma.cu ; run: ./fhfma
#include <cstdio>
#include <cstdint>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
__device__ __forceinline__ float fh_ll(uint32_t a, uint32_t b, float c) {
  float d;
  asm volatile("{\n\t.reg .b16 al, ah, bl, bh;\n\tmov.b32 {al, ah}, %1;\n\tmov.b32 {bl, bh}, %2;\n\tfma.rn.f32.f16 %0, al, bl, %3;\n\t}" : "=f"(d) : "r"(a), "r"(b), "f"(c));
  return d;
}
__device__ __forceinline__ float fh_hh(uint32_t a, uint32_t b, float c) {
  float d;
  asm volatile("{\n\t.reg .b16 al, ah, bl, bh;\n\tmov.b32 {al, ah}, %1;\n\tmov.b32 {bl, bh}, %2;\n\tfma.rn.f32.f16 %0, ah, bh, %3;\n\t}" : "=f"(d) : "r"(a), "r"(b), "f"(c));
  return d;
}
template <int MODE>
__global__ void __launch_bounds__(512) k(const uint32_t* in, float* out, int iters) {
  uint32_t a = in[threadIdx.x], b = in[threadIdx.x + 512];
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = (float)i;
  float fa = __uint_as_float(a), fb = __uint_as_float(b);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) { asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(acc[i]) : "f"(fa), "f"(fb)); }
      else if (MODE == 1) acc[i] = (i & 1) ? fh_hh(a, b, acc[i]) : fh_ll(a, b, acc[i]);
      else {
        float2 x = __half22float2(*reinterpret_cast<const __half2*>(&a));
        asm volatile("" : "+f"(x.x), "+f"(x.y));
        asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(acc[i]) : "f"(x.x), "f"(x.y));
        a += 1;
      }
    }
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += acc[i];
  out[blockIdx.x * 512 + threadIdx.x] = s;
}
template <int MODE>
void run(const char* name, const uint32_t* in, float* out) {
  const int iters = 20000, grid = 148 * 4;
  k<MODE><<<grid, 512>>>(in, out, 100);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<MODE><<<grid, 512>>>(in, out, iters);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double ops = (double)grid * 512 * iters * 8;
  printf("%-28s %.3f ms  %.1f G fma-lane-ops/s  (%.1f per clk per SM at 1.965 GHz)\n", name, ms, ops / ms / 1e6, ops / (ms * 1e-3) / 148 / 1.965e9);
}
int main() {
  uint32_t* in; float* out;
  cudaMalloc(&in, 4096 * 4); cudaMemset(in, 0x3c, 4096 * 4); cudaMalloc(&out, 148 * 4 * 512 * 4);
  run<0>("FFMA", in, out);
  run<1>("FHFMA", in, out);
  run<2>("2x HADD2.F32 + FFMA", in, out);
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
