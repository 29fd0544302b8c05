// Issue rate of FFMA vs HFMA2 vs FFMA2 (packed fp32) on one SM: 8 independent accumulator chains per thread, 16 warps.
// Prints warp-instructions per cycle per SM sub-partition.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 fma_rate.cu
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

constexpr int ITERS = 4096, CH = 8;

__global__ void k_ffma(float* out, float a, float b, long long* cyc) {
  float acc[CH];
  for (int i = 0; i < CH; ++i) acc[i] = threadIdx.x + i;
  const long long t0 = clock64();
  for (int it = 0; it < ITERS; ++it)
#pragma unroll
    for (int i = 0; i < CH; ++i) acc[i] = fmaf(acc[i], a, b);
  const long long t1 = clock64();
  float s = 0;
  for (int i = 0; i < CH; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
__global__ void k_hfma2(__half2* out, float a, float b, long long* cyc) {
  __half2 acc[CH];
  const __half2 a2 = __float2half2_rn(a), b2 = __float2half2_rn(b);
  for (int i = 0; i < CH; ++i) acc[i] = __float2half2_rn((float)(threadIdx.x + i) * 1e-3f);
  const long long t0 = clock64();
  for (int it = 0; it < ITERS; ++it)
#pragma unroll
    for (int i = 0; i < CH; ++i) acc[i] = __hfma2(acc[i], a2, b2);
  const long long t1 = clock64();
  __half2 s = acc[0];
  for (int i = 1; i < CH; ++i) s = __hadd2(s, acc[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
__global__ void k_ffma2(float2* out, float a, float b, long long* cyc) {
  float2 acc[CH];
  const float2 a2 = make_float2(a, a), b2 = make_float2(b, b);
  for (int i = 0; i < CH; ++i) acc[i] = make_float2(threadIdx.x + i, threadIdx.x - i);
  const long long t0 = clock64();
  for (int it = 0; it < ITERS; ++it)
#pragma unroll
    for (int i = 0; i < CH; ++i) acc[i] = __ffma2_rn(acc[i], a2, b2);
  const long long t1 = clock64();
  float2 s = acc[0];
  for (int i = 1; i < CH; ++i) { s.x += acc[i].x; s.y += acc[i].y; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

int main() {
  void* out;
  long long* cyc;
  cudaMalloc(&out, 1 << 20);
  cudaMalloc(&cyc, 1024);
  long long h;
  const int threads = 512;  // 16 warps on one SM = 4 per sub-partition
  for (int rep = 0; rep < 2; ++rep) {
    k_ffma<<<1, threads>>>((float*)out, 0.999f, 0.001f, cyc);
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    if (rep) printf("FFMA   %.3f warp-instr/cycle/SMSP\n", 4.0 * ITERS * CH / h);
    k_hfma2<<<1, threads>>>((__half2*)out, 0.999f, 0.001f, cyc);
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    if (rep) printf("HFMA2  %.3f warp-instr/cycle/SMSP\n", 4.0 * ITERS * CH / h);
    k_ffma2<<<1, threads>>>((float2*)out, 0.999f, 0.001f, cyc);
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    if (rep) printf("FFMA2  %.3f warp-instr/cycle/SMSP\n", 4.0 * ITERS * CH / h);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
