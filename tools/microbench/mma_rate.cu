// Micro-benchmark (developer aid): cycles per tcgen05.mma for different smem layouts / N / majors.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_rate mma_rate.cu ; run on a B200.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../extended-gan_b200/csrc/tc_common.cuh"
namespace cgat { char* last_error_buf() { static char b[512]; return b; } }
using namespace cgat;

__device__ __forceinline__ uint64_t desc_sw(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
  uint64_t d = make_smem_desc(addr, lbo, sbo);
  d |= (uint64_t)layout << 61;
  return d;
}

// mode: 0 = no swizzle K-major (A rows 16B planes), 1 = SW64 K-major, 2 = SW128 K-major, 3 = no-swizzle MN-major both
__global__ void __launch_bounds__(128, 1) k(int mode, int N, int iters, int same_acc, long long* out, int walk = 0) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint32_t tptr;
  __shared__ uint64_t bar;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  fence_proxy_async_smem();
  if (threadIdx.x < 32) tmem_alloc(&tptr, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = tptr;
  if (threadIdx.x == 0) {
    const uint32_t a = smem_u32(smem), b = a + 64 * 1024;
    uint64_t ad, bd;
    uint32_t idesc;
    if (mode == 0) { ad = desc_sw(a, 2944, 160, 0); bd = desc_sw(b, N * 16, 128, 0); idesc = make_idesc_bf16(128, N, 0, 0); }
    else if (mode == 1) { ad = desc_sw(a, 16, 640, 4); bd = desc_sw(b, 16, 512, 4); idesc = make_idesc_bf16(128, N, 0, 0); }
    else if (mode == 2) { ad = desc_sw(a, 16, 1024, 2); bd = desc_sw(b, 16, 1024, 2); idesc = make_idesc_bf16(128, N, 0, 0); }
    else if (mode == 3) { ad = desc_sw(a, 128, 2048, 0); bd = desc_sw(b, 160, 2944, 0); idesc = make_idesc_bf16(128, N, 1, 1); }
    else if (mode == 4) { ad = desc_sw(a, 2304, 128, 0); bd = desc_sw(b, N * 16, 128, 0); idesc = make_idesc_bf16(128, N, 0, 0); }  // layer_fused fprop (column planes)
    else if (mode == 5) { ad = desc_sw(a, 128, 2048, 0); bd = desc_sw(b, 128, 2304, 0); idesc = make_idesc_bf16(128, N, 1, 1); }    // layer_fused wgrad
    else if (mode == 6) { ad = desc_sw(a, 2048, 128, 0); bd = desc_sw(b, N * 16, 128, 0); idesc = make_idesc_bf16(128, N, 0, 0); }  // dense K-major, 128 rows contiguous
    else { ad = desc_sw(a, 2304, 128, 0); bd = desc_sw(b, N * 16, 128, 0); idesc = make_idesc_bf16(64, N, 0, 0); }  // M = 64
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      const uint32_t col = same_acc ? 0 : (uint32_t)((i % (512 / N)) * N);
      if (walk) break;
      uint64_t ai = ad, bi = bd;
      umma_bf16(tb + col, ai, bi, idesc, 1);
    }
    if (walk) {
      for (int i = 0; i < iters; i += 15) {
#pragma unroll
        for (int kk = 0; kk < 15; ++kk) {
          const int r = kk / 5, q = 2 * (kk % 5), jj = kk % 8;
          uint64_t ai = ad, bi = bd;
          if (walk == 1) { ai += (uint64_t)((q * 2304 + r * 128) >> 4); bi += (uint64_t)((kk * 2 * N * 16) >> 4); }
          if (walk == 2) { ai += (uint64_t)((jj * 256) >> 4); bi += (uint64_t)((jj * 256) >> 4); }
          if (walk == 3) ai += (uint64_t)((q * 2304 + r * 128) >> 4);
          if (walk == 4) bi += (uint64_t)((kk * 2 * N * 16) >> 4);
          if (walk == 5) ai += (uint64_t)((kk * 4096) >> 4);             // A walks by whole 4 KB blocks
          if (walk == 6) ai += (uint64_t)((q * 2304) >> 4);              // A walks by planes only (no row shift)
          if (walk == 7) ai += (uint64_t)((r * 128) >> 4);               // A walks by rows only
          umma_bf16(tb, ai, bi, idesc, 1);
        }
      }
    }
    long long t1 = clock64();
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    long long t2 = clock64();
    if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tb, 512);
}

int main() {
  long long* out;
  cudaMalloc(&out, 16);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
  const char* names[8] = {"none/K-major", "SW64/K-major", "SW128/K-major", "none/MN-major", "lf fprop", "lf wgrad", "none/K dense", "lf fprop M64"};
  for (int mode = 0; mode < 8; ++mode)
    for (int N : {32, 96, 128, 240, 256})
      for (int same : {1, 0}) {
        const int iters = 512;
        k<<<148, 128, 160 * 1024>>>(mode, N, iters, same, out);
        cudaError_t e = cudaDeviceSynchronize();
        long long h[2];
        cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
        printf("%-14s N=%3d same_acc=%d  issue %.1f cyc/mma  complete %.1f cyc/mma  (%s)\n", names[mode], N, same,
               (double)h[0] / iters, (double)h[1] / iters, cudaGetErrorString(e));
      }
  for (int walk = 1; walk <= 7; ++walk)
    for (int N : {80, 96, 128}) {
      const int mode = walk == 2 ? 5 : 4, iters = 480;
      k<<<148, 128, 160 * 1024>>>(mode, N, iters, 1, out, walk);
      cudaError_t e = cudaDeviceSynchronize();
      long long h[2];
      cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
      printf("walk %d mode %d N=%3d  issue %.1f cyc/mma  complete %.1f cyc/mma  (%s)\n", walk, mode, N, (double)h[0] / iters,
             (double)h[1] / iters, cudaGetErrorString(e));
    }
  return 0;
}
