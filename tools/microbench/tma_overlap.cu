// Developer probe: is a tensor map whose dimension 1 OVERLAPS dimension 0 (stride 16 B < 128-byte inner extent) accepted,
// and does the box land as expected?  x padded-planar [n][c][h][W+2][8] bf16; box (64 el, 3 shifts, 3 chunks, 18 rows, 1).
#include <cstdio>
#include <cstdint>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include "../../extended-gan_b200/csrc/tc_common.cuh"
namespace cgat { char* last_error_buf() { static char b[512]; return b; } }
using namespace cgat;
constexpr int H = 64, W = 64, N = 4, NCH = 3, WP = W + 2;

__device__ __forceinline__ void tma_load_5d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, int c4, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
               ::"r"(smem_u32(dst)), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4), "r"(smem_u32(bar)) : "memory");
}

__global__ void k(const __grid_constant__ CUtensorMap map, int tw, int th, int n, uint16_t* out, long long* cyc) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bar;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  __syncthreads();
  if (threadIdx.x == 0) {
    long long t0 = clock64();
    mbar_arrive_expect_tx(&bar, 18 * 9 * 128);
    tma_load_5d(smem, &map, tw * 8 * 8, 0, 0, th * 16 - 1, n, &bar);
    mbar_wait(&bar, 0);
    cyc[0] = clock64() - t0;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 18 * 9 * 64; i += blockDim.x) out[i] = reinterpret_cast<uint16_t*>(smem)[i];
}

int main() {
  cudaFree(0);
  const size_t elems = (size_t)N * NCH * H * WP * 8;
  std::vector<uint16_t> h(elems);
  // value encodes (c, h, wp, e) so the landing can be checked: bf16 bits = arbitrary 16-bit pattern
  for (int n = 0; n < N; ++n) for (int c = 0; c < NCH; ++c) for (int y = 0; y < H; ++y) for (int x = 0; x < WP; ++x) for (int e = 0; e < 8; ++e)
    h[((((size_t)n * NCH + c) * H + y) * WP + x) * 8 + e] = (uint16_t)((c << 14) | (y << 8) | (x << 1) | (e & 1)) ^ (uint16_t)(n * 7);
  uint16_t* d; cudaMalloc(&d, elems * 2); cudaMemcpy(d, h.data(), elems * 2, cudaMemcpyHostToDevice);
  CUtensorMap m;
  cuuint64_t dims[5] = {(cuuint64_t)WP * 8, 3, NCH, H, N};
  cuuint64_t strides[4] = {16, (cuuint64_t)H * WP * 16, (cuuint64_t)WP * 16, (cuuint64_t)NCH * H * WP * 16};
  cuuint32_t box[5] = {64, 3, NCH, 18, 1};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = cuTensorMapEncodeTiled(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, d, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                      CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode: CUresult %d\n", (int)r);
  if (r != CUDA_SUCCESS) return 1;
  uint16_t* out; cudaMalloc(&out, 18 * 9 * 64 * 2);
  long long* cyc; cudaMalloc(&cyc, 8);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  for (int trial = 0; trial < 3; ++trial) {
    const int tw = trial == 0 ? 0 : (trial == 1 ? 7 : 3), th = trial == 0 ? 0 : (trial == 1 ? 3 : 1), n = trial;
    k<<<1, 128, 64 * 1024>>>(m, tw, th, n, out, cyc);
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<uint16_t> o(18 * 9 * 64);
    long long c; cudaMemcpy(o.data(), out, o.size() * 2, cudaMemcpyDeviceToHost); cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    long bad = 0;
    // expected landing: [row 18][c 3][s 3][8 px][8 e];  source pixel column (padded) = tw*8 + s + px, row = th*16 - 1 + row
    for (int row = 0; row < 18; ++row) for (int c2 = 0; c2 < 3; ++c2) for (int s = 0; s < 3; ++s) for (int px = 0; px < 8; ++px) for (int e2 = 0; e2 < 8; ++e2) {
      const int y = th * 16 - 1 + row, x = tw * 8 + s + px;
      uint16_t want = 0;
      if (y >= 0 && y < H) want = (uint16_t)((c2 << 14) | (y << 8) | (x << 1) | (e2 & 1)) ^ (uint16_t)(n * 7);
      const uint16_t got = o[(((row * 3 + c2) * 3 + s) * 8 + px) * 8 + e2];
      if (got != want) { if (bad < 5) printf("  mismatch row %d c %d s %d px %d e %d: got %04x want %04x\n", row, c2, s, px, e2, got, want); ++bad; }
    }
    printf("trial %d (tw %d th %d n %d): %s, %ld mismatches, %lld cycles (%s)\n", trial, tw, th, n, bad ? "WRONG" : "ok", bad, c, cudaGetErrorString(e));
  }
  return 0;
}
