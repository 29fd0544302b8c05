// Micro-benchmark (developer aid): issue -> landed latency of the TMA boxes that fill one x stage of layer_fused.cu,
// for the record layout (9 boxes of 144 16-byte rows), the chunk-planar layout (3 boxes of 54 128-byte rows, or 9 of 18)
// and one raw halo box (18 rows of 480 B), cold (first touch: HBM) and warm (second pass: L2), 148 CTAs at once.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_box tma_box.cu -lcuda ; run on a B200.
#include <cstdio>
#include <cstdint>
#include <vector>
#include <algorithm>
#include <cuda.h>
#include <cuda_runtime.h>
#include "../../extended-gan_b200/csrc/tc_common.cuh"
namespace cgat { char* last_error_buf() { static char b[512]; return b; } }
using namespace cgat;

constexpr int H = 64, W = 64, N = 64, C = 24, NCH = 3;
constexpr int PLANE = 18 * 128;

__global__ void __launch_bounds__(32, 1) k(const __grid_constant__ CUtensorMap map, int mode, int iters, int depth, long long* out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bar[8];
  if (threadIdx.x == 0) {
    for (int i = 0; i < 8; ++i) mbar_init(&bar[i], 1);
    fence_mbar_init();
  }
  __syncthreads();
  if (threadIdx.x != 0) return;
  long long tsum = 0, tmax = 0, issue = 0;
  uint32_t phase = 0;
  const int tiles_w = W / 8, tiles_h = H / 16;
  for (int it = 0; it < iters; it += depth) {
    const long long t0 = clock64();
    for (int d = 0; d < depth; ++d) {  // `depth` tiles in flight, each on its own barrier
      const int tile = (blockIdx.x + (it + d) * gridDim.x) % (N * tiles_w * tiles_h);
      const int tw = tile % tiles_w, th = (tile / tiles_w) % tiles_h, n = tile / (tiles_w * tiles_h);
      unsigned char* dst = smem + d * 10 * PLANE;
      if (mode == 0) {  // records [n][h][w][24]: box (8 ch, 8 px, 18 rows, 1) per (shift, chunk)
        mbar_arrive_expect_tx(&bar[d], 9 * PLANE);
        for (int sh = 0; sh < 3; ++sh)
          for (int c = 0; c < NCH; ++c)
            tma_load_4d(dst + (sh * NCH + c) * PLANE, &map, c * 8, tw * 8 - 1 + sh, th * 16 - 1, n, &bar[d]);
      } else if (mode == 1) {  // planar [n][3][h][w*8]: box (64, 18, 3, 1) per shift
        mbar_arrive_expect_tx(&bar[d], 9 * PLANE);
        for (int sh = 0; sh < 3; ++sh) tma_load_4d(dst + sh * NCH * PLANE, &map, (tw * 8 - 1 + sh) * 8, th * 16 - 1, 0, n, &bar[d]);
      } else if (mode == 2) {  // planar, box (64, 18, 1, 1) per (shift, chunk)
        mbar_arrive_expect_tx(&bar[d], 9 * PLANE);
        for (int sh = 0; sh < 3; ++sh)
          for (int c = 0; c < NCH; ++c)
            tma_load_4d(dst + (sh * NCH + c) * PLANE, &map, (tw * 8 - 1 + sh) * 8, th * 16 - 1, c, n, &bar[d]);
      } else if (mode == 3) {  // planar, ONE raw halo box (80 = 10 px, 18, 3, 1): 160-byte rows, 8 640 B
        mbar_arrive_expect_tx(&bar[d], 160 * 18 * 3);
        tma_load_4d(dst, &map, (tw * 8 - 1) * 8, th * 16 - 1, 0, n, &bar[d]);
      } else {  // records, ONE raw halo box (24 ch, 10 px, 18 rows): 48-byte rows x 180
        mbar_arrive_expect_tx(&bar[d], 48 * 10 * 18);
        tma_load_4d(dst, &map, 0, tw * 8 - 1, th * 16 - 1, n, &bar[d]);
      }
    }
    const long long t1 = clock64();
    for (int d = 0; d < depth; ++d) mbar_wait(&bar[d], phase);
    const long long t2 = clock64();
    phase ^= 1u;
    issue += t1 - t0;
    tsum += t2 - t0;
    tmax = max(tmax, t2 - t0);
  }
  out[blockIdx.x * 3 + 0] = tsum;
  out[blockIdx.x * 3 + 1] = tmax;
  out[blockIdx.x * 3 + 2] = issue;
}

static CUtensorMap make_map(void* base, int mode) {
  CUtensorMap m;
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r;
  if (mode == 0 || mode == 4) {
    cuuint64_t dims[4] = {C, W, H, N};
    cuuint64_t strides[3] = {C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
    cuuint32_t box[4] = {8, 8, 18, 1};
    if (mode == 4) { box[0] = 24; box[1] = 10; }
    r = cuTensorMapEncodeTiled(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  } else {
    cuuint64_t dims[4] = {(cuuint64_t)W * 8, H, NCH, N};
    cuuint64_t strides[3] = {(cuuint64_t)W * 16, (cuuint64_t)H * W * 16, (cuuint64_t)NCH * H * W * 16};
    cuuint32_t box[4] = {64, 18, (cuuint32_t)(mode == 2 ? 1 : NCH), 1};
    if (mode == 3) box[0] = 80;
    r = cuTensorMapEncodeTiled(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  }
  if (r != CUDA_SUCCESS) printf("encode failed %d (mode %d)\n", (int)r, mode);
  return m;
}

int main() {
  cudaFree(0);
  void* x;
  const size_t bytes = (size_t)N * H * W * C * 2;
  cudaMalloc(&x, bytes);
  cudaMemset(x, 0, bytes);
  void* flush;
  cudaMalloc(&flush, 512u << 20);
  long long* out;
  cudaMalloc(&out, 148 * 3 * 8);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const char* names[5] = {"records 9 x (8ch,8,18)", "planar 3 x (64,18,3)", "planar 9 x (64,18,1)", "planar 1 raw (80,18,3)",
                          "records 1 raw (24,10,18)"};
  for (int mode = 0; mode < 5; ++mode) {
    CUtensorMap m = make_map(x, mode);
    for (int depth : {1, 2, 4})
      for (int warm = 0; warm < 2; ++warm) {
        const int iters = 12;  // 12 * 148 tiles < 2048: every tile is touched once per launch
        if (!warm) cudaMemset(flush, 1, 512u << 20);  // evict x from L2
        k<<<148, 32, 200 * 1024>>>(m, mode, iters, depth, out);
        cudaError_t e = cudaDeviceSynchronize();
        std::vector<long long> h(148 * 3);
        cudaMemcpy(h.data(), out, 148 * 3 * 8, cudaMemcpyDeviceToHost);
        double s = 0, is = 0;
        long long mx = 0;
        for (int i = 0; i < 148; ++i) { s += h[3 * i]; mx = std::max(mx, h[3 * i + 1]); is += h[3 * i + 2]; }
        printf("%-26s depth %d %s: issue->landed %.0f cycles per round of %d tile(s) (max %lld), issue alone %.0f  (%s)\n",
               names[mode], depth, warm ? "warm L2" : "cold   ", s / 148 / (iters / depth), depth, mx, is / 148 / (iters / depth),
               cudaGetErrorString(e));
      }
  }
  return 0;
}
