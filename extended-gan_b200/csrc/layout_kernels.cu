// Zero-pad + 2x2 space-to-depth in one pass (and its inverse): the activation side of the regrouping that turns the DCGAN
// discriminators' k = 4, stride 2, padding 1 convs (dcgan/model.py:150-165) into stride-1 2x2 convs over 4*C channels for
// the tcgen05 implicit-GEMM kernels (cgat.conv_layers.Conv2d._space_to_depth_route):
//   xs[n][i][j][(a, b, c)] = x[n][2i + a - 1][2j + b - 1][c]      (0 outside the image);  i < h/2 + 1, j < w/2 + 1
// Every x element appears exactly once in xs, so the backward is the inverse gather
//   dx[n][y][x][c] = dxs[n][(y+1)/2][(x+1)/2][((y+1)%2, (x+1)%2, c)].
// HBM-bound permutation: a thread moves one G-byte group of channels (G = 16, 8 or the element size, whatever divides
// the pixel's channel row); consecutive threads walk the DESTINATION contiguously.  PyTorch did this as F.pad (fill +
// strided copy) + view/permute/reshape (another strided copy) and two more copies in the backward.
#include "common.cuh"

namespace cgat {

template <typename G>
__global__ void __launch_bounds__(256) s2d_pad_kernel(const G* __restrict__ src, G* __restrict__ dst, long long n_dst, int h,
                                                        int w, int cg, int inverse) {
  griddep_wait();
  const int hs = h / 2 + 1, ws = w / 2 + 1;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_dst; i += (long long)gridDim.x * blockDim.x) {
    if (!inverse) {  // dst = xs [n][hs][ws][2][2][cg]
      const int c = (int)(i % cg);
      long long q = i / cg;
      const int b = (int)(q & 1), a = (int)((q >> 1) & 1);
      q >>= 2;
      const int j = (int)(q % ws);
      q /= ws;
      const int ii = (int)(q % hs);
      const long long n = q / hs;
      const int y = 2 * ii + a - 1, x = 2 * j + b - 1;
      G v{};
      if (y >= 0 && y < h && x >= 0 && x < w) v = src[((n * h + y) * w + x) * cg + c];
      dst[i] = v;
    } else {  // dst = dx [n][h][w][cg]
      const int c = (int)(i % cg);
      long long q = i / cg;
      const int x = (int)(q % w);
      q /= w;
      const int y = (int)(q % h);
      const long long n = q / h;
      const int ii = (y + 1) >> 1, a = (y + 1) & 1, j = (x + 1) >> 1, b = (x + 1) & 1;
      dst[i] = src[((((n * hs + ii) * ws + j) * 2 + a) * 2 + b) * cg + c];
    }
  }
}

}  // namespace cgat

using namespace cgat;

extern "C" int cgat_s2d_pad(const void* src, void* dst, int32_t dtype, int64_t n, int32_t h, int32_t w, int32_t c,
                            int32_t inverse, void* stream) {
  if (!src || !dst) return fail(CGAT_EINVAL, "null argument");
  if (n < 1 || h < 2 || w < 2 || (h & 1) || (w & 1) || c < 1) return fail(CGAT_EINVAL, "s2d_pad: even h, w >= 2 expected");
  if (dtype != CGAT_F32 && dtype != CGAT_BF16) return fail(CGAT_EINVAL, "bad dtype %d", dtype);
  const int esz = dtype == CGAT_F32 ? 4 : 2;
  const long long row = (long long)c * esz;  // bytes of a pixel's channel row
  const uintptr_t al = reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst);
  const int g = (row % 16 == 0 && al % 16 == 0) ? 16 : (row % 8 == 0 && al % 8 == 0) ? 8 : (row % 4 == 0 && al % 4 == 0) ? 4 : esz;
  const int cg = (int)(row / g);
  const long long pix = inverse ? (long long)n * h * w : (long long)n * (h / 2 + 1) * (w / 2 + 1) * 4;
  const long long n_dst = pix * cg;
  long long ctas = (n_dst + 255) / 256;
  if (ctas > 148 * 32) ctas = 148 * 32;
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e;
#define S2D(G) e = launch_pdl(s2d_pad_kernel<G>, dim3((unsigned)ctas), dim3(256), 0, st, (const G*)src, (G*)dst, n_dst, (int)h, (int)w, cg, (int)inverse)
  if (g == 16) S2D(uint4);
  else if (g == 8) S2D(uint2);
  else if (g == 4) S2D(uint32_t);
  else S2D(uint16_t);
#undef S2D
  if (e != cudaSuccess) return fail((int)e, "s2d_pad_kernel: %s", cudaGetErrorString(e));
  return check_launch("s2d_pad_kernel");
}
