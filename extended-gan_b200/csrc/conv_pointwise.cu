// Pointwise (1x1, stride 1) convolutions as shared-memory tiled GEMMs on the CUDA cores: the fallback for the shapes the
// tcgen05 kernels of conv_tc.cu do not serve yet (fp32 tensors; K*N too large for resident packed weights, e.g. the
// SmaAt-UNet pointwise convs with up to 1024 input channels, 93 % of that net's flops -- SURVEY.md 8a row a10).  Replaces
// the one-thread-per-output direct kernels for this class (their wgrad re-read x once per output channel).
//   fprop  y[p][co]  = act(sum_ci x[p][ci] w[co][ci] + bias[co])          C = A . B^T      (K = cin  contiguous in both)
//   dgrad  dx[p][ci] = sum_co dy[p][co] w[co][ci]                         C = A . B        (B is [K][N])
//   wgrad  dw[co][ci] = sum_p dy[p][co] x[p][ci]                          C = A^T . B      (K = pixels, split over CTAs,
//                                                                         fp32 atomicAdd into the zeroed dw)
// 64 x 64 x 16 tiles, 256 threads, 4 x 4 outputs per thread, fp32 accumulation.
#include "common.cuh"

namespace cgat {

constexpr int PW_BM = 64, PW_BN = 64, PW_BK = 16, PW_THREADS = 256;

__device__ __forceinline__ float pw_act(float v, int act) {
  switch (act) {
    case 1: return fmaxf(v, 0.f);
    case 2: return v > 0.f ? v : 0.2f * v;
    case 3: return 1.f / (1.f + __expf(-v));
    default: return v;
  }
}

// MODE 0: A[m][k] (lda), B[n][k] (ldb).  MODE 1: A[m][k], B[k][n].  MODE 2: A[k][m], B[k][n] (+ split-K, atomics).
template <typename T, int MODE>
__global__ void __launch_bounds__(PW_THREADS)
pointwise_gemm_kernel(const T* __restrict__ A, const T* __restrict__ B, void* __restrict__ Cv, const float* __restrict__ bias,
                      long long M, int N, long long K, long long lda, long long ldb, long long ldc, int act,
                      long long k_per_split) {
  __shared__ float As[PW_BK][PW_BM + 4];
  __shared__ float Bs[PW_BK][PW_BN + 4];
  const int tid = threadIdx.x;
  const long long m0 = (long long)blockIdx.y * PW_BM;
  const int n0 = blockIdx.x * PW_BN;
  const long long kbeg = MODE == 2 ? (long long)blockIdx.z * k_per_split : 0;
  const long long kend = MODE == 2 ? (kbeg + k_per_split < K ? kbeg + k_per_split : K) : K;
  const int ty = tid / 16, tx = tid % 16;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (long long k0 = kbeg; k0 < kend; k0 += PW_BK) {
    // ---- stage A: As[k][m] ----
    if (MODE == 2) {  // A[k][m]: m contiguous
      const int k = tid / 16, m4 = (tid % 16) * 4;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const long long kk = k0 + k, mm = m0 + m4 + j;
        As[k][m4 + j] = (kk < kend && mm < M) ? DT<T>::to_f(A[kk * lda + mm]) : 0.f;
      }
    } else {  // A[m][k]: k contiguous
      const int m = tid / 4, k4 = (tid % 4) * 4;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const long long kk = k0 + k4 + j, mm = m0 + m;
        As[k4 + j][m] = (kk < kend && mm < M) ? DT<T>::to_f(A[mm * lda + kk]) : 0.f;
      }
    }
    // ---- stage B: Bs[k][n] ----
    if (MODE == 0) {  // B[n][k]: k contiguous
      const int n = tid / 4, k4 = (tid % 4) * 4;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const long long kk = k0 + k4 + j;
        const int nn = n0 + n;
        Bs[k4 + j][n] = (kk < kend && nn < N) ? DT<T>::to_f(B[(long long)nn * ldb + kk]) : 0.f;
      }
    } else {  // B[k][n]: n contiguous
      const int k = tid / 16, n4 = (tid % 16) * 4;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const long long kk = k0 + k;
        const int nn = n0 + n4 + j;
        Bs[k][n4 + j] = (kk < kend && nn < N) ? DT<T>::to_f(B[kk * ldb + nn]) : 0.f;
      }
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < PW_BK; ++kk) {
      const float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long long mm = m0 + ty * 4 + i;
    if (mm >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int nn = n0 + tx * 4 + j;
      if (nn >= N) continue;
      if (MODE == 2) {
        atomicAdd(reinterpret_cast<float*>(Cv) + mm * ldc + nn, acc[i][j]);
      } else {
        float v = acc[i][j];
        if (MODE == 0) v = pw_act(v + (bias ? bias[nn] : 0.f), act);
        reinterpret_cast<T*>(Cv)[mm * ldc + nn] = DT<T>::from_f(v);
      }
    }
  }
}

int conv_is_pointwise(const cgat_conv_desc* d) {
  return d->kh == 1 && d->kw == 1 && d->stride == 1 && d->pad_top == 0 && d->pad_left == 0 && d->groups == 1 &&
         d->ho == d->h && d->wo == d->w;
}

template <typename T>
static int pw_launch(int which, const cgat_conv_desc* d, const void* a, const void* b, void* c, const float* bias,
                     cudaStream_t st) {
  const long long P = (long long)d->n * d->h * d->w;
  if (which == 0) {  // fprop: M = P, N = cout, K = cin
    dim3 grid((d->cout + PW_BN - 1) / PW_BN, (unsigned)((P + PW_BM - 1) / PW_BM), 1);
    pointwise_gemm_kernel<T, 0><<<grid, PW_THREADS, 0, st>>>((const T*)a, (const T*)b, c, bias, P, d->cout, d->cin, d->cin,
                                                             d->cin, d->cout, d->act, 0);
  } else if (which == 1) {  // dgrad: M = P, N = cin, K = cout; B = w [cout][cin]
    dim3 grid((d->cin + PW_BN - 1) / PW_BN, (unsigned)((P + PW_BM - 1) / PW_BM), 1);
    pointwise_gemm_kernel<T, 1><<<grid, PW_THREADS, 0, st>>>((const T*)a, (const T*)b, c, nullptr, P, d->cin, d->cout,
                                                             d->cout, d->cin, d->cin, 0, 0);
  } else {  // wgrad: M = cout, N = cin, K = P split over CTAs; A = dy [P][cout], B = x [P][cin]
    const int tiles = ((d->cout + PW_BM - 1) / PW_BM) * ((d->cin + PW_BN - 1) / PW_BN);
    long long splits = (148LL * 4 + tiles - 1) / tiles;
    const long long max_splits = (P + 4 * PW_BK - 1) / (4 * PW_BK);
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
    long long per = (P + splits - 1) / splits;
    per = (per + PW_BK - 1) / PW_BK * PW_BK;
    splits = (P + per - 1) / per;
    cudaMemsetAsync(c, 0, sizeof(float) * (size_t)d->cout * d->cin, st);
    dim3 grid((d->cin + PW_BN - 1) / PW_BN, (d->cout + PW_BM - 1) / PW_BM, (unsigned)splits);
    pointwise_gemm_kernel<T, 2><<<grid, PW_THREADS, 0, st>>>((const T*)a, (const T*)b, c, nullptr, d->cout, d->cin, P,
                                                             d->cout, d->cin, d->cin, 0, per);
  }
  return check_launch("pointwise_gemm_kernel");
}

// which: 0 fprop (a = x, b = w, c = y), 1 dgrad (a = dy, b = w, c = dx), 2 wgrad (a = dy, b = x, c = dw fp32)
int conv_pointwise_launch(int which, const cgat_conv_desc* d, const void* a, const void* b, void* c, const float* bias,
                          cudaStream_t st) {
  return d->dtype == CGAT_F32 ? pw_launch<float>(which, d, a, b, c, bias, st)
                              : pw_launch<__nv_bfloat16>(which, d, a, b, c, bias, st);
}

}  // namespace cgat
