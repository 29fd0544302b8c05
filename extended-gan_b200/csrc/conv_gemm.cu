// CUDA-core implicit-GEMM convolution (any kernel size / stride / padding, groups == 1, fp32 or bf16 tensors, fp32
// accumulate): the fallback for every shape the tcgen05 kernels of conv_tc.cu do not serve yet -- fp32 tensors, the
// DCGAN nets at ndf = 64 (packed weights too large to stay resident), stride 4, cin = 4.  Replaces the
// one-thread-per-output direct kernels for dense convs (conv_direct.cu keeps the grouped / depthwise cases): those
// re-read every input once per output channel (DCGAN adversarial step at N = 64: 400 ms; with this file: see DESIGN.md).
//
//   fprop  y[p][co]   = act(sum_{kh,kw,ci} x[n, ho*s+kh-pt, wo*s+kw-pl, ci] w[co][kh][kw][ci] + bias)   M = out pixels, K = kh*kw*cin
//   dgrad  dx[q][ci]  = sum_{kh,kw,co} dy[n, (hi+pt-kh)/s, (wi+pl-kw)/s, co] w[co][kh][kw][ci]          M = in pixels,  K = kh*kw*cout
//   wgrad  dw[co][kh][kw][ci] = sum_p dy[p][co] x[n, ho*s+kh-pt, wo*s+kw-pl, ci]                        M = cout, N = kh*kw*cin, K = out pixels
//          (K split over CTAs, fp32 atomicAdd into the zeroed dw)
// 64 x 64 x 16 tiles in shared memory, 256 threads, 4 x 4 outputs per thread; the im2col gather happens in the tile
// loaders (the per-thread row/column decode is hoisted out of the K loop).
#include "common.cuh"

namespace cgat {

constexpr int CG_BM = 64, CG_BN = 64, CG_BK = 16, CG_THREADS = 256;

__device__ __forceinline__ float cg_act(float v, int act) {
  switch (act) {
    case 1: return fmaxf(v, 0.f);
    case 2: return v > 0.f ? v : 0.2f * v;
    case 3: return 1.f / (1.f + __expf(-v));
    default: return v;
  }
}

struct CgPix { int n, h, w; bool ok; };
__device__ __forceinline__ CgPix cg_decode(long long p, long long total, int H, int W) {
  CgPix r;
  r.ok = p < total;
  const long long q = r.ok ? p : 0;
  r.w = (int)(q % W);
  r.h = (int)((q / W) % H);
  r.n = (int)(q / ((long long)W * H));
  return r;
}

template <typename T, int MODE>
__global__ void __launch_bounds__(CG_THREADS)
conv_gemm_kernel(const cgat_conv_desc d, const T* __restrict__ A, const T* __restrict__ B, void* __restrict__ Cv,
                 const float* __restrict__ bias, long long k_per_split) {
  __shared__ float As[CG_BK][CG_BM + 4];
  __shared__ float Bs[CG_BK][CG_BN + 4];
  const int tid = threadIdx.x, ty = tid / 16, tx = tid % 16;
  const int taps = d.kh * d.kw;
  const long long P_out = (long long)d.n * d.ho * d.wo, P_in = (long long)d.n * d.h * d.w;
  const long long M = MODE == 0 ? P_out : (MODE == 1 ? P_in : d.cout);
  const int N = MODE == 0 ? d.cout : (MODE == 1 ? d.cin : taps * d.cin);
  const long long K = MODE == 0 ? (long long)taps * d.cin : (MODE == 1 ? (long long)taps * d.cout : P_out);
  const long long m0 = (long long)blockIdx.y * CG_BM;
  const int n0 = blockIdx.x * CG_BN;
  const long long kbeg = MODE == 2 ? (long long)blockIdx.z * k_per_split : 0;
  const long long kend = MODE == 2 ? (kbeg + k_per_split < K ? kbeg + k_per_split : K) : K;

  // ---- per-thread decode that does not change along K ----
  const int a_row = MODE == 2 ? (tid % 16) * 4 : tid / 4;  // MODE 2: m index base (4 consecutive m), else the tile row
  CgPix apix{};                                            // MODE 0 / 1: the pixel of this thread's A row
  if (MODE != 2) apix = cg_decode(m0 + a_row, M, MODE == 0 ? d.ho : d.h, MODE == 0 ? d.wo : d.w);
  int bn_tap[4], bn_ci[4];                                 // MODE 2: (tap, ci) of this thread's 4 B columns
  if (MODE == 2) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int nn = n0 + (tid % 16) * 4 + j;
      bn_tap[j] = nn < N ? nn / d.cin : -1;
      bn_ci[j] = nn < N ? nn % d.cin : 0;
    }
  }

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (long long k0 = kbeg; k0 < kend; k0 += CG_BK) {
    if (MODE == 0) {
      // A[m][k] = x gathered: thread = (row m, 4 consecutive k)
      const int k4 = (tid % 4) * 4;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const long long kk = k0 + k4 + j;
        float v = 0.f;
        if (apix.ok && kk < kend) {
          const int tap = (int)(kk / d.cin), ci = (int)(kk % d.cin);
          const int hi = apix.h * d.stride + tap / d.kw - d.pad_top, wi = apix.w * d.stride + tap % d.kw - d.pad_left;
          if (hi >= 0 && hi < d.h && wi >= 0 && wi < d.w)
            v = DT<T>::to_f(A[(((long long)apix.n * d.h + hi) * d.w + wi) * d.cin + ci]);
        }
        As[k4 + j][a_row] = v;
      }
      // B[n][k] = w[co][k]: k contiguous
      const int n = tid / 4;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const long long kk = k0 + k4 + j;
        const int nn = n0 + n;
        Bs[k4 + j][n] = (kk < kend && nn < N) ? DT<T>::to_f(B[(long long)nn * K + kk]) : 0.f;
      }
    } else if (MODE == 1) {
      // A[m][k] = dy gathered at (kh, kw, co): thread = (input pixel m, 4 consecutive k)
      const int k4 = (tid % 4) * 4;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const long long kk = k0 + k4 + j;
        float v = 0.f;
        if (apix.ok && kk < kend) {
          const int tap = (int)(kk / d.cout), co = (int)(kk % d.cout);
          const int hn = apix.h + d.pad_top - tap / d.kw, wn = apix.w + d.pad_left - tap % d.kw;
          if (hn >= 0 && wn >= 0 && hn % d.stride == 0 && wn % d.stride == 0) {
            const int ho = hn / d.stride, wo = wn / d.stride;
            if (ho < d.ho && wo < d.wo) v = DT<T>::to_f(A[(((long long)apix.n * d.ho + ho) * d.wo + wo) * d.cout + co]);
          }
        }
        As[k4 + j][a_row] = v;
      }
      // B[k][n] = w[co][tap][ci]: n = ci contiguous
      const int k = tid / 16, n4 = (tid % 16) * 4;
      const long long kk = k0 + k;
      const int tap = kk < kend ? (int)(kk / d.cout) : 0, co = kk < kend ? (int)(kk % d.cout) : 0;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int nn = n0 + n4 + j;
        Bs[k][n4 + j] = (kk < kend && nn < N) ? DT<T>::to_f(B[((long long)co * taps + tap) * d.cin + nn]) : 0.f;
      }
    } else {
      // A[k][m] = dy[p][co]: m = co contiguous;  B[k][n] = x gathered at output pixel p, column (tap, ci)
      const int k = tid / 16;
      const long long pp = k0 + k;
      const CgPix op = cg_decode(pp, kend, d.ho, d.wo);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const long long mm = m0 + a_row + j;
        As[k][a_row + j] = (op.ok && mm < M) ? DT<T>::to_f(A[pp * d.cout + mm]) : 0.f;
        float v = 0.f;
        if (op.ok && bn_tap[j] >= 0) {
          const int hi = op.h * d.stride + bn_tap[j] / d.kw - d.pad_top, wi = op.w * d.stride + bn_tap[j] % d.kw - d.pad_left;
          if (hi >= 0 && hi < d.h && wi >= 0 && wi < d.w)
            v = DT<T>::to_f(B[(((long long)op.n * d.h + hi) * d.w + wi) * d.cin + bn_ci[j]]);
        }
        Bs[k][(tid % 16) * 4 + j] = v;
      }
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < CG_BK; ++kk) {
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
        atomicAdd(reinterpret_cast<float*>(Cv) + mm * N + nn, acc[i][j]);
      } else {
        float v = acc[i][j];
        if (MODE == 0) v = cg_act(v + (bias ? bias[nn] : 0.f), d.act);
        reinterpret_cast<T*>(Cv)[mm * N + nn] = DT<T>::from_f(v);
      }
    }
  }
}

int conv_gemm_served(const cgat_conv_desc* d) { return d->groups == 1; }

template <typename T>
static int cg_launch(int which, const cgat_conv_desc* d, const void* a, const void* b, void* c, const float* bias,
                     cudaStream_t st) {
  const int taps = d->kh * d->kw;
  const long long P_out = (long long)d->n * d->ho * d->wo, P_in = (long long)d->n * d->h * d->w;
  if (which == 0) {
    dim3 grid((d->cout + CG_BN - 1) / CG_BN, (unsigned)((P_out + CG_BM - 1) / CG_BM), 1);
    conv_gemm_kernel<T, 0><<<grid, CG_THREADS, 0, st>>>(*d, (const T*)a, (const T*)b, c, bias, 0);
  } else if (which == 1) {
    dim3 grid((d->cin + CG_BN - 1) / CG_BN, (unsigned)((P_in + CG_BM - 1) / CG_BM), 1);
    conv_gemm_kernel<T, 1><<<grid, CG_THREADS, 0, st>>>(*d, (const T*)a, (const T*)b, c, nullptr, 0);
  } else {
    const int N = taps * d->cin;
    const int tiles = ((d->cout + CG_BM - 1) / CG_BM) * ((N + CG_BN - 1) / CG_BN);
    long long splits = (148LL * 4 + tiles - 1) / tiles;
    const long long max_splits = (P_out + 4 * CG_BK - 1) / (4 * CG_BK);
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
    long long per = (P_out + splits - 1) / splits;
    per = (per + CG_BK - 1) / CG_BK * CG_BK;
    splits = (P_out + per - 1) / per;
    cudaMemsetAsync(c, 0, sizeof(float) * (size_t)d->cout * N, st);
    dim3 grid((N + CG_BN - 1) / CG_BN, (d->cout + CG_BM - 1) / CG_BM, (unsigned)splits);
    conv_gemm_kernel<T, 2><<<grid, CG_THREADS, 0, st>>>(*d, (const T*)a, (const T*)b, c, nullptr, per);
  }
  return check_launch("conv_gemm_kernel");
}

// which: 0 fprop (a = x, b = w, c = y), 1 dgrad (a = dy, b = w, c = dx), 2 wgrad (a = dy, b = x, c = dw fp32)
int conv_gemm_launch(int which, const cgat_conv_desc* d, const void* a, const void* b, void* c, const float* bias,
                     cudaStream_t st) {
  return d->dtype == CGAT_F32 ? cg_launch<float>(which, d, a, b, c, bias, st)
                              : cg_launch<__nv_bfloat16>(which, d, a, b, c, bias, st);
}

}  // namespace cgat
