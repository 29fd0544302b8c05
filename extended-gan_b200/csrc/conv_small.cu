// Dense stride-1 convs with FEW channels on both sides (the DCGAN generator's ends, dcgan/model.py:19-34: 4 -> 32 and
// 4 -> 4 channels with k = 4 "same" padding at 64 x 64; the 7x7 2 -> 1 conv of the SmaAt-UNet's CBAM spatial gates) -- fprop, and dgrad as the same kernel over dy with the taps
// flipped and the padding mirrored.  The implicit-GEMM tiles (conv_gemm.cu, 64 pixels x 64 couts x 16) waste 15/16 of a
// tile on 4 output channels and ran these at 130-165 us per launch; here a thread owns PX consecutive output pixels of a
// row and ALL output channels: per tap one vector load of the pixel's CI input channels, CI x CO FMAs against weights that
// sit in shared memory as fp32 [tap][ci][co] and are read as warp-wide broadcasts.  FP32-FMA bound (2 x taps x CI x CO
// flops per pixel), no tensor cores: K = CI <= 8 per tap cannot fill an MMA.
#include "common.cuh"

namespace cgat {

constexpr int CS_THREADS = 128;

__device__ __forceinline__ float cs_act(float v, int act) {
  switch (act) {
    case 1: return fmaxf(v, 0.f);
    case 2: return v > 0.f ? v : 0.2f * v;
    case 3: return 1.f / (1.f + __expf(-v));
    default: return v;
  }
}

template <typename T, int C>
__device__ __forceinline__ void cs_load(const T* p, float (&v)[C]) {
  if constexpr (C == 1) {
    v[0] = DT<T>::to_f(p[0]);
  } else if constexpr (C == 2 && sizeof(T) == 4) {
    const float2 f = *reinterpret_cast<const float2*>(p);
    v[0] = f.x; v[1] = f.y;
  } else if constexpr (C == 2) {
    const uint32_t u = *reinterpret_cast<const uint32_t*>(p);
    v[0] = __uint_as_float(u << 16); v[1] = __uint_as_float(u & 0xffff0000u);
  } else if constexpr (sizeof(T) == 4) {
#pragma unroll
    for (int q = 0; q < C / 4; ++q) {
      const float4 f = reinterpret_cast<const float4*>(p)[q];
      v[4 * q] = f.x; v[4 * q + 1] = f.y; v[4 * q + 2] = f.z; v[4 * q + 3] = f.w;
    }
  } else if constexpr (C == 4) {
    const uint2 u = *reinterpret_cast<const uint2*>(p);
    v[0] = __uint_as_float(u.x << 16); v[1] = __uint_as_float(u.x & 0xffff0000u);
    v[2] = __uint_as_float(u.y << 16); v[3] = __uint_as_float(u.y & 0xffff0000u);
  } else {
#pragma unroll
    for (int q = 0; q < C / 8; ++q) {
      const uint4 u = reinterpret_cast<const uint4*>(p)[q];
      const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        v[8 * q + 2 * i] = __uint_as_float(w[i] << 16);
        v[8 * q + 2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
      }
    }
  }
}
template <typename T, int C>
__device__ __forceinline__ void cs_store(T* p, const float (&v)[C]) {
  if constexpr (C == 1) {
    p[0] = DT<T>::from_f(v[0]);
  } else if constexpr (C == 2 && sizeof(T) == 4) {
    *reinterpret_cast<float2*>(p) = make_float2(v[0], v[1]);
  } else if constexpr (C == 2) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(v[0], v[1]);
    *reinterpret_cast<uint32_t*>(p) = *reinterpret_cast<const uint32_t*>(&h);
  } else if constexpr (sizeof(T) == 4) {
#pragma unroll
    for (int q = 0; q < C / 4; ++q) reinterpret_cast<float4*>(p)[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
  } else {
    uint32_t w[C / 2];
#pragma unroll
    for (int i = 0; i < C / 2; ++i) {
      const __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
      w[i] = *reinterpret_cast<const uint32_t*>(&h);
    }
    if constexpr (C == 4) {
      *reinterpret_cast<uint2*>(p) = make_uint2(w[0], w[1]);
    } else {
#pragma unroll
      for (int q = 0; q < C / 8; ++q) reinterpret_cast<uint4*>(p)[q] = make_uint4(w[4 * q], w[4 * q + 1], w[4 * q + 2], w[4 * q + 3]);
    }
  }
}

// in [n][hi][wi][CI] -> out [n][ho][wo][CO];  out[yo][xo] = sum_{r,s} in[yo - pt + r][xo - pl + s] . W[r][s]
// flip == 0 (fprop):  W[r][s][ci][co] = w[co][r][s][ci]                     (w is [cout = CO][kh][kw][cin = CI])
// flip == 1 (dgrad):  W[r][s][ci][co] = w[ci][kh-1-r][kw-1-s][co]           (w is [cout = CI][kh][kw][cin = CO])
template <typename T, int CI, int CO, int PX>
__global__ void __launch_bounds__(CS_THREADS) conv_small_kernel(const T* __restrict__ in, const T* __restrict__ w,
                                                                  const float* __restrict__ bias, T* __restrict__ out, int n,
                                                                  int hi, int wi, int ho, int wo, int kh, int kw, int pt,
                                                                  int pl, int act, int flip) {
  extern __shared__ __align__(16) float cs_w[];  // [taps][CI][CO]
  const int taps = kh * kw;
  for (int i = threadIdx.x; i < taps * CI * CO; i += CS_THREADS) {
    const int co = i % CO, ci = (i / CO) % CI, tap = i / (CO * CI);
    cs_w[i] = flip ? DT<T>::to_f(w[((long long)ci * taps + (taps - 1 - tap)) * CO + co])
                   : DT<T>::to_f(w[((long long)co * taps + tap) * CI + ci]);
  }
  __syncthreads();
  const int wq = (wo + PX - 1) / PX;  // pixel groups per output row
  const long long total = (long long)n * ho * wq;
  for (long long idx = (long long)blockIdx.x * CS_THREADS + threadIdx.x; idx < total; idx += (long long)gridDim.x * CS_THREADS) {
    const int xq = (int)(idx % wq);
    long long t = idx / wq;
    const int yo = (int)(t % ho);
    const long long img = t / ho;
    const int xo0 = xq * PX;
    float acc[PX][CO];
#pragma unroll
    for (int p = 0; p < PX; ++p)
#pragma unroll
      for (int c = 0; c < CO; ++c) acc[p][c] = bias ? bias[c] : 0.f;
    for (int r = 0; r < kh; ++r) {
      const int yi = yo - pt + r;
      if (yi < 0 || yi >= hi) continue;
      const T* rowp = in + ((img * hi + yi) * wi) * CI;
      for (int s = 0; s < kw; ++s) {
        float xv[PX][CI];
#pragma unroll
        for (int p = 0; p < PX; ++p) {
          const int xi = xo0 + p - pl + s;
          if (xi >= 0 && xi < wi) {
            cs_load<T, CI>(rowp + (long long)xi * CI, xv[p]);
          } else {
#pragma unroll
            for (int c = 0; c < CI; ++c) xv[p][c] = 0.f;
          }
        }
        const float* wt = cs_w + (r * kw + s) * CI * CO;
#pragma unroll
        for (int ci = 0; ci < CI; ++ci) {
          float wv[CO];
          if constexpr (CO % 4 == 0) {
#pragma unroll
            for (int q = 0; q < CO / 4; ++q) {
              const float4 f = reinterpret_cast<const float4*>(wt + ci * CO)[q];
              wv[4 * q] = f.x; wv[4 * q + 1] = f.y; wv[4 * q + 2] = f.z; wv[4 * q + 3] = f.w;
            }
          } else {
#pragma unroll
            for (int q = 0; q < CO; ++q) wv[q] = wt[ci * CO + q];
          }
#pragma unroll
          for (int p = 0; p < PX; ++p)
#pragma unroll
            for (int c = 0; c < CO; ++c) acc[p][c] = fmaf(xv[p][ci], wv[c], acc[p][c]);
        }
      }
    }
#pragma unroll
    for (int p = 0; p < PX; ++p) {
      if (xo0 + p >= wo) continue;
#pragma unroll
      for (int c = 0; c < CO; ++c) acc[p][c] = cs_act(acc[p][c], act);
      cs_store<T, CO>(out + ((img * ho + yo) * wo + xo0 + p) * CO, acc[p]);
    }
  }
}

static bool cs_pair(int ci, int co) {
  if ((ci == 2 && co == 1) || (ci == 1 && co == 2)) return true;  // CBAM's 7x7 spatial-gate conv (2 -> 1) and its dgrad
  return (ci == 4 || ci == 8) && (co == 4 || co == 8 || co == 16 || co == 32);
}

// which: 0 fprop, 1 dgrad
int conv_small_served(const cgat_conv_desc* d, int which) {
  if (d->groups != 1 || d->stride != 1 || (d->dtype != CGAT_F32 && d->dtype != CGAT_BF16)) return 0;
  if (d->kh * d->kw > 49) return 0;
  if (which == 0) return cs_pair(d->cin, d->cout);
  if (which == 1) return cs_pair(d->cout, d->cin);  // (the activation derivative is applied to dy by the caller)
  return 0;
}

template <typename T>
static int cs_launch(int ci, int co, const T* in, const T* w, const float* bias, T* out, int n, int hi, int wi, int ho, int wo,
                     int kh, int kw, int pt, int pl, int act, int flip, cudaStream_t st) {
  const size_t smem = (size_t)kh * kw * ci * co * sizeof(float);
  if (smem > 48 * 1024) return fail(CGAT_EUNSUPPORTED, "conv_small: %zu B of weights", smem);
#define CS_GO(CI, CO, PX)                                                                                              \
  do {                                                                                                                 \
    const long long total = (long long)n * ho * ((wo + PX - 1) / PX);                                                  \
    long long ctas = (total + CS_THREADS - 1) / CS_THREADS;                                                            \
    if (ctas > 148 * 16) ctas = 148 * 16;                                                                              \
    conv_small_kernel<T, CI, CO, PX><<<(unsigned)ctas, CS_THREADS, smem, st>>>(in, w, bias, out, n, hi, wi, ho, wo, kh, kw, pt, \
                                                                               pl, act, flip);                         \
  } while (0)
  if (ci == 2 && co == 1) CS_GO(2, 1, 4);
  else if (ci == 1 && co == 2) CS_GO(1, 2, 4);
  else if (ci == 4 && co == 4) CS_GO(4, 4, 4);
  else if (ci == 4 && co == 8) CS_GO(4, 8, 4);
  else if (ci == 4 && co == 16) CS_GO(4, 16, 2);
  else if (ci == 4 && co == 32) CS_GO(4, 32, 2);
  else if (ci == 8 && co == 4) CS_GO(8, 4, 4);
  else if (ci == 8 && co == 8) CS_GO(8, 8, 4);
  else if (ci == 8 && co == 16) CS_GO(8, 16, 2);
  else if (ci == 8 && co == 32) CS_GO(8, 32, 2);
  else return fail(CGAT_EUNSUPPORTED, "conv_small: %d -> %d channels", ci, co);
#undef CS_GO
  return check_launch("conv_small_kernel");
}

int conv_small_launch(int which, const cgat_conv_desc* d, const void* in, const void* w, const float* bias, void* out,
                      cudaStream_t st) {
  // fprop: in = x [n][h][w][cin] -> y [n][ho][wo][cout];  dgrad: in = dy [n][ho][wo][cout] -> dx [n][h][w][cin], mirrored padding
  const int ci = which == 0 ? d->cin : d->cout, co = which == 0 ? d->cout : d->cin;
  const int hi = which == 0 ? d->h : d->ho, wi = which == 0 ? d->w : d->wo;
  const int ho = which == 0 ? d->ho : d->h, wo = which == 0 ? d->wo : d->w;
  const int pt = which == 0 ? d->pad_top : d->kh - 1 - d->pad_top, pl = which == 0 ? d->pad_left : d->kw - 1 - d->pad_left;
  const int act = which == 0 ? d->act : 0;
  if (d->dtype == CGAT_F32)
    return cs_launch<float>(ci, co, (const float*)in, (const float*)w, bias, (float*)out, d->n, hi, wi, ho, wo, d->kh, d->kw, pt,
                            pl, act, which, st);
  return cs_launch<__nv_bfloat16>(ci, co, (const __nv_bfloat16*)in, (const __nv_bfloat16*)w, bias, (__nv_bfloat16*)out, d->n, hi,
                                  wi, ho, wo, d->kh, d->kw, pt, pl, act, which, st);
}

}  // namespace cgat
