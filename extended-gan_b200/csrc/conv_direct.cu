// Direct (CUDA-core) NHWC convolution: fprop / dgrad / wgrad for ANY shape.
//
// Role: (1) the small-channel convs where an implicit GEMM cannot fill a tensor-core tile (DCGAN
// generator, 4..32 channels, dcgan/model.py:55-76; conv1 of the discriminators, cin = 4/8), and
// (2) the on-device cross-check of the tcgen05 implicit-GEMM kernels in conv_tc.cu (tests compare
// both against the oracle).  fp32 accumulation, weights [cout][kh][kw][cin], leading padding given
// explicitly so that PyTorch's asymmetric padding="same" for even kernels (left 1 / right 2 for k=4,
// dcgan/model.py:61-72) is expressible.
#include "common.cuh"
#include "ew_vec.cuh"

namespace cgat {

constexpr int CD_THREADS = 256;

__device__ __forceinline__ float apply_act(float v, int act) {
  switch (act) {
    case 1: return fmaxf(v, 0.f);
    case 2: return v > 0.f ? v : 0.2f * v;
    case 3: return 1.f / (1.f + __expf(-v));
    default: return v;
  }
}

// one thread per output element, cout fastest (x is a warp broadcast, w rows stay in L1)
template <typename T>
__global__ void __launch_bounds__(CD_THREADS) conv_fprop_direct(const cgat_conv_desc d, const T* __restrict__ x,
                                                                const T* __restrict__ w,
                                                                const float* __restrict__ bias, T* __restrict__ y) {
  const long long total = (long long)d.n * d.ho * d.wo * d.cout;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int co = (int)(idx % d.cout);
    long long pix = idx / d.cout;
    const int wo = (int)(pix % d.wo);
    pix /= d.wo;
    const int ho = (int)(pix % d.ho);
    const int n = (int)(pix / d.ho);
    float acc = bias ? bias[co] : 0.f;
    const int cing = d.cin / d.groups;                 // input channels per group
    const int ci0 = (co / (d.cout / d.groups)) * cing;  // first input channel of this output's group
    for (int kh = 0; kh < d.kh; ++kh) {
      const int hi = ho * d.stride + kh - d.pad_top;
      if (hi < 0 || hi >= d.h) continue;
      for (int kw = 0; kw < d.kw; ++kw) {
        const int wi = wo * d.stride + kw - d.pad_left;
        if (wi < 0 || wi >= d.w) continue;
        const T* xp = x + (((long long)n * d.h + hi) * d.w + wi) * d.cin + ci0;
        const T* wp = w + (((long long)co * d.kh + kh) * d.kw + kw) * cing;
        for (int ci = 0; ci < cing; ++ci) acc = fmaf(DT<T>::to_f(xp[ci]), DT<T>::to_f(wp[ci]), acc);
      }
    }
    y[idx] = DT<T>::from_f(apply_act(acc, d.act));
  }
}

// one thread per input element, cin fastest
template <typename T>
__global__ void __launch_bounds__(CD_THREADS) conv_dgrad_direct(const cgat_conv_desc d, const T* __restrict__ dy,
                                                                const T* __restrict__ w, T* __restrict__ dx) {
  const long long total = (long long)d.n * d.h * d.w * d.cin;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int ci = (int)(idx % d.cin);
    long long pix = idx / d.cin;
    const int wi = (int)(pix % d.w);
    pix /= d.w;
    const int hi = (int)(pix % d.h);
    const int n = (int)(pix / d.h);
    float acc = 0.f;
    for (int kh = 0; kh < d.kh; ++kh) {
      const int hn = hi + d.pad_top - kh;
      if (hn < 0 || hn % d.stride) continue;
      const int ho = hn / d.stride;
      if (ho >= d.ho) continue;
      for (int kw = 0; kw < d.kw; ++kw) {
        const int wn = wi + d.pad_left - kw;
        if (wn < 0 || wn % d.stride) continue;
        const int wo = wn / d.stride;
        if (wo >= d.wo) continue;
        const int cing = d.cin / d.groups, coutg = d.cout / d.groups;
        const int g = ci / cing;
        const T* dyp = dy + (((long long)n * d.ho + ho) * d.wo + wo) * d.cout + g * coutg;
        const long long wstride = (long long)d.kh * d.kw * cing;
        const T* wp = w + (long long)g * coutg * wstride + ((long long)kh * d.kw + kw) * cing + (ci - g * cing);
        for (int co = 0; co < coutg; ++co) acc = fmaf(DT<T>::to_f(dyp[co]), DT<T>::to_f(wp[co * wstride]), acc);
      }
    }
    dx[idx] = DT<T>::from_f(acc);
  }
}

// one CTA per weight element group: block reduces over all output pixels.  dw is WRITTEN (fp32).
template <typename T>
__global__ void __launch_bounds__(CD_THREADS) conv_wgrad_direct(const cgat_conv_desc d, const T* __restrict__ x,
                                                                const T* __restrict__ dy, float* __restrict__ dw) {
  // blockIdx.x enumerates (co, kh, kw); threads stride over (pixel, ci-chunk)
  const int kw = blockIdx.x % d.kw;
  const int kh = (blockIdx.x / d.kw) % d.kh;
  const int co = blockIdx.x / (d.kw * d.kh);
  extern __shared__ float s_acc[];  // [cin / groups]
  const int cing = d.cin / d.groups;
  const int ci_base = (co / (d.cout / d.groups)) * cing;
  for (int ci = threadIdx.x; ci < cing; ci += blockDim.x) s_acc[ci] = 0.f;
  __syncthreads();
  const long long M = (long long)d.n * d.ho * d.wo;
  // thread -> (ci, pixel lane): ci fastest for coalesced x reads
  const int lanes_ci = cing < CD_THREADS ? cing : CD_THREADS;
  const int ci0 = threadIdx.x % lanes_ci;
  const int prow = threadIdx.x / lanes_ci;
  const int prows = CD_THREADS / lanes_ci;
  if (prow < prows) {
    for (int ci = ci0; ci < cing; ci += lanes_ci) {
      float acc = 0.f;
      for (long long m = prow; m < M; m += prows) {
        const int wo = (int)(m % d.wo);
        const int ho = (int)((m / d.wo) % d.ho);
        const int n = (int)(m / ((long long)d.wo * d.ho));
        const int hi = ho * d.stride + kh - d.pad_top;
        const int wi = wo * d.stride + kw - d.pad_left;
        if (hi < 0 || hi >= d.h || wi < 0 || wi >= d.w) continue;
        acc = fmaf(DT<T>::to_f(dy[m * d.cout + co]),
                   DT<T>::to_f(x[(((long long)n * d.h + hi) * d.w + wi) * d.cin + ci_base + ci]), acc);
      }
      atomicAdd(&s_acc[ci], acc);
    }
  }
  __syncthreads();
  for (int ci = threadIdx.x; ci < cing; ci += blockDim.x)
    dw[(((long long)co * d.kh + kh) * d.kw + kw) * cing + ci] = s_acc[ci];
}

// Depthwise wgrad (groups == cin, e.g. the SmaAt-UNet 3x3 depthwise convs): thread = output channel (coalesced along the
// NHWC channel axis), CTA = a slab of output pixels; kh*kw partial sums per thread in registers, one atomicAdd per
// (channel, tap, CTA) into the zeroed dw.  (The general kernel above walks every pixel once per output channel.)
constexpr int DW_MAX_TAPS = 49;
template <typename T>
__global__ void __launch_bounds__(CD_THREADS) conv_wgrad_depthwise(const cgat_conv_desc d, const T* __restrict__ x,
                                                                   const T* __restrict__ dy, float* __restrict__ dw,
                                                                   long long pix_per_cta) {
  const int co = blockIdx.x * CD_THREADS + threadIdx.x;
  if (co >= d.cout) return;
  const int ci = co / (d.cout / d.groups);  // one input channel per group
  const int taps = d.kh * d.kw;
  float acc[DW_MAX_TAPS];
#pragma unroll
  for (int t = 0; t < DW_MAX_TAPS; ++t) acc[t] = 0.f;
  const long long M = (long long)d.n * d.ho * d.wo;
  const long long p0 = (long long)blockIdx.y * pix_per_cta;
  const long long p1 = p0 + pix_per_cta < M ? p0 + pix_per_cta : M;
  for (long long m = p0; m < p1; ++m) {
    const int wo = (int)(m % d.wo);
    const int ho = (int)((m / d.wo) % d.ho);
    const int n = (int)(m / ((long long)d.wo * d.ho));
    const float g = DT<T>::to_f(dy[m * d.cout + co]);
#pragma unroll 1
    for (int kh = 0; kh < d.kh; ++kh) {
      const int hi = ho * d.stride + kh - d.pad_top;
      if (hi < 0 || hi >= d.h) continue;
      for (int kw = 0; kw < d.kw; ++kw) {
        const int wi = wo * d.stride + kw - d.pad_left;
        if (wi < 0 || wi >= d.w) continue;
        acc[kh * d.kw + kw] = fmaf(g, DT<T>::to_f(x[(((long long)n * d.h + hi) * d.w + wi) * d.cin + ci]), acc[kh * d.kw + kw]);
      }
    }
  }
  for (int t = 0; t < taps; ++t) atomicAdd(&dw[(long long)co * taps + t], acc[t]);
}

template <typename T>
__global__ void __launch_bounds__(CD_THREADS) conv_dbias_direct(const T* __restrict__ dy, float* __restrict__ db,
                                                                long long M, int cout) {
  const int co = blockIdx.x;
  float acc = 0.f;
  for (long long m = threadIdx.x; m < M; m += blockDim.x) acc += DT<T>::to_f(dy[m * cout + co]);
  __shared__ float s[CD_THREADS / 32];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float v = 0.f;
    for (int i = 0; i < CD_THREADS / 32; ++i) v += s[i];
    db[co] = v;
  }
}

// ---- dbias = column sums of dy [M][cout], coalesced and deterministic: the kernel above walks one channel per block with a
// stride of cout elements (57 us for the 33 MB dy of a UNet pointwise conv); here a CTA owns a slab of pixels, a thread
// (pixel lane, group of V channels) keeps DB_U 16-byte loads in flight, the lanes are folded through shared memory, every CTA
// writes ONE row of partial sums into the caller's workspace and a second tiny launch adds the rows in a fixed order.
constexpr int DB_THREADS = 256, DB_U = 4, DB_MAX_CTAS = 296;

template <typename T, int V>
__global__ void __launch_bounds__(DB_THREADS) dbias_partial_kernel(const T* __restrict__ dy, float* __restrict__ part,
                                                                     long long M, int cout, int pix_per_cta) {
  __shared__ float s_part[DB_THREADS * (V > 1 ? V : 1)];
  const int groups = cout / V;
  const int cg = groups >= DB_THREADS ? DB_THREADS : groups;
  int lanes = 1;
  while (2 * lanes * cg <= DB_THREADS) lanes *= 2;
  const long long p0 = (long long)blockIdx.x * pix_per_cta, p1 = min(M, p0 + pix_per_cta);
  for (int g0 = 0; g0 < groups; g0 += cg) {
    const int g = g0 + (int)(threadIdx.x % cg), lane = threadIdx.x / cg;
    float acc[V];
#pragma unroll
    for (int i = 0; i < V; ++i) acc[i] = 0.f;
    if (g < groups && lane < lanes) {
      for (long long pb = p0 + lane; pb < p1; pb += (long long)lanes * DB_U) {
        NaRaw<T, V> r[DB_U];
#pragma unroll
        for (int q = 0; q < DB_U; ++q) {
          const long long p = pb + (long long)q * lanes;
          if (p < p1) r[q] = na_load_raw<T, V>(dy + p * cout + (long long)g * V);
        }
#pragma unroll
        for (int q = 0; q < DB_U; ++q) {
          if (pb + (long long)q * lanes >= p1) continue;
          float v[V];
          na_unpack<T, V>(r[q], v);
#pragma unroll
          for (int i = 0; i < V; ++i) acc[i] += v[i];
        }
      }
    }
#pragma unroll
    for (int i = 0; i < V; ++i) s_part[threadIdx.x * V + i] = acc[i];
    __syncthreads();
    for (int half = lanes >> 1; half > 0; half >>= 1) {
      if (lane < half) {
        const int other = (threadIdx.x + half * cg) * V;
#pragma unroll
        for (int i = 0; i < V; ++i) s_part[threadIdx.x * V + i] += s_part[other + i];
      }
      __syncthreads();
    }
    if ((int)threadIdx.x < cg && g0 + (int)threadIdx.x < groups) {
#pragma unroll
      for (int i = 0; i < V; ++i) part[(long long)blockIdx.x * cout + (g0 + threadIdx.x) * V + i] = s_part[threadIdx.x * V + i];
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(DB_THREADS) dbias_final_kernel(const float* __restrict__ part, float* __restrict__ db, int ctas,
                                                                   int cout) {
  const int co = blockIdx.x * DB_THREADS + threadIdx.x;
  if (co >= cout) return;
  float v = 0.f;
  for (int c = 0; c < ctas; ++c) v += part[(long long)c * cout + co];
  db[co] = v;
}

size_t conv_dbias_workspace(const cgat_conv_desc* d) { return (size_t)DB_MAX_CTAS * d->cout * sizeof(float); }

int conv_dbias_launch(const cgat_conv_desc* d, const void* dy, float* dbias, cudaStream_t st);
// the same with a caller workspace of conv_dbias_workspace(d) bytes (NULL: the strided kernel)
int conv_dbias_ws_launch(const cgat_conv_desc* d, const void* dy, float* dbias, void* workspace, cudaStream_t st) {
  if (!workspace || !aligned16(workspace)) return conv_dbias_launch(d, dy, dbias, st);
  const long long M = (long long)d->n * d->ho * d->wo;
  const int v = ew_vec(d->dtype, d->cout, dy);
  const int groups = d->cout / v, cg = groups >= DB_THREADS ? DB_THREADS : groups;
  int lanes = 1;
  while (2 * lanes * cg <= DB_THREADS) lanes *= 2;
  const int ppc_min = lanes * DB_U;
  long long ctas = (M + ppc_min - 1) / ppc_min;
  if (ctas > DB_MAX_CTAS) ctas = DB_MAX_CTAS;
  const int ppc = (int)((M + ctas - 1) / ctas);
  ctas = (M + ppc - 1) / ppc;
  float* part = (float*)workspace;
#define DB_GO(T, V) dbias_partial_kernel<T, V><<<(unsigned)ctas, DB_THREADS, 0, st>>>((const T*)dy, part, M, d->cout, ppc)
  if (d->dtype == CGAT_F32) { if (v == 4) DB_GO(float, 4); else DB_GO(float, 1); }
  else { if (v == 8) DB_GO(__nv_bfloat16, 8); else DB_GO(__nv_bfloat16, 1); }
#undef DB_GO
  dbias_final_kernel<<<(d->cout + DB_THREADS - 1) / DB_THREADS, DB_THREADS, 0, st>>>(part, dbias, (int)ctas, d->cout);
  return check_launch("dbias kernels");
}


static int grid_for(long long total) {
  long long g = (total + CD_THREADS - 1) / CD_THREADS;
  const long long cap = 148LL * 32;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

int validate_conv(const cgat_conv_desc* d) {
  if (!d) return fail(CGAT_EINVAL, "null conv descriptor");
  if (d->n < 1 || d->h < 1 || d->w < 1 || d->cin < 1 || d->cout < 1 || d->kh < 1 || d->kw < 1 || d->stride < 1 ||
      d->ho < 1 || d->wo < 1 || d->pad_top < 0 || d->pad_left < 0)
    return fail(CGAT_EINVAL, "bad conv descriptor");
  if (d->dtype != CGAT_F32 && d->dtype != CGAT_BF16) return fail(CGAT_EINVAL, "bad conv dtype %d", d->dtype);
  if (d->groups < 1 || d->cin % d->groups || d->cout % d->groups)
    return fail(CGAT_EINVAL, "groups=%d must divide cin=%d and cout=%d", d->groups, d->cin, d->cout);
  // the last output row/col must start inside the padded input
  if ((d->ho - 1) * d->stride - d->pad_top >= d->h || (d->wo - 1) * d->stride - d->pad_left >= d->w)
    return fail(CGAT_EINVAL, "conv output %dx%d does not fit input %dx%d", d->ho, d->wo, d->h, d->w);
  return 0;
}

// ---- full-window conv: the kernel covers the whole (unpadded) input, one output pixel per image ---------------
// The last conv of both DCGAN discriminators (dcgan/model.py:166-169: 512 -> 1, k=4 on the 4x4 map) is a plain dot
// product of length h*w*cin per (image, cout): x[n] and w[co] are contiguous in the same (h, w, c) order.  One block
// per (image, cout) with a warp-shuffle + shared-memory tree; the tiled GEMM would run it in a single CTA.
template <typename T>
__global__ void __launch_bounds__(256) conv_fullwindow_fprop(const cgat_conv_desc d, const T* __restrict__ x,
                                                            const T* __restrict__ w, const float* __restrict__ bias,
                                                            T* __restrict__ y) {
  const int co = blockIdx.x % d.cout, n = blockIdx.x / d.cout;
  const long long len = (long long)d.h * d.w * d.cin;
  const T* xr = x + (long long)n * len;
  const T* wr = w + (long long)co * len;
  float acc = 0.f;
  for (long long i = threadIdx.x; i < len; i += blockDim.x) acc = fmaf(DT<T>::to_f(xr[i]), DT<T>::to_f(wr[i]), acc);
  acc = warp_sum(acc);
  __shared__ float part[8];
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float v = bias ? bias[co] : 0.f;
    for (int i = 0; i < 8; ++i) v += part[i];
    if (d.act == 1) v = fmaxf(v, 0.f);
    else if (d.act == 2) v = v > 0.f ? v : 0.2f * v;
    else if (d.act == 3) v = 1.f / (1.f + __expf(-v));
    y[(long long)n * d.cout + co] = DT<T>::from_f(v);
  }
}

int conv_is_fullwindow(const cgat_conv_desc* d) {
  return d->groups == 1 && d->ho == 1 && d->wo == 1 && d->pad_top == 0 && d->pad_left == 0 && d->kh == d->h &&
         d->kw == d->w && d->cout <= 16;
}

int conv_fullwindow_fprop_launch(const cgat_conv_desc* d, const void* x, const void* w, const float* bias, void* y,
                                 cudaStream_t st) {
  const int blocks = d->n * d->cout;
  if (d->dtype == CGAT_F32)
    conv_fullwindow_fprop<float><<<blocks, 256, 0, st>>>(*d, (const float*)x, (const float*)w, bias, (float*)y);
  else
    conv_fullwindow_fprop<__nv_bfloat16><<<blocks, 256, 0, st>>>(*d, (const __nv_bfloat16*)x, (const __nv_bfloat16*)w,
                                                                bias, (__nv_bfloat16*)y);
  return check_launch("conv_fullwindow_fprop");
}

int conv_fprop_direct_launch(const cgat_conv_desc* d, const void* x, const void* w, const float* bias, void* y,
                             cudaStream_t st) {
  const long long total = (long long)d->n * d->ho * d->wo * d->cout;
  if (d->dtype == CGAT_F32)
    conv_fprop_direct<float><<<grid_for(total), CD_THREADS, 0, st>>>(*d, (const float*)x, (const float*)w, bias, (float*)y);
  else
    conv_fprop_direct<__nv_bfloat16><<<grid_for(total), CD_THREADS, 0, st>>>(*d, (const __nv_bfloat16*)x,
                                                                             (const __nv_bfloat16*)w, bias,
                                                                             (__nv_bfloat16*)y);
  return check_launch("conv_fprop_direct");
}

int conv_dgrad_direct_launch(const cgat_conv_desc* d, const void* dy, const void* w, void* dx, cudaStream_t st) {
  const long long total = (long long)d->n * d->h * d->w * d->cin;
  if (d->dtype == CGAT_F32)
    conv_dgrad_direct<float><<<grid_for(total), CD_THREADS, 0, st>>>(*d, (const float*)dy, (const float*)w, (float*)dx);
  else
    conv_dgrad_direct<__nv_bfloat16><<<grid_for(total), CD_THREADS, 0, st>>>(*d, (const __nv_bfloat16*)dy,
                                                                             (const __nv_bfloat16*)w,
                                                                             (__nv_bfloat16*)dx);
  return check_launch("conv_dgrad_direct");
}

int conv_wgrad_direct_launch(const cgat_conv_desc* d, const void* x, const void* dy, float* dw, float* dbias,
                             cudaStream_t st) {
  const int blocks = d->cout * d->kh * d->kw;
  const size_t smem = sizeof(float) * (d->cin / d->groups);
  const long long M = (long long)d->n * d->ho * d->wo;
  if (d->groups == d->cin && d->groups > 1 && d->kh * d->kw <= DW_MAX_TAPS) {  // depthwise
    cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)d->cout * d->kh * d->kw, st);
    const int cblocks = (d->cout + CD_THREADS - 1) / CD_THREADS;
    long long slabs = (148LL * 8 + cblocks - 1) / cblocks;
    if (slabs > M) slabs = M;
    const long long per = (M + slabs - 1) / slabs;
    dim3 grid(cblocks, (unsigned)((M + per - 1) / per));
    if (d->dtype == CGAT_F32)
      conv_wgrad_depthwise<float><<<grid, CD_THREADS, 0, st>>>(*d, (const float*)x, (const float*)dy, dw, per);
    else
      conv_wgrad_depthwise<__nv_bfloat16><<<grid, CD_THREADS, 0, st>>>(*d, (const __nv_bfloat16*)x, (const __nv_bfloat16*)dy,
                                                                      dw, per);
    if (int rc = check_launch("conv_wgrad_depthwise")) return rc;
    return dbias ? conv_dbias_launch(d, dy, dbias, st) : 0;
  }
  if (d->dtype == CGAT_F32) {
    conv_wgrad_direct<float><<<blocks, CD_THREADS, smem, st>>>(*d, (const float*)x, (const float*)dy, dw);
    if (dbias) conv_dbias_direct<float><<<d->cout, CD_THREADS, 0, st>>>((const float*)dy, dbias, M, d->cout);
  } else {
    conv_wgrad_direct<__nv_bfloat16><<<blocks, CD_THREADS, smem, st>>>(*d, (const __nv_bfloat16*)x,
                                                                       (const __nv_bfloat16*)dy, dw);
    if (dbias)
      conv_dbias_direct<__nv_bfloat16><<<d->cout, CD_THREADS, 0, st>>>((const __nv_bfloat16*)dy, dbias, M, d->cout);
  }
  return check_launch("conv_wgrad_direct");
}

int conv_dbias_launch(const cgat_conv_desc* d, const void* dy, float* dbias, cudaStream_t st) {
  const long long M = (long long)d->n * d->ho * d->wo;
  if (d->dtype == CGAT_F32)
    conv_dbias_direct<float><<<d->cout, CD_THREADS, 0, st>>>((const float*)dy, dbias, M, d->cout);
  else
    conv_dbias_direct<__nv_bfloat16><<<d->cout, CD_THREADS, 0, st>>>((const __nv_bfloat16*)dy, dbias, M, d->cout);
  return check_launch("conv_dbias_direct");
}

}  // namespace cgat
