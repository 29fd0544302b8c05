// Depthwise 3x3 convolutions of the SmaAt-UNet encoder behind unet_model.py:20 (DepthwiseSeparableConv with
// kernels_per_layer = 2: groups = cin, cout = 2*cin, stride 1, pad 1) -- fprop, dgrad, wgrad + dbias, NHWC.
//
// These are HBM-bound (9 multiply-adds per output element), so the kernels are organised around 16-byte accesses
// along the contiguous channel axis: one thread owns 8 consecutive OUTPUT channels of one pixel (= 8/M input
// channels for channel multiplier M), the weights sit transposed in shared memory as [tap][cout] so a thread reads
// its 8 weights of a tap with vector loads, and the nine taps re-read x through L1/L2.  The generic direct
// kernels (conv_direct.cu: one thread per output element, 2-byte accesses) stay for every other grouped shape.
#include "common.cuh"

namespace cgat {

constexpr int DWK_THREADS = 256;
constexpr int DWK_TAPS = 9;

template <typename T, int N> struct Vec;
template <int N> struct Vec<float, N> {
  static __device__ __forceinline__ void load(const float* p, float (&v)[N]) {
#pragma unroll
    for (int i = 0; i < N; i += 4) {
      const float4 t = *reinterpret_cast<const float4*>(p + i);
      v[i] = t.x; v[i + 1] = t.y; v[i + 2] = t.z; v[i + 3] = t.w;
    }
  }
  static __device__ __forceinline__ void store(float* p, const float (&v)[N]) {
#pragma unroll
    for (int i = 0; i < N; i += 4) *reinterpret_cast<float4*>(p + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
  }
};
template <int N> struct Vec<__nv_bfloat16, N> {
  static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&v)[N]) {
    if constexpr (N == 8) {
      const uint4 t = *reinterpret_cast<const uint4*>(p);
      const uint32_t u[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        v[2 * i] = __uint_as_float(u[i] << 16);
        v[2 * i + 1] = __uint_as_float(u[i] & 0xffff0000u);
      }
    } else {
      const uint2 t = *reinterpret_cast<const uint2*>(p);
      const uint32_t u[2] = {t.x, t.y};
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        v[2 * i] = __uint_as_float(u[i] << 16);
        v[2 * i + 1] = __uint_as_float(u[i] & 0xffff0000u);
      }
    }
  }
  static __device__ __forceinline__ void store(__nv_bfloat16* p, const float (&v)[N]) {
    uint32_t u[N / 2];
#pragma unroll
    for (int i = 0; i < N / 2; ++i) {
      const __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
      u[i] = *reinterpret_cast<const uint32_t*>(&h);
    }
    if constexpr (N == 8) *reinterpret_cast<uint4*>(p) = make_uint4(u[0], u[1], u[2], u[3]);
    else *reinterpret_cast<uint2*>(p) = make_uint2(u[0], u[1]);
  }
};

__device__ __forceinline__ float dwk_act(float v, int act) {
  if (act == 1) return fmaxf(v, 0.f);
  if (act == 2) return v > 0.f ? v : 0.2f * v;
  if (act == 3) return 1.f / (1.f + __expf(-v));
  return v;
}

// weights [cout][3][3][1] -> shared [tap][cout] (fp32)
template <typename T>
__device__ __forceinline__ void dwk_stage_weights(const T* __restrict__ w, float* sw, int cout) {
  for (int i = threadIdx.x; i < cout * DWK_TAPS; i += blockDim.x) {
    const int co = i / DWK_TAPS, tap = i - co * DWK_TAPS;
    sw[tap * cout + co] = DT<T>::to_f(w[i]);
  }
  __syncthreads();
}

template <typename T, int M>
__global__ void __launch_bounds__(DWK_THREADS) dw3x3_fprop_kernel(const cgat_conv_desc d, const T* __restrict__ x,
                                                                  const T* __restrict__ w, const float* __restrict__ bias,
                                                                  T* __restrict__ y) {
  extern __shared__ float sw[];
  dwk_stage_weights(w, sw, d.cout);
  const int oct = d.cout / 8;
  const long long total = (long long)d.n * d.ho * d.wo * oct;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int o = (int)(idx % oct);
    const long long pix = idx / oct;
    const int wo = (int)(pix % d.wo), ho = (int)((pix / d.wo) % d.ho), n = (int)(pix / ((long long)d.wo * d.ho));
    const int co0 = o * 8, ci0 = co0 / M;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = bias ? bias[co0 + j] : 0.f;
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
      const int hi = ho + kh - d.pad_top;
      if (hi < 0 || hi >= d.h) continue;
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        const int wi = wo + kw - d.pad_left;
        if (wi < 0 || wi >= d.w) continue;
        float xv[8 / M], wv[8];
        Vec<T, 8 / M>::load(x + (((long long)n * d.h + hi) * d.w + wi) * d.cin + ci0, xv);
        Vec<float, 8>::load(sw + (kh * 3 + kw) * d.cout + co0, wv);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = fmaf(xv[j / M], wv[j], acc[j]);
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = dwk_act(acc[j], d.act);
    Vec<T, 8>::store(y + pix * d.cout + co0, acc);
  }
}

template <typename T, int M>
__global__ void __launch_bounds__(DWK_THREADS) dw3x3_dgrad_kernel(const cgat_conv_desc d, const T* __restrict__ dy,
                                                                  const T* __restrict__ w, T* __restrict__ dx) {
  extern __shared__ float sw[];
  dwk_stage_weights(w, sw, d.cout);
  const int oct = d.cout / 8;
  const long long total = (long long)d.n * d.h * d.w * oct;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int o = (int)(idx % oct);
    const long long pix = idx / oct;
    const int wi = (int)(pix % d.w), hi = (int)((pix / d.w) % d.h), n = (int)(pix / ((long long)d.w * d.h));
    const int co0 = o * 8, ci0 = co0 / M;
    float acc[8 / M];
#pragma unroll
    for (int j = 0; j < 8 / M; ++j) acc[j] = 0.f;
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
      const int ho = hi + d.pad_top - kh;
      if (ho < 0 || ho >= d.ho) continue;
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        const int wo = wi + d.pad_left - kw;
        if (wo < 0 || wo >= d.wo) continue;
        float gv[8], wv[8];
        Vec<T, 8>::load(dy + (((long long)n * d.ho + ho) * d.wo + wo) * d.cout + co0, gv);
        Vec<float, 8>::load(sw + (kh * 3 + kw) * d.cout + co0, wv);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j / M] = fmaf(gv[j], wv[j], acc[j / M]);
      }
    }
    Vec<T, 8 / M>::store(dx + pix * d.cin + ci0, acc);
  }
}

// wgrad + dbias: thread = (channel octet o, pixel lane pl); 72 + 8 partial sums in registers over the CTA's pixel
// slab, merged through shared-memory atomics, one global atomicAdd per value per CTA (dw / dbias zeroed by the
// launcher).  blockIdx.y walks octet groups when cout/8 > 256.
template <typename T, int M>
__global__ void __launch_bounds__(DWK_THREADS) dw3x3_wgrad_kernel(const cgat_conv_desc d, const T* __restrict__ x,
                                                                  const T* __restrict__ dy, float* __restrict__ dw,
                                                                  float* __restrict__ dbias, int ol, long long per) {
  extern __shared__ float sacc[];  // [ol][80]
  for (int i = threadIdx.x; i < ol * 80; i += blockDim.x) sacc[i] = 0.f;
  __syncthreads();
  const int pl_n = DWK_THREADS / ol;
  const int ot = threadIdx.x % ol, pl = threadIdx.x / ol;
  const int o = blockIdx.y * ol + ot;
  const int co0 = o * 8, ci0 = co0 / M;
  const long long Mpix = (long long)d.n * d.ho * d.wo;
  const long long p0 = (long long)blockIdx.x * per, p1 = p0 + per < Mpix ? p0 + per : Mpix;
  float acc[DWK_TAPS][8], accb[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    accb[j] = 0.f;
#pragma unroll
    for (int t = 0; t < DWK_TAPS; ++t) acc[t][j] = 0.f;
  }
  if (co0 < d.cout) {
    for (long long m = p0 + pl; m < p1; m += pl_n) {
      const int wo = (int)(m % d.wo), ho = (int)((m / d.wo) % d.ho), n = (int)(m / ((long long)d.wo * d.ho));
      float gv[8];
      Vec<T, 8>::load(dy + m * d.cout + co0, gv);
#pragma unroll
      for (int j = 0; j < 8; ++j) accb[j] += gv[j];
#pragma unroll
      for (int kh = 0; kh < 3; ++kh) {
        const int hi = ho + kh - d.pad_top;
        if (hi < 0 || hi >= d.h) continue;
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
          const int wi = wo + kw - d.pad_left;
          if (wi < 0 || wi >= d.w) continue;
          float xv[8 / M];
          Vec<T, 8 / M>::load(x + (((long long)n * d.h + hi) * d.w + wi) * d.cin + ci0, xv);
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[kh * 3 + kw][j] = fmaf(gv[j], xv[j / M], acc[kh * 3 + kw][j]);
        }
      }
    }
  }
  // lanes of a warp that own the same octet (ol < 32: lane = pl*ol + ot) merge by shuffles first, so the shared
  // atomics see one contender per warp instead of 32/ol
  if (ol < 32) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
#pragma unroll
      for (int t = 0; t < DWK_TAPS; ++t)
        for (int off = 16; off >= ol; off >>= 1) acc[t][j] += __shfl_xor_sync(0xffffffffu, acc[t][j], off);
      for (int off = 16; off >= ol; off >>= 1) accb[j] += __shfl_xor_sync(0xffffffffu, accb[j], off);
    }
  }
  if (co0 < d.cout && (ol >= 32 || (threadIdx.x & 31) < ol)) {
    float* s = sacc + ot * 80;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
#pragma unroll
      for (int t = 0; t < DWK_TAPS; ++t) atomicAdd(s + j * DWK_TAPS + t, acc[t][j]);
      atomicAdd(s + 72 + j, accb[j]);
    }
  }
  __syncthreads();
  // sacc[ot][j*9 + t] is dw[(co0 + j)*9 + t]: 72 consecutive values per octet
  for (int i = threadIdx.x; i < ol * 80; i += blockDim.x) {
    const int oo = i / 80, r = i - oo * 80;
    const int c0 = (blockIdx.y * ol + oo) * 8;
    if (c0 >= d.cout) continue;
    if (r < 72) atomicAdd(dw + (long long)c0 * DWK_TAPS + r, sacc[i]);
    else if (dbias) atomicAdd(dbias + c0 + (r - 72), sacc[i]);
  }
}

int conv_dw3x3_served(const cgat_conv_desc* d) {
  if (d->groups != d->cin || d->groups < 2 || d->kh != 3 || d->kw != 3 || d->stride != 1) return 0;
  const int m = d->cout / d->cin;
  if (d->cout != m * d->cin || (m != 1 && m != 2) || d->cout % 8) return 0;
  return (size_t)d->cout * DWK_TAPS * sizeof(float) <= 96 * 1024;
}

template <typename T, int M>
static int dw_launch_t(int which, const cgat_conv_desc* d, const void* a, const void* b, void* c, float* dbias,
                       const float* bias, cudaStream_t st) {
  const size_t wsm = (size_t)d->cout * DWK_TAPS * sizeof(float);
  const int oct = d->cout / 8;
  if (which == 0 || which == 1) {
    const long long total = (long long)d->n * (which == 0 ? d->ho * d->wo : d->h * d->w) * oct;
    long long blocks = (total + DWK_THREADS - 1) / DWK_THREADS;
    if (blocks > 148 * 8) blocks = 148 * 8;
    if (which == 0) {
      static bool attr = false;
      if (!attr) {
        cudaFuncSetAttribute(dw3x3_fprop_kernel<T, M>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
        attr = true;
      }
      dw3x3_fprop_kernel<T, M><<<(int)blocks, DWK_THREADS, wsm, st>>>(*d, (const T*)a, (const T*)b, bias, (T*)c);
      return check_launch("dw3x3_fprop_kernel");
    }
    static bool attr = false;
    if (!attr) {
      cudaFuncSetAttribute(dw3x3_dgrad_kernel<T, M>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
      attr = true;
    }
    dw3x3_dgrad_kernel<T, M><<<(int)blocks, DWK_THREADS, wsm, st>>>(*d, (const T*)a, (const T*)b, (T*)c);
    return check_launch("dw3x3_dgrad_kernel");
  }
  // wgrad: a = x, b = dy, c = dw
  int ol = 1;
  while (ol < oct && ol < DWK_THREADS) ol *= 2;  // octet lanes per CTA (power of two <= 256)
  const int ygroups = (oct + ol - 1) / ol;
  const int pl_n = DWK_THREADS / ol;
  const long long Mpix = (long long)d->n * d->ho * d->wo;
  long long slabs = (148LL * 4 + ygroups - 1) / ygroups;
  const long long min_per = (long long)pl_n * 16;  // at least 16 pixels per lane before paying the atomics
  if (slabs > (Mpix + min_per - 1) / min_per) slabs = (Mpix + min_per - 1) / min_per;
  if (slabs < 1) slabs = 1;
  const long long per = (Mpix + slabs - 1) / slabs;
  cudaMemsetAsync(c, 0, sizeof(float) * (size_t)d->cout * DWK_TAPS, st);
  if (dbias) cudaMemsetAsync(dbias, 0, sizeof(float) * (size_t)d->cout, st);
  dim3 grid((unsigned)((Mpix + per - 1) / per), (unsigned)ygroups);
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(dw3x3_wgrad_kernel<T, M>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    attr = true;
  }
  dw3x3_wgrad_kernel<T, M><<<grid, DWK_THREADS, (size_t)ol * 80 * sizeof(float), st>>>(*d, (const T*)a, (const T*)b,
                                                                                      (float*)c, dbias, ol, per);
  return check_launch("dw3x3_wgrad_kernel");
}

// which: 0 fprop (a = x, b = w, c = y), 1 dgrad (a = dy, b = w, c = dx), 2 wgrad (a = x, b = dy, c = dw; dbias optional)
int conv_dw3x3_launch(int which, const cgat_conv_desc* d, const void* a, const void* b, void* c, float* dbias,
                      const float* bias, cudaStream_t st) {
  if (!aligned16(a) || !aligned16(b) || !aligned16(c)) return fail(CGAT_EALIGN, "depthwise conv tensors must be 16-byte aligned");
  const int m = d->cout / d->cin;
  if (d->dtype == CGAT_F32)
    return m == 1 ? dw_launch_t<float, 1>(which, d, a, b, c, dbias, bias, st)
                  : dw_launch_t<float, 2>(which, d, a, b, c, dbias, bias, st);
  return m == 1 ? dw_launch_t<__nv_bfloat16, 1>(which, d, a, b, c, dbias, bias, st)
                : dw_launch_t<__nv_bfloat16, 2>(which, d, a, b, c, dbias, bias, st);
}

}  // namespace cgat
