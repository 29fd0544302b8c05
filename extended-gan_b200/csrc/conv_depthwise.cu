// Depthwise 3x3 convolutions of the SmaAt-UNet encoder behind unet_model.py:20 (DepthwiseSeparableConv with
// kernels_per_layer = 2: groups = cin, cout = 2*cin, stride 1, pad 1) -- fprop, dgrad, wgrad + dbias, NHWC.
//
// These are HBM-bound (9 multiply-adds per output element), so the kernels are organised around 16-byte accesses
// along the contiguous channel axis: one thread owns 8 consecutive OUTPUT channels of one pixel (= 8/M input
// channels for channel multiplier M), the weights sit transposed in shared memory as [tap][cout] so a thread reads
// its 8 weights of a tap with vector loads, and the nine taps re-read x through L1/L2.  The generic direct
// kernels (conv_direct.cu: one thread per output element, 2-byte accesses) stay for every other grouped shape.
#include "common.cuh"
#include <cstdlib>

namespace cgat {

constexpr int DWK_THREADS = 256;
constexpr int DWK_TAPS = 9;
constexpr int DWK_TW = 8;  // pixel-tile width of the fprop / dgrad kernels (the height follows the channel count)
constexpr int DWK_PX = 4;  // output pixels per thread (a strip of a row)

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

// IT: the integer type of the index arithmetic -- unsigned when every element offset fits 31 bits (the host checks): the
// three 64-bit divisions per output octet of the first version cost more instructions than the nine taps.
//
// Thread = (channel octet, STRIP of DWK_PX consecutive output pixels of a row): per kernel row the strip's DWK_PX + 2 input
// pixels are loaded ONCE and every tap's weights once, where the one-pixel-per-thread version issued 9 loads + 18 shared
// loads per pixel and ran at 97 % of the L1 pipe with DRAM at 7 % (ncu, profiles/r2z_depthwise.txt).  A CTA walks tiles of
// DWK_TW x th_rows pixels (octets of a strip on consecutive threads, then the strips of a tile row, then the rows).
template <typename T, int M, typename IT>
__global__ void __launch_bounds__(DWK_THREADS) dw3x3_fprop_kernel(const cgat_conv_desc d, const T* __restrict__ x,
                                                                  const T* __restrict__ w, const float* __restrict__ bias,
                                                                  T* __restrict__ y, int th_rows) {
  extern __shared__ float sw[];
  dwk_stage_weights(w, sw, d.cout);
  const int oct = d.cout / 8;
  constexpr int SPT = DWK_TW / DWK_PX;  // strips per tile row
  const int tiles_w = (d.wo + DWK_TW - 1) / DWK_TW, tiles_h = (d.ho + th_rows - 1) / th_rows;
  const IT n_tiles = (IT)d.n * tiles_h * tiles_w, per_tile = (IT)SPT * th_rows * oct;
  for (IT tile = blockIdx.x; tile < n_tiles; tile += gridDim.x)
  for (IT r = threadIdx.x; r < per_tile; r += blockDim.x) {
    const int tw = (int)(tile % (IT)tiles_w);
    const IT trest = tile / (IT)tiles_w;
    const int th = (int)(trest % (IT)tiles_h), n = (int)(trest / (IT)tiles_h);
    const int o = (int)(r % (IT)oct);
    const int q = (int)(r / (IT)oct);
    const int wo0 = tw * DWK_TW + (q % SPT) * DWK_PX, ho = th * th_rows + q / SPT;
    if (wo0 >= d.wo || ho >= d.ho) continue;
    const int co0 = o * 8, ci0 = co0 / M;
    float acc[DWK_PX][8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float b = bias ? bias[co0 + j] : 0.f;
#pragma unroll
      for (int px = 0; px < DWK_PX; ++px) acc[px][j] = b;
    }
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
      const int hi = ho + kh - d.pad_top;
      if (hi < 0 || hi >= d.h) continue;
      float xv[DWK_PX + 2][8 / M];
      const T* xrow = x + (((IT)n * d.h + hi) * d.w) * d.cin + ci0;
#pragma unroll
      for (int c = 0; c < DWK_PX + 2; ++c) {
        const int wi = wo0 + c - d.pad_left;
        if (wi >= 0 && wi < d.w) {
          Vec<T, 8 / M>::load(xrow + (IT)wi * d.cin, xv[c]);
        } else {
#pragma unroll
          for (int j = 0; j < 8 / M; ++j) xv[c][j] = 0.f;
        }
      }
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        float wv[8];
        Vec<float, 8>::load(sw + (kh * 3 + kw) * d.cout + co0, wv);
#pragma unroll
        for (int px = 0; px < DWK_PX; ++px)
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[px][j] = fmaf(xv[px + kw][j / M], wv[j], acc[px][j]);
      }
    }
    T* yrow = y + (((IT)n * d.ho + ho) * d.wo) * d.cout + co0;
#pragma unroll
    for (int px = 0; px < DWK_PX; ++px) {
      if (wo0 + px >= d.wo) continue;
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[px][j] = dwk_act(acc[px][j], d.act);
      Vec<T, 8>::store(yrow + (IT)(wo0 + px) * d.cout, acc[px]);
    }
  }
}

// dgrad: dx[hi][wi][ci] = sum_{kh,kw} dy[hi + pad_top - kh][wi + pad_left - kw][co] * w[co][kh][kw]; same strips over dx
template <typename T, int M, typename IT>
__global__ void __launch_bounds__(DWK_THREADS) dw3x3_dgrad_kernel(const cgat_conv_desc d, const T* __restrict__ dy,
                                                                  const T* __restrict__ w, T* __restrict__ dx, int th_rows) {
  extern __shared__ float sw[];
  dwk_stage_weights(w, sw, d.cout);
  const int oct = d.cout / 8;
  constexpr int SPT = DWK_TW / DWK_PX;
  const int tiles_w = (d.w + DWK_TW - 1) / DWK_TW, tiles_h = (d.h + th_rows - 1) / th_rows;
  const IT n_tiles = (IT)d.n * tiles_h * tiles_w, per_tile = (IT)SPT * th_rows * oct;
  for (IT tile = blockIdx.x; tile < n_tiles; tile += gridDim.x)
  for (IT r = threadIdx.x; r < per_tile; r += blockDim.x) {
    const int tw = (int)(tile % (IT)tiles_w);
    const IT trest = tile / (IT)tiles_w;
    const int th = (int)(trest % (IT)tiles_h), n = (int)(trest / (IT)tiles_h);
    const int o = (int)(r % (IT)oct);
    const int q = (int)(r / (IT)oct);
    const int wi0 = tw * DWK_TW + (q % SPT) * DWK_PX, hi = th * th_rows + q / SPT;
    if (wi0 >= d.w || hi >= d.h) continue;
    const int co0 = o * 8, ci0 = co0 / M;
    float acc[DWK_PX][8 / M];
#pragma unroll
    for (int px = 0; px < DWK_PX; ++px)
#pragma unroll
      for (int j = 0; j < 8 / M; ++j) acc[px][j] = 0.f;
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
      const int ho = hi + d.pad_top - kh;
      if (ho < 0 || ho >= d.ho) continue;
      float gv[DWK_PX + 2][8];  // column c: wo = wi0 + pad_left - 2 + c  (pixel px, tap kw reads c = px + 2 - kw)
      const T* grow = dy + (((IT)n * d.ho + ho) * d.wo) * d.cout + co0;
#pragma unroll
      for (int c = 0; c < DWK_PX + 2; ++c) {
        const int wo = wi0 + d.pad_left - 2 + c;
        if (wo >= 0 && wo < d.wo) {
          Vec<T, 8>::load(grow + (IT)wo * d.cout, gv[c]);
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) gv[c][j] = 0.f;
        }
      }
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        float wv[8];
        Vec<float, 8>::load(sw + (kh * 3 + kw) * d.cout + co0, wv);
#pragma unroll
        for (int px = 0; px < DWK_PX; ++px)
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[px][j / M] = fmaf(gv[px + 2 - kw][j], wv[j], acc[px][j / M]);
      }
    }
    T* drow = dx + (((IT)n * d.h + hi) * d.w) * d.cin + ci0;
#pragma unroll
    for (int px = 0; px < DWK_PX; ++px)
      if (wi0 + px < d.w) Vec<T, 8 / M>::store(drow + (IT)(wi0 + px) * d.cin, acc[px]);
  }
}

// wgrad + dbias: thread = (channel octet o, pixel lane pl); 72 + 8 partial sums in registers over the CTA's pixel
// slab, merged through shared-memory atomics, one global atomicAdd per value per CTA (dw / dbias zeroed by the
// launcher).  blockIdx.y walks octet groups when cout/8 > 256.
template <typename T, int M, bool STRIPS>
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
    if constexpr (STRIPS) {
      // STRIPS of DWK_PX consecutive pixels of a row per lane and trip (host: wo % DWK_PX == 0, slabs strip-aligned): the
      // strip's DWK_PX + 2 input pixels of a kernel row are loaded once -- 5.5 loads per pixel instead of 10
      long long m = p0 + (long long)pl * DWK_PX;
      int wo = (int)(m % d.wo), ho = (int)((m / d.wo) % d.ho), n = (int)(m / ((long long)d.wo * d.ho));
      const int step = pl_n * DWK_PX, s_wo = step % d.wo, s_row = step / d.wo;
      for (; m < p1; m += step) {
        float gv[DWK_PX][8];
#pragma unroll
        for (int px = 0; px < DWK_PX; ++px) {
          Vec<T, 8>::load(dy + (m + px) * d.cout + co0, gv[px]);
#pragma unroll
          for (int j = 0; j < 8; ++j) accb[j] += gv[px][j];
        }
#pragma unroll
        for (int kh = 0; kh < 3; ++kh) {
          const int hi = ho + kh - d.pad_top;
          if (hi < 0 || hi >= d.h) continue;
          float xv[DWK_PX + 2][8 / M];
          const T* xrow = x + (((long long)n * d.h + hi) * d.w) * d.cin + ci0;
#pragma unroll
          for (int c = 0; c < DWK_PX + 2; ++c) {
            const int wi = wo + c - d.pad_left;
            if (wi >= 0 && wi < d.w) {
              Vec<T, 8 / M>::load(xrow + (long long)wi * d.cin, xv[c]);
            } else {
#pragma unroll
              for (int j = 0; j < 8 / M; ++j) xv[c][j] = 0.f;
            }
          }
#pragma unroll
          for (int kw = 0; kw < 3; ++kw)
#pragma unroll
            for (int px = 0; px < DWK_PX; ++px)
#pragma unroll
              for (int j = 0; j < 8; ++j) acc[kh * 3 + kw][j] = fmaf(gv[px][j], xv[px + kw][j / M], acc[kh * 3 + kw][j]);
        }
        wo += s_wo;
        ho += s_row;
        if (wo >= d.wo) { wo -= d.wo; ++ho; }
        while (ho >= d.ho) { ho -= d.ho; ++n; }
      }
    } else {
    // pixel coordinates are carried from pixel to pixel (one division per thread, not three per pixel)
    long long m = p0 + pl;
    int wo = (int)(m % d.wo), ho = (int)((m / d.wo) % d.ho), n = (int)(m / ((long long)d.wo * d.ho));
    const int s_wo = pl_n % d.wo, s_row = pl_n / d.wo;  // pl_n pixels = s_row rows + s_wo columns
    for (; m < p1; m += pl_n) {
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
      wo += s_wo;
      ho += s_row;
      if (wo >= d.wo) { wo -= d.wo; ++ho; }
      while (ho >= d.ho) { ho -= d.ho; ++n; }
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
    // tile height: at least 4 rows, and at least four passes of the CTA's 256 threads per tile
    int th_rows = (4 * DWK_THREADS + (DWK_TW / DWK_PX) * oct - 1) / ((DWK_TW / DWK_PX) * oct);
    th_rows = th_rows < 4 ? 4 : (th_rows > 64 ? 64 : th_rows);
    const int oh = which == 0 ? d->ho : d->h, ow = which == 0 ? d->wo : d->w;
    long long blocks = (long long)d->n * ((oh + th_rows - 1) / th_rows) * ((ow + DWK_TW - 1) / DWK_TW);
    if (blocks > 148 * 8) blocks = 148 * 8;
    // 32-bit index arithmetic when every element offset (and the grid-stride overshoot) fits 31 bits
    const long long elems = (long long)d->n * (d->h > d->ho ? d->h : d->ho) * (d->w > d->wo ? d->w : d->wo) * d->cout;
    const bool small = elems < (1ll << 31) - (1 << 20);
    static bool attr = false;
    if (!attr) {
      cudaFuncSetAttribute(dw3x3_fprop_kernel<T, M, unsigned>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
      cudaFuncSetAttribute(dw3x3_fprop_kernel<T, M, long long>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
      cudaFuncSetAttribute(dw3x3_dgrad_kernel<T, M, unsigned>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
      cudaFuncSetAttribute(dw3x3_dgrad_kernel<T, M, long long>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
      attr = true;
    }
    if (which == 0) {
      if (small) dw3x3_fprop_kernel<T, M, unsigned><<<(int)blocks, DWK_THREADS, wsm, st>>>(*d, (const T*)a, (const T*)b, bias, (T*)c, th_rows);
      else dw3x3_fprop_kernel<T, M, long long><<<(int)blocks, DWK_THREADS, wsm, st>>>(*d, (const T*)a, (const T*)b, bias, (T*)c, th_rows);
      return check_launch("dw3x3_fprop_kernel");
    }
    if (small) dw3x3_dgrad_kernel<T, M, unsigned><<<(int)blocks, DWK_THREADS, wsm, st>>>(*d, (const T*)a, (const T*)b, (T*)c, th_rows);
    else dw3x3_dgrad_kernel<T, M, long long><<<(int)blocks, DWK_THREADS, wsm, st>>>(*d, (const T*)a, (const T*)b, (T*)c, th_rows);
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
  long long per = (Mpix + slabs - 1) / slabs;
  // strips of DWK_PX pixels per lane when the rows allow it (the slabs then hold whole strips) -- for fp32 activations only:
  // the strip kernel needs 160-250 registers (one CTA per SM); measured on UnetModel [2,128,128,4,8]: fp32 17.9 -> 17.3 ms,
  // bf16 8.6 -> 9.0 ms (its loads are half the size, the lost occupancy costs more than the saved requests)
  const bool strips = d->wo % DWK_PX == 0 && sizeof(T) == 4 && !getenv("CGAT_DW_NO_STRIPS");
  if (strips) per = (per + DWK_PX - 1) / DWK_PX * DWK_PX;
  cudaMemsetAsync(c, 0, sizeof(float) * (size_t)d->cout * DWK_TAPS, st);
  if (dbias) cudaMemsetAsync(dbias, 0, sizeof(float) * (size_t)d->cout, st);
  dim3 grid((unsigned)((Mpix + per - 1) / per), (unsigned)ygroups);
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(dw3x3_wgrad_kernel<T, M, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    cudaFuncSetAttribute(dw3x3_wgrad_kernel<T, M, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    attr = true;
  }
  if (strips)
    dw3x3_wgrad_kernel<T, M, true><<<grid, DWK_THREADS, (size_t)ol * 80 * sizeof(float), st>>>(*d, (const T*)a, (const T*)b,
                                                                                              (float*)c, dbias, ol, per);
  else
    dw3x3_wgrad_kernel<T, M, false><<<grid, DWK_THREADS, (size_t)ol * 80 * sizeof(float), st>>>(*d, (const T*)a, (const T*)b,
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
