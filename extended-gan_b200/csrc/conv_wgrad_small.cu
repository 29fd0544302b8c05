// Weight gradient of the small-channel, large-image convs of the DCGAN generator (dcgan/model.py:55-76: k = 4,
// padding="same", 4 -> 32 -> 16 -> 8 -> 4 -> 4 channels on 64x64 frames), NHWC, stride 1, fp32 accumulation.
//
// dW[co][tap][ci] has at most a few thousand entries while the contraction runs over every pixel of the batch
// (262 144 at N = 64): a tensor-core tile would be 1/8 full (the streamed wgrad of conv_tc_big.cu needs 0.26 ms for
// 32 -> 16), and the generic direct kernel walks all pixels once per output channel with 2-byte loads (1.15 ms).
// Here a persistent CTA stages a tile of 8 x 32 output pixels (dY) and its input halo (X) in shared memory, a thread
// owns one or two (tap, ci) pairs and keeps the sums for ALL cout of them in registers: per pixel one scalar X read,
// cout/4 broadcast 16-byte dY reads (fp32, converted once when the tile is staged) and cout FMAs.  With fewer than 256 pairs the threads split the tile's pixels
// into groups.  Partial sums leave the CTA once, at the end, through shared-memory then global atomics (dw zeroed by
// the launcher); dbias rides along in the threads that own pair 0.
#include "common.cuh"

namespace cgat {

constexpr int WS_THREADS = 256;
constexpr int WS_TH = 8, WS_TW = 32;  // output pixels per tile

template <typename T> __device__ __forceinline__ float ws_ld(const T* p);
template <> __device__ __forceinline__ float ws_ld<float>(const float* p) { return *p; }
template <> __device__ __forceinline__ float ws_ld<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }

// COUT consecutive elements of shared memory as floats, 16-byte (8-byte for 4 bf16) loads
template <typename T, int COUT> struct WsVec;
template <int COUT> struct WsVec<float, COUT> {
  static __device__ __forceinline__ void load(const float* p, float (&g)[COUT]) {
#pragma unroll
    for (int i = 0; i < COUT; i += 4) {
      const float4 t = *reinterpret_cast<const float4*>(p + i);
      g[i] = t.x; g[i + 1] = t.y; g[i + 2] = t.z; g[i + 3] = t.w;
    }
  }
};
template <int COUT> struct WsVec<__nv_bfloat16, COUT> {
  static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&g)[COUT]) {
    if constexpr (COUT == 4) {
      const uint2 t = *reinterpret_cast<const uint2*>(p);
      g[0] = __uint_as_float(t.x << 16); g[1] = __uint_as_float(t.x & 0xffff0000u);
      g[2] = __uint_as_float(t.y << 16); g[3] = __uint_as_float(t.y & 0xffff0000u);
    } else {
#pragma unroll
      for (int i = 0; i < COUT; i += 8) {
        const uint4 t = *reinterpret_cast<const uint4*>(p + i);
        const uint32_t u[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          g[i + 2 * k] = __uint_as_float(u[k] << 16);
          g[i + 2 * k + 1] = __uint_as_float(u[k] & 0xffff0000u);
        }
      }
    }
  }
};

template <typename T, int COUT>
__global__ void __launch_bounds__(WS_THREADS)
conv_wgrad_small_kernel(const cgat_conv_desc d, const T* __restrict__ x, const T* __restrict__ dy, float* __restrict__ dw,
                        float* __restrict__ dbias) {
  extern __shared__ __align__(16) unsigned char ws_smem[];
  const int taps = d.kh * d.kw, cin = d.cin;
  const int P = taps * cin;                       // (tap, ci) pairs
  const int O = COUT * P;
  const int XH = WS_TH + d.kh - 1, XW = WS_TW + d.kw - 1;
  float* red = reinterpret_cast<float*>(ws_smem);                    // [O + COUT]
  float* ys = red + ((O + COUT + 3) & ~3);                           // [TH*TW][COUT], converted once at staging
  T* xs = reinterpret_cast<T*>(ys + WS_TH * WS_TW * COUT);           // [XH][XW][cin]
  for (int i = threadIdx.x; i < O + COUT; i += WS_THREADS) red[i] = 0.f;

  // thread -> pairs and pixel group
  const int G = P >= WS_THREADS ? 1 : WS_THREADS / P;                // pixel groups
  const int grp = P >= WS_THREADS ? 0 : threadIdx.x / P;
  const int pr0 = P >= WS_THREADS ? threadIdx.x : threadIdx.x - grp * P;
  const bool act0 = grp < G && pr0 < P;
  const int pr1 = pr0 + WS_THREADS;
  const bool two = P > WS_THREADS;  // block-uniform: a second pair per thread
  const bool act1 = two && pr1 < P;
  const int tap0 = pr0 / cin, ci0 = pr0 - tap0 * cin, r0 = tap0 / d.kw, s0 = tap0 - r0 * d.kw;
  const int tap1 = act1 ? pr1 / cin : 0, ci1 = act1 ? pr1 - tap1 * cin : 0, r1 = tap1 / d.kw, s1 = tap1 - r1 * d.kw;
  const int xoff0 = (r0 * XW + s0) * cin + ci0, xoff1 = (r1 * XW + s1) * cin + ci1;
  const bool bias_thread = dbias != nullptr && act0 && pr0 == 0;

  float acc0[COUT], acc1[COUT], accb[COUT];
#pragma unroll
  for (int j = 0; j < COUT; ++j) acc0[j] = acc1[j] = accb[j] = 0.f;

  const int tiles_w = (d.wo + WS_TW - 1) / WS_TW, tiles_h = (d.ho + WS_TH - 1) / WS_TH;
  const int tiles = d.n * tiles_h * tiles_w;
  for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const int tw = tile % tiles_w, th = (tile / tiles_w) % tiles_h, n = tile / (tiles_w * tiles_h);
    const int ho0 = th * WS_TH, wo0 = tw * WS_TW;
    __syncthreads();  // previous tile fully consumed (and `red` zeroed on the first pass)
    // dY tile (zero outside the image)
    for (int i = threadIdx.x; i < WS_TH * WS_TW * COUT; i += WS_THREADS) {
      const int co = i % COUT, p = i / COUT, px = p % WS_TW, py = p / WS_TW;
      const int ho = ho0 + py, wo = wo0 + px;
      ys[i] = (ho < d.ho && wo < d.wo) ? DT<T>::to_f(dy[(((long long)n * d.ho + ho) * d.wo + wo) * COUT + co]) : 0.f;
    }
    // X halo tile (zero = conv padding)
    for (int i = threadIdx.x; i < XH * XW * cin; i += WS_THREADS) {
      const int ci = i % cin, q = i / cin, qx = q % XW, qy = q / XW;
      const int hi = ho0 + qy - d.pad_top, wi = wo0 + qx - d.pad_left;
      xs[i] = (hi >= 0 && hi < d.h && wi >= 0 && wi < d.w) ? x[(((long long)n * d.h + hi) * d.w + wi) * cin + ci]
                                                           : DT<T>::from_f(0.f);
    }
    __syncthreads();
    if (act0) {
#pragma unroll 4
      for (int p = grp; p < WS_TH * WS_TW; p += G) {
        const int px = p % WS_TW, py = p / WS_TW;
        const int base = (py * XW + px) * cin;
        const float xv0 = ws_ld(xs + base + xoff0);
        const float xv1 = act1 ? ws_ld(xs + base + xoff1) : 0.f;
        float g[COUT];
        WsVec<float, COUT>::load(ys + p * COUT, g);
#pragma unroll
        for (int j = 0; j < COUT; ++j) acc0[j] = fmaf(g[j], xv0, acc0[j]);
        if (two) {
#pragma unroll
          for (int j = 0; j < COUT; ++j) acc1[j] = fmaf(g[j], xv1, acc1[j]);
        }
        if (bias_thread) {
#pragma unroll
          for (int j = 0; j < COUT; ++j) accb[j] += g[j];
        }
      }
    }
  }
  __syncthreads();
  if (act0) {
#pragma unroll
    for (int j = 0; j < COUT; ++j) {
      atomicAdd(red + j * P + pr0, acc0[j]);      // dw index (co*taps + tap)*cin + ci = co*P + pair
      if (act1) atomicAdd(red + j * P + pr1, acc1[j]);
      if (bias_thread) atomicAdd(red + O + j, accb[j]);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < O; i += WS_THREADS) atomicAdd(dw + i, red[i]);
  if (dbias)
    for (int i = threadIdx.x; i < COUT; i += WS_THREADS) atomicAdd(dbias + i, red[O + i]);
}

static size_t ws_smem_bytes(const cgat_conv_desc* d) {
  const size_t esz = d->dtype == CGAT_F32 ? 4 : 2;
  const int P = d->kh * d->kw * d->cin, O = d->cout * P;
  const size_t red = ((size_t)(O + d->cout + 3) & ~(size_t)3) * 4;
  return red + (size_t)WS_TH * WS_TW * d->cout * 4 + (size_t)(WS_TH + d->kh - 1) * (WS_TW + d->kw - 1) * d->cin * esz;
}

int conv_wgrad_small_served(const cgat_conv_desc* d) {
  if (d->groups != 1 || d->stride != 1) return 0;
  if (d->cout != 4 && d->cout != 8 && d->cout != 16 && d->cout != 32) return 0;
  if (d->cin > 32 || d->kh * d->kw * d->cin > 2 * WS_THREADS || d->kh > 7 || d->kw > 7) return 0;
  if ((long long)d->n * d->ho * d->wo < 4096) return 0;  // few pixels: the other kernels are fine
  return ws_smem_bytes(d) <= 96 * 1024;
}

template <typename T, int COUT>
static int ws_launch(const cgat_conv_desc* d, const void* x, const void* dy, float* dw, float* dbias, cudaStream_t st) {
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(conv_wgrad_small_kernel<T, COUT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    attr = true;
  }
  const int tiles = d->n * ((d->ho + WS_TH - 1) / WS_TH) * ((d->wo + WS_TW - 1) / WS_TW);
  const int grid = tiles < 148 * 2 ? tiles : 148 * 2;
  cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)d->cout * d->kh * d->kw * d->cin, st);
  if (dbias) cudaMemsetAsync(dbias, 0, sizeof(float) * (size_t)d->cout, st);
  conv_wgrad_small_kernel<T, COUT><<<grid, WS_THREADS, ws_smem_bytes(d), st>>>(*d, (const T*)x, (const T*)dy, dw, dbias);
  return check_launch("conv_wgrad_small_kernel");
}

int conv_wgrad_small_launch(const cgat_conv_desc* d, const void* x, const void* dy, float* dw, float* dbias,
                            cudaStream_t st) {
#define WS_CASE(C)                                                                        \
  case C:                                                                                 \
    return d->dtype == CGAT_F32 ? ws_launch<float, C>(d, x, dy, dw, dbias, st)             \
                                : ws_launch<__nv_bfloat16, C>(d, x, dy, dw, dbias, st)
  switch (d->cout) {
    WS_CASE(4);
    WS_CASE(8);
    WS_CASE(16);
    WS_CASE(32);
  }
#undef WS_CASE
  return fail(CGAT_EUNSUPPORTED, "small-channel wgrad serves cout in {4, 8, 16, 32}");
}

}  // namespace cgat
