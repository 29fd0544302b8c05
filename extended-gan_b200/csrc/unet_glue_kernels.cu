// The non-conv ops of the SmaAt-UNet that convolutional_gat/unet_model.py:20-29 applies per vertex (public
// architecture: double depthwise-separable convs, CBAM, 2x2 max-pooling, bilinear x2 up-sampling + pad + concat), NHWC,
// fp32 or bf16, forward and backward -- SURVEY.md section 8(f) rank 2.  All of them are HBM-bound passes:
//
//   max-pool 2x2        fwd: one read of x, y + a byte of arg-max per output;  bwd: one read of dy + arg-max, one write
//   up-sample + concat  fwd: out[..., :C2] = x2, out[..., C2:] = pad(bilinear_x2(x1), align_corners=True) in ONE pass (the
//                       reference materialises the up-sampled, the padded and the concatenated tensor);  bwd: d(x2) is a
//                       channel slice, d(x1) GATHERS its <= 5x5 contributions (deterministic, no atomics)
//   CBAM channel gate   pool_hw (mean / max / first arg-max over the pixels, split over CTAs, the last CTA of an image
//                       combines) -> cbam_mlp (both MLP passes + sigmoid, one CTA per image) -> gate (y = x * s[n][c]);
//                       bwd: sum(dy * x) per (n, c) -> cbam_mlp_bwd -> dx = dy * s + d(avg)/HW + d(max) at the arg-max
//   CBAM spatial gate   chan_pool (mean / max / arg-max over the channels of a pixel) -> 7x7 conv + BatchNorm + sigmoid
//                       (conv / norm_act kernels) -> gate (y = x * s[n][pixel]);  bwd: sum over channels of dy * x,
//                       dx = dy * s, and the pooled branch's  d(mean)/C + d(max) at the arg-max
#include "common.cuh"
#include "ew_vec.cuh"

namespace cgat {

constexpr int UG_THREADS = 256;

#define UG_LAUNCH(KERN, GRID, ST, DTYPE, V, ...)                                                                              \
  do {                                                                                                                        \
    cudaError_t e_;                                                                                                           \
    if ((DTYPE) == CGAT_F32) {                                                                                                \
      if ((V) == 4) e_ = launch_pdl(KERN<float, 4>, GRID, dim3(UG_THREADS), 0, ST, __VA_ARGS__);                              \
      else e_ = launch_pdl(KERN<float, 1>, GRID, dim3(UG_THREADS), 0, ST, __VA_ARGS__);                                       \
    } else {                                                                                                                  \
      if ((V) == 8) e_ = launch_pdl(KERN<__nv_bfloat16, 8>, GRID, dim3(UG_THREADS), 0, ST, __VA_ARGS__);                      \
      else e_ = launch_pdl(KERN<__nv_bfloat16, 1>, GRID, dim3(UG_THREADS), 0, ST, __VA_ARGS__);                               \
    }                                                                                                                         \
    if (e_ != cudaSuccess) return fail((int)e_, #KERN ": %s", cudaGetErrorString(e_));                                        \
    return check_launch(#KERN);                                                                                               \
  } while (0)

static inline dim3 ug_grid(long long work) {
  long long g = (work + UG_THREADS - 1) / UG_THREADS;
  if (g > 148 * 8 * 4) g = 148 * 8 * 4;
  if (g < 1) g = 1;
  return dim3((unsigned)g);
}

// ---- max-pool 2x2 (nn.MaxPool2d(2): floor) -----------------------------------------------------------------------------
template <typename T, int V>
__global__ void __launch_bounds__(UG_THREADS) maxpool_fwd_kernel(const T* __restrict__ x, T* __restrict__ y,
                                                                 uint8_t* __restrict__ idx, long long n, int h, int w, int c) {
  griddep_wait();
  const int ho = h / 2, wo = w / 2, groups = c / V;
  const long long total = n * ho * wo * groups;
  for (long long i = (long long)blockIdx.x * UG_THREADS + threadIdx.x; i < total; i += (long long)gridDim.x * UG_THREADS) {
    const int g = (int)(i % groups);
    long long p = i / groups;
    const int xo = (int)(p % wo);
    p /= wo;
    const int yo = (int)(p % ho);
    const long long img = p / ho;
    const T* base = x + ((img * h + 2 * yo) * w + 2 * xo) * c + (long long)g * V;
    float v[4][V];
    na_load<T, V>(base, v[0]);
    na_load<T, V>(base + c, v[1]);
    na_load<T, V>(base + (long long)w * c, v[2]);
    na_load<T, V>(base + (long long)w * c + c, v[3]);
    float best[V];
    uint8_t bi[V];
#pragma unroll
    for (int k = 0; k < V; ++k) {
      best[k] = v[0][k];
      bi[k] = 0;
#pragma unroll
      for (int q = 1; q < 4; ++q)
        if (v[q][k] > best[k]) { best[k] = v[q][k]; bi[k] = (uint8_t)q; }  // strict: the first maximum wins, as torch
    }
    const long long o = ((img * ho + yo) * wo + xo) * c + (long long)g * V;
    na_store<T, V>(y + o, best);
#pragma unroll
    for (int k = 0; k < V; ++k) idx[o + k] = bi[k];
  }
}

template <typename T, int V>
__global__ void __launch_bounds__(UG_THREADS) maxpool_bwd_kernel(const T* __restrict__ dy, const uint8_t* __restrict__ idx,
                                                                 T* __restrict__ dx, long long n, int h, int w, int c) {
  griddep_wait();
  const int ho = h / 2, wo = w / 2, groups = c / V;
  const long long total = n * h * w * groups;
  for (long long i = (long long)blockIdx.x * UG_THREADS + threadIdx.x; i < total; i += (long long)gridDim.x * UG_THREADS) {
    const int g = (int)(i % groups);
    long long p = i / groups;
    const int xi = (int)(p % w);
    p /= w;
    const int yi = (int)(p % h);
    const long long img = p / h;
    float o[V];
#pragma unroll
    for (int k = 0; k < V; ++k) o[k] = 0.f;
    const int yo = yi >> 1, xo = xi >> 1;
    if (yo < ho && xo < wo) {
      const long long q = ((img * ho + yo) * wo + xo) * c + (long long)g * V;
      float d[V];
      na_load<T, V>(dy + q, d);
      const int local = (yi & 1) * 2 + (xi & 1);
#pragma unroll
      for (int k = 0; k < V; ++k) o[k] = idx[q + k] == local ? d[k] : 0.f;
    }
    na_store<T, V>(dx + ((img * h + yi) * w + xi) * c + (long long)g * V, o);
  }
}

// ---- bilinear x2 (align_corners=True) + zero pad + channel concat --------------------------------------------------------
struct UpArgs {
  const void* x1;   // [n][h1][w1][c1]    the coarse map (up-sampled)
  const void* x2;   // [n][h][w][c2]      the skip connection
  void* out;        // fwd: [n][h][w][c2 + c1]
  const void* dout; // bwd
  void* dx1;
  void* dx2;
  long long n;
  int h1, w1, c1, h, w, c2, pad_t, pad_l;
  float ry, rx;     // (h1 - 1) / (2 h1 - 1), (w1 - 1) / (2 w1 - 1): torch's align_corners source-index scale
};

__device__ __forceinline__ void up_src(int u, float r, int n_in, int& i0, int& i1, float& f) {
  const float s = r * (float)u;
  i0 = (int)s;
  if (i0 > n_in - 1) i0 = n_in - 1;
  i1 = i0 + (i0 < n_in - 1 ? 1 : 0);
  f = s - (float)i0;
}

template <typename T, int V>
__global__ void __launch_bounds__(UG_THREADS) upcat_fwd_kernel(const UpArgs A) {
  griddep_wait();
  const int ct = A.c1 + A.c2, groups = ct / V, g2 = A.c2 / V;
  const long long total = A.n * A.h * A.w * groups;
  const T* x1 = reinterpret_cast<const T*>(A.x1);
  const T* x2 = reinterpret_cast<const T*>(A.x2);
  T* out = reinterpret_cast<T*>(A.out);
  for (long long i = (long long)blockIdx.x * UG_THREADS + threadIdx.x; i < total; i += (long long)gridDim.x * UG_THREADS) {
    const int g = (int)(i % groups);
    long long p = i / groups;
    const int X = (int)(p % A.w);
    p /= A.w;
    const int Y = (int)(p % A.h);
    const long long img = p / A.h;
    float o[V];
    if (g < g2) {
      na_load<T, V>(x2 + ((img * A.h + Y) * A.w + X) * A.c2 + (long long)g * V, o);
    } else {
      const int uy = Y - A.pad_t, ux = X - A.pad_l;
#pragma unroll
      for (int k = 0; k < V; ++k) o[k] = 0.f;
      if (uy >= 0 && uy < 2 * A.h1 && ux >= 0 && ux < 2 * A.w1) {
        int y0, y1, x0, x1i;
        float fy, fx;
        up_src(uy, A.ry, A.h1, y0, y1, fy);
        up_src(ux, A.rx, A.w1, x0, x1i, fx);
        const long long cg = (long long)(g - g2) * V;
        float a[V], b[V], c[V], d[V];
        na_load<T, V>(x1 + ((img * A.h1 + y0) * A.w1 + x0) * A.c1 + cg, a);
        na_load<T, V>(x1 + ((img * A.h1 + y0) * A.w1 + x1i) * A.c1 + cg, b);
        na_load<T, V>(x1 + ((img * A.h1 + y1) * A.w1 + x0) * A.c1 + cg, c);
        na_load<T, V>(x1 + ((img * A.h1 + y1) * A.w1 + x1i) * A.c1 + cg, d);
#pragma unroll
        for (int k = 0; k < V; ++k)
          o[k] = (1.f - fy) * ((1.f - fx) * a[k] + fx * b[k]) + fy * ((1.f - fx) * c[k] + fx * d[k]);
      }
    }
    na_store<T, V>(out + ((img * A.h + Y) * A.w + X) * ct + (long long)g * V, o);
  }
}

template <typename T, int V>
__global__ void __launch_bounds__(UG_THREADS) upcat_bwd_kernel(const UpArgs A) {
  griddep_wait();
  const int ct = A.c1 + A.c2, g2 = A.c2 / V, g1 = A.c1 / V;
  const long long n2 = A.n * A.h * A.w * g2, n1 = A.n * A.h1 * A.w1 * g1;
  const T* dout = reinterpret_cast<const T*>(A.dout);
  T* dx1 = reinterpret_cast<T*>(A.dx1);
  T* dx2 = reinterpret_cast<T*>(A.dx2);
  for (long long i = (long long)blockIdx.x * UG_THREADS + threadIdx.x; i < n2 + n1; i += (long long)gridDim.x * UG_THREADS) {
    float o[V];
    if (i < n2) {  // d(x2): the first c2 channels of d(out)
      const int g = (int)(i % g2);
      const long long p = i / g2;
      na_load<T, V>(dout + p * ct + (long long)g * V, o);
      na_store<T, V>(dx2 + p * A.c2 + (long long)g * V, o);
      continue;
    }
    // d(x1)[y][x]: every up-sampled pixel whose 2x2 interpolation footprint contains (y, x), gathered
    const long long j = i - n2;
    const int g = (int)(j % g1);
    long long p = j / g1;
    const int x = (int)(p % A.w1);
    p /= A.w1;
    const int y = (int)(p % A.h1);
    const long long img = p / A.h1;
#pragma unroll
    for (int k = 0; k < V; ++k) o[k] = 0.f;
    const int uy_lo = A.ry > 0.f ? max(0, (int)floorf((float)(y - 1) / A.ry)) : 0;
    const int uy_hi = A.ry > 0.f ? min(2 * A.h1 - 1, (int)ceilf((float)(y + 1) / A.ry)) : 2 * A.h1 - 1;
    const int ux_lo = A.rx > 0.f ? max(0, (int)floorf((float)(x - 1) / A.rx)) : 0;
    const int ux_hi = A.rx > 0.f ? min(2 * A.w1 - 1, (int)ceilf((float)(x + 1) / A.rx)) : 2 * A.w1 - 1;
    for (int uy = uy_lo; uy <= uy_hi; ++uy) {
      const int Y = uy + A.pad_t;
      if (Y < 0 || Y >= A.h) continue;
      int y0, y1;
      float fy;
      up_src(uy, A.ry, A.h1, y0, y1, fy);
      const float wy = (y0 == y ? 1.f - fy : 0.f) + (y1 == y ? fy : 0.f);
      if (wy == 0.f) continue;
      for (int ux = ux_lo; ux <= ux_hi; ++ux) {
        const int X = ux + A.pad_l;
        if (X < 0 || X >= A.w) continue;
        int x0, x1i;
        float fx;
        up_src(ux, A.rx, A.w1, x0, x1i, fx);
        const float wx = (x0 == x ? 1.f - fx : 0.f) + (x1i == x ? fx : 0.f);
        if (wx == 0.f) continue;
        float d[V];
        na_load<T, V>(dout + ((img * A.h + Y) * A.w + X) * ct + A.c2 + (long long)g * V, d);
#pragma unroll
        for (int k = 0; k < V; ++k) o[k] = fmaf(wy * wx, d[k], o[k]);
      }
    }
    na_store<T, V>(dx1 + ((img * A.h1 + y) * A.w1 + x) * A.c1 + (long long)g * V, o);
  }
}

// ---- per-(image, channel) reductions over the pixels --------------------------------------------------------------------
// OP 0: sum, max, first arg-max of x.   OP 1: sum of dy * x.
// grid (chunks, n): a CTA reduces one pixel chunk of one image into workspace [n][chunks][3][c]; the last CTA of an image
// (counter[n]) combines the chunks in order: deterministic, first occurrence of the maximum as torch's adaptive_max_pool2d.
struct PoolArgs {
  const void* x;
  const void* dy;
  long long hw;
  int c, chunks;
  float* ws;              // [n][chunks][3][c]
  unsigned int* counter;  // [n], zeroed by the caller
  float* out_sum;         // [n][c]  OP 0: mean;  OP 1: sum(dy * x)
  float* out_max;         // [n][c]
  int* out_arg;           // [n][c]
};

template <typename T, int V, int OP>
__device__ __forceinline__ void pool_hw_body(const PoolArgs& A) {
  __shared__ float s_sum[UG_THREADS * (V > 1 ? V : 1)];
  __shared__ float s_max[OP == 0 ? UG_THREADS * (V > 1 ? V : 1) : 1];
  __shared__ int s_arg[OP == 0 ? UG_THREADS * (V > 1 ? V : 1) : 1];
  __shared__ int s_last;
  const int groups = A.c / V;
  const int cg = groups >= UG_THREADS ? UG_THREADS : groups;
  int lanes = 1;
  while (2 * lanes * cg <= UG_THREADS) lanes *= 2;
  const long long img = blockIdx.y;
  const long long per = (A.hw + A.chunks - 1) / A.chunks;
  const long long p0 = (long long)blockIdx.x * per, p1 = min(A.hw, p0 + per);
  const T* x = reinterpret_cast<const T*>(A.x) + img * A.hw * A.c;
  const T* dy = reinterpret_cast<const T*>(A.dy) + img * A.hw * A.c;
  float* wsl = A.ws + ((img * A.chunks + blockIdx.x) * 3) * A.c;
  for (int g0 = 0; g0 < groups; g0 += cg) {
    const int g = g0 + (int)(threadIdx.x % cg), lane = threadIdx.x / cg;
    float sum[V], mx[V];
    int arg[V];
#pragma unroll
    for (int k = 0; k < V; ++k) { sum[k] = 0.f; mx[k] = -INFINITY; arg[k] = 0x7fffffff; }
    if (g < groups && lane < lanes) {
      for (long long p = p0 + lane; p < p1; p += lanes) {
        float xv[V];
        na_load<T, V>(x + p * A.c + (long long)g * V, xv);
        if constexpr (OP == 0) {
#pragma unroll
          for (int k = 0; k < V; ++k) {
            sum[k] += xv[k];
            if (xv[k] > mx[k]) { mx[k] = xv[k]; arg[k] = (int)p; }
          }
        } else {
          float dv[V];
          na_load<T, V>(dy + p * A.c + (long long)g * V, dv);
#pragma unroll
          for (int k = 0; k < V; ++k) sum[k] = fmaf(dv[k], xv[k], sum[k]);
        }
      }
    }
#pragma unroll
    for (int k = 0; k < V; ++k) {
      s_sum[threadIdx.x * V + k] = sum[k];
      if constexpr (OP == 0) { s_max[threadIdx.x * V + k] = mx[k]; s_arg[threadIdx.x * V + k] = arg[k]; }
    }
    __syncthreads();
    for (int half = lanes >> 1; half > 0; half >>= 1) {
      if (lane < half) {
        const int me = threadIdx.x * V, other = (threadIdx.x + half * cg) * V;
#pragma unroll
        for (int k = 0; k < V; ++k) {
          s_sum[me + k] += s_sum[other + k];
          if constexpr (OP == 0) {
            const float ov = s_max[other + k];
            const int oa = s_arg[other + k];
            if (ov > s_max[me + k] || (ov == s_max[me + k] && oa < s_arg[me + k])) { s_max[me + k] = ov; s_arg[me + k] = oa; }
          }
        }
      }
      __syncthreads();
    }
    if ((int)threadIdx.x < cg && g0 + (int)threadIdx.x < groups) {
#pragma unroll
      for (int k = 0; k < V; ++k) {
        const int ch = (g0 + threadIdx.x) * V + k;
        wsl[ch] = s_sum[threadIdx.x * V + k];
        if constexpr (OP == 0) {
          wsl[A.c + ch] = s_max[threadIdx.x * V + k];
          reinterpret_cast<int*>(wsl)[2 * A.c + ch] = s_arg[threadIdx.x * V + k];
        }
      }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {  // one cumulative fence behind the loop's closing __syncthreads, by the signalling thread
    __threadfence();
    s_last = atomicAdd(A.counter + img, 1u) == (unsigned)A.chunks - 1;
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  for (int ch = threadIdx.x; ch < A.c; ch += UG_THREADS) {
    float sum = 0.f, mx = -INFINITY;
    int arg = 0;
    for (int q = 0; q < A.chunks; ++q) {
      const float* wq = A.ws + ((img * A.chunks + q) * 3) * A.c;
      sum += __ldcg(wq + ch);
      if constexpr (OP == 0) {
        const float v = __ldcg(wq + A.c + ch);
        if (v > mx) { mx = v; arg = __ldcg(reinterpret_cast<const int*>(wq) + 2 * A.c + ch); }  // chunks are in pixel order
      }
    }
    if constexpr (OP == 0) {
      A.out_sum[img * A.c + ch] = sum / (float)A.hw;
      A.out_max[img * A.c + ch] = mx;
      A.out_arg[img * A.c + ch] = arg;
    } else {
      A.out_sum[img * A.c + ch] = sum;
    }
  }
}
template <typename T, int V>
__global__ void __launch_bounds__(UG_THREADS) pool_hw_kernel(const PoolArgs A) {
  griddep_wait();
  pool_hw_body<T, V, 0>(A);
}
template <typename T, int V>
__global__ void __launch_bounds__(UG_THREADS) dot_hw_kernel(const PoolArgs A) {
  griddep_wait();
  pool_hw_body<T, V, 1>(A);
}

// ---- CBAM channel-gate MLP: Linear(c -> hid), ReLU, Linear(hid -> c) on the mean and the max vector, summed, sigmoid ----
constexpr int MLP_MAX_C = 2048, MLP_MAX_HID = 128;
struct MlpArgs {
  const float *avg, *mx;              // [n][c]
  const float *w1, *b1, *w2, *b2;     // nn.Linear layout: w1 [hid][c], w2 [c][hid]
  int c, hid, n;
  float *pre;                         // [n][2][hid]  pre-activations of the hidden layer (mean branch, max branch)
  float *scale;                       // [n][c]       sigmoid(MLP(avg) + MLP(max))
  // backward
  const float* dscale;                // [n][c]
  float *ds;                          // [n][c]       d(pre-sigmoid)
  float *dpre;                        // [n][2][hid]
  float *davg, *dmax;                 // [n][c]
  float *dw1, *db1, *dw2, *db2;
};

__global__ void __launch_bounds__(UG_THREADS) cbam_mlp_fwd_kernel(const MlpArgs A) {
  griddep_wait();
  __shared__ float s_v[2][MLP_MAX_C];
  __shared__ float s_h[MLP_MAX_HID];
  const int img = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < A.c; i += UG_THREADS) {
    s_v[0][i] = A.avg[(size_t)img * A.c + i];
    s_v[1][i] = A.mx[(size_t)img * A.c + i];
  }
  __syncthreads();
  for (int j = warp; j < A.hid; j += UG_THREADS / 32) {
    float a = 0.f, m = 0.f;
    for (int i = lane; i < A.c; i += 32) {
      const float w = A.w1[(size_t)j * A.c + i];
      a = fmaf(w, s_v[0][i], a);
      m = fmaf(w, s_v[1][i], m);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      a += __shfl_xor_sync(0xffffffffu, a, off);
      m += __shfl_xor_sync(0xffffffffu, m, off);
    }
    if (lane == 0) {
      a += A.b1[j];
      m += A.b1[j];
      A.pre[((size_t)img * 2 + 0) * A.hid + j] = a;
      A.pre[((size_t)img * 2 + 1) * A.hid + j] = m;
      s_h[j] = fmaxf(a, 0.f) + fmaxf(m, 0.f);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < A.c; i += UG_THREADS) {
    float o = 2.f * A.b2[i];
    for (int j = 0; j < A.hid; ++j) o = fmaf(A.w2[(size_t)i * A.hid + j], s_h[j], o);
    A.scale[(size_t)img * A.c + i] = 1.f / (1.f + __expf(-o));
  }
}

// per image: d(pre-sigmoid), the hidden layer's gradients, d(avg), d(max)
__global__ void __launch_bounds__(UG_THREADS) cbam_mlp_bwd_kernel(const MlpArgs A) {
  griddep_wait();
  __shared__ float s_ds[MLP_MAX_C];
  __shared__ float s_dp[2][MLP_MAX_HID];
  const int img = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < A.c; i += UG_THREADS) {
    const float s = A.scale[(size_t)img * A.c + i];
    const float d = A.dscale[(size_t)img * A.c + i] * s * (1.f - s);
    s_ds[i] = d;
    A.ds[(size_t)img * A.c + i] = d;
  }
  __syncthreads();
  for (int j = warp; j < A.hid; j += UG_THREADS / 32) {
    float dh = 0.f;
    for (int i = lane; i < A.c; i += 32) dh = fmaf(A.w2[(size_t)i * A.hid + j], s_ds[i], dh);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) dh += __shfl_xor_sync(0xffffffffu, dh, off);
    if (lane == 0) {
      const float pa = A.pre[((size_t)img * 2 + 0) * A.hid + j], pm = A.pre[((size_t)img * 2 + 1) * A.hid + j];
      const float da = pa > 0.f ? dh : 0.f, dm = pm > 0.f ? dh : 0.f;
      s_dp[0][j] = da;
      s_dp[1][j] = dm;
      A.dpre[((size_t)img * 2 + 0) * A.hid + j] = da;
      A.dpre[((size_t)img * 2 + 1) * A.hid + j] = dm;
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < A.c; i += UG_THREADS) {
    float da = 0.f, dm = 0.f;
    for (int j = 0; j < A.hid; ++j) {
      const float w = A.w1[(size_t)j * A.c + i];
      da = fmaf(w, s_dp[0][j], da);
      dm = fmaf(w, s_dp[1][j], dm);
    }
    A.davg[(size_t)img * A.c + i] = da;
    A.dmax[(size_t)img * A.c + i] = dm;
  }
}

// the weight gradients: one thread per element, summed over the images in a fixed order
__global__ void __launch_bounds__(UG_THREADS) cbam_mlp_wgrad_kernel(const MlpArgs A) {
  griddep_wait();
  const int nw = A.c * A.hid;
  const int total = 2 * nw + A.c + A.hid;
  for (int e = blockIdx.x * UG_THREADS + threadIdx.x; e < total; e += gridDim.x * UG_THREADS) {
    float acc = 0.f;
    if (e < nw) {  // dw2[i][j] = sum_n ds[n][i] * (relu(pre_avg) + relu(pre_max))[n][j]
      const int i = e / A.hid, j = e - i * A.hid;
      for (int n = 0; n < A.n; ++n) {
        const float h = fmaxf(A.pre[((size_t)n * 2) * A.hid + j], 0.f) + fmaxf(A.pre[((size_t)n * 2 + 1) * A.hid + j], 0.f);
        acc = fmaf(A.ds[(size_t)n * A.c + i], h, acc);
      }
      A.dw2[e] = acc;
    } else if (e < 2 * nw) {  // dw1[j][i] = sum_n dpre_avg[n][j] * avg[n][i] + dpre_max[n][j] * max[n][i]
      const int q = e - nw, j = q / A.c, i = q - j * A.c;
      for (int n = 0; n < A.n; ++n) {
        acc = fmaf(A.dpre[((size_t)n * 2) * A.hid + j], A.avg[(size_t)n * A.c + i], acc);
        acc = fmaf(A.dpre[((size_t)n * 2 + 1) * A.hid + j], A.mx[(size_t)n * A.c + i], acc);
      }
      A.dw1[q] = acc;
    } else if (e < 2 * nw + A.c) {  // db2[i] = 2 * sum_n ds[n][i]   (b2 enters through both branches)
      const int i = e - 2 * nw;
      for (int n = 0; n < A.n; ++n) acc += A.ds[(size_t)n * A.c + i];
      A.db2[i] = 2.f * acc;
    } else {
      const int j = e - 2 * nw - A.c;
      for (int n = 0; n < A.n; ++n) acc += A.dpre[((size_t)n * 2) * A.hid + j] + A.dpre[((size_t)n * 2 + 1) * A.hid + j];
      A.db1[j] = acc;
    }
  }
}

// ---- gates --------------------------------------------------------------------------------------------------------------
// MODE 0: s[n][c] fp32 (channel gate).   MODE 1: s[n][pixel] of the tensor's dtype (spatial gate).
struct GateArgs {
  const void* x;
  const void* dy;
  void* out;
  const void* s;
  long long n, hw;
  int c;
  // channel gate backward: the pooled branches' gradients
  const float* davg;
  const float* dmax;
  const int* arg;
};

template <typename T, int V, int MODE, bool BWD>
__device__ __forceinline__ void gate_body(const GateArgs& A) {
  const int groups = A.c / V;
  const long long total = A.n * A.hw * groups;
  const T* x = reinterpret_cast<const T*>(A.x);
  const T* dy = reinterpret_cast<const T*>(A.dy);
  T* out = reinterpret_cast<T*>(A.out);
  const float inv_hw = 1.f / (float)A.hw;
  for (long long i = (long long)blockIdx.x * UG_THREADS + threadIdx.x; i < total; i += (long long)gridDim.x * UG_THREADS) {
    const int g = (int)(i % groups);
    const long long p = i / groups;  // pixel over all images
    const long long img = p / A.hw;
    float v[V], o[V];
    na_load<T, V>((BWD ? dy : x) + p * A.c + (long long)g * V, v);
    if constexpr (MODE == 0) {
      const float* s = reinterpret_cast<const float*>(A.s) + img * A.c + g * V;
      const int pin = (int)(p - img * A.hw);
#pragma unroll
      for (int k = 0; k < V; ++k) {
        o[k] = v[k] * s[k];
        if constexpr (BWD) {
          const long long q = img * A.c + g * V + k;
          o[k] += A.davg[q] * inv_hw + (A.arg[q] == pin ? A.dmax[q] : 0.f);
        }
      }
    } else {
      const float s = DT<T>::to_f(reinterpret_cast<const T*>(A.s)[p]);
#pragma unroll
      for (int k = 0; k < V; ++k) o[k] = v[k] * s;
    }
    na_store<T, V>(out + p * A.c + (long long)g * V, o);
  }
}
template <typename T, int V>
__global__ void __launch_bounds__(UG_THREADS) gate_c_fwd_kernel(const GateArgs A) { griddep_wait(); gate_body<T, V, 0, false>(A); }
template <typename T, int V>
__global__ void __launch_bounds__(UG_THREADS) gate_c_bwd_kernel(const GateArgs A) { griddep_wait(); gate_body<T, V, 0, true>(A); }
template <typename T, int V>
__global__ void __launch_bounds__(UG_THREADS) gate_p_kernel(const GateArgs A) { griddep_wait(); gate_body<T, V, 1, false>(A); }

// ---- per-pixel reductions over the channels: a warp per pixel -----------------------------------------------------------
// OP 0: mean, max, first arg-max of x -> o[pixel][2], arg[pixel].   OP 1: sum_c dy * x -> o[pixel] (dtype T).
template <typename T, int V, int OP>
__device__ __forceinline__ void chan_body(const T* __restrict__ x, const T* __restrict__ dy, T* __restrict__ o,
                                          int* __restrict__ arg, long long npix, int c) {
  const int lane = threadIdx.x & 31, groups = c / V;
  const long long warp0 = ((long long)blockIdx.x * UG_THREADS + threadIdx.x) >> 5, nwarps = ((long long)gridDim.x * UG_THREADS) >> 5;
  for (long long p = warp0; p < npix; p += nwarps) {
    float sum = 0.f, mx = -INFINITY;
    int am = 0x7fffffff;
    for (int g = lane; g < groups; g += 32) {
      float xv[V];
      na_load<T, V>(x + p * c + (long long)g * V, xv);
      if constexpr (OP == 0) {
#pragma unroll
        for (int k = 0; k < V; ++k) {
          sum += xv[k];
          if (xv[k] > mx) { mx = xv[k]; am = g * V + k; }
        }
      } else {
        float dv[V];
        na_load<T, V>(dy + p * c + (long long)g * V, dv);
#pragma unroll
        for (int k = 0; k < V; ++k) sum = fmaf(dv[k], xv[k], sum);
      }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      sum += __shfl_xor_sync(0xffffffffu, sum, off);
      if constexpr (OP == 0) {
        const float om = __shfl_xor_sync(0xffffffffu, mx, off);
        const int oa = __shfl_xor_sync(0xffffffffu, am, off);
        if (om > mx || (om == mx && oa < am)) { mx = om; am = oa; }
      }
    }
    if (lane == 0) {
      if constexpr (OP == 0) {
        o[2 * p] = DT<T>::from_f(sum / (float)c);
        o[2 * p + 1] = DT<T>::from_f(mx);
        arg[p] = am;
      } else {
        o[p] = DT<T>::from_f(sum);
      }
    }
  }
}
template <typename T, int V>
__global__ void __launch_bounds__(UG_THREADS) chan_pool_kernel(const T* x, T* o, int* arg, long long npix, int c) {
  griddep_wait();
  chan_body<T, V, 0>(x, nullptr, o, arg, npix, c);
}
template <typename T, int V>
__global__ void __launch_bounds__(UG_THREADS) chan_dot_kernel(const T* x, const T* dy, T* o, long long npix, int c) {
  griddep_wait();
  chan_body<T, V, 1>(x, dy, o, nullptr, npix, c);
}
// d(x) of the channel pooling: d(mean)/C everywhere + d(max) at the arg-max channel
template <typename T, int V>
__global__ void __launch_bounds__(UG_THREADS) chan_pool_bwd_kernel(const T* __restrict__ dpool, const int* __restrict__ arg,
                                                                   T* __restrict__ dx, long long npix, int c) {
  griddep_wait();
  const int groups = c / V;
  const float inv_c = 1.f / (float)c;
  for (long long i = (long long)blockIdx.x * UG_THREADS + threadIdx.x; i < npix * groups; i += (long long)gridDim.x * UG_THREADS) {
    const int g = (int)(i % groups);
    const long long p = i / groups;
    const float dm = DT<T>::to_f(dpool[2 * p]) * inv_c, dmx = DT<T>::to_f(dpool[2 * p + 1]);
    const int a = arg[p];
    float o[V];
#pragma unroll
    for (int k = 0; k < V; ++k) o[k] = dm + (a == g * V + k ? dmx : 0.f);
    na_store<T, V>(dx + p * c + (long long)g * V, o);
  }
}

static int ug_check(const void* p, int dtype, long long n, int c) {
  if (!p) return fail(CGAT_EINVAL, "null tensor");
  if (n < 1 || c < 1) return fail(CGAT_EINVAL, "bad geometry");
  if (dtype != CGAT_F32 && dtype != CGAT_BF16) return fail(CGAT_EINVAL, "bad dtype %d", dtype);
  return 0;
}

}  // namespace cgat

using namespace cgat;

extern "C" int cgat_maxpool2_fwd(const void* x, void* y, uint8_t* idx, int32_t dtype, int64_t n, int32_t h, int32_t w,
                                 int32_t c, void* stream) {
  if (int rc = ug_check(x, dtype, n, c)) return rc;
  if (!y || !idx || h < 2 || w < 2) return fail(CGAT_EINVAL, "null argument or map smaller than the window");
  const int v = ew_vec(dtype, c, x, nullptr, y);
  const long long work = (long long)n * (h / 2) * (w / 2) * (c / v);
#define ARGS(T) (const T*)x, (T*)y, idx, (long long)n, (int)h, (int)w, (int)c
  if (dtype == CGAT_F32) UG_LAUNCH(maxpool_fwd_kernel, ug_grid(work), (cudaStream_t)stream, dtype, v, ARGS(float));
  UG_LAUNCH(maxpool_fwd_kernel, ug_grid(work), (cudaStream_t)stream, dtype, v, ARGS(__nv_bfloat16));
#undef ARGS
}

extern "C" int cgat_maxpool2_bwd(const void* dy, const uint8_t* idx, void* dx, int32_t dtype, int64_t n, int32_t h, int32_t w,
                                 int32_t c, void* stream) {
  if (int rc = ug_check(dy, dtype, n, c)) return rc;
  if (!dx || !idx || h < 2 || w < 2) return fail(CGAT_EINVAL, "null argument or map smaller than the window");
  const int v = ew_vec(dtype, c, dy, nullptr, dx);
  const long long work = (long long)n * h * w * (c / v);
#define ARGS(T) (const T*)dy, idx, (T*)dx, (long long)n, (int)h, (int)w, (int)c
  if (dtype == CGAT_F32) UG_LAUNCH(maxpool_bwd_kernel, ug_grid(work), (cudaStream_t)stream, dtype, v, ARGS(float));
  UG_LAUNCH(maxpool_bwd_kernel, ug_grid(work), (cudaStream_t)stream, dtype, v, ARGS(__nv_bfloat16));
#undef ARGS
}

static int up_args(UpArgs& A, int64_t n, int32_t h1, int32_t w1, int32_t c1, int32_t h, int32_t w, int32_t c2) {
  if (n < 1 || h1 < 1 || w1 < 1 || c1 < 1 || c2 < 1 || h < 2 * h1 || w < 2 * w1)
    return fail(CGAT_EINVAL, "up-sample + concat: the skip map must be at least twice the coarse map");
  A.n = n; A.h1 = h1; A.w1 = w1; A.c1 = c1; A.h = h; A.w = w; A.c2 = c2;
  A.pad_t = (h - 2 * h1) / 2;
  A.pad_l = (w - 2 * w1) / 2;
  A.ry = h1 > 1 ? (float)(h1 - 1) / (float)(2 * h1 - 1) : 0.f;
  A.rx = w1 > 1 ? (float)(w1 - 1) / (float)(2 * w1 - 1) : 0.f;
  return 0;
}

extern "C" int cgat_upcat_fwd(const void* x1, const void* x2, void* out, int32_t dtype, int64_t n, int32_t h1, int32_t w1,
                              int32_t c1, int32_t h, int32_t w, int32_t c2, void* stream) {
  if (!x1 || !x2 || !out) return fail(CGAT_EINVAL, "null argument");
  if (dtype != CGAT_F32 && dtype != CGAT_BF16) return fail(CGAT_EINVAL, "bad dtype %d", dtype);
  UpArgs A{};
  if (int rc = up_args(A, n, h1, w1, c1, h, w, c2)) return rc;
  A.x1 = x1; A.x2 = x2; A.out = out;
  int v = ew_vec(dtype, c1, x1, x2, out);
  if (c2 % v) v = 1;
  UG_LAUNCH(upcat_fwd_kernel, ug_grid((long long)n * h * w * ((c1 + c2) / v)), (cudaStream_t)stream, dtype, v, A);
}

extern "C" int cgat_upcat_bwd(const void* dout, void* dx1, void* dx2, int32_t dtype, int64_t n, int32_t h1, int32_t w1,
                              int32_t c1, int32_t h, int32_t w, int32_t c2, void* stream) {
  if (!dout || !dx1 || !dx2) return fail(CGAT_EINVAL, "null argument");
  if (dtype != CGAT_F32 && dtype != CGAT_BF16) return fail(CGAT_EINVAL, "bad dtype %d", dtype);
  UpArgs A{};
  if (int rc = up_args(A, n, h1, w1, c1, h, w, c2)) return rc;
  A.dout = dout; A.dx1 = dx1; A.dx2 = dx2;
  int v = ew_vec(dtype, c1, dout, dx1, dx2);
  if (c2 % v) v = 1;
  UG_LAUNCH(upcat_bwd_kernel, ug_grid((long long)n * (h * w * (c2 / v) + h1 * w1 * (c1 / v))), (cudaStream_t)stream, dtype, v, A);
}

extern "C" int64_t cgat_pool_hw_workspace_bytes(int64_t n, int64_t hw, int32_t c) {
  if (n < 1 || hw < 1 || c < 1) return 0;
  long long chunks = (hw + 255) / 256;
  if (chunks > 64) chunks = 64;
  return (int64_t)n * chunks * 3 * c * 4 + n * 4;
}

static int pool_launch(bool dot, const void* x, const void* dy, int32_t dtype, int64_t n, int64_t hw, int32_t c, void* workspace,
                       float* out_sum, float* out_max, int32_t* out_arg, cudaStream_t st) {
  if (int rc = ug_check(x, dtype, n, c)) return rc;
  if (!workspace || !out_sum || (!dot && (!out_max || !out_arg)) || (dot && !dy)) return fail(CGAT_EINVAL, "null argument");
  if (n > 65535) return fail(CGAT_EUNSUPPORTED, "pool_hw: n > 65535");
  PoolArgs A{};
  long long chunks = (hw + 255) / 256;
  if (chunks > 64) chunks = 64;
  A.x = x; A.dy = dy; A.hw = hw; A.c = c; A.chunks = (int)chunks;
  A.ws = (float*)workspace;
  A.counter = reinterpret_cast<unsigned int*>((char*)workspace + (size_t)n * chunks * 3 * c * 4);
  A.out_sum = out_sum; A.out_max = out_max; A.out_arg = out_arg;
  cudaError_t e = cudaMemsetAsync(A.counter, 0, (size_t)n * 4, st);
  if (e != cudaSuccess) return fail((int)e, "pool counter memset: %s", cudaGetErrorString(e));
  const int v = ew_vec(dtype, c, x, dy, nullptr);
  const dim3 grid((unsigned)chunks, (unsigned)n);
  if (dot) UG_LAUNCH(dot_hw_kernel, grid, st, dtype, v, A);
  UG_LAUNCH(pool_hw_kernel, grid, st, dtype, v, A);
}

extern "C" int cgat_pool_hw(const void* x, int32_t dtype, int64_t n, int64_t hw, int32_t c, void* workspace, float* avg,
                            float* mx, int32_t* argmax, void* stream) {
  return pool_launch(false, x, nullptr, dtype, n, hw, c, workspace, avg, mx, argmax, (cudaStream_t)stream);
}

extern "C" int cgat_dot_hw(const void* x, const void* dy, int32_t dtype, int64_t n, int64_t hw, int32_t c, void* workspace,
                           float* out, void* stream) {
  return pool_launch(true, x, dy, dtype, n, hw, c, workspace, out, nullptr, nullptr, (cudaStream_t)stream);
}

extern "C" int cgat_cbam_mlp_fwd(const float* avg, const float* mx, const float* w1, const float* b1, const float* w2,
                                 const float* b2, int32_t n, int32_t c, int32_t hid, float* pre, float* scale, void* stream) {
  if (!avg || !mx || !w1 || !b1 || !w2 || !b2 || !pre || !scale) return fail(CGAT_EINVAL, "null argument");
  if (n < 1 || c < 1 || c > MLP_MAX_C || hid < 1 || hid > MLP_MAX_HID)
    return fail(CGAT_EUNSUPPORTED, "CBAM MLP serves c <= %d, hidden <= %d", MLP_MAX_C, MLP_MAX_HID);
  MlpArgs A{};
  A.avg = avg; A.mx = mx; A.w1 = w1; A.b1 = b1; A.w2 = w2; A.b2 = b2; A.n = n; A.c = c; A.hid = hid; A.pre = pre; A.scale = scale;
  cudaError_t e = launch_pdl(cbam_mlp_fwd_kernel, dim3(n), dim3(UG_THREADS), 0, (cudaStream_t)stream, A);
  if (e != cudaSuccess) return fail((int)e, "cbam_mlp_fwd_kernel: %s", cudaGetErrorString(e));
  return check_launch("cbam_mlp_fwd_kernel");
}

extern "C" int cgat_cbam_mlp_bwd(const float* dscale, const float* scale, const float* pre, const float* avg, const float* mx,
                                 const float* w1, const float* w2, int32_t n, int32_t c, int32_t hid, float* ds, float* dpre,
                                 float* davg, float* dmax, float* dw1, float* db1, float* dw2, float* db2, void* stream) {
  if (!dscale || !scale || !pre || !avg || !mx || !w1 || !w2 || !ds || !dpre || !davg || !dmax || !dw1 || !db1 || !dw2 || !db2)
    return fail(CGAT_EINVAL, "null argument");
  if (n < 1 || c < 1 || c > MLP_MAX_C || hid < 1 || hid > MLP_MAX_HID)
    return fail(CGAT_EUNSUPPORTED, "CBAM MLP serves c <= %d, hidden <= %d", MLP_MAX_C, MLP_MAX_HID);
  MlpArgs A{};
  A.dscale = dscale; A.scale = const_cast<float*>(scale); A.pre = const_cast<float*>(pre); A.avg = avg; A.mx = mx; A.w1 = w1; A.w2 = w2; A.n = n; A.c = c; A.hid = hid;
  A.ds = ds; A.dpre = dpre; A.davg = davg; A.dmax = dmax; A.dw1 = dw1; A.db1 = db1; A.dw2 = dw2; A.db2 = db2;
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = launch_pdl(cbam_mlp_bwd_kernel, dim3(n), dim3(UG_THREADS), 0, st, A);
  if (e != cudaSuccess) return fail((int)e, "cbam_mlp_bwd_kernel: %s", cudaGetErrorString(e));
  const int total = 2 * c * hid + c + hid;
  e = launch_pdl(cbam_mlp_wgrad_kernel, dim3((total + UG_THREADS - 1) / UG_THREADS), dim3(UG_THREADS), 0, st, A);
  if (e != cudaSuccess) return fail((int)e, "cbam_mlp_wgrad_kernel: %s", cudaGetErrorString(e));
  return check_launch("cbam_mlp_bwd");
}

static GateArgs gate_args(const void* x, const void* dy, void* out, const void* s, int64_t n, int64_t hw, int32_t c) {
  GateArgs A{};
  A.x = x; A.dy = dy; A.out = out; A.s = s; A.n = n; A.hw = hw; A.c = c;
  return A;
}

extern "C" int cgat_gate_channels_fwd(const void* x, const float* scale, void* y, int32_t dtype, int64_t n, int64_t hw,
                                      int32_t c, void* stream) {
  if (int rc = ug_check(x, dtype, n, c)) return rc;
  if (!scale || !y) return fail(CGAT_EINVAL, "null argument");
  const GateArgs A = gate_args(x, nullptr, y, scale, n, hw, c);
  const int v = ew_vec(dtype, c, x, nullptr, y);
  UG_LAUNCH(gate_c_fwd_kernel, ug_grid((long long)n * hw * (c / v)), (cudaStream_t)stream, dtype, v, A);
}

extern "C" int cgat_gate_channels_bwd(const void* dy, const float* scale, const float* davg, const float* dmax,
                                      const int32_t* argmax, void* dx, int32_t dtype, int64_t n, int64_t hw, int32_t c,
                                      void* stream) {
  if (int rc = ug_check(dy, dtype, n, c)) return rc;
  if (!scale || !davg || !dmax || !argmax || !dx) return fail(CGAT_EINVAL, "null argument");
  GateArgs A = gate_args(nullptr, dy, dx, scale, n, hw, c);
  A.davg = davg; A.dmax = dmax; A.arg = argmax;
  const int v = ew_vec(dtype, c, dy, nullptr, dx);
  UG_LAUNCH(gate_c_bwd_kernel, ug_grid((long long)n * hw * (c / v)), (cudaStream_t)stream, dtype, v, A);
}

extern "C" int cgat_gate_pixels(const void* x, const void* s, void* y, int32_t dtype, int64_t n, int64_t hw, int32_t c,
                                void* stream) {
  if (int rc = ug_check(x, dtype, n, c)) return rc;
  if (!s || !y) return fail(CGAT_EINVAL, "null argument");
  const GateArgs A = gate_args(x, nullptr, y, s, n, hw, c);
  const int v = ew_vec(dtype, c, x, nullptr, y);
  UG_LAUNCH(gate_p_kernel, ug_grid((long long)n * hw * (c / v)), (cudaStream_t)stream, dtype, v, A);
}

extern "C" int cgat_chan_pool_fwd(const void* x, void* pooled, int32_t* argmax, int32_t dtype, int64_t npix, int32_t c,
                                  void* stream) {
  if (int rc = ug_check(x, dtype, npix, c)) return rc;
  if (!pooled || !argmax) return fail(CGAT_EINVAL, "null argument");
  const int v = ew_vec(dtype, c, x);
  const dim3 grid = ug_grid(npix * 32);
#define ARGS(T) (const T*)x, (T*)pooled, argmax, (long long)npix, (int)c
  if (dtype == CGAT_F32) UG_LAUNCH(chan_pool_kernel, grid, (cudaStream_t)stream, dtype, v, ARGS(float));
  UG_LAUNCH(chan_pool_kernel, grid, (cudaStream_t)stream, dtype, v, ARGS(__nv_bfloat16));
#undef ARGS
}

extern "C" int cgat_chan_pool_bwd(const void* dpooled, const int32_t* argmax, void* dx, int32_t dtype, int64_t npix, int32_t c,
                                  void* stream) {
  if (int rc = ug_check(dpooled, dtype, npix, c)) return rc;
  if (!argmax || !dx) return fail(CGAT_EINVAL, "null argument");
  const int v = ew_vec(dtype, c, dx);
#define ARGS(T) (const T*)dpooled, argmax, (T*)dx, (long long)npix, (int)c
  if (dtype == CGAT_F32) UG_LAUNCH(chan_pool_bwd_kernel, ug_grid(npix * (c / v)), (cudaStream_t)stream, dtype, v, ARGS(float));
  UG_LAUNCH(chan_pool_bwd_kernel, ug_grid(npix * (c / v)), (cudaStream_t)stream, dtype, v, ARGS(__nv_bfloat16));
#undef ARGS
}

extern "C" int cgat_chan_dot(const void* x, const void* dy, void* out, int32_t dtype, int64_t npix, int32_t c, void* stream) {
  if (int rc = ug_check(x, dtype, npix, c)) return rc;
  if (!dy || !out) return fail(CGAT_EINVAL, "null argument");
  const int v = ew_vec(dtype, c, x, dy);
  const dim3 grid = ug_grid(npix * 32);
#define ARGS(T) (const T*)x, (const T*)dy, (T*)out, (long long)npix, (int)c
  if (dtype == CGAT_F32) UG_LAUNCH(chan_dot_kernel, grid, (cudaStream_t)stream, dtype, v, ARGS(float));
  UG_LAUNCH(chan_dot_kernel, grid, (cudaStream_t)stream, dtype, v, ARGS(__nv_bfloat16));
#undef ARGS
}
