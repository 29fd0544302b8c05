// a12: the pieces of the reference train step that sit either side of the layer, as single launches.
//   loss   = MSE(y_hat, y) - 0.0005 * sum(y_hat)/numel          convolutional_gat/train.py:131
//   update = torch.optim.Adam(lr, weight_decay=0.01)             convolutional_gat/train.py:212
// All three kernels are streaming, HBM-bound: 16-byte vector accesses, grid sized to the SM count.
#include "common.cuh"

namespace cgat {

constexpr int EW_THREADS = 256;

template <typename T> struct Vec;  // 16-byte vector of T
template <> struct Vec<float> { static constexpr int N = 4; };
template <> struct Vec<__nv_bfloat16> { static constexpr int N = 8; };

template <typename T>
__device__ __forceinline__ void load_vec(const T* p, float* f) {
  uint4 v = *reinterpret_cast<const uint4*>(p);
  if constexpr (sizeof(T) == 4) {
    f[0] = __uint_as_float(v.x); f[1] = __uint_as_float(v.y); f[2] = __uint_as_float(v.z); f[3] = __uint_as_float(v.w);
  } else {
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      f[2 * k] = __uint_as_float(w[k] << 16);
      f[2 * k + 1] = __uint_as_float(w[k] & 0xffff0000u);
    }
  }
}
template <typename T>
__device__ __forceinline__ void store_vec(T* p, const float* f) {
  uint4 v;
  if constexpr (sizeof(T) == 4) {
    v.x = __float_as_uint(f[0]); v.y = __float_as_uint(f[1]); v.z = __float_as_uint(f[2]); v.w = __float_as_uint(f[3]);
  } else {
    __nv_bfloat162 b0 = __floats2bfloat162_rn(f[0], f[1]), b1 = __floats2bfloat162_rn(f[2], f[3]);
    __nv_bfloat162 b2 = __floats2bfloat162_rn(f[4], f[5]), b3 = __floats2bfloat162_rn(f[6], f[7]);
    v.x = *reinterpret_cast<uint32_t*>(&b0); v.y = *reinterpret_cast<uint32_t*>(&b1);
    v.z = *reinterpret_cast<uint32_t*>(&b2); v.w = *reinterpret_cast<uint32_t*>(&b3);
  }
  *reinterpret_cast<uint4*>(p) = v;
}

// d loss / d yhat = (2 (yhat - y) - lambda) / n * grad_scale
template <typename T>
__global__ void __launch_bounds__(EW_THREADS) loss_kernel(const T* __restrict__ yhat, const T* __restrict__ y,
                                                          T* __restrict__ dyhat, float* __restrict__ loss_out,
                                                          float* __restrict__ mse_out, long long n, float lambda,
                                                          float grad_scale) {
  constexpr int V = Vec<T>::N;
  const float inv_n = 1.f / (float)n;
  float acc = 0.f, acc2 = 0.f;
  const long long nvec = n / V;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (long long)gridDim.x * blockDim.x) {
    float a[V], b[V], g[V];
    load_vec<T>(yhat + i * V, a);
    load_vec<T>(y + i * V, b);
#pragma unroll
    for (int k = 0; k < V; ++k) {
      const float d = a[k] - b[k];
      acc += d * d - lambda * a[k];
      acc2 += d * d;
      g[k] = (2.f * d - lambda) * inv_n * grad_scale;
    }
    store_vec<T>(dyhat + i * V, g);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {  // tail (n % V elements)
    for (long long i = nvec * V; i < n; ++i) {
      const float a = DT<T>::to_f(yhat[i]), d = a - DT<T>::to_f(y[i]);
      acc += d * d - lambda * a;
      acc2 += d * d;
      dyhat[i] = DT<T>::from_f((2.f * d - lambda) * inv_n * grad_scale);
    }
  }
  __shared__ float s[2][EW_THREADS / 32];
  acc = warp_sum(acc);
  acc2 = warp_sum(acc2);
  if ((threadIdx.x & 31) == 0) { s[0][threadIdx.x >> 5] = acc; s[1][threadIdx.x >> 5] = acc2; }
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < EW_THREADS / 32 ? s[0][threadIdx.x] : 0.f;
    float v2 = threadIdx.x < EW_THREADS / 32 ? s[1][threadIdx.x] : 0.f;
    v = warp_sum(v);
    v2 = warp_sum(v2);
    if (threadIdx.x == 0) {
      atomicAdd(loss_out, v * inv_n);
      if (mse_out != nullptr) atomicAdd(mse_out, v2 * inv_n);
    }
  }
}

// torch.optim.Adam semantics (L2 weight decay folded into the gradient, bias-corrected moments).
__global__ void __launch_bounds__(EW_THREADS) adam_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                          float* __restrict__ m, float* __restrict__ v,
                                                          const long long* __restrict__ step_dev, long long step_host,
                                                          long long n, float lr, float b1, float b2, float eps, float wd,
                                                          float gscale) {
  griddep_wait();  // (launched with programmatic serialisation behind the parameter-gradient kernel, common.cuh)
  const float step = (float)(step_dev != nullptr ? *step_dev : step_host);
  const float bc1 = 1.f - powf(b1, step);
  const float bc2 = 1.f - powf(b2, step);
  const float step_size = lr / bc1;
  const float inv_sqrt_bc2 = rsqrtf(bc2);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float pi = p[i];
    const float gi = fmaf(wd, pi, g[i] * gscale);
    const float mi = fmaf(b1, m[i], (1.f - b1) * gi);
    const float vi = fmaf(b2, v[i], (1.f - b2) * gi * gi);
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) * inv_sqrt_bc2 + eps;
    p[i] = pi - step_size * (mi / denom);
  }
}

template <typename S, typename D>
__global__ void __launch_bounds__(EW_THREADS) cast_kernel(const S* __restrict__ src, D* __restrict__ dst, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    dst[i] = DT<D>::from_f(DT<S>::to_f(src[i]));
}

static int ew_grid(long long work_items) {
  long long g = (work_items + EW_THREADS - 1) / EW_THREADS;
  const long long cap = 148 * 8;  // 8 resident CTAs of 256 threads per SM on B200
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace cgat

using namespace cgat;

extern "C" int cgat_loss_fwd_bwd(const void* yhat, const void* y, void* dyhat, float* loss_out, float* mse_out,
                                 int64_t n, float lambda, float grad_scale, int dtype, void* stream) {
  if (!yhat || !y || !dyhat || !loss_out || n <= 0) return fail(CGAT_EINVAL, "null argument or n <= 0");
  if (!aligned16(yhat) || !aligned16(y) || !aligned16(dyhat)) return fail(CGAT_EALIGN, "pointers must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == CGAT_F32) {
    loss_kernel<float><<<ew_grid(n / 4 + 1), EW_THREADS, 0, st>>>((const float*)yhat, (const float*)y, (float*)dyhat,
                                                                   loss_out, mse_out, n, lambda, grad_scale);
  } else if (dtype == CGAT_BF16) {
    loss_kernel<__nv_bfloat16><<<ew_grid(n / 8 + 1), EW_THREADS, 0, st>>>(
        (const __nv_bfloat16*)yhat, (const __nv_bfloat16*)y, (__nv_bfloat16*)dyhat, loss_out, mse_out, n, lambda, grad_scale);
  } else {
    return fail(CGAT_EINVAL, "bad dtype %d", dtype);
  }
  return check_launch("loss_kernel");
}

extern "C" int cgat_adam_step(float* param, const float* grad, float* m, float* v, const int64_t* step_dev, int64_t n,
                              float lr, float beta1, float beta2, float eps, float weight_decay, float grad_scale,
                              void* stream) {
  if (!param || !grad || !m || !v || !step_dev || n <= 0) return fail(CGAT_EINVAL, "null argument or n <= 0");
  adam_kernel<<<ew_grid(n), EW_THREADS, 0, (cudaStream_t)stream>>>(param, grad, m, v, (const long long*)step_dev, 0, n, lr,
                                                                   beta1, beta2, eps, weight_decay, grad_scale);
  return check_launch("adam_kernel");
}

extern "C" int cgat_adam_step_at(float* param, const float* grad, float* m, float* v, int64_t step, int64_t n, float lr,
                                 float beta1, float beta2, float eps, float weight_decay, float grad_scale,
                                 void* stream) {
  if (!param || !grad || !m || !v || n <= 0 || step < 1) return fail(CGAT_EINVAL, "null argument, n <= 0 or step < 1");
  cudaError_t e = launch_pdl(adam_kernel, dim3(ew_grid(n)), dim3(EW_THREADS), 0, (cudaStream_t)stream, param, grad, m, v,
                             (const long long*)nullptr, (long long)step, (long long)n, lr, beta1, beta2, eps, weight_decay,
                             grad_scale);
  if (e != cudaSuccess) return fail((int)e, "adam_kernel: %s", cudaGetErrorString(e));
  return check_launch("adam_kernel");
}

extern "C" int cgat_cast(const void* src, int src_dtype, void* dst, int dst_dtype, int64_t n, void* stream) {
  if (!src || !dst || n <= 0) return fail(CGAT_EINVAL, "null argument or n <= 0");
  cudaStream_t st = (cudaStream_t)stream;
  const int g = ew_grid(n);
  if (src_dtype == CGAT_F32 && dst_dtype == CGAT_BF16)
    cast_kernel<float, __nv_bfloat16><<<g, EW_THREADS, 0, st>>>((const float*)src, (__nv_bfloat16*)dst, n);
  else if (src_dtype == CGAT_BF16 && dst_dtype == CGAT_F32)
    cast_kernel<__nv_bfloat16, float><<<g, EW_THREADS, 0, st>>>((const __nv_bfloat16*)src, (float*)dst, n);
  else
    return fail(CGAT_EINVAL, "unsupported cast %d -> %d", src_dtype, dst_dtype);
  return check_launch("cast_kernel");
}
