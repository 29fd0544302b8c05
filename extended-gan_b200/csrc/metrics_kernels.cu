// f4: the validation metrics of convolutional_gat/train.py:28-91 on the device.
//
// Per batch the reference undoes the loader's power transform (:54-55), sums the squared error (:56-58) and its
// de-normalised version (:72-75) and, after cloning y and y_hat to the CPU, binarises them at a threshold and counts
// agreement / true positives / false positives / false negatives (utils.py:135-167).  One pass over y and y_hat does
// all of it: per-thread partial sums in double, warp shuffle + shared-memory reduction, one atomicAdd per value per CTA.
//   out[0] = sum (y' - yh')^2            y' = y^(1/power), yh' = yh^(1/power)
//   out[1] = sum ((y' - yh') * norm_max)^2
//   out[2] = TP   out[3] = FP   out[4] = FN   out[5] = #(bin(y') == bin(yh'))
// bin(v): v < thr -> 0, then v >= thr -> 1 (the two in-place assignments of utils.py:138-141, in that order).
#include "common.cuh"

namespace cgat {

constexpr int MT_THREADS = 256;

__device__ __forceinline__ float metric_bin(float v, float thr) {
  if (v < thr) v = 0.f;   // y[y < mean] = 0
  if (v >= thr) v = 1.f;  // y[y >= mean] = 1
  return v;
}

template <typename T>
__global__ void __launch_bounds__(MT_THREADS) val_metrics_kernel(const T* __restrict__ y, const T* __restrict__ yh,
                                                                 long long n, float inv_power, float thr, float nmax,
                                                                 double* __restrict__ out) {
  double acc[6] = {0, 0, 0, 0, 0, 0};
  const bool unit = inv_power == 1.0f;
  for (long long i = (long long)blockIdx.x * MT_THREADS + threadIdx.x; i < n; i += (long long)gridDim.x * MT_THREADS) {
    float a = DT<T>::to_f(y[i]), b = DT<T>::to_f(yh[i]);
    if (!unit) { a = powf(a, inv_power); b = powf(b, inv_power); }
    const float d = a - b;
    acc[0] += (double)(d * d);
    const float dn = d * nmax;
    acc[1] += (double)(dn * dn);
    const float ba = metric_bin(a, thr), bb = metric_bin(b, thr);
    acc[2] += (bb == 1.f && ba == 1.f) ? 1.0 : 0.0;
    acc[3] += (bb == 1.f && ba == 0.f) ? 1.0 : 0.0;
    acc[4] += (bb == 0.f && ba == 1.f) ? 1.0 : 0.0;
    acc[5] += (ba == bb) ? 1.0 : 0.0;
  }
  __shared__ double red[6][MT_THREADS / 32];
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    double v = acc[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) red[k][threadIdx.x >> 5] = v;
  }
  __syncthreads();
  if (threadIdx.x < 6) {
    double v = 0;
    for (int w = 0; w < MT_THREADS / 32; ++w) v += red[threadIdx.x][w];
    atomicAdd(out + threadIdx.x, v);
  }
}

}  // namespace cgat

using namespace cgat;

extern "C" int cgat_val_metrics(const void* y, const void* y_hat, int64_t n, float power, float threshold,
                                float normalizing_max, int32_t dtype, double* out6, void* stream) {
  if (!y || !y_hat || !out6 || n <= 0) return fail(CGAT_EINVAL, "null argument or n <= 0");
  if (!(power > 0.f)) return fail(CGAT_EINVAL, "power must be positive");
  long long blocks = (n + MT_THREADS - 1) / MT_THREADS;
  if (blocks > 148 * 8) blocks = 148 * 8;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == CGAT_F32)
    val_metrics_kernel<float><<<(unsigned)blocks, MT_THREADS, 0, st>>>((const float*)y, (const float*)y_hat, n, 1.f / power,
                                                                       threshold, normalizing_max, out6);
  else if (dtype == CGAT_BF16)
    val_metrics_kernel<__nv_bfloat16><<<(unsigned)blocks, MT_THREADS, 0, st>>>(
        (const __nv_bfloat16*)y, (const __nv_bfloat16*)y_hat, n, 1.f / power, threshold, normalizing_max, out6);
  else
    return fail(CGAT_EINVAL, "bad dtype %d", dtype);
  return check_launch("val_metrics_kernel");
}
