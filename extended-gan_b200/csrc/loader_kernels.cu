// f3: the KNMI loader's windowing / normalisation / layout change on the device.
//
// The reference loader (convolutional_gat/data_loaders/kmni_data_loader.py:72-127) turns a file of raw integer
// frames data[L, V, H, W] (values 0..254, preprocessing/kmni_dataset/__main__.py:76-110) into overlapping windows of 8
// frames with stride 1 (:79-85), divides by 254 (:75), applies pow(., power) (:76), splits every window into 4 input and
// 4 target frames (:92-94), crops (:95-96), and per batch copies 2 x N x 4 x V x H x W fp32 values to the GPU before
// permuting them to [N, H, W, T, V] (:114-126).  N consecutive windows share all but N+7 of their frames, so the host
// copy is 8 x 4 bytes per value of redundancy.  Here the raw frames cross PCIe ONCE as uint8 and one kernel gathers
//   x[n, h, w, t, v] = pow(frame[start[n] + t    , v, h, w] / 254, power)
//   y[n, h, w, t, v] = pow(frame[start[n] + 4 + t, v, h, w] / 254, power)
// straight into the pixel-record layout the conv-GAT kernels read.  One thread per output pixel: its T*V record is
// written with 16-byte stores; reads are coalesced along w.  fp32 output is bit-exact with the reference for power = 1
// (IEEE division); bf16 output rounds that value once.
#include "common.cuh"

namespace cgat {

constexpr int LD_THREADS = 128;
constexpr int LD_MAX_REC = 64;  // T * V elements per pixel record

// REC = steps * V when known at compile time (record kept in registers, 16-byte stores), 0 = generic
template <typename T, int REC>
__global__ void __launch_bounds__(LD_THREADS)
loader_gather_kernel(const uint8_t* __restrict__ frames, const int32_t* __restrict__ start, T* __restrict__ x,
                     T* __restrict__ y, int n, int V, int H, int W, int crop_h, int crop_w, int steps, float nmax,
                     float power) {
  const long long pix = (long long)blockIdx.x * LD_THREADS + threadIdx.x;
  const long long total = (long long)n * crop_h * crop_w;
  if (pix >= total) return;
  const int w = (int)(pix % crop_w);
  const int h = (int)((pix / crop_w) % crop_h);
  const int s = (int)(pix / ((long long)crop_w * crop_h));
  const int f0 = start[s];
  const int rec = REC ? REC : steps * V;
  const size_t plane = (size_t)H * W;
  const bool unit = power == 1.0f;
#pragma unroll 1
  for (int half = 0; half < 2; ++half) {
    T* dst = (half ? y : x) + pix * rec;
    // e = t * V + v: frame f0 (+ steps for the target) + t, vertex v  ->  source plane (f0 + ..) * V + e
    const uint8_t* src = frames + ((size_t)(f0 + half * steps) * V) * plane + (size_t)h * W + w;
    if constexpr (REC != 0) {
      float vals[REC];
#pragma unroll
      for (int e = 0; e < REC; ++e) {
        float v = __fdiv_rn((float)src[(size_t)e * plane], nmax);  // IEEE division, as torch's data / 254 (:75)
        if (!unit) v = powf(v, power);                              // t.pow(norm_data, power) (:76)
        vals[e] = v;
      }
      if constexpr (sizeof(T) == 4) {
#pragma unroll
        for (int q = 0; q < REC / 4; ++q)
          reinterpret_cast<float4*>(dst)[q] = make_float4(vals[4 * q], vals[4 * q + 1], vals[4 * q + 2], vals[4 * q + 3]);
      } else {
#pragma unroll
        for (int q = 0; q < REC / 8; ++q) {
          uint4 o;
          __nv_bfloat162 t2;
#define PK(a, b) (t2 = __floats2bfloat162_rn(a, b), *reinterpret_cast<uint32_t*>(&t2))
          o.x = PK(vals[8 * q], vals[8 * q + 1]); o.y = PK(vals[8 * q + 2], vals[8 * q + 3]);
          o.z = PK(vals[8 * q + 4], vals[8 * q + 5]); o.w = PK(vals[8 * q + 6], vals[8 * q + 7]);
#undef PK
          reinterpret_cast<uint4*>(dst)[q] = o;
        }
      }
    } else {
      for (int e = 0; e < rec; ++e) {
        float v = __fdiv_rn((float)src[(size_t)e * plane], nmax);
        if (!unit) v = powf(v, power);
        dst[e] = DT<T>::from_f(v);
      }
    }
  }
}

template <typename T>
static void loader_launch(int rec, unsigned grid, cudaStream_t st, const uint8_t* frames, const int32_t* start, T* x, T* y,
                          int n, int V, int H, int W, int ch, int cw, int steps, float nmax, float power) {
  if (rec == 24) loader_gather_kernel<T, 24><<<grid, LD_THREADS, 0, st>>>(frames, start, x, y, n, V, H, W, ch, cw, steps, nmax, power);
  else if (rec == 32) loader_gather_kernel<T, 32><<<grid, LD_THREADS, 0, st>>>(frames, start, x, y, n, V, H, W, ch, cw, steps, nmax, power);
  else loader_gather_kernel<T, 0><<<grid, LD_THREADS, 0, st>>>(frames, start, x, y, n, V, H, W, ch, cw, steps, nmax, power);
}

}  // namespace cgat

using namespace cgat;

extern "C" int cgat_loader_gather(const uint8_t* frames, int64_t n_frames, const int32_t* start, void* x, void* y,
                                  int32_t n, int32_t vertices, int32_t h, int32_t w, int32_t crop_h, int32_t crop_w,
                                  int32_t steps, float normalizing_max, float power, int32_t dtype, void* stream) {
  if (!frames || !start || !x || !y) return fail(CGAT_EINVAL, "null argument");
  if (n < 1 || vertices < 1 || h < 1 || w < 1 || steps < 1 || crop_h < 1 || crop_w < 1 || crop_h > h || crop_w > w ||
      n_frames < 2 * steps)
    return fail(CGAT_EINVAL, "bad loader geometry");
  if (steps * vertices > LD_MAX_REC) return fail(CGAT_EUNSUPPORTED, "pixel record of %d elements (max %d)", steps * vertices, LD_MAX_REC);
  if (dtype != CGAT_F32 && dtype != CGAT_BF16) return fail(CGAT_EINVAL, "bad dtype %d", dtype);
  if (!(normalizing_max > 0.f)) return fail(CGAT_EINVAL, "normalizing_max must be positive");
  const long long total = (long long)n * crop_h * crop_w;
  const unsigned grid = (unsigned)((total + LD_THREADS - 1) / LD_THREADS);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == CGAT_F32)
    loader_launch<float>(steps * vertices, grid, st, frames, start, (float*)x, (float*)y, n, vertices, h, w, crop_h, crop_w,
                         steps, normalizing_max, power);
  else
    loader_launch<__nv_bfloat16>(steps * vertices, grid, st, frames, start, (__nv_bfloat16*)x, (__nv_bfloat16*)y, n, vertices,
                                 h, w, crop_h, crop_w, steps, normalizing_max, power);
  return check_launch("loader_gather_kernel");
}
