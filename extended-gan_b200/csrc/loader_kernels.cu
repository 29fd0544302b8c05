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
// straight into the pixel-record layout the conv-GAT kernels read.  fp32 output is bit-exact with the reference for
// power = 1 (IEEE division); bf16 output rounds that value once.
#include "common.cuh"
#include <cstdlib>

namespace cgat {

constexpr int LD_THREADS = 128;
constexpr int LD_MAX_REC = 64;  // T * V elements per pixel record
constexpr int LDQ_PIX = 256;    // pixels per CTA of the quad-organised fast path

// One CTA = LD_THREADS consecutive output pixels of one sample.  Phase 1: the 2*steps*V source planes of the window
// are staged in shared memory with 4-byte loads where the 4 pixels lie in one image row (always, for the usual
// crop_w % 4 == 0), 12 loads per thread instead of 48 single bytes.  A 256-entry table holds pow(k / max, power) in
// the output dtype (computed once per CTA with the IEEE division the reference's torch ops perform), so phase 2 is
// two shared-memory reads per element and 16-byte record stores.
// RECT = steps * V at compile time (fully unrolled record loops), 0 = generic.  Launch bounds ask for 16 resident CTAs
// per SM (32 registers): the 2 048 CTAs of the bench shape then fit in ONE wave (at 40 registers 12 fit and a thin second
// wave doubled the kernel's latency-bound run time: 17.2 -> 15.3 us).
template <typename T, int RECT>
__global__ void __launch_bounds__(LD_THREADS, 16)
loader_gather_kernel(const uint8_t* __restrict__ frames, const int32_t* __restrict__ start, T* __restrict__ x,
                     T* __restrict__ y, int V, int H, int W, int crop_h, int crop_w, int steps, float nmax, float power,
                     int x_planar, int mul_exact) {
  extern __shared__ __align__(16) unsigned char ld_smem[];
  T* lut = reinterpret_cast<T*>(ld_smem);                       // [256]
  uint8_t* tile = ld_smem + 256 * sizeof(T);                    // [2*steps*V][LD_THREADS]
  const int rec = RECT ? RECT : steps * V, planes = 2 * rec;
  const long long per_sample = (long long)crop_h * crop_w;
  const long long blocks_per_sample = (per_sample + LD_THREADS - 1) / LD_THREADS;
  const int s = (int)(blockIdx.x / blocks_per_sample);
  const long long p0 = (long long)(blockIdx.x % blocks_per_sample) * LD_THREADS;  // first pixel of this CTA in the sample
  const int npix = (int)min((long long)LD_THREADS, per_sample - p0);
  const int f0 = start[s];
  const size_t plane = (size_t)H * W;
  for (int k = threadIdx.x; k < 256; k += LD_THREADS) {
    float v = __fdiv_rn((float)k, nmax);          // torch: data / 254          (kmni_data_loader.py:75)
    if (power != 1.0f) v = powf(v, power);        //        t.pow(norm, power)  (:76)
    lut[k] = DT<T>::from_f(v);
  }
  const uint8_t* src0 = frames + (size_t)f0 * V * plane;  // plane e of the window = frame f0 + e / V, vertex e % V
  {
    // thread t stages pixels 4q .. 4q+3 (q = t % 32) of planes e = t / 32, t / 32 + 4, ...: the pixel geometry is
    // computed once, the loop runs over planes only
    const int q = threadIdx.x % (LD_THREADS / 4), e0 = threadIdx.x / (LD_THREADS / 4);
    const int lp = 4 * q;
    if (lp < npix) {
      const int p = (int)p0 + lp;
      const int h = p / crop_w, w = p - h * crop_w;
      const size_t off = (size_t)h * W + w;
      const bool fast = w + 3 < crop_w && lp + 3 < npix && ((reinterpret_cast<uintptr_t>(src0) + off) & 3u) == 0 &&
                        (plane & 3u) == 0;
      int offj[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int pj = p + j, hj = pj / crop_w;
        offj[j] = lp + j < npix ? hj * W + (pj - hj * crop_w) : -1;
      }
      for (int e = e0; e < planes; e += LD_THREADS / (LD_THREADS / 4)) {
        const uint8_t* sp = src0 + (size_t)e * plane;
        uint32_t word;
        if (fast) {
          word = *reinterpret_cast<const uint32_t*>(sp + off);
        } else {
          word = 0;
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (offj[j] >= 0) word |= (uint32_t)sp[offj[j]] << (8 * j);
        }
        *reinterpret_cast<uint32_t*>(tile + (size_t)e * LD_THREADS + lp) = word;
      }
    }
  }
  __syncthreads();
  if ((int)threadIdx.x >= npix) return;
  const long long pix = (long long)s * per_sample + p0 + threadIdx.x;
#pragma unroll 1
  for (int half = 0; half < 2; ++half) {
    T* dst = (half ? y : x) + pix * rec;
    // padded chunk-planar x (CGAT_X_PLANAR, the fused layer kernels' input format): 16-byte chunk q of the record goes
    // to plane q of the sample, [n][rec/8][crop_h][wp][8], image column w at padded column w + 1 -- consecutive threads
    // write consecutive 16-byte slots; the padding columns are never written (the buffer is zeroed once)
    size_t qstride = 1;  // distance between a record's 16-byte chunks, in uint4
    if (half == 0 && x_planar) {
      const int p = (int)p0 + threadIdx.x, hh = p / crop_w, ww = p - hh * crop_w, wp = lf_padded_width(crop_w);
      qstride = (size_t)crop_h * wp;
      dst = x + ((size_t)s * (rec * sizeof(T) / 16) * qstride + (size_t)hh * wp + ww + 1) * (16 / sizeof(T));
    }
    const uint8_t* col = tile + (size_t)(half * rec) * LD_THREADS + threadIdx.x;
    if constexpr (RECT != 0 && sizeof(T) == 2) {
      if (mul_exact) {
        // bf16 output, power = 1, and k * (1/max) rounds to the same bf16 as k / max for all 256 byte values (checked on the
        // host at launch): convert arithmetically -- the table costs one more shared-memory access per element, with bank
        // conflicts (the kernel was bound by the shared-memory pipe: ncu mio_throttle 4.0 per issue)
        const float inv = 1.f / nmax;
        uint8_t k[RECT];
#pragma unroll
        for (int e = 0; e < RECT; ++e) k[e] = col[(size_t)e * LD_THREADS];
#pragma unroll
        for (int q = 0; q < RECT / 8; ++q) {
          uint32_t w[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const __nv_bfloat162 h = __floats2bfloat162_rn((float)k[q * 8 + 2 * j] * inv, (float)k[q * 8 + 2 * j + 1] * inv);
            w[j] = *reinterpret_cast<const uint32_t*>(&h);
          }
          reinterpret_cast<uint4*>(dst)[q * qstride] = make_uint4(w[0], w[1], w[2], w[3]);
        }
        continue;
      }
    }
    if constexpr (RECT != 0 && (RECT * sizeof(T)) % 16 == 0) {
      constexpr int PER = 16 / sizeof(T);
      uint8_t k[RECT];
#pragma unroll
      for (int e = 0; e < RECT; ++e) k[e] = col[(size_t)e * LD_THREADS];  // all byte loads first, then the table
#pragma unroll
      for (int q = 0; q < RECT / PER; ++q) {
        T v[PER];
#pragma unroll
        for (int j = 0; j < PER; ++j) v[j] = lut[k[q * PER + j]];
        reinterpret_cast<uint4*>(dst)[q * qstride] = *reinterpret_cast<const uint4*>(v);
      }
    } else if ((rec * sizeof(T)) % 16 == 0) {
      constexpr int PER = 16 / sizeof(T);
      for (int q = 0; q < rec / PER; ++q) {
        T v[PER];
#pragma unroll
        for (int j = 0; j < PER; ++j) v[j] = lut[col[(size_t)(q * PER + j) * LD_THREADS]];
        reinterpret_cast<uint4*>(dst)[q * qstride] = *reinterpret_cast<const uint4*>(v);
      }
    } else {
      for (int e = 0; e < rec; ++e) dst[e] = lut[col[(size_t)e * LD_THREADS]];
    }
  }
}

// The bench shape's fast path (bf16 output, power = 1 with the exact reciprocal, whole 16-byte chunks, image rows a multiple
// of 4 pixels): a CTA still owns LD_THREADS consecutive pixels of a sample, but phase 2 is organised by (pixel QUAD, 16-byte
// record chunk) instead of by pixel: warp c of the CTA handles chunk c (c < NCH: x, else y) and its lane a quad of 4
// consecutive pixels, so a thread reads its 8 source planes with EIGHT 4-byte shared-memory loads (4 pixels each; the
// per-pixel kernel issues 2 * rec single-byte loads per thread), converts the bytes in place (I2F with a byte selector, one
// multiply, packed bf16 conversion) and writes four 16-byte chunks.  665 -> ~150 instructions per thread; the kernel sits
// on the critical path of the end-to-end step behind the train kernel (it cannot share the SMs with it).
template <int NCH>
__global__ void __launch_bounds__(64 * NCH)
loader_gather_quads_kernel(const uint8_t* __restrict__ frames, const int32_t* __restrict__ start, __nv_bfloat16* __restrict__ x,
                           __nv_bfloat16* __restrict__ y, int V, int H, int W, int crop_h, int crop_w, float inv, int x_planar) {
  // LDQ_PIX pixels per CTA = QPT quads per lane: the bench shape then runs as 1 024 CTAs of 192 threads, ONE wave at 7 CTAs
  // per SM (2 048 CTAs of 128 pixels were 1.4 waves of a latency-bound kernel), with 16 loads in flight per thread
  constexpr int REC = 8 * NCH, PLANES = 2 * REC, NT = 64 * NCH, QUADS = LDQ_PIX / 4, QPT = QUADS / 32;
  __shared__ __align__(16) uint32_t tile[PLANES][QUADS];  // [plane][pixel quad]: 4 pixel bytes per word
  const long long per_sample = (long long)crop_h * crop_w;
  const long long blocks_per_sample = (per_sample + LDQ_PIX - 1) / LDQ_PIX;
  const int s = (int)(blockIdx.x / blocks_per_sample);
  const long long p0 = (long long)(blockIdx.x % blocks_per_sample) * LDQ_PIX;
  const int npix = (int)min((long long)LDQ_PIX, per_sample - p0);  // a multiple of 4 (host-checked)
  const size_t plane = (size_t)H * W;
  const uint8_t* src0 = frames + (size_t)start[s] * V * plane;
  const int lane = threadIdx.x & 31, c = threadIdx.x >> 5;  // warp c: chunk c of the record (c < NCH: x, else y)
  int hh[QPT], ww[QPT];
#pragma unroll
  for (int k = 0; k < QPT; ++k) {  // quad lane + 32 k: 4 consecutive pixels of one image row (crop_w % 4 == 0)
    const int p = (int)p0 + 4 * (lane + 32 * k);
    hh[k] = p / crop_w;
    ww[k] = p - hh[k] * crop_w;
  }
  // stage: warp c takes planes c, c + 2 NCH, ... (8 of them), each lane its QPT quads of the plane
#pragma unroll
  for (int i = 0; i < PLANES / (2 * NCH); ++i) {
    const int e = c + i * 2 * NCH;
#pragma unroll
    for (int k = 0; k < QPT; ++k)
      if (4 * (lane + 32 * k) < npix)
        tile[e][lane + 32 * k] = *reinterpret_cast<const uint32_t*>(src0 + (size_t)e * plane + (size_t)hh[k] * W + ww[k]);
  }
  __syncthreads();
  const bool is_y = c >= NCH;
  const int cc = is_y ? c - NCH : c;
#pragma unroll
  for (int k = 0; k < QPT; ++k) {
    const int lp = 4 * (lane + 32 * k);
    if (lp >= npix) continue;
    uint32_t wv[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) wv[j] = tile[c * 8 + j][lane + 32 * k];
    uint4* dst;  // (inv = 1 / normalizing_max from the host: the value the launch checked for bit-exactness)
    size_t pstride;  // distance between consecutive pixels' chunks, in uint4
    if (!is_y && x_planar) {
      const int wp = lf_padded_width(crop_w);
      dst = reinterpret_cast<uint4*>(x) + ((size_t)s * NCH + cc) * crop_h * wp + (size_t)hh[k] * wp + ww[k] + 1;
      pstride = 1;
    } else {
      dst = reinterpret_cast<uint4*>(is_y ? y : x) + ((size_t)s * per_sample + p0 + lp) * NCH + cc;
      pstride = NCH;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {  // pixel i of the quad = byte i of every word
      uint32_t o[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float a = (float)((wv[2 * j] >> (8 * i)) & 0xffu) * inv, b = (float)((wv[2 * j + 1] >> (8 * i)) & 0xffu) * inv;
        const __nv_bfloat162 h2 = __floats2bfloat162_rn(a, b);
        o[j] = *reinterpret_cast<const uint32_t*>(&h2);
      }
      dst[(size_t)i * pstride] = make_uint4(o[0], o[1], o[2], o[3]);
    }
  }
}

// pixel records [n][h][w][rec] -> padded chunk-planar [n][rec/8][h][wp][8] (bf16): for x tensors that did not come from
// the loader kernel.  One thread per pixel: rec/8 16-byte loads (a warp reads 32 * rec * 2 contiguous bytes), rec/8
// coalesced stores.  Padding columns are not written.
__global__ void __launch_bounds__(256) records_to_planar_kernel(const uint4* __restrict__ in, uint4* __restrict__ out,
                                                                  long long n_pix, int h, int w, int nchunk) {
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_pix) return;
  const int wp = lf_padded_width(w);
  const long long per_sample = (long long)h * w, smp = p / per_sample, q = p - smp * per_sample;
  const int hh = (int)(q / w), ww = (int)(q - (long long)hh * w);
  const long long plane = (long long)h * wp;
  uint4 v[8];
  for (int c0 = 0; c0 < nchunk; c0 += 8) {
#pragma unroll
    for (int c = 0; c < 8; ++c)
      if (c0 + c < nchunk) v[c] = in[p * nchunk + c0 + c];
#pragma unroll
    for (int c = 0; c < 8; ++c)
      if (c0 + c < nchunk) out[(smp * nchunk + c0 + c) * plane + (long long)hh * wp + ww + 1] = v[c];
  }
}

// ARAI loader (convolutional_gat/data_loaders/arai_data_loader.py:57-93,117-119): the preprocessed files hold FLOAT frames
// data[L][regions][1][H][W]; __batchify cuts windows i .. i+2*steps-1 (stride 1), splits them into steps input and steps
// target frames, fix_sizes squeezes the singleton axis and permutes to [N, H, W, T, V].  No normalisation (norm_max /
// norm_min are stored and never applied).  The reference materialises every overlapping window on the host and copies
// 2*N*steps*V*H*W floats per batch; here a file's frames cross PCIe once and a batch is one launch:
//   x[n, h, w, t, v] = frames[start[n] + t, v, h, w]        y[...] = frames[start[n] + steps + t, v, h, w]
// HBM-bound transpose: a CTA stages the rec = steps*V source planes of LD_THREADS consecutive pixels (coalesced along the
// image row) in shared memory with a one-float skew, then streams its CONTIGUOUS npix*rec output block out with
// consecutive threads on consecutive elements (fp32: bit-exact copy; bf16: one rounding).
template <typename T>
__global__ void __launch_bounds__(LD_THREADS)
loader_gather_f32_kernel(const float* __restrict__ frames, const int32_t* __restrict__ start, T* __restrict__ x,
                         T* __restrict__ y, int V, int H, int W, int crop_h, int crop_w, int steps) {
  extern __shared__ __align__(16) unsigned char ld_smem[];
  float* tile = reinterpret_cast<float*>(ld_smem);  // [rec][LD_THREADS + 1]
  const int rec = steps * V;
  const long long per_sample = (long long)crop_h * crop_w;
  const long long blocks_per_sample = (per_sample + LD_THREADS - 1) / LD_THREADS;
  const int s = (int)(blockIdx.x / blocks_per_sample);
  const long long p0 = (long long)(blockIdx.x % blocks_per_sample) * LD_THREADS;
  const int npix = (int)min((long long)LD_THREADS, per_sample - p0);
  const size_t plane = (size_t)H * W;
  const int p = (int)p0 + threadIdx.x, h = p / crop_w, w = p - h * crop_w;
  const size_t off = (size_t)h * W + w;
#pragma unroll 1
  for (int half = 0; half < 2; ++half) {
    const float* src0 = frames + ((size_t)start[s] + (size_t)half * steps) * V * plane;
    if ((int)threadIdx.x < npix)
      for (int e = 0; e < rec; ++e) tile[e * (LD_THREADS + 1) + threadIdx.x] = src0[(size_t)e * plane + off];
    __syncthreads();
    T* dst = (half ? y : x) + ((long long)s * per_sample + p0) * rec;
    for (int i = threadIdx.x; i < npix * rec; i += LD_THREADS) {
      const int px = i / rec, e = i - px * rec;
      dst[i] = DT<T>::from_f(tile[e * (LD_THREADS + 1) + px]);
    }
    __syncthreads();
  }
}

static int loader_gather_impl(const uint8_t* frames, int64_t n_frames, const int32_t* start, void* x, void* y,
                              int32_t n, int32_t vertices, int32_t h, int32_t w, int32_t crop_h, int32_t crop_w,
                              int32_t steps, float normalizing_max, float power, int32_t dtype, int x_planar, void* stream);

}  // namespace cgat

using namespace cgat;

extern "C" int cgat_loader_gather(const uint8_t* frames, int64_t n_frames, const int32_t* start, void* x, void* y,
                                  int32_t n, int32_t vertices, int32_t h, int32_t w, int32_t crop_h, int32_t crop_w,
                                  int32_t steps, float normalizing_max, float power, int32_t dtype, void* stream) {
  return loader_gather_impl(frames, n_frames, start, x, y, n, vertices, h, w, crop_h, crop_w, steps, normalizing_max, power,
                            dtype, 0, stream);
}

extern "C" int cgat_loader_gather_planar(const uint8_t* frames, int64_t n_frames, const int32_t* start, void* x_planar,
                                         void* y, int32_t n, int32_t vertices, int32_t h, int32_t w, int32_t crop_h,
                                         int32_t crop_w, int32_t steps, float normalizing_max, float power, void* stream) {
  if ((steps * vertices) % 8) return fail(CGAT_EUNSUPPORTED, "chunk-planar x needs steps*vertices to be a multiple of 8");
  return loader_gather_impl(frames, n_frames, start, x_planar, y, n, vertices, h, w, crop_h, crop_w, steps, normalizing_max,
                            power, CGAT_BF16, 1, stream);
}

extern "C" int cgat_loader_gather_f32(const float* frames, int64_t n_frames, const int32_t* start, void* x, void* y,
                                      int32_t n, int32_t vertices, int32_t h, int32_t w, int32_t crop_h, int32_t crop_w,
                                      int32_t steps, int32_t dtype, void* stream) {
  if (!frames || !start || !x || !y) return fail(CGAT_EINVAL, "null argument");
  if (n < 1 || vertices < 1 || h < 1 || w < 1 || steps < 1 || crop_h < 1 || crop_w < 1 || crop_h > h || crop_w > w ||
      n_frames < 2 * steps)
    return fail(CGAT_EINVAL, "bad loader geometry");
  if (steps * vertices > LD_MAX_REC) return fail(CGAT_EUNSUPPORTED, "pixel record of %d elements (max %d)", steps * vertices, LD_MAX_REC);
  if (dtype != CGAT_F32 && dtype != CGAT_BF16) return fail(CGAT_EINVAL, "bad dtype %d", dtype);
  const long long per_sample = (long long)crop_h * crop_w;
  const unsigned grid = (unsigned)(n * ((per_sample + LD_THREADS - 1) / LD_THREADS));
  const size_t smem = (size_t)steps * vertices * (LD_THREADS + 1) * sizeof(float);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == CGAT_F32)
    loader_gather_f32_kernel<float><<<grid, LD_THREADS, smem, st>>>(frames, start, (float*)x, (float*)y, vertices, h, w,
                                                                    crop_h, crop_w, steps);
  else
    loader_gather_f32_kernel<__nv_bfloat16><<<grid, LD_THREADS, smem, st>>>(frames, start, (__nv_bfloat16*)x,
                                                                            (__nv_bfloat16*)y, vertices, h, w, crop_h, crop_w, steps);
  return check_launch("loader_gather_f32_kernel");
}

extern "C" int cgat_records_to_planar(const void* x, void* x_planar, int64_t n, int32_t h, int32_t w, int32_t rec,
                                      void* stream) {
  if (!x || !x_planar) return fail(CGAT_EINVAL, "null argument");
  if (n < 1 || h < 1 || w < 1 || rec < 8 || rec % 8) return fail(CGAT_EINVAL, "records_to_planar: rec must be a multiple of 8");
  if (!aligned16(x) || !aligned16(x_planar)) return fail(CGAT_EALIGN, "records_to_planar: 16-byte aligned tensors");
  const long long n_pix = (long long)n * h * w;
  records_to_planar_kernel<<<(unsigned)((n_pix + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      (const uint4*)x, (uint4*)x_planar, n_pix, h, w, rec / 8);
  return check_launch("records_to_planar_kernel");
}

static int cgat::loader_gather_impl(const uint8_t* frames, int64_t n_frames, const int32_t* start, void* x, void* y,
                                    int32_t n, int32_t vertices, int32_t h, int32_t w, int32_t crop_h, int32_t crop_w,
                                    int32_t steps, float normalizing_max, float power, int32_t dtype, int x_planar,
                                    void* stream) {
  if (!frames || !start || !x || !y) return fail(CGAT_EINVAL, "null argument");
  if (n < 1 || vertices < 1 || h < 1 || w < 1 || steps < 1 || crop_h < 1 || crop_w < 1 || crop_h > h || crop_w > w ||
      n_frames < 2 * steps)
    return fail(CGAT_EINVAL, "bad loader geometry");
  if (steps * vertices > LD_MAX_REC) return fail(CGAT_EUNSUPPORTED, "pixel record of %d elements (max %d)", steps * vertices, LD_MAX_REC);
  if (dtype != CGAT_F32 && dtype != CGAT_BF16) return fail(CGAT_EINVAL, "bad dtype %d", dtype);
  if (!(normalizing_max > 0.f)) return fail(CGAT_EINVAL, "normalizing_max must be positive");
  const long long per_sample = (long long)crop_h * crop_w;
  const unsigned grid = (unsigned)(n * ((per_sample + LD_THREADS - 1) / LD_THREADS));
  cudaStream_t st = (cudaStream_t)stream;
  const size_t esz = dtype == CGAT_F32 ? 4 : 2;
  const size_t smem = 256 * esz + (size_t)2 * steps * vertices * LD_THREADS;
  if (smem > 48 * 1024) return fail(CGAT_EUNSUPPORTED, "loader tile needs %zu B of shared memory", smem);
  // may the bf16 path multiply by the reciprocal?  Only if that is bit-identical to the reference's division (then rounded
  // to bf16) for every byte value
  int mul_exact = 0;
  if (dtype == CGAT_BF16 && power == 1.0f) {
    mul_exact = 1;
    const float inv = 1.f / normalizing_max;
    for (int k = 0; k < 256 && mul_exact; ++k) {
      const __nv_bfloat16 a = __float2bfloat16_rn((float)k / normalizing_max), b = __float2bfloat16_rn((float)k * inv);
      if (__bfloat16_as_ushort(a) != __bfloat16_as_ushort(b)) mul_exact = 0;
    }
  }
  // fast path: (pixel quad, record chunk) work items -- see loader_gather_quads_kernel
  if (dtype == CGAT_BF16 && mul_exact && (steps * vertices) % 8 == 0 && steps * vertices <= 32 && crop_w % 4 == 0 && w % 4 == 0 &&
      (reinterpret_cast<uintptr_t>(frames) & 3u) == 0 && !getenv("CGAT_LOADER_NO_QUADS")) {
    const int nch = steps * vertices / 8;
    const unsigned qgrid = (unsigned)(n * ((per_sample + LDQ_PIX - 1) / LDQ_PIX));
#define LDQ_LAUNCH(NCH)                                                                                                        \
  loader_gather_quads_kernel<NCH><<<qgrid, 64 * NCH, 0, st>>>(frames, start, (__nv_bfloat16*)x, (__nv_bfloat16*)y, vertices, h, w, \
                                                             crop_h, crop_w, 1.f / normalizing_max, x_planar)
    if (nch == 1) LDQ_LAUNCH(1);
    else if (nch == 2) LDQ_LAUNCH(2);
    else if (nch == 3) LDQ_LAUNCH(3);
    else LDQ_LAUNCH(4);
#undef LDQ_LAUNCH
    return check_launch("loader_gather_quads_kernel");
  }
#define LD_LAUNCH(T, R)                                                                                                   \
  loader_gather_kernel<T, R><<<grid, LD_THREADS, smem, st>>>(frames, start, (T*)x, (T*)y, vertices, h, w, crop_h, crop_w, steps, \
                                                             normalizing_max, power, x_planar, mul_exact)
  const int rec = steps * vertices;
  if (dtype == CGAT_F32) {
    if (rec == 24) LD_LAUNCH(float, 24); else LD_LAUNCH(float, 0);
  } else {
    if (rec == 24) LD_LAUNCH(__nv_bfloat16, 24); else LD_LAUNCH(__nv_bfloat16, 0);
  }
#undef LD_LAUNCH
  return check_launch("loader_gather_kernel");
}
