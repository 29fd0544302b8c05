// BatchNorm2d (train / eval) + activation + Dropout2d around the conv kernels, NHWC, as four HBM-bound launches:
// the glue of the reference's ConvBlock (dcgan/model.py:35-52: conv -> BatchNorm2d -> Dropout2d(0.01) -> activation) and
// of the SmaAt-UNet double convs (conv -> BatchNorm2d -> ReLU; convolutional_gat/unet_model.py:20 builds that net).
//
//   forward :  bn_stats     per-channel sum / sum of squares over the N*H*W pixels (fp32 per thread, fp64 across CTAs; the
//                           last CTA to arrive finishes mean, 1/sqrt(var + eps) and the running statistics:
//                           running = (1 - momentum) * running + momentum * batch, UNBIASED batch variance, as torch)
//              bn_act_fwd   y = act(((x - mean) * rstd * gamma + beta)) * mask[n][c]            one pass over x
//   backward:  bn_act_bwd_reduce   dz = dy * act'(z) * mask;  sum(dz), sum(dz * xhat) per channel  (= dbeta, dgamma)
//              bn_act_bwd_apply    dx = gamma * rstd * (dz - sum(dz)/M - xhat * sum(dz * xhat)/M)   (train)
//                                  dx = gamma * rstd * dz                                           (eval / no norm)
// The pre-activation z is recomputed from x in the backward (nothing but x, the two [C] statistics and the [N][C]
// dropout mask is kept): 2 reads + 1 write per element and direction, the HBM minimum for an op that needs batch statistics
// before it can apply them.  mean == NULL means "no normalisation" (activation / dropout only blocks).
// Dropout2d zeroes whole channels of a sample: the mask is [N][C] floats, 0 or 1/(1-p), from a counter-based generator
// (Philox4x32-10) keyed by (seed, a device-resident call counter, n*C + c): replayable from a CUDA graph.
//
// Statistic SETS (the *_sets entry points): the batch is `sets` equal groups of consecutive images, each normalised with ITS
// OWN batch statistics, the running statistics updated once per set IN SET ORDER -- exactly what the reference's
// UnetModel does when it pushes one vertex after the other through the same UNet (convolutional_gat/unet_model.py:25-26),
// but as one launch per pass over all vertices.  mean / rstd / the backward's per-set sums are [sets][C].
#include "common.cuh"
#include "ew_vec.cuh"

namespace cgat {

constexpr int NA_THREADS = 256;
constexpr int NA_U = 4;  // pixels a reduction thread loads per trip

__device__ __forceinline__ float na_act(float z, int act, float slope) {
  switch (act) {
    case 1: return fmaxf(z, 0.f);
    case 2: return z > 0.f ? z : slope * z;
    case 3: return 1.f / (1.f + __expf(-z));
    default: return z;
  }
}
__device__ __forceinline__ float na_act_grad(float z, int act, float slope) {
  switch (act) {
    case 1: return z > 0.f ? 1.f : 0.f;
    case 2: return z > 0.f ? 1.f : slope;
    case 3: {
      const float s = 1.f / (1.f + __expf(-z));
      return s * (1.f - s);
    }
    default: return 1.f;
  }
}

struct NaArgs {
  const void* x;
  const void* dy;
  void* out;              // y (forward) or dx (backward apply)
  long long n, hw;        // images, pixels per image
  int c;
  const float* mean;      // [C] or NULL (no normalisation)
  const float* rstd;      // [C]
  const float* gamma;     // [C] or NULL (1)
  const float* beta;      // [C] or NULL (0)
  const float* mask;      // [N][C] or NULL (1)
  int act;
  float slope;
  double* sums;           // [2C] fp64 accumulators (+ one counter word behind them), zeroed by the caller
  float* out_a;           // stats: mean      | bwd reduce: dbeta   (sum dz)
  float* out_b;           // stats: rstd      | bwd reduce: dgamma  (sum dz * xhat)
  const float* sum_dz;    // bwd apply
  const float* sum_dzx;   // bwd apply
  float* running_mean;    // stats, optional
  float* running_var;
  long long* num_batches;
  float momentum, eps;
  int training;           // bwd apply: 1 = batch statistics took part in the forward (mean terms), 0 = constants
  int accumulate;         // bwd reduce: add into dgamma / dbeta instead of overwriting
  int sets;               // statistic sets (>= 1): images [s * n_set, (s + 1) * n_set) share statistics; mean, rstd, sum_dz,
  long long n_set;        //   sum_dzx, out_a, out_b are [sets][C]; sums is [sets][2C + 2] doubles + one counter word
  float* tot_a;           // bwd reduce, optional: out_a / out_b summed over the sets ([C]; = dbeta, dgamma of the layer)
  float* tot_b;
};

// ---- per-channel reductions: a CTA owns a slab of pixels; thread = (pixel lane, channel group) -----------------------
// MODE 0: sum x, sum x^2.   MODE 1: sum dz, sum dz * xhat.
template <typename T, int V, int MODE>
__global__ void __launch_bounds__(NA_THREADS) na_reduce_kernel(const NaArgs A, int pix_per_cta) {
  griddep_wait();
  __shared__ float s_part[2][NA_THREADS * (V > 1 ? V : 1)];
  __shared__ int s_last;
  const int groups = A.c / V;                         // channel groups per pixel
  const int cg_per_pass = groups >= NA_THREADS ? NA_THREADS : groups;
  int lanes = 1;                                      // pixel lanes per CTA: the largest power of two that fits
  while (2 * lanes * cg_per_pass <= NA_THREADS) lanes *= 2;
  const int set = blockIdx.y;                        // statistic set of this CTA
  const long long P = A.n_set * A.hw;                // pixels of one set
  const long long p0 = (long long)blockIdx.x * pix_per_cta, p1 = min(P, p0 + pix_per_cta);
  const T* x = reinterpret_cast<const T*>(A.x) + (long long)set * P * A.c;
  const T* dy = reinterpret_cast<const T*>(A.dy) + (long long)set * P * A.c;
  const int so = set * A.c;                          // this set's row of the [sets][C] statistics
  double* sums = A.sums + (long long)set * (2 * A.c + 2);
  for (int g0 = 0; g0 < groups; g0 += cg_per_pass) {
    const int g = g0 + (int)(threadIdx.x % cg_per_pass), lane = threadIdx.x / cg_per_pass;
    float a0[V], a1[V];
#pragma unroll
    for (int i = 0; i < V; ++i) a0[i] = a1[i] = 0.f;
    if (g < groups && lane < lanes) {
      float mu[V], rs[V], ga[V], be[V];
#pragma unroll
      for (int i = 0; i < V; ++i) {
        const int ch = g * V + i;
        mu[i] = (MODE == 1 && A.mean) ? A.mean[so + ch] : 0.f;
        rs[i] = (MODE == 1 && A.mean) ? A.rstd[so + ch] : 1.f;
        ga[i] = (MODE == 1 && A.gamma) ? A.gamma[ch] : 1.f;
        be[i] = (MODE == 1 && A.beta) ? A.beta[ch] : 0.f;
      }
      // NA_U pixels per lane and trip, every load issued before the first use (a thread otherwise has ONE 16-byte
      // request in flight and the pass runs at DRAM latency: 31 us for the 1 MB tensor of a 4x4x512 map)
      for (long long pb = p0 + lane; pb < p1; pb += (long long)lanes * NA_U) {
        NaRaw<T, V> rx[NA_U], rd[NA_U];
#pragma unroll
        for (int q = 0; q < NA_U; ++q) {
          const long long p = pb + (long long)q * lanes;
          if (p < p1) {
            rx[q] = na_load_raw<T, V>(x + p * A.c + (long long)g * V);
            if constexpr (MODE == 1) rd[q] = na_load_raw<T, V>(dy + p * A.c + (long long)g * V);
          }
        }
#pragma unroll
        for (int q = 0; q < NA_U; ++q) {
          const long long p = pb + (long long)q * lanes;
          if (p >= p1) continue;
          float xv[V];
          na_unpack<T, V>(rx[q], xv);
          if constexpr (MODE == 0) {
#pragma unroll
            for (int i = 0; i < V; ++i) { a0[i] += xv[i]; a1[i] = fmaf(xv[i], xv[i], a1[i]); }
          } else {
            float dv[V];
            na_unpack<T, V>(rd[q], dv);
            const long long img = (long long)set * A.n_set + p / A.hw;
#pragma unroll
            for (int i = 0; i < V; ++i) {
              const float xh = (xv[i] - mu[i]) * rs[i];
              const float z = fmaf(xh, ga[i], be[i]);
              float dz = dv[i] * na_act_grad(z, A.act, A.slope);
              if (A.mask) dz *= A.mask[img * A.c + g * V + i];
              a0[i] += dz;
              a1[i] = fmaf(dz, xh, a1[i]);
            }
          }
        }
      }
    }
    // pixel lanes of one channel group -> one value: a tree over the lanes in shared memory; then ONE fp64 atomic per
    // channel and CTA
#pragma unroll
    for (int i = 0; i < V; ++i) {
      s_part[0][threadIdx.x * V + i] = a0[i];
      s_part[1][threadIdx.x * V + i] = a1[i];
    }
    __syncthreads();
    for (int half = lanes >> 1; half > 0; half >>= 1) {
      if (lane < half) {
        const int other = (threadIdx.x + half * cg_per_pass) * V;
#pragma unroll
        for (int i = 0; i < V; ++i) {
          s_part[0][threadIdx.x * V + i] += s_part[0][other + i];
          s_part[1][threadIdx.x * V + i] += s_part[1][other + i];
        }
      }
      __syncthreads();
    }
    if ((int)threadIdx.x < cg_per_pass && g0 + (int)threadIdx.x < groups) {
#pragma unroll
      for (int i = 0; i < V; ++i) {
        const int ch = (g0 + threadIdx.x) * V + i;
        atomicAdd(sums + ch, (double)s_part[0][threadIdx.x * V + i]);
        atomicAdd(sums + A.c + ch, (double)s_part[1][threadIdx.x * V + i]);
      }
    }
    __syncthreads();
  }
  // the last CTA of a set finishes that set's per-channel results; the last SET to finish then does what needs every set
  // in order: the running statistics (one momentum update per set, set 0 first) / the sums over the sets
  // (ONE fence per CTA, by the thread that signals, behind the loop's closing __syncthreads: fences are cumulative, and a
  // MEMBAR.GPU in all 256 threads cost more than the reduction itself -- ncu: membar stall 24 per issue, 21 us per launch)
  if (threadIdx.x == 0) {
    __threadfence();
    unsigned long long* counter = reinterpret_cast<unsigned long long*>(sums + 2 * A.c);
    s_last = atomicAdd(counter, 1ull) == (unsigned long long)gridDim.x - 1;
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  const double M = (double)P;
  for (int ch = threadIdx.x; ch < A.c; ch += NA_THREADS) {
    const double t0 = __ldcg(sums + ch), t1 = __ldcg(sums + A.c + ch);
    if constexpr (MODE == 0) {
      const double mean = t0 / M;
      double var = t1 / M - mean * mean;
      var = var > 0.0 ? var : 0.0;
      A.out_a[so + ch] = (float)mean;
      A.out_b[so + ch] = (float)(1.0 / sqrt(var + (double)A.eps));
      sums[ch] = mean;                                            // kept for the in-order running update below
      sums[A.c + ch] = M > 1.0 ? var * M / (M - 1.0) : var;       // unbiased, as torch's running_var
    } else {
      A.out_a[so + ch] = (float)t0;
      A.out_b[so + ch] = (float)t1;
    }
  }
  __syncthreads();
  if (A.sets > 1) {  // (one set: this CTA is also the last of all -- no second signalling round, ~3 us of fence + atomic latency)
    if (threadIdx.x == 0) {
      __threadfence();
      unsigned long long* gcounter = reinterpret_cast<unsigned long long*>(A.sums + (long long)A.sets * (2 * A.c + 2));
      s_last = atomicAdd(gcounter, 1ull) == (unsigned long long)A.sets - 1;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
  }
  for (int ch = threadIdx.x; ch < A.c; ch += NA_THREADS) {
    if constexpr (MODE == 0) {
      if (A.running_mean != nullptr) {
        double rm = (double)A.running_mean[ch], rv = (double)A.running_var[ch];
        for (int q = 0; q < A.sets; ++q) {
          const double* sq = A.sums + (long long)q * (2 * A.c + 2);
          // (each update rounded to fp32 like the reference's V separate BatchNorm calls)
          rm = (double)(float)((1.0 - A.momentum) * rm + A.momentum * __ldcg(sq + ch));
          rv = (double)(float)((1.0 - A.momentum) * rv + A.momentum * __ldcg(sq + A.c + ch));
        }
        A.running_mean[ch] = (float)rm;
        A.running_var[ch] = (float)rv;
      }
    } else if (A.tot_a != nullptr) {
      double ta = 0.0, tb = 0.0;
      for (int q = 0; q < A.sets; ++q) {
        const double* sq = A.sums + (long long)q * (2 * A.c + 2);
        ta += __ldcg(sq + ch);
        tb += __ldcg(sq + A.c + ch);
      }
      A.tot_a[ch] = A.accumulate ? A.tot_a[ch] + (float)ta : (float)ta;
      A.tot_b[ch] = A.accumulate ? A.tot_b[ch] + (float)tb : (float)tb;
    }
  }
  if constexpr (MODE == 0)
    if (threadIdx.x == 0 && A.num_batches != nullptr) *A.num_batches += A.sets;
}

// ---- elementwise passes ------------------------------------------------------------------------------------------------
// MODE 0: forward apply.   MODE 1: backward apply.
template <typename T, int V, int MODE>
__global__ void __launch_bounds__(NA_THREADS) na_apply_kernel(const NaArgs A) {
  griddep_wait();
  const int groups = A.c / V;
  const long long img = blockIdx.y;                       // one grid row per image: no division to find the dropout mask
  const unsigned per_img = (unsigned)(A.hw * groups);     // (host: hw * groups < 2^31)
  const T* x = reinterpret_cast<const T*>(A.x) + img * A.hw * A.c;
  const T* dy = reinterpret_cast<const T*>(A.dy) + img * A.hw * A.c;
  T* out = reinterpret_cast<T*>(A.out) + img * A.hw * A.c;
  const float invM = 1.f / (float)(A.n_set * A.hw);
  const int so = (int)(img / A.n_set) * A.c;              // this image's row of the [sets][C] statistics
  for (unsigned i = blockIdx.x * NA_THREADS + threadIdx.x; i < per_img; i += gridDim.x * NA_THREADS) {
    const unsigned p = i / (unsigned)groups;
    const int g = (int)(i - p * (unsigned)groups);
    const size_t off = (size_t)p * A.c + (size_t)g * V;
    float xv[V], ov[V];
    na_load<T, V>(x + off, xv);
    float dv[V];
    if constexpr (MODE == 1) na_load<T, V>(dy + off, dv);
#pragma unroll
    for (int k = 0; k < V; ++k) {
      const int ch = g * V + k;
      const float mu = A.mean ? A.mean[so + ch] : 0.f, rs = A.mean ? A.rstd[so + ch] : 1.f;
      const float ga = A.gamma ? A.gamma[ch] : 1.f, be = A.beta ? A.beta[ch] : 0.f;
      const float mk = A.mask ? A.mask[img * A.c + ch] : 1.f;
      const float xh = (xv[k] - mu) * rs;
      const float z = fmaf(xh, ga, be);
      if constexpr (MODE == 0) {
        ov[k] = na_act(z, A.act, A.slope) * mk;
      } else {
        const float dz = dv[k] * na_act_grad(z, A.act, A.slope) * mk;
        if (A.training && A.mean) ov[k] = ga * rs * (dz - A.sum_dz[so + ch] * invM - xh * A.sum_dzx[so + ch] * invM);
        else ov[k] = ga * rs * dz;
      }
    }
    na_store<T, V>(out + off, ov);
  }
}

// ---- Dropout2d mask ----------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t mulhilo32(uint32_t a, uint32_t b, uint32_t* hi) {
  const unsigned long long p = (unsigned long long)a * b;
  *hi = (uint32_t)(p >> 32);
  return (uint32_t)p;
}
// Philox4x32-10 (Salmon et al., SC'11): counter (c0..c3), key (k0, k1) -> 4 random words
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0, hi1;
    const uint32_t lo0 = mulhilo32(0xD2511F53u, c.x, &hi0), lo1 = mulhilo32(0xCD9E8D57u, c.z, &hi1);
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += 0x9E3779B9u;
    k.y += 0xBB67AE85u;
  }
  return c;
}

__global__ void __launch_bounds__(1024) dropout2d_mask_kernel(float* __restrict__ mask, long long n, float p,
                                                              unsigned long long seed, unsigned long long* counter) {
  griddep_wait();
  const unsigned long long call = *counter;
  __syncthreads();
  const float keep = 1.f - p, scale = 1.f / keep;
  for (long long i = threadIdx.x; i < (n + 3) / 4; i += blockDim.x) {
    const uint4 r = philox4x32_10(make_uint4((uint32_t)i, (uint32_t)(i >> 32), (uint32_t)call, (uint32_t)(call >> 32)),
                                  make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (4 * i + j < n) mask[4 * i + j] = ((float)(w[j] >> 8) * (1.f / 16777216.f)) < keep ? scale : 0.f;
  }
  __syncthreads();
  if (threadIdx.x == 0) *counter = call + 1;
}

// ---- host ----------------------------------------------------------------------------------------------------------------
static int na_check(const void* x, int dtype, long long n, long long hw, int c) {
  if (!x) return fail(CGAT_EINVAL, "null tensor");
  if (n < 1 || hw < 1 || c < 1) return fail(CGAT_EINVAL, "bad geometry n=%lld hw=%lld c=%d", n, hw, c);
  if (dtype != CGAT_F32 && dtype != CGAT_BF16) return fail(CGAT_EINVAL, "bad dtype %d", dtype);
  return 0;
}
static int na_vec(int dtype, int c, const void* a, const void* b, const void* o) { return ew_vec(dtype, c, a, b, o); }
template <int MODE>
static int na_launch_reduce(const NaArgs& A, int dtype, cudaStream_t st) {
  const long long P = A.n_set * A.hw;
  // at least one unrolled trip per pixel lane and CTA, at most two CTAs per SM (over all sets): every CTA ends with 2C same-address fp64 atomics,
  // and a thousand CTAs queueing on each address cost more than the reduction itself
  const int v = na_vec(dtype, A.c, A.x, A.dy, nullptr);
  const int cg = A.c / v >= NA_THREADS ? NA_THREADS : A.c / v;  // channel groups per pass (as in the kernel)
  int lanes = 1;
  while (2 * lanes * cg <= NA_THREADS) lanes *= 2;
  const int ppc_min = lanes * NA_U;                             // one unrolled trip per lane
  long long ctas = (P + ppc_min - 1) / ppc_min;
  const long long cap = (148 * 2 + A.sets - 1) / A.sets;
  if (ctas > cap) ctas = cap;
  const int ppc = (int)((P + ctas - 1) / ctas);
  ctas = (P + ppc - 1) / ppc;
  cudaError_t e;
#define NA_RED(T, V) e = launch_pdl(na_reduce_kernel<T, V, MODE>, dim3((unsigned)ctas, (unsigned)A.sets), dim3(NA_THREADS), 0, st, A, ppc)
  if (dtype == CGAT_F32) { if (v == 4) NA_RED(float, 4); else NA_RED(float, 1); }
  else { if (v == 8) NA_RED(__nv_bfloat16, 8); else NA_RED(__nv_bfloat16, 1); }
#undef NA_RED
  if (e != cudaSuccess) return fail((int)e, "na_reduce_kernel: %s", cudaGetErrorString(e));
  return check_launch("na_reduce_kernel");
}
template <int MODE>
static int na_launch_apply(const NaArgs& A, int dtype, cudaStream_t st) {
  const int v = na_vec(dtype, A.c, A.x, A.dy, A.out);
  const long long per_img = A.hw * (A.c / v);
  if (per_img >= (1ll << 31) || A.n > 65535) return fail(CGAT_EUNSUPPORTED, "norm/act apply: image too large or n > 65535");
  long long ctas = (per_img + NA_THREADS - 1) / NA_THREADS;   // per image; ~4 waves of 8 CTAs per SM over the whole grid
  const long long cap = (148 * 8 * 4 + A.n - 1) / A.n;
  if (ctas > cap) ctas = cap < 1 ? 1 : cap;
  cudaError_t e;
#define NA_APP(T, V) e = launch_pdl(na_apply_kernel<T, V, MODE>, dim3((unsigned)ctas, (unsigned)A.n), dim3(NA_THREADS), 0, st, A)
  if (dtype == CGAT_F32) { if (v == 4) NA_APP(float, 4); else NA_APP(float, 1); }
  else { if (v == 8) NA_APP(__nv_bfloat16, 8); else NA_APP(__nv_bfloat16, 1); }
#undef NA_APP
  if (e != cudaSuccess) return fail((int)e, "na_apply_kernel: %s", cudaGetErrorString(e));
  return check_launch("na_apply_kernel");
}

}  // namespace cgat

using namespace cgat;

extern "C" int64_t cgat_bn_workspace_bytes_sets(int32_t c, int32_t sets) {
  return c < 1 || sets < 1 ? 0 : (int64_t)sets * (2 * c + 2) * 8 + 16;
}
extern "C" int64_t cgat_bn_workspace_bytes(int32_t c) { return cgat_bn_workspace_bytes_sets(c, 1); }

static int na_sets_check(int64_t n, int32_t sets) {
  if (sets < 1 || sets > 65535 || n % sets) return fail(CGAT_EINVAL, "statistic sets: %d sets do not divide %lld images", sets, (long long)n);
  return 0;
}

extern "C" int cgat_bn_stats_sets(const void* x, int32_t dtype, int64_t n, int64_t hw, int32_t c, int32_t sets, void* workspace,
                                  float* mean, float* rstd, float* running_mean, float* running_var,
                                  int64_t* num_batches_tracked, float momentum, float eps, void* stream) {
  if (int rc = na_check(x, dtype, n, hw, c)) return rc;
  if (int rc = na_sets_check(n, sets)) return rc;
  if (!workspace || !mean || !rstd) return fail(CGAT_EINVAL, "null argument");
  if ((running_mean == nullptr) != (running_var == nullptr)) return fail(CGAT_EINVAL, "running_mean and running_var go together");
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaMemsetAsync(workspace, 0, (size_t)cgat_bn_workspace_bytes_sets(c, sets), st);
  if (e != cudaSuccess) return fail((int)e, "bn workspace memset: %s", cudaGetErrorString(e));
  NaArgs A{};
  A.x = x; A.n = n; A.hw = hw; A.c = c; A.sums = (double*)workspace; A.out_a = mean; A.out_b = rstd;
  A.running_mean = running_mean; A.running_var = running_var; A.num_batches = (long long*)num_batches_tracked;
  A.momentum = momentum; A.eps = eps; A.sets = sets; A.n_set = n / sets;
  return na_launch_reduce<0>(A, dtype, st);
}

extern "C" int cgat_bn_stats(const void* x, int32_t dtype, int64_t n, int64_t hw, int32_t c, void* workspace, float* mean,
                             float* rstd, float* running_mean, float* running_var, int64_t* num_batches_tracked,
                             float momentum, float eps, void* stream) {
  return cgat_bn_stats_sets(x, dtype, n, hw, c, 1, workspace, mean, rstd, running_mean, running_var, num_batches_tracked,
                            momentum, eps, stream);
}

extern "C" int cgat_bn_act_fwd_sets(const void* x, void* y, int32_t dtype, int64_t n, int64_t hw, int32_t c, int32_t sets,
                                    const float* mean, const float* rstd, const float* gamma, const float* beta,
                                    const float* mask, int32_t act, float slope, void* stream) {
  if (int rc = na_check(x, dtype, n, hw, c)) return rc;
  if (int rc = na_sets_check(n, sets)) return rc;
  if (!y || (mean != nullptr && rstd == nullptr)) return fail(CGAT_EINVAL, "null argument");
  NaArgs A{};
  A.x = x; A.out = y; A.n = n; A.hw = hw; A.c = c; A.mean = mean; A.rstd = rstd; A.gamma = gamma; A.beta = beta; A.mask = mask;
  A.act = act; A.slope = slope; A.sets = sets; A.n_set = n / sets;
  return na_launch_apply<0>(A, dtype, (cudaStream_t)stream);
}

extern "C" int cgat_bn_act_fwd(const void* x, void* y, int32_t dtype, int64_t n, int64_t hw, int32_t c, const float* mean,
                               const float* rstd, const float* gamma, const float* beta, const float* mask, int32_t act,
                               float slope, void* stream) {
  return cgat_bn_act_fwd_sets(x, y, dtype, n, hw, c, 1, mean, rstd, gamma, beta, mask, act, slope, stream);
}

extern "C" int cgat_bn_act_bwd_sets(const void* x, const void* dy, void* dx, int32_t dtype, int64_t n, int64_t hw, int32_t c,
                                    int32_t sets, const float* mean, const float* rstd, const float* gamma, const float* beta,
                                    const float* mask, int32_t act, float slope, int32_t training, void* workspace,
                                    float* set_sums, float* dgamma, float* dbeta, int32_t accumulate, void* stream) {
  if (int rc = na_check(x, dtype, n, hw, c)) return rc;
  if (int rc = na_sets_check(n, sets)) return rc;
  if (!dy || !dx || (mean != nullptr && rstd == nullptr)) return fail(CGAT_EINVAL, "null argument");
  cudaStream_t st = (cudaStream_t)stream;
  NaArgs A{};
  A.x = x; A.dy = dy; A.out = dx; A.n = n; A.hw = hw; A.c = c; A.mean = mean; A.rstd = rstd; A.gamma = gamma; A.beta = beta;
  A.mask = mask; A.act = act; A.slope = slope; A.training = training; A.sets = sets; A.n_set = n / sets;
  const bool need_sums = dgamma != nullptr || (training && mean != nullptr);
  if (need_sums) {
    if (!workspace || !set_sums || !dgamma || !dbeta)
      return fail(CGAT_EINVAL, "the reduction needs workspace, set_sums [2][sets][C], dgamma and dbeta");
    cudaError_t e = cudaMemsetAsync(workspace, 0, (size_t)cgat_bn_workspace_bytes_sets(c, sets), st);
    if (e != cudaSuccess) return fail((int)e, "bn workspace memset: %s", cudaGetErrorString(e));
    A.sums = (double*)workspace;
    A.out_a = set_sums; A.out_b = set_sums + (size_t)sets * c;  // per-set sum dz | sum dz * xhat (the apply pass reads them)
    A.tot_a = dbeta; A.tot_b = dgamma; A.accumulate = accumulate;
    if (int rc = na_launch_reduce<1>(A, dtype, st)) return rc;
  }
  A.sum_dz = set_sums; A.sum_dzx = set_sums ? set_sums + (size_t)sets * c : nullptr;
  return na_launch_apply<1>(A, dtype, st);
}

extern "C" int cgat_bn_act_bwd(const void* x, const void* dy, void* dx, int32_t dtype, int64_t n, int64_t hw, int32_t c,
                               const float* mean, const float* rstd, const float* gamma, const float* beta, const float* mask,
                               int32_t act, float slope, int32_t training, void* workspace, float* dgamma, float* dbeta,
                               int32_t accumulate, void* stream) {
  if (int rc = na_check(x, dtype, n, hw, c)) return rc;
  if (!dy || !dx || (mean != nullptr && rstd == nullptr)) return fail(CGAT_EINVAL, "null argument");
  cudaStream_t st = (cudaStream_t)stream;
  NaArgs A{};
  A.x = x; A.dy = dy; A.out = dx; A.n = n; A.hw = hw; A.c = c; A.mean = mean; A.rstd = rstd; A.gamma = gamma; A.beta = beta;
  A.mask = mask; A.act = act; A.slope = slope; A.training = training; A.sets = 1; A.n_set = n;
  const bool need_sums = dgamma != nullptr || (training && mean != nullptr);
  if (need_sums) {
    if (!workspace || !dgamma || !dbeta) return fail(CGAT_EINVAL, "the reduction needs workspace, dgamma and dbeta");
    cudaError_t e = cudaMemsetAsync(workspace, 0, (size_t)cgat_bn_workspace_bytes(c), st);
    if (e != cudaSuccess) return fail((int)e, "bn workspace memset: %s", cudaGetErrorString(e));
    A.sums = (double*)workspace; A.out_a = dbeta; A.out_b = dgamma; A.accumulate = 0;
    if (int rc = na_launch_reduce<1>(A, dtype, st)) return rc;
    // (the apply pass needs THIS call's sums: with accumulate the caller's buffers would mix calls, so the sums are
    // always overwritten here and `accumulate` is honoured by the caller's own add)
    (void)accumulate;
  }
  A.sum_dz = dbeta; A.sum_dzx = dgamma;
  return na_launch_apply<1>(A, dtype, st);
}

extern "C" int cgat_dropout2d_mask(float* mask, int64_t n, float p, uint64_t seed, uint64_t* counter, void* stream) {
  if (!mask || !counter || n < 1) return fail(CGAT_EINVAL, "null argument or n < 1");
  if (!(p >= 0.f && p < 1.f)) return fail(CGAT_EINVAL, "dropout probability %g outside [0, 1)", p);
  cudaError_t e = launch_pdl(dropout2d_mask_kernel, dim3(1), dim3(1024), 0, (cudaStream_t)stream, mask, (long long)n, p,
                             (unsigned long long)seed, (unsigned long long*)counter);
  if (e != cudaSuccess) return fail((int)e, "dropout2d_mask_kernel: %s", cudaGetErrorString(e));
  return check_launch("dropout2d_mask_kernel");
}
