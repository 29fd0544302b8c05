// K4 / K5, packed-half2 path: bf16 in HBM, two pixels per thread in __half2 lanes.
//
// Why: at V = 6 nodes the fused attention kernels are bound by FP32 instruction issue, not by HBM (ncu:
// issue-active 67 % forward, register-limited 14 % warp occupancy backward, DRAM < 20 %).  Packing two pixels
// into one half2 register halves the instructions per pixel and doubles the pixels in flight per register.
// Inputs are bf16 (8-bit mantissa) so fp16's 11-bit mantissa loses nothing on load; fp16's narrow exponent is
// handled in the backward by scaling every pixel's upstream gradient with its own power of two (exact) and
// undoing it in fp32 before the results leave the thread.  Parameter-gradient sums across pixels stay fp32.
// Same data movement as the fp32-math kernels: one TMA bulk copy per operand tile in, one bulk store out; the
// backward writes d(in) IN PLACE over the input tile in shared memory.
#include "attn_common.cuh"

namespace cgat {

constexpr int TILE2 = 256;  // pixels per CTA; thread t owns pixels t and t + 128
constexpr int H2_THREADS = 128;

template <int NODES, int CI, int CO>
struct SmemParamsH2 {
  __half2 W[MAX_HEADS][CI * CO];
  __half2 a[MAX_HEADS][2 * CO];
  __half2 adj[MAX_HEADS][NODES * NODES];
  uint64_t maskrow[NODES];
};

template <int NODES, int CI, int CO>
__device__ __forceinline__ void load_params_h2(SmemParamsH2<NODES, CI, CO>& sp, const AttnArgs& A, bool need_W) {
  const int tid = threadIdx.x, nt = blockDim.x;
  if (need_W)
    for (int i = tid; i < A.heads * CI * CO; i += nt) sp.W[i / (CI * CO)][i % (CI * CO)] = __float2half2_rn(A.W[i]);
  for (int i = tid; i < A.heads * 2 * CO; i += nt) sp.a[i / (2 * CO)][i % (2 * CO)] = __float2half2_rn(A.a[i]);
  for (int i = tid; i < A.heads * NODES * NODES; i += nt)
    sp.adj[i / (NODES * NODES)][i % (NODES * NODES)] = __float2half2_rn(A.adj[i]);
  if (tid < NODES) {
    uint64_t m = 0;
    for (int j = 0; j < NODES; ++j)
      if (A.mask == nullptr || A.mask[tid * NODES + j] != 0) m |= (1ull << j);
    sp.maskrow[tid] = m;
  }
}

// two bf16 records (pixel A -> low lanes, pixel B -> high lanes), each scaled in fp32 before the conversion
template <int N>
__device__ __forceinline__ void load_pair(const __nv_bfloat16* __restrict__ pa, const __nv_bfloat16* __restrict__ pb,
                                          float sa, float sb, __half2 (&h)[N]) {
  static_assert(N % 8 == 0, "record must be a multiple of 16 bytes");
  const uint4* qa = reinterpret_cast<const uint4*>(pa);
  const uint4* qb = reinterpret_cast<const uint4*>(pb);
#pragma unroll
  for (int i = 0; i < N / 8; ++i) {
    const uint4 va = qa[i], vb = qb[i];
    const uint32_t wa[4] = {va.x, va.y, va.z, va.w}, wb[4] = {vb.x, vb.y, vb.z, vb.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      h[8 * i + 2 * k] = __floats2half2_rn(__uint_as_float(wa[k] << 16) * sa, __uint_as_float(wb[k] << 16) * sb);
      h[8 * i + 2 * k + 1] =
          __floats2half2_rn(__uint_as_float(wa[k] & 0xffff0000u) * sa, __uint_as_float(wb[k] & 0xffff0000u) * sb);
    }
  }
}

// max |x| of one bf16 record (fp32)
template <int N>
__device__ __forceinline__ float rec_amax(const __nv_bfloat16* __restrict__ p) {
  const uint4* q = reinterpret_cast<const uint4*>(p);
  uint32_t m = 0;  // max over the magnitudes as integers (sign bit cleared): monotone for non-negative floats
#pragma unroll
  for (int i = 0; i < N / 8; ++i) {
    const uint4 v = q[i];
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      m = max(m, (w[k] << 16) & 0x7fff0000u);
      m = max(m, w[k] & 0x7fff0000u);
    }
  }
  return __uint_as_float(m);
}

// power-of-two scale that brings amax into [1,2), and its inverse (both exact); identity for tiny / zero amax
__device__ __forceinline__ void pow2_scale(float amax, float& scale, float& unscale) {
  const uint32_t e = (__float_as_uint(amax) >> 23) & 0xffu;
  if (e < 16u || e > 240u) { scale = 1.f; unscale = 1.f; return; }
  scale = __uint_as_float((254u - e) << 23);
  unscale = __uint_as_float(e << 23);
}

template <int N>
__device__ __forceinline__ void store_pair(__nv_bfloat16* __restrict__ pa, __nv_bfloat16* __restrict__ pb, bool wa,
                                           bool wb, float ua, float ub, const __half2 (&h)[N]) {
  uint4* qa = reinterpret_cast<uint4*>(pa);
  uint4* qb = reinterpret_cast<uint4*>(pb);
#pragma unroll
  for (int i = 0; i < N / 8; ++i) {
    uint32_t oa[4], ob[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 f0 = __half22float2(h[8 * i + 2 * k]), f1 = __half22float2(h[8 * i + 2 * k + 1]);
      oa[k] = pack_bf16x2(f0.x * ua, f1.x * ua);
      ob[k] = pack_bf16x2(f0.y * ub, f1.y * ub);
    }
    if (wa) qa[i] = make_uint4(oa[0], oa[1], oa[2], oa[3]);
    if (wb) qb[i] = make_uint4(ob[0], ob[1], ob[2], ob[3]);
  }
}

template <int NODES, int C, bool SPATIAL>
__device__ __forceinline__ void rec_to_mat_h2(const __half2 (&r)[NODES * C], __half2 (&m)[NODES][C]) {
#pragma unroll
  for (int n = 0; n < NODES; ++n)
#pragma unroll
    for (int c = 0; c < C; ++c) m[n][c] = r[rec_off<NODES, C, SPATIAL>(n, c)];
}
template <int NODES, int C, bool SPATIAL>
__device__ __forceinline__ void mat_to_rec_h2(const __half2 (&m)[NODES][C], __half2 (&r)[NODES * C]) {
#pragma unroll
  for (int n = 0; n < NODES; ++n)
#pragma unroll
    for (int c = 0; c < C; ++c) r[rec_off<NODES, C, SPATIAL>(n, c)] = m[n][c];
}

// ===================================================================================================
// forward
// ===================================================================================================
template <int NODES, int CI, int CO, bool SPATIAL, bool PRE>
__global__ void __launch_bounds__(H2_THREADS) attn_fwd_h2_kernel(const AttnArgs A) {
  using T = __nv_bfloat16;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  constexpr int IN_SUB = PRE ? NODES * CO : NODES * CI;
  constexpr int OUT_SUB = NODES * CO;
  const int heads = A.heads;
  const int in_rec = PRE ? heads * IN_SUB : IN_SUB;
  const bool mean = A.merge == CGAT_MERGE_MEAN;
  const int out_rec = mean ? OUT_SUB : heads * OUT_SUB;

  using SP = SmemParamsH2<NODES, CI, CO>;
  SP& sp = *reinterpret_cast<SP*>(smem_raw);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw + ((sizeof(SP) + 15) & ~15));
  T* s_in = reinterpret_cast<T*>(smem_raw + ((sizeof(SP) + 15) & ~15) + 128);
  T* s_out = s_in + (size_t)TILE2 * in_rec;

  const int tid = threadIdx.x;
  const long long pix0 = (long long)blockIdx.x * TILE2;
  const int npix = (int)min((long long)TILE2, A.n_pix - pix0);

  if (tid == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  __syncthreads();
  if (tid == 0) {
    const uint32_t bytes = (uint32_t)npix * in_rec * sizeof(T);
    mbar_arrive_expect_tx(bar, bytes);
    bulk_g2s(s_in, reinterpret_cast<const T*>(A.in) + pix0 * in_rec, bytes, bar);
  }
  load_params_h2(sp, A, !PRE);
  __syncthreads();
  mbar_wait(bar, 0);

  const bool va = tid < npix, vb = tid + H2_THREADS < npix;
  if (va) {
    const int pa = tid, pb = vb ? tid + H2_THREADS : tid;  // an absent second pixel shadows the first (not stored)
    const __half2 alpha = __float2half2_rn(A.alpha);
    __half2 X[NODES][PRE ? 1 : CI];
    if constexpr (!PRE) {
      __half2 r[NODES * CI];
      load_pair<NODES * CI>(s_in + (size_t)pa * in_rec, s_in + (size_t)pb * in_rec, 1.f, 1.f, r);
      rec_to_mat_h2<NODES, CI, SPATIAL>(r, X);
    }
    __half2 acc[NODES][CO];
#pragma unroll
    for (int v = 0; v < NODES; ++v)
#pragma unroll
      for (int u = 0; u < CO; ++u) acc[v][u] = H2::zero();
    const __half2 inv_heads = __float2half2_rn(1.f / (float)heads);

    for (int k = 0; k < heads; ++k) {
      __half2 Wh[NODES][CO];
      if constexpr (PRE) {
        __half2 r[NODES * CO];
        load_pair<NODES * CO>(s_in + (size_t)pa * in_rec + k * IN_SUB, s_in + (size_t)pb * in_rec + k * IN_SUB, 1.f,
                              1.f, r);
        rec_to_mat_h2<NODES, CO, SPATIAL>(r, Wh);
      } else {
        project_linear<H2, NODES, CI, CO>(X, sp.W[k], Wh);
      }
      __half2 z[NODES][CO];
#pragma unroll
      for (int v = 0; v < NODES; ++v)
#pragma unroll
        for (int u = 0; u < CO; ++u) z[v][u] = H2::zero();
      attn_forward_pixel<H2, NODES, CO, false>(Wh, sp.a[k], sp.adj[k], sp.maskrow, alpha, nullptr, nullptr, z);
      if (A.apply_elu) {
#pragma unroll
        for (int v = 0; v < NODES; ++v)
#pragma unroll
          for (int u = 0; u < CO; ++u) z[v][u] = elu_fwd<H2>(z[v][u]);
      }
      if (mean) {
#pragma unroll
        for (int v = 0; v < NODES; ++v)
#pragma unroll
          for (int u = 0; u < CO; ++u) acc[v][u] = __hfma2(z[v][u], inv_heads, acc[v][u]);
      } else if (SPATIAL) {
        __half2 r[NODES * CO];
        mat_to_rec_h2<NODES, CO, true>(z, r);
        store_pair<NODES * CO>(s_out + (size_t)pa * out_rec + k * OUT_SUB, s_out + (size_t)pb * out_rec + k * OUT_SUB,
                               true, vb, 1.f, 1.f, r);
      } else {
        T* oa = s_out + (size_t)pa * out_rec;
        T* ob = s_out + (size_t)pb * out_rec;
#pragma unroll
        for (int v = 0; v < NODES; ++v)
#pragma unroll
          for (int u = 0; u < CO; ++u) {
            const float2 f = __half22float2(z[v][u]);
            oa[v * (heads * CO) + k * CO + u] = __float2bfloat16_rn(f.x);
            if (vb) ob[v * (heads * CO) + k * CO + u] = __float2bfloat16_rn(f.y);
          }
      }
    }
    if (mean) {
      __half2 r[NODES * CO];
      mat_to_rec_h2<NODES, CO, SPATIAL>(acc, r);
      store_pair<NODES * CO>(s_out + (size_t)pa * out_rec, s_out + (size_t)pb * out_rec, true, vb, 1.f, 1.f, r);
    }
  }
  fence_proxy_async_smem();
  __syncthreads();
  if (tid == 0) {
    bulk_s2g(reinterpret_cast<T*>(A.out) + pix0 * out_rec, s_out, (uint32_t)npix * out_rec * sizeof(T));
    bulk_commit();
    bulk_wait_read0();
  }
}

// ===================================================================================================
// backward
// ===================================================================================================
template <int NODES, int CI, int CO, bool SPATIAL, bool PRE>
__global__ void __launch_bounds__(H2_THREADS) attn_bwd_h2_kernel(const AttnArgs A) {
  using T = __nv_bfloat16;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  constexpr int IN_SUB = PRE ? NODES * CO : NODES * CI;
  constexpr int OUT_SUB = NODES * CO;
  constexpr int R_ADJ = 0, R_A = NODES * NODES, R_W = R_A + 2 * CO, R = R_W + (PRE ? 0 : CI * CO);
  const int heads = A.heads;
  const int in_rec = PRE ? heads * IN_SUB : IN_SUB;
  const bool mean = A.merge == CGAT_MERGE_MEAN;
  const int out_rec = mean ? OUT_SUB : heads * OUT_SUB;

  using SP = SmemParamsH2<NODES, CI, CO>;
  SP& sp = *reinterpret_cast<SP*>(smem_raw);
  size_t off = (sizeof(SP) + 15) & ~size_t(15);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw + off);
  off += 128;
  float* red = reinterpret_cast<float*>(smem_raw + off);
  off += sizeof(float) * R * H2_THREADS;
  T* s_in = reinterpret_cast<T*>(smem_raw + off);  // input tile; overwritten in place with d(in)
  T* s_dout = s_in + (size_t)TILE2 * in_rec;

  const int tid = threadIdx.x;
  const long long pix0 = (long long)blockIdx.x * TILE2;
  const int npix = (int)min((long long)TILE2, A.n_pix - pix0);

  if (tid == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  __syncthreads();
  if (tid == 0) {
    const uint32_t b_in = (uint32_t)npix * in_rec * sizeof(T);
    const uint32_t b_do = (uint32_t)npix * out_rec * sizeof(T);
    mbar_arrive_expect_tx(bar, b_in + b_do);
    bulk_g2s(s_in, reinterpret_cast<const T*>(A.in) + pix0 * in_rec, b_in, bar);
    bulk_g2s(s_dout, reinterpret_cast<const T*>(A.dout) + pix0 * out_rec, b_do, bar);
  }
  load_params_h2(sp, A, !PRE);
  __syncthreads();
  mbar_wait(bar, 0);

  const bool va = tid < npix, vb = tid + H2_THREADS < npix;
  const int pa = va ? tid : 0, pb = vb ? tid + H2_THREADS : pa;
  const __half2 alpha = __float2half2_rn(A.alpha);
  const float gscale = mean ? 1.f / (float)heads : 1.f;

  // per-pixel power-of-two scale of the upstream gradient (whole dout record of the pixel)
  float sa = 1.f, ua = 1.f, sb = 1.f, ub = 1.f;
  if (va) {
    float ama = 0.f, amb = 0.f;
    for (int q = 0; q < out_rec; q += OUT_SUB) {
      ama = fmaxf(ama, rec_amax<OUT_SUB>(s_dout + (size_t)pa * out_rec + q));
      amb = fmaxf(amb, rec_amax<OUT_SUB>(s_dout + (size_t)pb * out_rec + q));
    }
    pow2_scale(ama * gscale, sa, ua);
    pow2_scale(amb * gscale, sb, ub);
  }
  // results of absent pixels are discarded: their un-scale factor is zero
  const float fa = va ? ua : 0.f, fb = vb ? ub : 0.f;

  __half2 X[NODES][PRE ? 1 : CI];
  __half2 dX[NODES][PRE ? 1 : CI];
  if constexpr (!PRE) {
    __half2 r[NODES * CI];
    load_pair<NODES * CI>(s_in + (size_t)pa * in_rec, s_in + (size_t)pb * in_rec, 1.f, 1.f, r);
    rec_to_mat_h2<NODES, CI, SPATIAL>(r, X);
#pragma unroll
    for (int n = 0; n < NODES; ++n)
#pragma unroll
      for (int c = 0; c < CI; ++c) dX[n][c] = H2::zero();
  }

  for (int k = 0; k < heads; ++k) {
    float* mycol = red + tid;  // red[r * H2_THREADS + tid]
    {
      __half2 Wh[NODES][CO];
      if constexpr (PRE) {
        __half2 r[NODES * CO];
        load_pair<NODES * CO>(s_in + (size_t)pa * in_rec + k * IN_SUB, s_in + (size_t)pb * in_rec + k * IN_SUB, 1.f,
                              1.f, r);
        rec_to_mat_h2<NODES, CO, SPATIAL>(r, Wh);
      } else {
        project_linear<H2, NODES, CI, CO>(X, sp.W[k], Wh);
      }
      // ---- recompute z, then dz = (scaled dout) * ELU'(z) ----
      __half2 z[NODES][CO];
#pragma unroll
      for (int v = 0; v < NODES; ++v)
#pragma unroll
        for (int u = 0; u < CO; ++u) z[v][u] = H2::zero();
      attn_forward_pixel<H2, NODES, CO, false>(Wh, sp.a[k], sp.adj[k], sp.maskrow, alpha, nullptr, nullptr, z);
      {
        __half2 dz[NODES][CO];
        if (mean || SPATIAL) {
          __half2 r[NODES * CO];
          const int q = mean ? 0 : k * OUT_SUB;
          load_pair<NODES * CO>(s_dout + (size_t)pa * out_rec + q, s_dout + (size_t)pb * out_rec + q, sa * gscale,
                                sb * gscale, r);
          rec_to_mat_h2<NODES, CO, SPATIAL>(r, dz);
        } else {
          const T* oa = s_dout + (size_t)pa * out_rec;
          const T* ob = s_dout + (size_t)pb * out_rec;
#pragma unroll
          for (int v = 0; v < NODES; ++v)
#pragma unroll
            for (int u = 0; u < CO; ++u)
              dz[v][u] = __floats2half2_rn(__bfloat162float(oa[v * (heads * CO) + k * CO + u]) * sa * gscale,
                                           __bfloat162float(ob[v * (heads * CO) + k * CO + u]) * sb * gscale);
        }
#pragma unroll
        for (int v = 0; v < NODES; ++v)
#pragma unroll
          for (int u = 0; u < CO; ++u)
            z[v][u] = A.apply_elu ? __hmul2(dz[v][u], elu_grad<H2>(z[v][u])) : dz[v][u];
      }
      // z now holds dz (scaled)
      __half2 dWh[NODES][CO];
#pragma unroll
      for (int v = 0; v < NODES; ++v)
#pragma unroll
        for (int u = 0; u < CO; ++u) dWh[v][u] = H2::zero();
      __half2 g_a[2 * CO];
#pragma unroll
      for (int u = 0; u < 2 * CO; ++u) g_a[u] = H2::zero();
      __half2 g_adj[NODES * NODES];
#pragma unroll
      for (int i = 0; i < NODES * NODES; ++i) g_adj[i] = H2::zero();
      attn_backward_pixel<H2, NODES, CO, false, 0>(Wh, z, sp.a[k], sp.adj[k], sp.maskrow, alpha, nullptr, nullptr,
                                                   nullptr, dWh, g_a, g_adj, nullptr);
#pragma unroll
      for (int i = 0; i < NODES * NODES; ++i) {
        const float2 f = __half22float2(g_adj[i]);
        mycol[(R_ADJ + i) * H2_THREADS] = f.x * fa + f.y * fb;
      }
#pragma unroll
      for (int u = 0; u < 2 * CO; ++u) {
        const float2 f = __half22float2(g_a[u]);
        mycol[(R_A + u) * H2_THREADS] = f.x * fa + f.y * fb;
      }
      if constexpr (PRE) {
        __half2 r[NODES * CO];
        mat_to_rec_h2<NODES, CO, SPATIAL>(dWh, r);
        store_pair<NODES * CO>(s_in + (size_t)pa * in_rec + k * IN_SUB, s_in + (size_t)pb * in_rec + k * IN_SUB, va, vb,
                               ua, ub, r);
      } else {
        __half2 g_W[CI * CO];
#pragma unroll
        for (int i = 0; i < CI * CO; ++i) g_W[i] = H2::zero();
        project_linear_bwd<H2, NODES, CI, CO>(X, dWh, sp.W[k], dX, g_W);
#pragma unroll
        for (int i = 0; i < CI * CO; ++i) {
          const float2 f = __half22float2(g_W[i]);
          mycol[(R_W + i) * H2_THREADS] = f.x * fa + f.y * fb;
        }
      }
    }
    __syncthreads();
    {
      const int warp = tid >> 5, lane = tid & 31;
      for (int r = warp; r < R; r += H2_THREADS / 32) {
        float s = 0.f;
#pragma unroll
        for (int q = 0; q < H2_THREADS / 32; ++q) s += red[r * H2_THREADS + lane + 32 * q];
        s = warp_sum(s);
        if (lane == 0) {
          if (r < R_A)
            atomicAdd(A.gadj + (size_t)k * NODES * NODES + r, s);
          else if (r < R_W)
            atomicAdd(A.ga + (size_t)k * 2 * CO + (r - R_A), s);
          else
            atomicAdd(A.gW + (size_t)k * CI * CO + (r - R_W), s);
        }
      }
    }
    __syncthreads();
  }

  if constexpr (!PRE) {
    __half2 r[NODES * CI];
    mat_to_rec_h2<NODES, CI, SPATIAL>(dX, r);
    store_pair<NODES * CI>(s_in + (size_t)pa * in_rec, s_in + (size_t)pb * in_rec, va, vb, ua, ub, r);
  }
  fence_proxy_async_smem();
  __syncthreads();
  if (tid == 0) {
    bulk_s2g(reinterpret_cast<T*>(A.out) + pix0 * in_rec, s_in, (uint32_t)npix * in_rec * sizeof(T));
    bulk_commit();
    bulk_wait_read0();
  }
}

// ===================================================================================================
template <int NODES, int CI, int CO, bool SPATIAL, bool PRE>
static int launch_h2(AttnOp op, const cgat_attn_desc* d, const AttnArgs& A, cudaStream_t st) {
  const int in_rec = PRE ? d->heads * NODES * CO : NODES * CI;
  const int out_rec = d->merge == CGAT_MERGE_MEAN ? NODES * CO : d->heads * NODES * CO;
  const size_t base = ((sizeof(SmemParamsH2<NODES, CI, CO>) + 15) & ~size_t(15)) + 128;
  const unsigned grid = (unsigned)((d->n_pix + TILE2 - 1) / TILE2);
  if (op == OP_FWD) {
    auto kern = attn_fwd_h2_kernel<NODES, CI, CO, SPATIAL, PRE>;
    const size_t smem = base + (size_t)TILE2 * (in_rec + out_rec) * 2;
    if (smem > 227 * 1024) return CGAT_EUNSUPPORTED;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return fail((int)e, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    kern<<<grid, H2_THREADS, smem, st>>>(A);
    return check_launch("attn_fwd_h2_kernel");
  }
  auto kern = attn_bwd_h2_kernel<NODES, CI, CO, SPATIAL, PRE>;
  const int R = NODES * NODES + 2 * CO + (PRE ? 0 : CI * CO);
  const size_t smem = base + sizeof(float) * R * H2_THREADS + (size_t)TILE2 * (in_rec + out_rec) * 2;
  if (smem > 227 * 1024) return CGAT_EUNSUPPORTED;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return fail((int)e, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
  kern<<<grid, H2_THREADS, smem, st>>>(A);
  return check_launch("attn_bwd_h2_kernel");
}

int attn_h2_launch(AttnOp op, const cgat_attn_desc* d, const AttnArgs& A, cudaStream_t st) {
  if (op != OP_FWD && op != OP_BWD) return CGAT_EUNSUPPORTED;
  const bool sp = d->layout == CGAT_LAYOUT_SPATIAL;
  const bool pre = d->proj == CGAT_PROJ_PRE;
  const int ci = pre ? d->co : d->ci;
  if (sp && d->nodes == 6 && ci == 4 && d->co == 4)
    return pre ? launch_h2<6, 4, 4, true, true>(op, d, A, st) : launch_h2<6, 4, 4, true, false>(op, d, A, st);
  if (!sp && d->nodes == 4 && ci == 6 && d->co == 6)
    return pre ? launch_h2<4, 6, 6, false, true>(op, d, A, st) : launch_h2<4, 6, 6, false, false>(op, d, A, st);
  return CGAT_EUNSUPPORTED;
}

}  // namespace cgat
