// Per-pixel graph-attention arithmetic shared by the fused forward (K4) and backward (K5) kernels.
//
// Everything here is register-resident math on compile-time sized arrays, written once as templates over an
// arithmetic policy P:
//   F32  one pixel per thread in fp32 (exact-parity path; also compiled by the HOST compiler into
//        tests/host_harness so the CPU-only suite checks this very code against the oracle);
//   H2   two pixels per thread packed in __half2 lanes (bf16 I/O fast path: the kernels are FP32-issue bound at
//        V = 6, packing halves the instruction count per pixel; inputs are bf16 so fp16's 11-bit mantissa does
//        not lose input precision, and the backward rescales every pixel's upstream gradient by a power of two
//        to stay inside fp16's exponent range).
//
// Math (reference convolutional_gat/baseline_model.py:119-160):
//   s1[i] = sum_u Wh[i][u] a[u]          s2[j] = sum_u Wh[j][u] a[CO+u]            (:128-129,162-169)
//   e[i][j] = LeakyReLU_alpha(s1[i] + s2[j]);  masked entries -> fill              (:130)
//   att = softmax_j(e)  (neighbour)   or   exp(e - max_p) / sum_p  (pixel, :131)
//   hp[i][u] = sum_j att[i][j] Wh[j][u]                                             (:145-152)
//   z[v][u]  = sum_i hp[i][u] adj[i][v]                                             (:154-158)
//   out = ELU(z)                                                                    (:160)
#pragma once
#include <cstdint>
#include <cmath>

#if defined(__CUDACC__)
#include <cuda_fp16.h>
#define CGAT_HD __host__ __device__ __forceinline__
#define CGAT_D __device__ __forceinline__
#else
#define CGAT_HD inline
#endif

namespace cgat {

constexpr float kMaskFill = -9e15f;  // pyGAT convention

CGAT_HD float fast_exp(float x) {
#if defined(__CUDA_ARCH__)
  return exp2f(x * 1.4426950408889634f);  // ex2.approx after -use_fast_math; accurate enough for rtol 1e-4
#else
  return expf(x);
#endif
}

// ---- arithmetic policies ----------------------------------------------------------------------------------
struct F32 {
  using T = float;
  static CGAT_HD T bc(float v) { return v; }
  static CGAT_HD T zero() { return 0.f; }
  static CGAT_HD T neg_inf() { return -INFINITY; }
  static CGAT_HD T mask_fill() { return kMaskFill; }
  static CGAT_HD T fma(T a, T b, T c) { return fmaf(a, b, c); }
  static CGAT_HD T add(T a, T b) { return a + b; }
  static CGAT_HD T sub(T a, T b) { return a - b; }
  static CGAT_HD T mul(T a, T b) { return a * b; }
  static CGAT_HD T max(T a, T b) { return fmaxf(a, b); }
  static CGAT_HD T min(T a, T b) { return fminf(a, b); }
  static CGAT_HD T exp(T a) { return fast_exp(a); }
  // e^(a - m) given nml = -m * log2(e): one FMA + one ex2 on the device
  static CGAT_HD T neg_l2e(T m) { return m * -1.4426950408889634f; }
  static CGAT_HD T exp_sub(T a, T nml) {
#if defined(__CUDA_ARCH__)
    return exp2f(fmaf(a, 1.4426950408889634f, nml));
#else
    return exp2f(fmaf(a, 1.4426950408889634f, nml));
#endif
  }
  static CGAT_HD T rcp(T a) { return 1.f / a; }
  static CGAT_HD T gt0(T a) { return a > 0.f ? 1.f : 0.f; }  // 1 where a > 0 else 0
  static CGAT_HD T abs(T a) { return fabsf(a); }
  static CGAT_HD T neg_where_neg(T a, T s) { return s < 0.f ? -a : a; }  // a with its sign flipped where s < 0
};

#if defined(__CUDACC__)
struct H2 {
  using T = __half2;
  static CGAT_D T bc(float v) { return __float2half2_rn(v); }
  static CGAT_D T zero() { return __float2half2_rn(0.f); }
  static CGAT_D T neg_inf() { return __float2half2_rn(-65504.f); }
  static CGAT_D T mask_fill() { return __float2half2_rn(-60000.f); }  // exp(fill - max) underflows to 0
  static CGAT_D T fma(T a, T b, T c) { return __hfma2(a, b, c); }
  static CGAT_D T add(T a, T b) { return __hadd2(a, b); }
  static CGAT_D T sub(T a, T b) { return __hsub2(a, b); }
  static CGAT_D T mul(T a, T b) { return __hmul2(a, b); }
  static CGAT_D T max(T a, T b) { return __hmax2(a, b); }
  static CGAT_D T min(T a, T b) { return __hmin2(a, b); }
  // e^x = 2^(x*log2 e) with ONE ex2.approx.f16x2 (h2exp() wraps the same instruction in ~10 fix-up instructions
  // for special values that cannot occur here: arguments are <= 0 after the max subtraction / min(z,0))
  static CGAT_D T exp(T a) {
    const T y = __hmul2(a, __float2half2_rn(1.4426950408889634f));
    uint32_t in = *reinterpret_cast<const uint32_t*>(&y), out;
    asm("ex2.approx.f16x2 %0, %1;" : "=r"(out) : "r"(in));
    return *reinterpret_cast<T*>(&out);
  }
  static CGAT_D T neg_l2e(T m) { return __hmul2(m, __float2half2_rn(-1.4426950408889634f)); }
  static CGAT_D T exp_sub(T a, T nml) {
    const T y = __hfma2(a, __float2half2_rn(1.4426950408889634f), nml);
    uint32_t in = *reinterpret_cast<const uint32_t*>(&y), out;
    asm("ex2.approx.f16x2 %0, %1;" : "=r"(out) : "r"(in));
    return *reinterpret_cast<T*>(&out);
  }
  static CGAT_D T rcp(T a) { return h2rcp(a); }
  static CGAT_D T gt0(T a) { return __hgt2(a, __float2half2_rn(0.f)); }
  static CGAT_D T abs(T a) { return __habs2(a); }
  static CGAT_D T neg_where_neg(T a, T s) {  // one LOP3 on the integer pipe: a ^ (s & sign bits)
    const uint32_t ua = *reinterpret_cast<const uint32_t*>(&a), us = *reinterpret_cast<const uint32_t*>(&s);
    const uint32_t r = ua ^ (us & 0x80008000u);
    return *reinterpret_cast<const T*>(&r);
  }
};
#endif

// ELU and its derivative, branch-free:  ELU(z) = max(z, exp(min(z,0)) - 1)  (e^t - 1 >= t, with equality of the two
// branches at 0) ;  ELU'(z) = exp(min(z,0))
template <typename P>
CGAT_HD typename P::T elu_fwd(typename P::T z) {
  return P::max(z, P::sub(P::exp(P::min(z, P::zero())), P::bc(1.f)));
}
template <typename P>
CGAT_HD typename P::T elu_grad(typename P::T z) {
  return P::exp(P::min(z, P::zero()));
}

// ---------------------------------------------------------------------------------------------
// forward of one head on one pixel (H2: two pixels).  z is ACCUMULATED INTO (caller zeroes it), pre-ELU.
// a [2*CO], adj [NODES*NODES] row-major [i][v] (already transposed by the caller for the 1-D layer's
// convention) are policy-typed broadcasts.  maskrow[i] bit j set  <=>  edge (i,j) present.
// st_max / st_rinv: pixel-axis soft-max statistics (PIXEL mode only, F32 policy only).
// ---------------------------------------------------------------------------------------------
template <typename P, int NODES, int CO, bool PIXEL, bool MASKED = true>
CGAT_HD void attn_forward_pixel(const typename P::T (&Wh)[NODES][CO], const typename P::T* __restrict__ a,
                                const typename P::T* __restrict__ adj, const uint64_t* __restrict__ maskrow,
                                typename P::T alpha, const typename P::T* __restrict__ st_max,
                                const typename P::T* __restrict__ st_rinv, typename P::T (&z)[NODES][CO]) {
  using T = typename P::T;
  T s1[NODES], s2[NODES];
#pragma unroll
  for (int i = 0; i < NODES; ++i) {
    T p = P::zero(), q = P::zero();
#pragma unroll
    for (int u = 0; u < CO; ++u) {
      p = P::fma(Wh[i][u], a[u], p);
      q = P::fma(Wh[i][u], a[CO + u], q);
    }
    s1[i] = p;
    s2[i] = q;
  }
#pragma unroll
  for (int i = 0; i < NODES; ++i) {
    T att[NODES];
    const uint64_t mrow = MASKED ? maskrow[i] : ~0ull;
    if (PIXEL) {
#pragma unroll
      for (int j = 0; j < NODES; ++j) {
        const T pre = P::add(s1[i], s2[j]);
        T e = P::max(pre, P::mul(alpha, pre));
        if (MASKED && !((mrow >> j) & 1ull)) e = P::mask_fill();
        att[j] = P::mul(P::exp(P::sub(e, st_max[i * NODES + j])), st_rinv[i * NODES + j]);
      }
    } else {
      T m = P::neg_inf();
#pragma unroll
      for (int j = 0; j < NODES; ++j) {
        const T pre = P::add(s1[i], s2[j]);
        T e = P::max(pre, P::mul(alpha, pre));
        if (MASKED && !((mrow >> j) & 1ull)) e = P::mask_fill();
        att[j] = e;
        m = j == 0 ? e : P::max(m, e);
      }
      const T nml = P::neg_l2e(m);
      T sum = P::zero();
#pragma unroll
      for (int j = 0; j < NODES; ++j) {
        // (the FMA form loses exp(fill - fill) = 1 on fully masked rows: masked kernels subtract exactly)
        att[j] = MASKED ? P::exp(P::sub(att[j], m)) : P::exp_sub(att[j], nml);
        sum = j == 0 ? att[j] : P::add(sum, att[j]);
      }
      const T r = P::rcp(sum);
#pragma unroll
      for (int j = 0; j < NODES; ++j) att[j] = P::mul(att[j], r);
    }
    T hp[CO];
#pragma unroll
    for (int u = 0; u < CO; ++u) {
      T acc = P::zero();
#pragma unroll
      for (int j = 0; j < NODES; ++j) acc = P::fma(att[j], Wh[j][u], acc);
      hp[u] = acc;
    }
#pragma unroll
    for (int v = 0; v < NODES; ++v) {
      const T w = adj[i * NODES + v];
#pragma unroll
      for (int u = 0; u < CO; ++u) z[v][u] = P::fma(hp[u], w, z[v][u]);
    }
  }
}

// logits only (pixel-softmax statistics pre-pass, F32): e[i][j]
template <int NODES, int CO>
CGAT_HD void attn_logits_pixel(const float (&Wh)[NODES][CO], const float* __restrict__ a,
                               const uint64_t* __restrict__ maskrow, float alpha, float (&e)[NODES][NODES]) {
  float s1[NODES], s2[NODES];
#pragma unroll
  for (int i = 0; i < NODES; ++i) {
    float p = 0.f, q = 0.f;
#pragma unroll
    for (int u = 0; u < CO; ++u) {
      p = fmaf(Wh[i][u], a[u], p);
      q = fmaf(Wh[i][u], a[CO + u], q);
    }
    s1[i] = p;
    s2[i] = q;
  }
#pragma unroll
  for (int i = 0; i < NODES; ++i)
#pragma unroll
    for (int j = 0; j < NODES; ++j) {
      float pre = s1[i] + s2[j];
      float v = pre > 0.f ? pre : alpha * pre;
      if (!((maskrow[i] >> j) & 1ull)) v = kMaskFill;
      e[i][j] = v;
    }
}

// ---------------------------------------------------------------------------------------------
// backward of one head on one pixel (forward recomputed from Wh).
//   dz[v][u]   gradient w.r.t. the pre-ELU z, i.e. the caller already multiplied by ELU'(z).
//   dWh        ACCUMULATED INTO.
//   g_a[2CO], g_adj[NODES*NODES]  ACCUMULATED INTO (per-thread partial sums of the parameter grads).
// pixel mode: st_max/st_rinv as in the forward, st_dot[i][j] = sum_p att[i][j][p] dAtt[i][j][p].
// MODE 0: full backward.  MODE 1: only accumulate dot[i][j] += att*dAtt (pixel-mode pre-pass).
// ---------------------------------------------------------------------------------------------
template <typename P, int NODES, int CO, bool PIXEL, int MODE, bool MASKED = true>
CGAT_HD void attn_backward_pixel(const typename P::T (&Wh)[NODES][CO], const typename P::T (&dz)[NODES][CO],
                                 const typename P::T* __restrict__ a, const typename P::T* __restrict__ adj,
                                 const uint64_t* __restrict__ maskrow, typename P::T alpha,
                                 const typename P::T* __restrict__ st_max, const typename P::T* __restrict__ st_rinv,
                                 const typename P::T* __restrict__ st_dot, typename P::T (&dWh)[NODES][CO],
                                 typename P::T* __restrict__ g_a, typename P::T* __restrict__ g_adj,
                                 typename P::T* __restrict__ dot_out) {
  using T = typename P::T;
  T s1[NODES], s2[NODES];
#pragma unroll
  for (int i = 0; i < NODES; ++i) {
    T p = P::zero(), q = P::zero();
#pragma unroll
    for (int u = 0; u < CO; ++u) {
      p = P::fma(Wh[i][u], a[u], p);
      q = P::fma(Wh[i][u], a[CO + u], q);
    }
    s1[i] = p;
    s2[i] = q;
  }
  T ds2[NODES];
#pragma unroll
  for (int j = 0; j < NODES; ++j) ds2[j] = P::zero();
  const T one_minus_alpha = P::sub(P::bc(1.f), alpha);

#pragma unroll
  for (int i = 0; i < NODES; ++i) {
    // ---- recompute row i of the attention ----
    T att[NODES];
    T slope[NODES];
    const uint64_t mrow = MASKED ? maskrow[i] : ~0ull;
    if (PIXEL) {
#pragma unroll
      for (int j = 0; j < NODES; ++j) {
        const T pre = P::add(s1[i], s2[j]);
        slope[j] = P::fma(one_minus_alpha, P::gt0(pre), alpha);
        T e = P::mul(pre, slope[j]);
        if (MASKED && !((mrow >> j) & 1ull)) { e = P::mask_fill(); slope[j] = P::zero(); }
        att[j] = P::mul(P::exp(P::sub(e, st_max[i * NODES + j])), st_rinv[i * NODES + j]);
      }
    } else {
      T m = P::neg_inf();
#pragma unroll
      for (int j = 0; j < NODES; ++j) {
        const T pre = P::add(s1[i], s2[j]);
        slope[j] = P::fma(one_minus_alpha, P::gt0(pre), alpha);
        T e = P::mul(pre, slope[j]);
        if (MASKED && !((mrow >> j) & 1ull)) { e = P::mask_fill(); slope[j] = P::zero(); }
        att[j] = e;
        m = j == 0 ? e : P::max(m, e);
      }
      const T nml = P::neg_l2e(m);
      T sum = P::zero();
#pragma unroll
      for (int j = 0; j < NODES; ++j) {
        // (the FMA form loses exp(fill - fill) = 1 on fully masked rows: masked kernels subtract exactly)
        att[j] = MASKED ? P::exp(P::sub(att[j], m)) : P::exp_sub(att[j], nml);
        sum = j == 0 ? att[j] : P::add(sum, att[j]);
      }
      const T r = P::rcp(sum);
#pragma unroll
      for (int j = 0; j < NODES; ++j) att[j] = P::mul(att[j], r);
    }
    // ---- dhp[u] = sum_v dz[v][u] adj[i][v];   g_adj[i][v] += sum_u hp[u] dz[v][u] ----
    T dhp[CO];
#pragma unroll
    for (int u = 0; u < CO; ++u) dhp[u] = P::zero();
    if (MODE == 0) {
      T hp[CO];
#pragma unroll
      for (int u = 0; u < CO; ++u) {
        T acc = P::zero();
#pragma unroll
        for (int j = 0; j < NODES; ++j) acc = P::fma(att[j], Wh[j][u], acc);
        hp[u] = acc;
      }
#pragma unroll
      for (int v = 0; v < NODES; ++v) {
        T g = g_adj[i * NODES + v];
#pragma unroll
        for (int u = 0; u < CO; ++u) g = P::fma(hp[u], dz[v][u], g);
        g_adj[i * NODES + v] = g;
      }
    }
#pragma unroll
    for (int v = 0; v < NODES; ++v) {
      const T w = adj[i * NODES + v];
#pragma unroll
      for (int u = 0; u < CO; ++u) dhp[u] = P::fma(dz[v][u], w, dhp[u]);
    }
    T de[NODES];
    if (PIXEL) {
      // ---- dAtt[j] = sum_u dhp[u] Wh[j][u];  pixel-axis soft-max backward uses the per-sample dot ----
      T datt[NODES];
#pragma unroll
      for (int j = 0; j < NODES; ++j) {
        T acc = P::zero();
#pragma unroll
        for (int u = 0; u < CO; ++u) acc = P::fma(dhp[u], Wh[j][u], acc);
        datt[j] = acc;
      }
      if (MODE == 1) {
#pragma unroll
        for (int j = 0; j < NODES; ++j) dot_out[i * NODES + j] = P::fma(att[j], datt[j], dot_out[i * NODES + j]);
        continue;
      }
#pragma unroll
      for (int j = 0; j < NODES; ++j) de[j] = P::mul(att[j], P::sub(datt[j], st_dot[i * NODES + j]));
    } else {
      // ---- neighbour soft-max backward in CENTRED form:
      //      de[j] = att[j] (dAtt[j] - sum_j' att[j'] dAtt[j']) = att[j] sum_u dhp[u] (Wh[j][u] - hp[u])
      //      (hp = sum_j att[j] Wh[j] is the aggregated feature).  Subtracting before the multiply-accumulate
      //      avoids the cancellation of two large dot products, which matters in the half2 path. ----
      T hpc[CO];
#pragma unroll
      for (int u = 0; u < CO; ++u) {
        T acc = P::zero();
#pragma unroll
        for (int j = 0; j < NODES; ++j) acc = P::fma(att[j], Wh[j][u], acc);
        hpc[u] = acc;
      }
#pragma unroll
      for (int j = 0; j < NODES; ++j) {
        T acc = P::zero();
#pragma unroll
        for (int u = 0; u < CO; ++u) acc = P::fma(dhp[u], P::sub(Wh[j][u], hpc[u]), acc);
        de[j] = P::mul(att[j], acc);
      }
    }
    // ---- dWh[j][u] += att[j] dhp[u] ----
#pragma unroll
    for (int j = 0; j < NODES; ++j)
#pragma unroll
      for (int u = 0; u < CO; ++u) dWh[j][u] = P::fma(att[j], dhp[u], dWh[j][u]);
    // ---- LeakyReLU backward, ds1 / ds2 ----
    T ds1 = P::zero();
#pragma unroll
    for (int j = 0; j < NODES; ++j) {
      const T dp = P::mul(de[j], slope[j]);
      ds1 = P::add(ds1, dp);
      ds2[j] = P::add(ds2[j], dp);
    }
#pragma unroll
    for (int u = 0; u < CO; ++u) {
      dWh[i][u] = P::fma(ds1, a[u], dWh[i][u]);
      g_a[u] = P::fma(ds1, Wh[i][u], g_a[u]);
    }
  }
  if (MODE == 0) {
#pragma unroll
    for (int j = 0; j < NODES; ++j)
#pragma unroll
      for (int u = 0; u < CO; ++u) {
        dWh[j][u] = P::fma(ds2[j], a[CO + u], dWh[j][u]);
        g_a[CO + u] = P::fma(ds2[j], Wh[j][u], g_a[CO + u]);
      }
  }
}


// =============================================================================================
// Neighbour-soft-max attention restructured through  M = adj^T . att  (used by the fused layer kernels):
//   z[v][u] = sum_i adj[i][v] sum_j att[i][j] Wh[j][u] = sum_j M[v][j] Wh[j][u],   M[v][j] = sum_i adj[i][v] att[i][j]
// costs NODES^3 + NODES^2 CO multiply-adds instead of 2 NODES^2 CO, and its backward needs no second pass over
// the aggregated features:
//   dM[v][j]   = sum_u dz[v][u] Wh[j][u]            dWh[j][u] = sum_v M[v][j] dz[v][u]   (+ the score terms below)
//   g_adj[i][v] += sum_j att[i][j] dM[v][j]          dAtt[i][j] = sum_v adj[i][v] dM[v][j]
//   de[i][j]   = att[i][j] (dAtt[i][j] - sum_j' att[i][j'] dAtt[i][j'])                  (soft-max backward)
//   dp[i][j]   = de[i][j] LeakyReLU'(s1[i] + s2[j])   (0 on masked edges)
//   ds1[i] = sum_j dp[i][j]   ds2[j] = sum_i dp[i][j]
//   dWh[i][u] += ds1[i] a[u] + ds2[i] a[CO+u]        g_a[u] += sum_i ds1[i] Wh[i][u]   g_a[CO+u] += sum_j ds2[j] Wh[j][u]
// The forward leaves s1, s2, att and M in `st`, so the backward kernel recomputes nothing twice.
// Same semantics as attn_forward_pixel / attn_backward_pixel with PIXEL = false (reference baseline_model.py:127-160).
// =============================================================================================
template <typename P, int NODES>
struct NbState {
  typename P::T s1[NODES], s2[NODES];
  typename P::T att[NODES][NODES];
  typename P::T M[NODES][NODES];  // [v][j]
};

// EXT_S: the scores s1 = Wh.a[:CO], s2 = Wh.a[CO:] are linear in the layer input, so the fused train kernel lets the
// fprop MMA produce them as extra output columns (conv weights W.a): the caller fills st.s1 / st.s2 and `a` is unused.
// SIGN_ATT (with EXT_S): st.att leaves the function carrying the SIGN of its logit's pre-activation s1[i] + s2[j]
// (|att| is the attention; LeakyReLU'(pre) = slope(att > 0)), so the backward needs neither s1 nor s2 nor the add.
template <typename P, int NODES, int CO, bool MASKED, bool EXT_S = false, bool SIGN_ATT = false>
CGAT_HD void attn_nb_forward(const typename P::T (&Wh)[NODES][CO], const typename P::T* __restrict__ a,
                             const typename P::T* __restrict__ adj, const uint64_t* __restrict__ maskrow,
                             typename P::T alpha, NbState<P, NODES>& st, typename P::T (&z)[NODES][CO]) {
  using T = typename P::T;
  if (!EXT_S) {
#pragma unroll
    for (int i = 0; i < NODES; ++i) {
      T p = P::mul(Wh[i][0], a[0]), q = P::mul(Wh[i][0], a[CO]);
#pragma unroll
      for (int u = 1; u < CO; ++u) {
        p = P::fma(Wh[i][u], a[u], p);
        q = P::fma(Wh[i][u], a[CO + u], q);
      }
      st.s1[i] = p;
      st.s2[i] = q;
    }
  }
#pragma unroll
  for (int i = 0; i < NODES; ++i) {
    const uint64_t mrow = MASKED ? maskrow[i] : ~0ull;
    T m = P::zero();
#pragma unroll
    for (int j = 0; j < NODES; ++j) {
      const T pre = P::add(st.s1[i], st.s2[j]);
      T e = P::max(pre, P::mul(alpha, pre));
      if (MASKED && !((mrow >> j) & 1ull)) e = P::mask_fill();
      st.att[i][j] = e;
      m = j == 0 ? e : P::max(m, e);
    }
    const T nml = P::neg_l2e(m);
    T sum = P::zero();
#pragma unroll
    for (int j = 0; j < NODES; ++j) {
      // (the FMA form loses exp(fill - fill) = 1 on fully masked rows: masked kernels subtract exactly)
      st.att[i][j] = MASKED ? P::exp(P::sub(st.att[i][j], m)) : P::exp_sub(st.att[i][j], nml);
      sum = j == 0 ? st.att[i][j] : P::add(sum, st.att[i][j]);
    }
    const T r = P::rcp(sum);
#pragma unroll
    for (int j = 0; j < NODES; ++j) st.att[i][j] = P::mul(st.att[i][j], r);
#pragma unroll
    for (int v = 0; v < NODES; ++v) {
      const T w = adj[i * NODES + v];
#pragma unroll
      for (int j = 0; j < NODES; ++j) st.M[v][j] = i == 0 ? P::mul(w, st.att[i][j]) : P::fma(w, st.att[i][j], st.M[v][j]);
    }
    if (SIGN_ATT) {
#pragma unroll
      for (int j = 0; j < NODES; ++j) st.att[i][j] = P::neg_where_neg(st.att[i][j], P::add(st.s1[i], st.s2[j]));
    }
  }
#pragma unroll
  for (int v = 0; v < NODES; ++v)
#pragma unroll
    for (int u = 0; u < CO; ++u) {
      T acc = P::mul(st.M[v][0], Wh[0][u]);
#pragma unroll
      for (int j = 1; j < NODES; ++j) acc = P::fma(st.M[v][j], Wh[j][u], acc);
      z[v][u] = acc;
    }
}

// dz: gradient w.r.t. the pre-ELU z.  dWh is OVERWRITTEN; g_a [2CO] and g_adj [NODES*NODES] are accumulated into.
// EXT_S: the score gradients are RETURNED in ds_out[0..NODES) = ds1, ds_out[NODES..2 NODES) = ds2 instead of being
// folded into dWh and g_a (the wgrad MMA takes them as extra rows: d(W.a) = ds^T . im2col(x); the parameter-gradient
// kernel then forms dW += a (x) d(W.a) and g_a = <W, d(W.a)>); `a` and `g_a` are unused.
template <typename P, int NODES, int CO, bool MASKED, bool EXT_S = false>
CGAT_HD void attn_nb_backward(const typename P::T (&Wh)[NODES][CO], const typename P::T (&dz)[NODES][CO],
                              const typename P::T* __restrict__ a, const typename P::T* __restrict__ adj,
                              const uint64_t* __restrict__ maskrow, typename P::T alpha, const NbState<P, NODES>& st,
                              typename P::T (&dWh)[NODES][CO], typename P::T* __restrict__ g_a,
                              typename P::T* __restrict__ g_adj, typename P::T* __restrict__ ds_out = nullptr) {
  using T = typename P::T;
  T datt[NODES][NODES];
#pragma unroll
  for (int v = 0; v < NODES; ++v) {
    T dMv[NODES];
#pragma unroll
    for (int j = 0; j < NODES; ++j) {
      T acc = P::mul(dz[v][0], Wh[j][0]);
#pragma unroll
      for (int u = 1; u < CO; ++u) acc = P::fma(dz[v][u], Wh[j][u], acc);
      dMv[j] = acc;
    }
#pragma unroll
    for (int j = 0; j < NODES; ++j)
#pragma unroll
      for (int u = 0; u < CO; ++u)
        dWh[j][u] = v == 0 ? P::mul(st.M[v][j], dz[v][u]) : P::fma(st.M[v][j], dz[v][u], dWh[j][u]);
#pragma unroll
    for (int i = 0; i < NODES; ++i) {
      T g = g_adj[i * NODES + v];
      const T w = adj[i * NODES + v];
#pragma unroll
      for (int j = 0; j < NODES; ++j) {
        g = P::fma(st.att[i][j], dMv[j], g);
        datt[i][j] = v == 0 ? P::mul(w, dMv[j]) : P::fma(w, dMv[j], datt[i][j]);
      }
      g_adj[i * NODES + v] = g;
    }
  }
  const T one_minus_alpha = P::sub(P::bc(1.f), alpha);
  T ds2[NODES];
#pragma unroll
  for (int i = 0; i < NODES; ++i) {
    const uint64_t mrow = MASKED ? maskrow[i] : ~0ull;
    T dot = P::mul(st.att[i][0], datt[i][0]);
#pragma unroll
    for (int j = 1; j < NODES; ++j) dot = P::fma(st.att[i][j], datt[i][j], dot);
    T ds1 = P::zero();
#pragma unroll
    for (int j = 0; j < NODES; ++j) {
      const T de = P::mul(st.att[i][j], P::sub(datt[i][j], dot));
      const T pre = P::add(st.s1[i], st.s2[j]);
      T slope = P::fma(one_minus_alpha, P::gt0(pre), alpha);
      if (MASKED && !((mrow >> j) & 1ull)) slope = P::zero();
      const T dp = P::mul(de, slope);
      ds1 = j == 0 ? dp : P::add(ds1, dp);
      ds2[j] = i == 0 ? dp : P::add(ds2[j], dp);
    }
    if (EXT_S) {
      ds_out[i] = ds1;
    } else {
#pragma unroll
      for (int u = 0; u < CO; ++u) {
        dWh[i][u] = P::fma(ds1, a[u], dWh[i][u]);
        g_a[u] = P::fma(ds1, Wh[i][u], g_a[u]);
      }
    }
  }
  if (EXT_S) {
#pragma unroll
    for (int j = 0; j < NODES; ++j) ds_out[NODES + j] = ds2[j];
  } else {
#pragma unroll
    for (int j = 0; j < NODES; ++j)
#pragma unroll
      for (int u = 0; u < CO; ++u) {
        dWh[j][u] = P::fma(ds2[j], a[CO + u], dWh[j][u]);
        g_a[CO + u] = P::fma(ds2[j], Wh[j][u], g_a[CO + u]);
      }
  }
}

// The EXT_S backward with a small register footprint (the fused train kernel runs two pixels per thread in 160
// registers; its CTA's shared memory leaves ~8 KB of L1, so every spilled word is an L2 round trip):
//   * dz is overwritten IN PLACE by d(Wh) -- column u of d(Wh) needs only column u of dz once dM is known;
//   * the attention-gradient rows datt[i][.] exist one row at a time.
//   * st.att carries the sign of the pre-activation (attn_nb_forward<..., SIGN_ATT>): st.s1 / st.s2 are not read.
// dzw: in dz (gradient w.r.t. the pre-ELU z), out d(Wh) WITHOUT the score terms; ds_out[0..NODES) = ds1,
// ds_out[NODES..2 NODES) = ds2; g_adj [NODES*NODES] accumulated into.  Masked edges have att = 0, hence de = 0.
template <typename P, int NODES, int CO, bool MASKED>
CGAT_HD void attn_nb_backward_inplace(const typename P::T (&Wh)[NODES][CO], typename P::T (&dzw)[NODES][CO],
                                      const typename P::T* __restrict__ adj, const uint64_t* __restrict__ maskrow,
                                      typename P::T alpha, const NbState<P, NODES>& st,
                                      typename P::T* __restrict__ g_adj, typename P::T* __restrict__ ds_out) {
  using T = typename P::T;
  T dM[NODES][NODES];  // [v][j] = sum_u dz[v][u] Wh[j][u]
#pragma unroll
  for (int v = 0; v < NODES; ++v)
#pragma unroll
    for (int j = 0; j < NODES; ++j) {
      T acc = P::mul(dzw[v][0], Wh[j][0]);
#pragma unroll
      for (int u = 1; u < CO; ++u) acc = P::fma(dzw[v][u], Wh[j][u], acc);
      dM[v][j] = acc;
    }
#pragma unroll
  for (int u = 0; u < CO; ++u) {  // d(Wh)[j][u] = sum_v M[v][j] dz[v][u], one column at a time, in place
    T col[NODES];
#pragma unroll
    for (int v = 0; v < NODES; ++v) col[v] = dzw[v][u];
#pragma unroll
    for (int j = 0; j < NODES; ++j) {
      T acc = P::mul(st.M[0][j], col[0]);
#pragma unroll
      for (int v = 1; v < NODES; ++v) acc = P::fma(st.M[v][j], col[v], acc);
      dzw[j][u] = acc;
    }
  }
  const T one_minus_alpha = P::sub(P::bc(1.f), alpha);
#pragma unroll
  for (int i = 0; i < NODES; ++i) {
    const uint64_t mrow = MASKED ? maskrow[i] : ~0ull;
    (void)mrow;
    T datt[NODES], att[NODES];
#pragma unroll
    for (int j = 0; j < NODES; ++j) {
      T acc = P::mul(adj[i * NODES], dM[0][j]);
#pragma unroll
      for (int v = 1; v < NODES; ++v) acc = P::fma(adj[i * NODES + v], dM[v][j], acc);
      datt[j] = acc;
      att[j] = P::abs(st.att[i][j]);
    }
#pragma unroll
    for (int v = 0; v < NODES; ++v) {
      T gsum = g_adj[i * NODES + v];
#pragma unroll
      for (int j = 0; j < NODES; ++j) gsum = P::fma(att[j], dM[v][j], gsum);
      g_adj[i * NODES + v] = gsum;
    }
    T dot = P::mul(att[0], datt[0]);
#pragma unroll
    for (int j = 1; j < NODES; ++j) dot = P::fma(att[j], datt[j], dot);
    T ds1 = P::zero();
#pragma unroll
    for (int j = 0; j < NODES; ++j) {
      const T de = P::mul(att[j], P::sub(datt[j], dot));
      const T slope = P::fma(one_minus_alpha, P::gt0(st.att[i][j]), alpha);
      const T dp = P::mul(de, slope);
      ds1 = j == 0 ? dp : P::add(ds1, dp);
      ds_out[NODES + j] = i == 0 ? dp : P::add(ds_out[NODES + j], dp);
    }
    ds_out[i] = ds1;
  }
}

// linear projection helpers (reference baseline_model.py:127  Wh = h @ W, W is [CI][CO] row-major)
template <typename P, int NODES, int CI, int CO>
CGAT_HD void project_linear(const typename P::T (&X)[NODES][CI], const typename P::T* __restrict__ W,
                            typename P::T (&Wh)[NODES][CO]) {
#pragma unroll
  for (int j = 0; j < NODES; ++j)
#pragma unroll
    for (int u = 0; u < CO; ++u) {
      typename P::T acc = P::zero();
#pragma unroll
      for (int t = 0; t < CI; ++t) acc = P::fma(X[j][t], W[t * CO + u], acc);
      Wh[j][u] = acc;
    }
}

// dX[j][t] += sum_u dWh[j][u] W[t][u];   g_W[t][u] += sum_j X[j][t] dWh[j][u]
template <typename P, int NODES, int CI, int CO>
CGAT_HD void project_linear_bwd(const typename P::T (&X)[NODES][CI], const typename P::T (&dWh)[NODES][CO],
                                const typename P::T* __restrict__ W, typename P::T (&dX)[NODES][CI],
                                typename P::T* __restrict__ g_W) {
#pragma unroll
  for (int j = 0; j < NODES; ++j)
#pragma unroll
    for (int t = 0; t < CI; ++t) {
      typename P::T acc = dX[j][t];
#pragma unroll
      for (int u = 0; u < CO; ++u) acc = P::fma(dWh[j][u], W[t * CO + u], acc);
      dX[j][t] = acc;
    }
#pragma unroll
  for (int t = 0; t < CI; ++t)
#pragma unroll
    for (int u = 0; u < CO; ++u) {
      typename P::T acc = g_W[t * CO + u];
#pragma unroll
      for (int j = 0; j < NODES; ++j) acc = P::fma(X[j][t], dWh[j][u], acc);
      g_W[t * CO + u] = acc;
    }
}

}  // namespace cgat
