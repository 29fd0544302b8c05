// Per-pixel graph-attention arithmetic shared by the fused forward (K4) and backward (K5) kernels.
//
// Everything here is register-resident math on compile-time sized arrays, written once as
// __host__ __device__ templates so that the very same code can be compiled by the host compiler
// into the test harness (tests/host_harness) and checked against the oracle on a CPU-only box
// before any GPU time is spent.  The host instantiation is test infrastructure; the product
// kernels live in attn_kernels.cu.
//
// Math (reference convolutional_gat/baseline_model.py:119-160):
//   s1[i] = sum_u Wh[i][u] a[u]          s2[j] = sum_u Wh[j][u] a[CO+u]            (:128-129,162-169)
//   e[i][j] = LeakyReLU_alpha(s1[i] + s2[j]);  masked entries -> -9e15             (:130)
//   att = softmax_j(e)  (neighbour)   or   exp(e - max_p) / sum_p  (pixel, :131)
//   hp[i][u] = sum_j att[i][j] Wh[j][u]                                             (:145-152)
//   z[v][u]  = sum_i hp[i][u] adj[i][v]                                             (:154-158)
//   out = ELU(z)                                                                    (:160)
#pragma once
#include <cstdint>
#include <cmath>

#if defined(__CUDACC__)
#define CGAT_HD __host__ __device__ __forceinline__
#else
#define CGAT_HD inline
#endif

namespace cgat {

constexpr float kMaskFill = -9e15f;

CGAT_HD float fast_exp(float x) {
#if defined(__CUDA_ARCH__)
  return exp2f(x * 1.4426950408889634f);  // ex2.approx after -use_fast_math; accurate enough for rtol 1e-4
#else
  return expf(x);
#endif
}

CGAT_HD float elu_fwd(float z) { return z > 0.f ? z : (fast_exp(z) - 1.f); }
// derivative of ELU expressed through z
CGAT_HD float elu_grad(float z) { return z > 0.f ? 1.f : fast_exp(z); }

// Per-(sample, head) statistics of the pixel-axis soft-max: [i][j] -> (max, 1/sum)
template <int NODES>
struct PixelStats {
  const float* mx;    // [NODES*NODES]
  const float* rinv;  // [NODES*NODES]
};

// ---------------------------------------------------------------------------------------------
// forward of one head on one pixel.  z is ACCUMULATED INTO (caller zeroes it), pre-ELU.
// adj is row-major [i][v] (already transposed by the caller for the 1-D layer's convention).
// maskrow[i] bit j set  <=>  edge (i,j) present.
// att_out (optional, may be nullptr) receives att[i][j] for the backward.
// ---------------------------------------------------------------------------------------------
template <int NODES, int CO, bool PIXEL>
CGAT_HD void attn_forward_pixel(const float (&Wh)[NODES][CO], const float* __restrict__ a,
                                const float* __restrict__ adj, const uint64_t* __restrict__ maskrow,
                                float alpha, const float* __restrict__ st_max,
                                const float* __restrict__ st_rinv, float (&z)[NODES][CO]) {
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
  for (int i = 0; i < NODES; ++i) {
    float att[NODES];
    const uint64_t mrow = maskrow[i];
    if (PIXEL) {
#pragma unroll
      for (int j = 0; j < NODES; ++j) {
        float pre = s1[i] + s2[j];
        float e = pre > 0.f ? pre : alpha * pre;
        if (!((mrow >> j) & 1ull)) e = kMaskFill;
        att[j] = fast_exp(e - st_max[i * NODES + j]) * st_rinv[i * NODES + j];
      }
    } else {
      float m = -INFINITY;
#pragma unroll
      for (int j = 0; j < NODES; ++j) {
        float pre = s1[i] + s2[j];
        float e = pre > 0.f ? pre : alpha * pre;
        if (!((mrow >> j) & 1ull)) e = kMaskFill;
        att[j] = e;
        m = fmaxf(m, e);
      }
      float sum = 0.f;
#pragma unroll
      for (int j = 0; j < NODES; ++j) {
        att[j] = fast_exp(att[j] - m);
        sum += att[j];
      }
      const float r = 1.f / sum;
#pragma unroll
      for (int j = 0; j < NODES; ++j) att[j] *= r;
    }
    float hp[CO];
#pragma unroll
    for (int u = 0; u < CO; ++u) {
      float acc = 0.f;
#pragma unroll
      for (int j = 0; j < NODES; ++j) acc = fmaf(att[j], Wh[j][u], acc);
      hp[u] = acc;
    }
#pragma unroll
    for (int v = 0; v < NODES; ++v) {
      const float w = adj[i * NODES + v];
#pragma unroll
      for (int u = 0; u < CO; ++u) z[v][u] = fmaf(hp[u], w, z[v][u]);
    }
  }
}

// logits only (used by the pixel-softmax statistics pre-pass): e[i][j]
template <int NODES, int CO>
CGAT_HD void attn_logits_pixel(const float (&Wh)[NODES][CO], const float* __restrict__ a,
                               const uint64_t* __restrict__ maskrow, float alpha,
                               float (&e)[NODES][NODES]) {
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
template <int NODES, int CO, bool PIXEL, int MODE>
CGAT_HD void attn_backward_pixel(const float (&Wh)[NODES][CO], const float (&dz)[NODES][CO],
                                 const float* __restrict__ a, const float* __restrict__ adj,
                                 const uint64_t* __restrict__ maskrow, float alpha,
                                 const float* __restrict__ st_max, const float* __restrict__ st_rinv,
                                 const float* __restrict__ st_dot, float (&dWh)[NODES][CO],
                                 float* __restrict__ g_a, float* __restrict__ g_adj,
                                 float* __restrict__ dot_out) {
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
  float ds2[NODES];
#pragma unroll
  for (int j = 0; j < NODES; ++j) ds2[j] = 0.f;

#pragma unroll
  for (int i = 0; i < NODES; ++i) {
    // ---- recompute row i of the attention ----
    float att[NODES];
    float slope[NODES];
    const uint64_t mrow = maskrow[i];
    if (PIXEL) {
#pragma unroll
      for (int j = 0; j < NODES; ++j) {
        float pre = s1[i] + s2[j];
        slope[j] = pre > 0.f ? 1.f : alpha;
        float e = pre * slope[j];
        if (!((mrow >> j) & 1ull)) { e = kMaskFill; slope[j] = 0.f; }
        att[j] = fast_exp(e - st_max[i * NODES + j]) * st_rinv[i * NODES + j];
      }
    } else {
      float m = -INFINITY;
#pragma unroll
      for (int j = 0; j < NODES; ++j) {
        float pre = s1[i] + s2[j];
        slope[j] = pre > 0.f ? 1.f : alpha;
        float e = pre * slope[j];
        if (!((mrow >> j) & 1ull)) { e = kMaskFill; slope[j] = 0.f; }
        att[j] = e;
        m = fmaxf(m, e);
      }
      float sum = 0.f;
#pragma unroll
      for (int j = 0; j < NODES; ++j) {
        att[j] = fast_exp(att[j] - m);
        sum += att[j];
      }
      const float r = 1.f / sum;
#pragma unroll
      for (int j = 0; j < NODES; ++j) att[j] *= r;
    }
    // ---- dhp[u] = sum_v dz[v][u] adj[i][v];   g_adj[i][v] += sum_u hp[u] dz[v][u] ----
    float dhp[CO];
#pragma unroll
    for (int u = 0; u < CO; ++u) dhp[u] = 0.f;
    if (MODE == 0) {
      float hp[CO];
#pragma unroll
      for (int u = 0; u < CO; ++u) {
        float acc = 0.f;
#pragma unroll
        for (int j = 0; j < NODES; ++j) acc = fmaf(att[j], Wh[j][u], acc);
        hp[u] = acc;
      }
#pragma unroll
      for (int v = 0; v < NODES; ++v) {
        float g = 0.f;
#pragma unroll
        for (int u = 0; u < CO; ++u) g = fmaf(hp[u], dz[v][u], g);
        g_adj[i * NODES + v] += g;
      }
    }
#pragma unroll
    for (int v = 0; v < NODES; ++v) {
      const float w = adj[i * NODES + v];
#pragma unroll
      for (int u = 0; u < CO; ++u) dhp[u] = fmaf(dz[v][u], w, dhp[u]);
    }
    // ---- dAtt[j] = sum_u dhp[u] Wh[j][u] ----
    float datt[NODES];
#pragma unroll
    for (int j = 0; j < NODES; ++j) {
      float acc = 0.f;
#pragma unroll
      for (int u = 0; u < CO; ++u) acc = fmaf(dhp[u], Wh[j][u], acc);
      datt[j] = acc;
    }
    if (MODE == 1) {
#pragma unroll
      for (int j = 0; j < NODES; ++j) dot_out[i * NODES + j] += att[j] * datt[j];
      continue;
    }
    // ---- dWh[j][u] += att[j] dhp[u] ----
#pragma unroll
    for (int j = 0; j < NODES; ++j)
#pragma unroll
      for (int u = 0; u < CO; ++u) dWh[j][u] = fmaf(att[j], dhp[u], dWh[j][u]);
    // ---- soft-max backward ----
    float de[NODES];
    if (PIXEL) {
#pragma unroll
      for (int j = 0; j < NODES; ++j) de[j] = att[j] * (datt[j] - st_dot[i * NODES + j]);
    } else {
      float dot = 0.f;
#pragma unroll
      for (int j = 0; j < NODES; ++j) dot = fmaf(att[j], datt[j], dot);
#pragma unroll
      for (int j = 0; j < NODES; ++j) de[j] = att[j] * (datt[j] - dot);
    }
    // ---- LeakyReLU backward, ds1 / ds2 ----
    float ds1 = 0.f;
#pragma unroll
    for (int j = 0; j < NODES; ++j) {
      const float dp = de[j] * slope[j];
      ds1 += dp;
      ds2[j] += dp;
    }
#pragma unroll
    for (int u = 0; u < CO; ++u) {
      dWh[i][u] = fmaf(ds1, a[u], dWh[i][u]);
      g_a[u] = fmaf(ds1, Wh[i][u], g_a[u]);
    }
  }
  if (MODE == 0) {
#pragma unroll
    for (int j = 0; j < NODES; ++j)
#pragma unroll
      for (int u = 0; u < CO; ++u) {
        dWh[j][u] = fmaf(ds2[j], a[CO + u], dWh[j][u]);
        g_a[CO + u] = fmaf(ds2[j], Wh[j][u], g_a[CO + u]);
      }
  }
}

// linear projection helpers (reference baseline_model.py:127  Wh = h @ W, W is [CI][CO] row-major)
template <int NODES, int CI, int CO>
CGAT_HD void project_linear(const float (&X)[NODES][CI], const float* __restrict__ W, float (&Wh)[NODES][CO]) {
#pragma unroll
  for (int j = 0; j < NODES; ++j)
#pragma unroll
    for (int u = 0; u < CO; ++u) {
      float acc = 0.f;
#pragma unroll
      for (int t = 0; t < CI; ++t) acc = fmaf(X[j][t], W[t * CO + u], acc);
      Wh[j][u] = acc;
    }
}

// dX[j][t] += sum_u dWh[j][u] W[t][u];   g_W[t][u] += sum_j X[j][t] dWh[j][u]
template <int NODES, int CI, int CO>
CGAT_HD void project_linear_bwd(const float (&X)[NODES][CI], const float (&dWh)[NODES][CO],
                                const float* __restrict__ W, float (&dX)[NODES][CI], float* __restrict__ g_W) {
#pragma unroll
  for (int j = 0; j < NODES; ++j)
#pragma unroll
    for (int t = 0; t < CI; ++t) {
      float acc = dX[j][t];
#pragma unroll
      for (int u = 0; u < CO; ++u) acc = fmaf(dWh[j][u], W[t * CO + u], acc);
      dX[j][t] = acc;
    }
#pragma unroll
  for (int t = 0; t < CI; ++t)
#pragma unroll
    for (int u = 0; u < CO; ++u) {
      float acc = g_W[t * CO + u];
#pragma unroll
      for (int j = 0; j < NODES; ++j) acc = fmaf(X[j][t], dWh[j][u], acc);
      g_W[t * CO + u] = acc;
    }
}

}  // namespace cgat
