// Host instantiation of the per-pixel attention math (extended-gan_b200/csrc/attn_math.cuh).
// TEST INFRASTRUCTURE: lets the CPU-only suite check the exact arithmetic the CUDA kernels run
// (forward and hand-derived backward) against the oracle before any GPU time is spent.  It is never
// loaded by the product package.
#include <cstdint>
#include <cstring>
#include <vector>
#include "../../extended-gan_b200/csrc/attn_math.cuh"

using namespace cgat;

template <int NODES, int CI, int CO>
static void run_fwd(int pixel_mode, long n_pix, const float* X, const float* W, const float* a, const float* adj,
                    const uint64_t* maskrow, float alpha, const float* st_max, const float* st_rinv, float* z_out) {
  for (long p = 0; p < n_pix; ++p) {
    float x[NODES][CI], Wh[NODES][CO], z[NODES][CO];
    std::memcpy(x, X + p * NODES * CI, sizeof(x));
    project_linear<F32, NODES, CI, CO>(x, W, Wh);
    for (int v = 0; v < NODES; ++v)
      for (int u = 0; u < CO; ++u) z[v][u] = 0.f;
    if (pixel_mode)
      attn_forward_pixel<F32, NODES, CO, true>(Wh, a, adj, maskrow, alpha, st_max, st_rinv, z);
    else
      attn_forward_pixel<F32, NODES, CO, false>(Wh, a, adj, maskrow, alpha, nullptr, nullptr, z);
    std::memcpy(z_out + p * NODES * CO, z, sizeof(z));
  }
}

template <int NODES, int CI, int CO>
static void run_logits(long n_pix, const float* X, const float* W, const float* a, const uint64_t* maskrow,
                       float alpha, float* e_out) {
  for (long p = 0; p < n_pix; ++p) {
    float x[NODES][CI], Wh[NODES][CO], e[NODES][NODES];
    std::memcpy(x, X + p * NODES * CI, sizeof(x));
    project_linear<F32, NODES, CI, CO>(x, W, Wh);
    attn_logits_pixel<NODES, CO>(Wh, a, maskrow, alpha, e);
    std::memcpy(e_out + p * NODES * NODES, e, sizeof(e));
  }
}

// mode 0: full backward (dX per pixel; gW, ga, gadj summed over pixels).  mode 1: dot[i][j] summed over pixels.
template <int NODES, int CI, int CO>
static void run_bwd(int pixel_mode, int mode, long n_pix, const float* X, const float* dZ, const float* W,
                    const float* a, const float* adj, const uint64_t* maskrow, float alpha, const float* st_max,
                    const float* st_rinv, const float* st_dot, float* dX_out, float* gW, float* ga, float* gadj,
                    float* dot_out) {
  for (long p = 0; p < n_pix; ++p) {
    float x[NODES][CI], Wh[NODES][CO], dz[NODES][CO], dWh[NODES][CO], dX[NODES][CI];
    std::memcpy(x, X + p * NODES * CI, sizeof(x));
    std::memcpy(dz, dZ + p * NODES * CO, sizeof(dz));
    project_linear<F32, NODES, CI, CO>(x, W, Wh);
    for (int v = 0; v < NODES; ++v)
      for (int u = 0; u < CO; ++u) dWh[v][u] = 0.f;
    for (int v = 0; v < NODES; ++v)
      for (int t = 0; t < CI; ++t) dX[v][t] = 0.f;
    if (mode == 1) {
      attn_backward_pixel<F32, NODES, CO, true, 1>(Wh, dz, a, adj, maskrow, alpha, st_max, st_rinv, nullptr, dWh, nullptr,
                                              nullptr, dot_out);
      continue;
    }
    float g_a[2 * CO] = {0}, g_adj[NODES * NODES] = {0}, g_W[CI * CO] = {0};
    if (pixel_mode)
      attn_backward_pixel<F32, NODES, CO, true, 0>(Wh, dz, a, adj, maskrow, alpha, st_max, st_rinv, st_dot, dWh, g_a, g_adj,
                                              nullptr);
    else
      attn_backward_pixel<F32, NODES, CO, false, 0>(Wh, dz, a, adj, maskrow, alpha, nullptr, nullptr, nullptr, dWh, g_a,
                                               g_adj, nullptr);
    project_linear_bwd<F32, NODES, CI, CO>(x, dWh, W, dX, g_W);
    std::memcpy(dX_out + p * NODES * CI, dX, sizeof(dX));
    for (int i = 0; i < 2 * CO; ++i) ga[i] += g_a[i];
    for (int i = 0; i < NODES * NODES; ++i) gadj[i] += g_adj[i];
    for (int i = 0; i < CI * CO; ++i) gW[i] += g_W[i];
  }
}

// the restructured neighbour-soft-max math of the fused layer kernels (attn_nb_forward / attn_nb_backward)
template <int NODES, int CI, int CO>
static void run_nb(int masked, long n_pix, const float* X, const float* dZ, const float* W, const float* a,
                   const float* adj, const uint64_t* maskrow, float alpha, float* z_out, float* dX_out, float* gW,
                   float* ga, float* gadj) {
  for (long p = 0; p < n_pix; ++p) {
    float x[NODES][CI], Wh[NODES][CO], z[NODES][CO], dz[NODES][CO], dWh[NODES][CO], dX[NODES][CI];
    std::memcpy(x, X + p * NODES * CI, sizeof(x));
    std::memcpy(dz, dZ + p * NODES * CO, sizeof(dz));
    project_linear<F32, NODES, CI, CO>(x, W, Wh);
    NbState<F32, NODES> st;
    float g_a[2 * CO] = {0}, g_adj[NODES * NODES] = {0}, g_W[CI * CO] = {0};
    if (masked) {
      attn_nb_forward<F32, NODES, CO, true>(Wh, a, adj, maskrow, alpha, st, z);
      attn_nb_backward<F32, NODES, CO, true>(Wh, dz, a, adj, maskrow, alpha, st, dWh, g_a, g_adj);
    } else {
      attn_nb_forward<F32, NODES, CO, false>(Wh, a, adj, maskrow, alpha, st, z);
      attn_nb_backward<F32, NODES, CO, false>(Wh, dz, a, adj, maskrow, alpha, st, dWh, g_a, g_adj);
    }
    for (int v = 0; v < NODES; ++v)
      for (int t = 0; t < CI; ++t) dX[v][t] = 0.f;
    project_linear_bwd<F32, NODES, CI, CO>(x, dWh, W, dX, g_W);
    std::memcpy(z_out + p * NODES * CO, z, sizeof(z));
    std::memcpy(dX_out + p * NODES * CI, dX, sizeof(dX));
    for (int i = 0; i < 2 * CO; ++i) ga[i] += g_a[i];
    for (int i = 0; i < NODES * NODES; ++i) gadj[i] += g_adj[i];
    for (int i = 0; i < CI * CO; ++i) gW[i] += g_W[i];
  }
}

#define DISPATCH(CALL)                                                   \
  if (nodes == 6 && ci == 4 && co == 4) { CALL(6, 4, 4); return 0; }     \
  if (nodes == 4 && ci == 6 && co == 6) { CALL(4, 6, 6); return 0; }     \
  if (nodes == 8 && ci == 4 && co == 4) { CALL(8, 4, 4); return 0; }     \
  return -2;

extern "C" int hh_fwd(int nodes, int ci, int co, int pixel_mode, long n_pix, const float* X, const float* W,
                      const float* a, const float* adj, const uint64_t* maskrow, float alpha, const float* st_max,
                      const float* st_rinv, float* z_out) {
#define CALL(N, I, O) run_fwd<N, I, O>(pixel_mode, n_pix, X, W, a, adj, maskrow, alpha, st_max, st_rinv, z_out)
  DISPATCH(CALL)
#undef CALL
}

extern "C" int hh_logits(int nodes, int ci, int co, long n_pix, const float* X, const float* W, const float* a,
                         const uint64_t* maskrow, float alpha, float* e_out) {
#define CALL(N, I, O) run_logits<N, I, O>(n_pix, X, W, a, maskrow, alpha, e_out)
  DISPATCH(CALL)
#undef CALL
}

extern "C" int hh_bwd(int nodes, int ci, int co, int pixel_mode, int mode, long n_pix, const float* X, const float* dZ,
                      const float* W, const float* a, const float* adj, const uint64_t* maskrow, float alpha,
                      const float* st_max, const float* st_rinv, const float* st_dot, float* dX_out, float* gW,
                      float* ga, float* gadj, float* dot_out) {
#define CALL(N, I, O)                                                                                              \
  run_bwd<N, I, O>(pixel_mode, mode, n_pix, X, dZ, W, a, adj, maskrow, alpha, st_max, st_rinv, st_dot, dX_out, gW, \
                   ga, gadj, dot_out)
  DISPATCH(CALL)
#undef CALL
}

extern "C" int hh_nb(int nodes, int ci, int co, int masked, long n_pix, const float* X, const float* dZ, const float* W,
                     const float* a, const float* adj, const uint64_t* maskrow, float alpha, float* z_out,
                     float* dX_out, float* gW, float* ga, float* gadj) {
#define CALL(N, I, O) run_nb<N, I, O>(masked, n_pix, X, dZ, W, a, adj, maskrow, alpha, z_out, dX_out, gW, ga, gadj)
  DISPATCH(CALL)
#undef CALL
}
