// The 1-D graph-attention layer of convolutional_gat/baseline_model.py:27-56 (`GraphAttentionLayer`), used by
// `BaselineModel` (:236-270): node features are flattened to F = H*W*T per vertex, so the layer is a GEMM
// (Wh = h.W, done by the conv kernels as a 1x1 convolution) followed by attention over the V vertices with
// F'-dimensional features.  This file holds everything after the GEMM, forward and backward:
//   s1[n,i] = Wh[n,i,:].a[:F'],  s2[n,j] = Wh[n,j,:].a[F':]                          (:36-37, :58-65)
//   e = LeakyReLU(s1_i + s2_j), att = softmax_j(e)                                    (:38-39)
//   M = A_hat . att                                                                   (:53)
//   out[n,v,:] = ELU(sum_j M[n,v,j] Wh[n,j,:])                                         (:54-56)
// Shapes here are "few vertices (V <= 32), many features": kernels parallelise over features and reduce with
// warp shuffles; everything is fp32 (the reference runs this path in fp32; F can be 16k).
#include "common.cuh"

namespace cgat {

constexpr int G1_THREADS = 256;
constexpr int G1_MAXV = 32;

__device__ __forceinline__ float block_sum(float v, float* scratch) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = 0.f;
  for (int w = 0; w < G1_THREADS / 32; ++w) r += scratch[w];
  return r;
}

// grid = N*V blocks: s1, s2 of one (sample, vertex)
__global__ void __launch_bounds__(G1_THREADS) gat1d_scores_kernel(const float* __restrict__ Wh, const float* __restrict__ a,
                                                                  float* __restrict__ s1, float* __restrict__ s2, int F) {
  __shared__ float scratch[G1_THREADS / 32];
  const float* row = Wh + (size_t)blockIdx.x * F;
  float p = 0.f, q = 0.f;
  for (int f = threadIdx.x; f < F; f += G1_THREADS) {
    const float w = row[f];
    p = fmaf(w, a[f], p);
    q = fmaf(w, a[F + f], q);
  }
  p = block_sum(p, scratch);
  q = block_sum(q, scratch);
  if (threadIdx.x == 0) { s1[blockIdx.x] = p; s2[blockIdx.x] = q; }
}

// grid = N blocks of V*V threads (<= 1024): att[n][i][j] and M[n][v][j] = sum_i A_hat[v][i] att[i][j]
__global__ void gat1d_attn_kernel(const float* __restrict__ s1, const float* __restrict__ s2,
                                  const float* __restrict__ adj /*A_hat [v][i]*/, const uint8_t* __restrict__ mask,
                                  float* __restrict__ att, float* __restrict__ M, int V, float alpha) {
  __shared__ float s_att[G1_MAXV * G1_MAXV];
  const int n = blockIdx.x;
  const int i = threadIdx.x / V, j = threadIdx.x % V;
  const float pre = s1[n * V + i] + s2[n * V + j];
  float e = pre > 0.f ? pre : alpha * pre;
  if (mask != nullptr && mask[i * V + j] == 0) e = -9e15f;
  s_att[i * V + j] = e;
  __syncthreads();
  float m = -INFINITY;
  for (int k = 0; k < V; ++k) m = fmaxf(m, s_att[i * V + k]);
  float sum = 0.f;
  for (int k = 0; k < V; ++k) sum += expf(s_att[i * V + k] - m);
  const float av = expf(e - m) / sum;
  __syncthreads();
  s_att[i * V + j] = av;
  att[(size_t)n * V * V + i * V + j] = av;
  __syncthreads();
  // thread (i, j) now plays (v = i, j)
  float acc = 0.f;
  for (int k = 0; k < V; ++k) acc = fmaf(adj[i * V + k], s_att[k * V + j], acc);
  M[(size_t)n * V * V + i * V + j] = acc;
}

// grid = (ceil(F/256), N): out[n,v,f] = ELU(sum_j M[n,v,j] Wh[n,j,f])
__global__ void __launch_bounds__(G1_THREADS) gat1d_aggregate_kernel(const float* __restrict__ Wh,
                                                                     const float* __restrict__ M,
                                                                     float* __restrict__ out, int V, int F) {
  __shared__ float s_M[G1_MAXV * G1_MAXV];
  const int n = blockIdx.y;
  for (int q = threadIdx.x; q < V * V; q += G1_THREADS) s_M[q] = M[(size_t)n * V * V + q];
  __syncthreads();
  const int f = blockIdx.x * G1_THREADS + threadIdx.x;
  if (f >= F) return;
  float wh[G1_MAXV];
  for (int j = 0; j < V; ++j) wh[j] = Wh[((size_t)n * V + j) * F + f];
  for (int v = 0; v < V; ++v) {
    float z = 0.f;
    for (int j = 0; j < V; ++j) z = fmaf(s_M[v * V + j], wh[j], z);
    out[((size_t)n * V + v) * F + f] = z > 0.f ? z : expm1f(z);
  }
}

// backward 1: grid = (ceil(F/256), N):  dz = dout * ELU'(z) (from the saved output), then
//   dWh[n,j,f] = sum_v M[n,v,j] dz[n,v,f]        (written)
//   dM[n,v,j] += sum_f dz[n,v,f] Wh[n,j,f]       (block partial sums -> atomicAdd)
__global__ void __launch_bounds__(G1_THREADS) gat1d_bwd_aggregate_kernel(const float* __restrict__ Wh,
                                                                         const float* __restrict__ M,
                                                                         const float* __restrict__ out,
                                                                         const float* __restrict__ dout,
                                                                         float* __restrict__ dWh, float* __restrict__ dM,
                                                                         int V, int F) {
  __shared__ float s_M[G1_MAXV * G1_MAXV];
  __shared__ float scratch[G1_THREADS / 32];
  const int n = blockIdx.y;
  for (int q = threadIdx.x; q < V * V; q += G1_THREADS) s_M[q] = M[(size_t)n * V * V + q];
  __syncthreads();
  const int f = blockIdx.x * G1_THREADS + threadIdx.x;
  const bool ok = f < F;
  float wh[G1_MAXV], dz[G1_MAXV];
  for (int j = 0; j < V; ++j) {
    const size_t idx = ((size_t)n * V + j) * F + f;
    wh[j] = ok ? Wh[idx] : 0.f;
    const float o = ok ? out[idx] : 0.f;
    dz[j] = ok ? dout[idx] * (o > 0.f ? 1.f : o + 1.f) : 0.f;  // ELU'(z) = 1 (z>0) or exp(z) = out + 1
  }
  if (ok)
    for (int j = 0; j < V; ++j) {
      float acc = 0.f;
      for (int v = 0; v < V; ++v) acc = fmaf(s_M[v * V + j], dz[v], acc);
      dWh[((size_t)n * V + j) * F + f] = acc;
    }
  for (int v = 0; v < V; ++v)
    for (int j = 0; j < V; ++j) {
      const float s = block_sum(dz[v] * wh[j], scratch);
      if (threadIdx.x == 0) atomicAdd(&dM[(size_t)n * V * V + v * V + j], s);
    }
}

// backward 2: grid = N blocks of V*V threads: through M = A_hat.att, the soft-max and LeakyReLU.
//   dadj[v][i] += sum_j dM[v][j] att[i][j]  (atomicAdd over samples);  ds1[n,i], ds2[n,j] written
__global__ void gat1d_bwd_attn_kernel(const float* __restrict__ s1, const float* __restrict__ s2,
                                      const float* __restrict__ adj, const uint8_t* __restrict__ mask,
                                      const float* __restrict__ att, const float* __restrict__ dM,
                                      float* __restrict__ dadj, float* __restrict__ ds1, float* __restrict__ ds2, int V,
                                      float alpha) {
  __shared__ float s_att[G1_MAXV * G1_MAXV], s_dM[G1_MAXV * G1_MAXV], s_dp[G1_MAXV * G1_MAXV];
  const int n = blockIdx.x;
  const int i = threadIdx.x / V, j = threadIdx.x % V;
  s_att[i * V + j] = att[(size_t)n * V * V + i * V + j];
  s_dM[i * V + j] = dM[(size_t)n * V * V + i * V + j];
  __syncthreads();
  // thread (v=i, k=j): dadj[v][k] += sum_jj dM[v][jj] att[k][jj]
  {
    float acc = 0.f;
    for (int jj = 0; jj < V; ++jj) acc = fmaf(s_dM[i * V + jj], s_att[j * V + jj], acc);
    atomicAdd(&dadj[i * V + j], acc);
  }
  // datt[i][j] = sum_v adj[v][i] dM[v][j]
  float datt = 0.f;
  for (int v = 0; v < V; ++v) datt = fmaf(adj[v * V + i], s_dM[v * V + j], datt);
  s_dp[i * V + j] = datt;
  __syncthreads();
  float dot = 0.f;
  for (int k = 0; k < V; ++k) dot = fmaf(s_att[i * V + k], s_dp[i * V + k], dot);
  const float de = s_att[i * V + j] * (datt - dot);
  const float pre = s1[n * V + i] + s2[n * V + j];
  float slope = pre > 0.f ? 1.f : alpha;
  if (mask != nullptr && mask[i * V + j] == 0) slope = 0.f;
  __syncthreads();
  s_dp[i * V + j] = de * slope;
  __syncthreads();
  if (j == 0) {
    float acc = 0.f;
    for (int k = 0; k < V; ++k) acc += s_dp[i * V + k];
    ds1[n * V + i] = acc;
  }
  if (i == 0) {
    float acc = 0.f;
    for (int k = 0; k < V; ++k) acc += s_dp[k * V + j];
    ds2[n * V + j] = acc;
  }
}

// backward 3: grid = ceil(F/256) blocks:  dWh[n,i,f] += ds1[n,i] a[f] + ds2[n,i] a[F+f];
//   da[f] = sum_{n,i} ds1 Wh[n,i,f],  da[F+f] = sum_{n,i} ds2 Wh[n,i,f]     (written)
__global__ void __launch_bounds__(G1_THREADS) gat1d_bwd_scores_kernel(const float* __restrict__ Wh,
                                                                      const float* __restrict__ a,
                                                                      const float* __restrict__ ds1,
                                                                      const float* __restrict__ ds2,
                                                                      float* __restrict__ dWh, float* __restrict__ da,
                                                                      int NV, int F) {
  const int f = blockIdx.x * G1_THREADS + threadIdx.x;
  if (f >= F) return;
  const float a1 = a[f], a2 = a[F + f];
  float g1 = 0.f, g2 = 0.f;
  for (int r = 0; r < NV; ++r) {
    const size_t idx = (size_t)r * F + f;
    const float w = Wh[idx], d1 = ds1[r], d2 = ds2[r];
    dWh[idx] += d1 * a1 + d2 * a2;
    g1 = fmaf(d1, w, g1);
    g2 = fmaf(d2, w, g2);
  }
  da[f] = g1;
  da[F + f] = g2;
}

}  // namespace cgat

using namespace cgat;

extern "C" int cgat_gat1d_fwd(const float* Wh, const float* a, const float* adj, const uint8_t* mask, float* s1, float* s2,
                              float* att, float* M, float* out, int n, int v, int f, float alpha, void* stream) {
  if (!Wh || !a || !adj || !s1 || !s2 || !att || !M || !out) return fail(CGAT_EINVAL, "null argument");
  if (n < 1 || v < 1 || v > G1_MAXV || f < 1) return fail(CGAT_EINVAL, "gat1d: n=%d v=%d (<= %d) f=%d", n, v, G1_MAXV, f);
  cudaStream_t st = (cudaStream_t)stream;
  gat1d_scores_kernel<<<n * v, G1_THREADS, 0, st>>>(Wh, a, s1, s2, f);
  gat1d_attn_kernel<<<n, v * v, 0, st>>>(s1, s2, adj, mask, att, M, v, alpha);
  gat1d_aggregate_kernel<<<dim3((f + G1_THREADS - 1) / G1_THREADS, n), G1_THREADS, 0, st>>>(Wh, M, out, v, f);
  return check_launch("gat1d forward");
}

extern "C" int cgat_gat1d_bwd(const float* Wh, const float* a, const float* adj, const uint8_t* mask, const float* s1,
                              const float* s2, const float* att, const float* M, const float* out, const float* dout,
                              float* dWh, float* da, float* dadj /*zeroed by caller*/, float* dM /*zeroed by caller*/,
                              float* ds1, float* ds2, int n, int v, int f, float alpha, void* stream) {
  if (!Wh || !a || !adj || !s1 || !s2 || !att || !M || !out || !dout || !dWh || !da || !dadj || !dM || !ds1 || !ds2)
    return fail(CGAT_EINVAL, "null argument");
  if (n < 1 || v < 1 || v > G1_MAXV || f < 1) return fail(CGAT_EINVAL, "gat1d: n=%d v=%d (<= %d) f=%d", n, v, G1_MAXV, f);
  cudaStream_t st = (cudaStream_t)stream;
  gat1d_bwd_aggregate_kernel<<<dim3((f + G1_THREADS - 1) / G1_THREADS, n), G1_THREADS, 0, st>>>(Wh, M, out, dout, dWh, dM,
                                                                                                 v, f);
  gat1d_bwd_attn_kernel<<<n, v * v, 0, st>>>(s1, s2, adj, mask, att, dM, dadj, ds1, ds2, v, alpha);
  gat1d_bwd_scores_kernel<<<(f + G1_THREADS - 1) / G1_THREADS, G1_THREADS, 0, st>>>(Wh, a, ds1, ds2, dWh, da, n * v, f);
  return check_launch("gat1d backward");
}
