// a5: learnable adjacency normalisation and its backward (one CTA per head).
//
// Reference: convolutional_gat/baseline_model.py:41-50 and :133-142
//   adj = B + I;  adj = (adj - min adj) / (max adj - min adj);
//   D = diag(rowsum(adj)) detached (legacy Variable(..., requires_grad=False), :48/:140);
//   A_hat = sqrt(inverse(D)) . adj . sqrt(inverse(D)).
// The reference spends ~10 tiny ATen launches (incl. an LU inverse of a diagonal matrix) per layer
// per step on this; here it is one launch for all heads.  Backward follows torch autograd: the
// gradient of a full min()/max() reduction is distributed evenly over tied elements.
#include "common.cuh"

namespace cgat {

constexpr int ADJ_THREADS = 256;
constexpr int ADJ_MAX_NODES = 64;

__device__ __forceinline__ float block_reduce(float v, int op, float* scratch) {
  // op 0: sum, 1: min, 2: max
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float w = __shfl_xor_sync(0xffffffffu, v, o);
    v = op == 0 ? v + w : (op == 1 ? fminf(v, w) : fmaxf(v, w));
  }
  __syncthreads();
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  float r = scratch[0];
  for (int w = 1; w < ADJ_THREADS / 32; ++w) {
    const float x = scratch[w];
    r = op == 0 ? r + x : (op == 1 ? fminf(r, x) : fmaxf(r, x));
  }
  return r;
}

__global__ void __launch_bounds__(ADJ_THREADS) adj_norm_fwd_kernel(const float* __restrict__ B,
                                                                   float* __restrict__ out, int nodes,
                                                                   int transpose) {
  __shared__ float scratch[ADJ_THREADS / 32];
  __shared__ float rs[ADJ_MAX_NODES];
  const int nn = nodes * nodes;
  const float* Bh = B + (size_t)blockIdx.x * nn;
  float* oh = out + (size_t)blockIdx.x * nn;
  float mn = INFINITY, mx = -INFINITY;
  for (int q = threadIdx.x; q < nn; q += ADJ_THREADS) {
    const float m = Bh[q] + ((q / nodes == q % nodes) ? 1.f : 0.f);
    mn = fminf(mn, m);
    mx = fmaxf(mx, m);
  }
  mn = block_reduce(mn, 1, scratch);
  mx = block_reduce(mx, 2, scratch);
  const float inv = 1.f / (mx - mn);
  // row sums -> r_i = sqrt(1/d_i)
  for (int i = threadIdx.x; i < nodes; i += ADJ_THREADS) {
    float d = 0.f;
    for (int k = 0; k < nodes; ++k) d += (Bh[i * nodes + k] + (i == k ? 1.f : 0.f) - mn) * inv;
    rs[i] = sqrtf(1.f / d);
  }
  __syncthreads();
  for (int q = threadIdx.x; q < nn; q += ADJ_THREADS) {
    const int i = q / nodes, k = q % nodes;
    const float adj = (Bh[q] + (i == k ? 1.f : 0.f) - mn) * inv;
    const float v = rs[i] * adj * rs[k];
    oh[transpose ? (k * nodes + i) : q] = v;
  }
}

__global__ void __launch_bounds__(ADJ_THREADS) adj_norm_bwd_kernel(const float* __restrict__ B,
                                                                   const float* __restrict__ g,
                                                                   float* __restrict__ gB, int nodes,
                                                                   int transpose) {
  __shared__ float scratch[ADJ_THREADS / 32];
  __shared__ float rs[ADJ_MAX_NODES];
  const int nn = nodes * nodes;
  const float* Bh = B + (size_t)blockIdx.x * nn;
  const float* gh = g + (size_t)blockIdx.x * nn;
  float* oh = gB + (size_t)blockIdx.x * nn;
  float mn = INFINITY, mx = -INFINITY;
  for (int q = threadIdx.x; q < nn; q += ADJ_THREADS) {
    const float m = Bh[q] + ((q / nodes == q % nodes) ? 1.f : 0.f);
    mn = fminf(mn, m);
    mx = fmaxf(mx, m);
  }
  mn = block_reduce(mn, 1, scratch);
  mx = block_reduce(mx, 2, scratch);
  const float inv = 1.f / (mx - mn);
  for (int i = threadIdx.x; i < nodes; i += ADJ_THREADS) {
    float d = 0.f;
    for (int k = 0; k < nodes; ++k) d += (Bh[i * nodes + k] + (i == k ? 1.f : 0.f) - mn) * inv;
    rs[i] = sqrtf(1.f / d);
  }
  __syncthreads();
  // d_mn = sum dadj (adj-1)/S ; d_mx = -sum dadj adj / S ; tie counts
  float dmn = 0.f, dmx = 0.f, cmn = 0.f, cmx = 0.f;
  for (int q = threadIdx.x; q < nn; q += ADJ_THREADS) {
    const int i = q / nodes, k = q % nodes;
    const float m = Bh[q] + (i == k ? 1.f : 0.f);
    const float adj = (m - mn) * inv;
    const float dadj = gh[transpose ? (k * nodes + i) : q] * rs[i] * rs[k];
    dmn += dadj * (adj - 1.f) * inv;
    dmx -= dadj * adj * inv;
    cmn += (m == mn) ? 1.f : 0.f;
    cmx += (m == mx) ? 1.f : 0.f;
  }
  dmn = block_reduce(dmn, 0, scratch);
  dmx = block_reduce(dmx, 0, scratch);
  cmn = block_reduce(cmn, 0, scratch);
  cmx = block_reduce(cmx, 0, scratch);
  for (int q = threadIdx.x; q < nn; q += ADJ_THREADS) {
    const int i = q / nodes, k = q % nodes;
    const float m = Bh[q] + (i == k ? 1.f : 0.f);
    float v = gh[transpose ? (k * nodes + i) : q] * rs[i] * rs[k] * inv;
    if (m == mn) v += dmn / cmn;
    if (m == mx) v += dmx / cmx;
    oh[q] = v;
  }
}

}  // namespace cgat

using namespace cgat;

extern "C" int cgat_adj_norm_fwd(const float* B, float* adj_hat, int heads, int nodes, int transpose, void* stream) {
  if (!B || !adj_hat) return fail(CGAT_EINVAL, "null argument");
  if (heads < 1 || nodes < 1 || nodes > ADJ_MAX_NODES)
    return fail(CGAT_EINVAL, "heads=%d nodes=%d out of range (nodes <= %d)", heads, nodes, ADJ_MAX_NODES);
  adj_norm_fwd_kernel<<<heads, ADJ_THREADS, 0, (cudaStream_t)stream>>>(B, adj_hat, nodes, transpose);
  return check_launch("adj_norm_fwd_kernel");
}

extern "C" int cgat_adj_norm_bwd(const float* B, const float* gadj, float* gB, int heads, int nodes, int transpose,
                                 void* stream) {
  if (!B || !gadj || !gB) return fail(CGAT_EINVAL, "null argument");
  if (heads < 1 || nodes < 1 || nodes > ADJ_MAX_NODES)
    return fail(CGAT_EINVAL, "heads=%d nodes=%d out of range (nodes <= %d)", heads, nodes, ADJ_MAX_NODES);
  adj_norm_bwd_kernel<<<heads, ADJ_THREADS, 0, (cudaStream_t)stream>>>(B, gadj, gB, nodes, transpose);
  return check_launch("adj_norm_bwd_kernel");
}
