// Device-side bodies of the learnable-adjacency normalisation (shared by adj_kernels.cu and stream_ops.cu).
#pragma once
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
  for (int w = 1; w < (int)(blockDim.x >> 5); ++w) {
    const float x = scratch[w];
    r = op == 0 ? r + x : (op == 1 ? fminf(r, x) : fmaxf(r, x));
  }
  return r;
}

// one CTA (any multiple of 32 threads up to 1024; every thread must call) normalises head `head`
__device__ __forceinline__ void adj_norm_fwd_block(const float* __restrict__ B, float* __restrict__ out, int nodes,
                                                   int transpose, int head) {
  __shared__ float scratch[32];  // one slot per warp of a block of up to 1024 threads
  __shared__ float rs[ADJ_MAX_NODES];
  const int nn = nodes * nodes;
  const float* Bh = B + (size_t)head * nn;
  float* oh = out + (size_t)head * nn;
  float mn = INFINITY, mx = -INFINITY;
  for (int q = threadIdx.x; q < nn; q += blockDim.x) {
    const float m = Bh[q] + ((q / nodes == q % nodes) ? 1.f : 0.f);
    mn = fminf(mn, m);
    mx = fmaxf(mx, m);
  }
  mn = block_reduce(mn, 1, scratch);
  mx = block_reduce(mx, 2, scratch);
  const float inv = 1.f / (mx - mn);
  // row sums -> r_i = sqrt(1/d_i)
  for (int i = threadIdx.x; i < nodes; i += blockDim.x) {
    float d = 0.f;
    for (int k = 0; k < nodes; ++k) d += (Bh[i * nodes + k] + (i == k ? 1.f : 0.f) - mn) * inv;
    rs[i] = sqrtf(1.f / d);
  }
  __syncthreads();
  for (int q = threadIdx.x; q < nn; q += blockDim.x) {
    const int i = q / nodes, k = q % nodes;
    const float adj = (Bh[q] + (i == k ? 1.f : 0.f) - mn) * inv;
    const float v = rs[i] * adj * rs[k];
    oh[transpose ? (k * nodes + i) : q] = v;
  }
}

// backward of the above for head `head`; accumulate != 0 adds into gB instead of overwriting
__device__ __forceinline__ void adj_norm_bwd_block(const float* __restrict__ B, const float* __restrict__ g,
                                                   float* __restrict__ gB, int nodes, int transpose, int head,
                                                   int accumulate) {
  __shared__ float scratch[32];  // one slot per warp of a block of up to 1024 threads
  __shared__ float rs[ADJ_MAX_NODES];
  const int nn = nodes * nodes;
  const float* Bh = B + (size_t)head * nn;
  const float* gh = g + (size_t)head * nn;
  float* oh = gB + (size_t)head * nn;
  float mn = INFINITY, mx = -INFINITY;
  for (int q = threadIdx.x; q < nn; q += blockDim.x) {
    const float m = Bh[q] + ((q / nodes == q % nodes) ? 1.f : 0.f);
    mn = fminf(mn, m);
    mx = fmaxf(mx, m);
  }
  mn = block_reduce(mn, 1, scratch);
  mx = block_reduce(mx, 2, scratch);
  const float inv = 1.f / (mx - mn);
  for (int i = threadIdx.x; i < nodes; i += blockDim.x) {
    float d = 0.f;
    for (int k = 0; k < nodes; ++k) d += (Bh[i * nodes + k] + (i == k ? 1.f : 0.f) - mn) * inv;
    rs[i] = sqrtf(1.f / d);
  }
  __syncthreads();
  // d_mn = sum dadj (adj-1)/S ; d_mx = -sum dadj adj / S ; tie counts
  float dmn = 0.f, dmx = 0.f, cmn = 0.f, cmx = 0.f;
  for (int q = threadIdx.x; q < nn; q += blockDim.x) {
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
  for (int q = threadIdx.x; q < nn; q += blockDim.x) {
    const int i = q / nodes, k = q % nodes;
    const float m = Bh[q] + (i == k ? 1.f : 0.f);
    float v = gh[transpose ? (k * nodes + i) : q] * rs[i] * rs[k] * inv;
    if (m == mn) v += dmn / cmn;
    if (m == mx) v += dmx / cmx;
    oh[q] = accumulate ? oh[q] + v : v;
  }
}

}  // namespace cgat
