// Shared declarations of the fused attention kernels (attn_kernels.cu: fp32-math path; attn_h2.cu: packed-half2 path).
#pragma once
#include "common.cuh"
#include "attn_math.cuh"

namespace cgat {

constexpr int TILE = 128;  // pixels per CTA == threads per CTA
constexpr int MAX_HEADS = 8;

struct AttnArgs {
  const void* in;
  void* out;         // fwd: out;  bwd: din
  const void* dout;  // bwd only
  const float* W;
  const float* a;
  const float* adj;
  const uint8_t* mask;
  const float* stats;
  const float* bstats;
  float* gW;
  float* ga;
  float* gadj;
  float* stats_out;  // pixstats kernels
  long long n_pix;
  long long pix_per_sample;
  int heads;
  int merge;
  int apply_elu;
  float alpha;
};

// ---- record <-> register helpers ---------------------------------------------------------------
// A "record" is NODES*C consecutive elements; (node, c) lives at c*NODES+node (spatial) or node*C+c.
template <int NODES, int C, bool SPATIAL>
__device__ __forceinline__ constexpr int rec_off(int node, int c) {
  return SPATIAL ? (c * NODES + node) : (node * C + c);
}

template <int N, typename T>
__device__ __forceinline__ void load_rec(const T* __restrict__ p, float (&r)[N]) {
  constexpr int PER = 16 / sizeof(T);
  static_assert(N % PER == 0, "record must be a multiple of 16 bytes");
  const uint4* q = reinterpret_cast<const uint4*>(p);
#pragma unroll
  for (int i = 0; i < N / PER; ++i) {
    uint4 v = q[i];
    if constexpr (sizeof(T) == 4) {
      r[4 * i + 0] = __uint_as_float(v.x);
      r[4 * i + 1] = __uint_as_float(v.y);
      r[4 * i + 2] = __uint_as_float(v.z);
      r[4 * i + 3] = __uint_as_float(v.w);
    } else {
      const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        r[8 * i + 2 * k + 0] = __uint_as_float(w[k] << 16);
        r[8 * i + 2 * k + 1] = __uint_as_float(w[k] & 0xffff0000u);
      }
    }
  }
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

template <int N, typename T>
__device__ __forceinline__ void store_rec(T* __restrict__ p, const float (&r)[N]) {
  constexpr int PER = 16 / sizeof(T);
  static_assert(N % PER == 0, "record must be a multiple of 16 bytes");
  uint4* q = reinterpret_cast<uint4*>(p);
#pragma unroll
  for (int i = 0; i < N / PER; ++i) {
    uint4 v;
    if constexpr (sizeof(T) == 4) {
      v.x = __float_as_uint(r[4 * i + 0]);
      v.y = __float_as_uint(r[4 * i + 1]);
      v.z = __float_as_uint(r[4 * i + 2]);
      v.w = __float_as_uint(r[4 * i + 3]);
    } else {
      v.x = pack_bf16x2(r[8 * i + 0], r[8 * i + 1]);
      v.y = pack_bf16x2(r[8 * i + 2], r[8 * i + 3]);
      v.z = pack_bf16x2(r[8 * i + 4], r[8 * i + 5]);
      v.w = pack_bf16x2(r[8 * i + 6], r[8 * i + 7]);
    }
    q[i] = v;
  }
}

template <int NODES, int C, bool SPATIAL>
__device__ __forceinline__ void rec_to_mat(const float (&r)[NODES * C], float (&m)[NODES][C]) {
#pragma unroll
  for (int n = 0; n < NODES; ++n)
#pragma unroll
    for (int c = 0; c < C; ++c) m[n][c] = r[rec_off<NODES, C, SPATIAL>(n, c)];
}
template <int NODES, int C, bool SPATIAL>
__device__ __forceinline__ void mat_to_rec(const float (&m)[NODES][C], float (&r)[NODES * C]) {
#pragma unroll
  for (int n = 0; n < NODES; ++n)
#pragma unroll
    for (int c = 0; c < C; ++c) r[rec_off<NODES, C, SPATIAL>(n, c)] = m[n][c];
}

// ---- shared-memory carve-up ----------------------------------------------------------------------
template <int NODES, int CI, int CO>
struct SmemParams {
  float W[MAX_HEADS][CI * CO];
  float a[MAX_HEADS][2 * CO];
  float adj[MAX_HEADS][NODES * NODES];
  uint64_t maskrow[NODES];
};

template <int NODES, int CI, int CO>
__device__ __forceinline__ void load_params(SmemParams<NODES, CI, CO>& sp, const AttnArgs& A, bool need_W,
                                            bool need_adj) {
  const int tid = threadIdx.x, nt = blockDim.x;
  if (need_W)
    for (int i = tid; i < A.heads * CI * CO; i += nt) sp.W[i / (CI * CO)][i % (CI * CO)] = A.W[i];
  for (int i = tid; i < A.heads * 2 * CO; i += nt) sp.a[i / (2 * CO)][i % (2 * CO)] = A.a[i];
  if (need_adj)
    for (int i = tid; i < A.heads * NODES * NODES; i += nt)
      sp.adj[i / (NODES * NODES)][i % (NODES * NODES)] = A.adj[i];
  if (tid < NODES) {
    uint64_t m = 0;
    for (int j = 0; j < NODES; ++j)
      if (A.mask == nullptr || A.mask[tid * NODES + j] != 0) m |= (1ull << j);
    sp.maskrow[tid] = m;
  }
}


enum AttnOp { OP_FWD, OP_BWD, OP_STATS, OP_BSTATS };

// packed-half2 fast path (attn_h2.cu): bf16 I/O, neighbour soft-max.  Returns CGAT_EUNSUPPORTED for other shapes.
int attn_h2_launch(AttnOp op, const cgat_attn_desc* d, const AttnArgs& A, cudaStream_t st);
// row-of-threads-per-pixel kernels for any node count <= 64 (attn_generic.cu): neighbour soft-max, fwd / bwd only.
int attn_generic_launch(AttnOp op, const cgat_attn_desc* d, const AttnArgs& A, cudaStream_t st);

}  // namespace cgat
