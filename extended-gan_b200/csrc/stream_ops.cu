// Parameter plumbing of one conv-GAT stream as TWO launches per step instead of ~25 framework kernels.
//
// The reference keeps per-head parameters (attention_{k}.W / .conv.weight, .a, .B; baseline_model.py:111-116,
// 191-192).  The fused attention / conv kernels want them stacked, normalised and -- for the conv mapping --
// expanded to the block-diagonal dense conv and packed in UMMA core-matrix order.  `stream_prepare_kernel` does all
// of that from the raw per-head pointers in one launch; `stream_param_grads_kernel` does the reverse for the
// gradients (reduction of the wgrad partial sums over CTAs and over the block-diagonal, adjacency-normalisation
// backward, un-stacking), optionally ACCUMULATING straight into the parameters' .grad buffers so that no
// per-parameter accumulate kernel runs.
#include "adj_math.cuh"

namespace cgat {

constexpr int SO_MAX_HEADS = 8;
struct PtrArr { const float* p[SO_MAX_HEADS]; };
struct MutPtrArr { float* p[SO_MAX_HEADS]; };

struct StreamGeom {
  int nodes, ci, co, heads, spatial;
  int cin, cout;          // dense conv channels: nodes*ci, heads*nodes*co
  int taps;               // 9 (3x3) for the conv mapping
  int nchunk, npairs, npad;        // fprop packing (conv_tc.cu: chunk-major K order)
  int d_nchunk, d_npairs, d_npad;  // dgrad packing (GEMM-K = cout, GEMM-N = cin)
};

__host__ __device__ inline StreamGeom make_geom(const cgat_stream_desc& d) {
  StreamGeom g;
  g.nodes = d.nodes; g.ci = d.ci; g.co = d.co; g.heads = d.heads; g.spatial = d.layout == CGAT_LAYOUT_SPATIAL;
  g.cin = d.nodes * d.ci;
  g.cout = d.heads * d.nodes * d.co;
  g.taps = 9;
  g.nchunk = g.cin / 8;
  g.npairs = (g.nchunk * g.taps + 1) / 2;
  g.npad = (g.cout + 15) & ~15;
  g.d_nchunk = g.cout / 8;
  g.d_npairs = (g.d_nchunk * g.taps + 1) / 2;
  g.d_npad = (g.cin + 15) & ~15;
  return g;
}

// pixel-record offset of (node, c) for C channels per node
__device__ __forceinline__ int rec_of(int spatial, int nodes, int C, int node, int c) {
  return spatial ? c * nodes + node : node * C + c;
}
__device__ __forceinline__ void rec_inv(int spatial, int nodes, int C, int idx, int& node, int& c) {
  if (spatial) { c = idx / nodes; node = idx - c * nodes; } else { node = idx / C; c = idx - node * C; }
}

// dense conv weight  Wd[row][tap][cin_idx] = w[k][u][c][tap] * delta(node, node')   (row = (k, node, u), cin_idx = (node', c))
__device__ __forceinline__ float dense_w(const StreamGeom& g, const PtrArr& w, int row, int tap, int cin_idx) {
  if (row >= g.cout || cin_idx >= g.cin) return 0.f;
  const int per_head = g.nodes * g.co;
  const int k = row / per_head;
  int node, u, node2, c;
  rec_inv(g.spatial, g.nodes, g.co, row - k * per_head, node, u);
  rec_inv(g.spatial, g.nodes, g.ci, cin_idx, node2, c);
  if (node != node2) return 0.f;
  return w.p[k][(u * g.ci + c) * g.taps + tap];  // conv.weight [co][ci][3][3]
}

struct PrepArgs {
  cgat_stream_desc d;
  PtrArr w, bias, a, B;
  __nv_bfloat16* wpack;        // conv: fprop packing
  __nv_bfloat16* wpack_dgrad;  // conv: dgrad packing or NULL
  float* w_stacked;            // linear: [heads][ci][co]
  float* bias_dense;           // conv: [cout]
  float* a_stacked;            // [heads][2co]
  float* adj;                  // [heads][nodes][nodes] normalised
};

__global__ void __launch_bounds__(ADJ_THREADS) stream_prepare_kernel(const PrepArgs A) {
  const StreamGeom g = make_geom(A.d);
  if ((int)blockIdx.x < g.heads) {  // adjacency normalisation of head blockIdx.x (baseline_model.py:41-50)
    adj_norm_fwd_block(A.B.p[blockIdx.x], A.adj + (size_t)blockIdx.x * g.nodes * g.nodes, g.nodes, A.d.transpose_adj, 0);
    return;
  }
  const int nb = gridDim.x - g.heads;
  const int b = blockIdx.x - g.heads;
  const long long tid = (long long)b * blockDim.x + threadIdx.x, nt = (long long)nb * blockDim.x;
  for (long long i = tid; i < g.heads * 2 * g.co; i += nt) A.a_stacked[i] = A.a.p[i / (2 * g.co)][i % (2 * g.co)];
  if (A.d.mapping == 0) {
    for (long long i = tid; i < (long long)g.heads * g.ci * g.co; i += nt)
      A.w_stacked[i] = A.w.p[i / (g.ci * g.co)][i % (g.ci * g.co)];
    return;
  }
  for (long long i = tid; i < g.cout; i += nt) {
    const int per_head = g.nodes * g.co;
    const int k = (int)i / per_head;
    int node, u;
    rec_inv(g.spatial, g.nodes, g.co, (int)i - k * per_head, node, u);
    A.bias_dense[i] = A.bias.p[k] ? A.bias.p[k][u] : 0.f;
  }
  if (A.d.wgrad_cols) {
    // fused layer kernels (layer_fused.cu): K-chunk kk = r*nq + q, q = s*nchunk + c for the column plane of horizontal
    // tap s and channel chunk c, vertical tap r; q = nq-1 multiplies the plane of ones: bias (hi + lo bf16 parts) at r = 0
    const int nq = 3 * g.nchunk + 1;
    const long long n_f = (long long)3 * nq * g.npad * 8;
    for (long long i = tid; i < n_f; i += nt) {
      const int e = (int)(i & 7);
      const long long qq = i >> 3;
      const int row = (int)(qq % g.npad);
      const int kk = (int)(qq / g.npad);
      const int r = kk / nq, q = kk - r * nq;
      float v = 0.f;
      if (q < nq - 1) {
        const int s = q / g.nchunk, c = q - s * g.nchunk;
        v = dense_w(g, A.w, row, r * 3 + s, c * 8 + e);
      } else if (r == 0 && row < g.cout && e < 2) {
        const int per_head = g.nodes * g.co;
        const int k = row / per_head;
        int node, u;
        rec_inv(g.spatial, g.nodes, g.co, row - k * per_head, node, u);
        const float b = A.bias.p[k] ? A.bias.p[k][u] : 0.f;
        const float hi = __bfloat162float(__float2bfloat16_rn(b));
        v = e == 0 ? hi : b - hi;
      }
      A.wpack[i] = __float2bfloat16_rn(v);
    }
  } else {
  // fprop packing: out[(ki*npad + row)*8 + e], ki = cchunk*taps + tap, value Wd[row][tap][cchunk*8+e]
  const long long n_f = (long long)g.npairs * 2 * g.npad * 8;
  for (long long i = tid; i < n_f; i += nt) {
    const int e = (int)(i & 7);
    const long long q = i >> 3;
    const int row = (int)(q % g.npad);
    const int ki = (int)(q / g.npad);
    const int cch = ki / g.taps, tap = ki - cch * g.taps;
    const float v = cch < g.nchunk ? dense_w(g, A.w, row, tap, cch * 8 + e) : 0.f;
    A.wpack[i] = __float2bfloat16_rn(v);
  }
  }
  // dgrad packing: GEMM-K = dense cout, GEMM-N = cin, kernel rotated: value Wd[k][taps-1-tap][row]
  if (A.wpack_dgrad != nullptr) {
    const long long n_d = (long long)g.d_npairs * 2 * g.d_npad * 8;
    for (long long i = tid; i < n_d; i += nt) {
      const int e = (int)(i & 7);
      const long long q = i >> 3;
      const int row = (int)(q % g.d_npad);
      const int ki = (int)(q / g.d_npad);
      const int cch = ki / g.taps, tap = ki - cch * g.taps;
      const float v = cch < g.d_nchunk ? dense_w(g, A.w, cch * 8 + e, g.taps - 1 - tap, row) : 0.f;
      A.wpack_dgrad[i] = __float2bfloat16_rn(v);
    }
  }
}

struct GradArgs {
  cgat_stream_desc d;
  const float* wg_partial;  // conv: [ncta][128][nt] wgrad partial sums (cout rows, taps*cin + ones columns)
  int ncta, nt;
  const float* gW_lin;      // linear: [heads][ci][co]
  const float* ga;          // [heads][2co]
  const float* gadj;        // [heads][nodes][nodes]
  PtrArr B;
  MutPtrArr g_w, g_bias, g_a, g_B;
  int accumulate;
};

__global__ void __launch_bounds__(ADJ_THREADS) stream_param_grads_kernel(const GradArgs A) {
  const StreamGeom g = make_geom(A.d);
  if ((int)blockIdx.x < g.heads) {
    const int k = blockIdx.x;
    adj_norm_bwd_block(A.B.p[k], A.gadj + (size_t)k * g.nodes * g.nodes, A.g_B.p[k], g.nodes, A.d.transpose_adj, 0,
                       A.accumulate);
    return;
  }
  const int nb = gridDim.x - g.heads;
  const int b = blockIdx.x - g.heads;
  const long long tid = (long long)b * blockDim.x + threadIdx.x, nt = (long long)nb * blockDim.x;
  for (long long i = tid; i < g.heads * 2 * g.co; i += nt) {
    float* dst = A.g_a.p[i / (2 * g.co)] + i % (2 * g.co);
    *dst = A.accumulate ? *dst + A.ga[i] : A.ga[i];
  }
  if (A.d.mapping == 0) {
    for (long long i = tid; i < (long long)g.heads * g.ci * g.co; i += nt) {
      float* dst = A.g_w.p[i / (g.ci * g.co)] + i % (g.ci * g.co);
      *dst = A.accumulate ? *dst + A.gW_lin[i] : A.gW_lin[i];
    }
    return;
  }
  // conv.weight grad [k][u][c][tap] = sum over nodes and CTAs of the dense partials (the block-diagonal's delta).
  // One BLOCK per output: warp w takes the CTAs w, w+8, ... -- with 148 partial tiles that is at most one load per
  // lane and node, all issued before anything is added, so a block pays the L2 latency once (a warp per output walked
  // it once per node and 32-CTA batch: 10.6 us under ncu) -- then the 8 warp sums are added in a fixed order:
  // deterministic, identical on every rank.
  __shared__ float wsum[ADJ_THREADS / 32];
  const int nwe = g.co * g.ci * g.taps;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int NW = ADJ_THREADS / 32;
  for (long long i = b; i < (long long)g.heads * (nwe + g.co); i += nb) {
    const int k = (int)(i / (nwe + g.co));
    const int r = (int)(i - (long long)k * (nwe + g.co));
    int u, c = 0, tap = 0;
    float* dst;
    if (r < nwe) {
      tap = r % g.taps;
      c = (r / g.taps) % g.ci;
      u = r / (g.taps * g.ci);
      dst = A.g_w.p[k] + r;
    } else {
      u = r - nwe;
      dst = A.g_bias.p[k] ? A.g_bias.p[k] + u : nullptr;
    }
    if (dst == nullptr) continue;  // block-uniform
    float acc = 0.f;
    const size_t cta_stride = (size_t)128 * A.nt;
    for (int node0 = 0; node0 < g.nodes; node0 += 4) {
      float v[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        v[j] = 0.f;
        const int node = node0 + j;
        if (node >= g.nodes) continue;
        const int row = k * g.nodes * g.co + rec_of(g.spatial, g.nodes, g.co, node, u);
        int col;
        if (A.d.wgrad_cols) {  // layer_fused.cu: [r][(s, cin chunk) | ones][8]; dbias is the ones column of r = 0
          const int nq = 3 * g.nchunk + 1;
          col = (nq - 1) * 8;
          if (r < nwe) {
            const int ci_idx = rec_of(g.spatial, g.nodes, g.ci, node, c);
            col = ((tap / 3) * nq + (tap % 3) * g.nchunk + (ci_idx >> 3)) * 8 + (ci_idx & 7);
          }
        } else {
          col = g.taps * g.cin;
          if (r < nwe) col = tap * g.cin + rec_of(g.spatial, g.nodes, g.ci, node, c);
        }
        const float* src = A.wg_partial + (size_t)row * A.nt + col;
        for (int cta = warp + NW * lane; cta < A.ncta; cta += NW * 32) v[j] += __ldcg(src + (size_t)cta * cta_stride);
      }
      acc += (v[0] + v[1]) + (v[2] + v[3]);
    }
    acc = warp_sum(acc);
    if (lane == 0) wsum[warp] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < NW; ++w) t += wsum[w];
      *dst = A.accumulate ? *dst + t : t;
    }
    __syncthreads();
  }
}

static int check_desc(const cgat_stream_desc* d) {
  if (!d) return fail(CGAT_EINVAL, "null stream descriptor");
  if (d->heads < 1 || d->heads > SO_MAX_HEADS || d->nodes < 1 || d->nodes > ADJ_MAX_NODES || d->ci < 1 || d->co < 1)
    return fail(CGAT_EINVAL, "bad stream descriptor");
  if (d->mapping != 0 && d->mapping != 1) return fail(CGAT_EINVAL, "mapping must be 0 (linear) or 1 (conv)");
  if (d->mapping == 1 && ((d->nodes * d->ci) % 8 || (d->heads * d->nodes * d->co) % 8))
    return fail(CGAT_EUNSUPPORTED, "conv mapping needs nodes*ci and heads*nodes*co to be multiples of 8");
  return 0;
}

}  // namespace cgat

using namespace cgat;

extern "C" int64_t cgat_stream_wpack_bytes(const cgat_stream_desc* d, int dgrad) {
  if (check_desc(d) || d->mapping != 1) return 0;
  const StreamGeom g = make_geom(*d);
  if (!dgrad && d->wgrad_cols) return (int64_t)3 * (3 * g.nchunk + 1) * g.npad * 16;
  return dgrad ? (int64_t)g.d_npairs * 2 * g.d_npad * 16 : (int64_t)g.npairs * 2 * g.npad * 16;
}

extern "C" int cgat_stream_prepare(const cgat_stream_desc* d, const float* const* w, const float* const* bias,
                                   const float* const* a, const float* const* B, void* wpack, void* wpack_dgrad,
                                   float* w_stacked, float* bias_dense, float* a_stacked, float* adj, void* stream) {
  if (int rc = check_desc(d)) return rc;
  if (!w || !a || !B || !a_stacked || !adj) return fail(CGAT_EINVAL, "null argument");
  if (d->mapping == 1 && (!wpack || !bias_dense)) return fail(CGAT_EINVAL, "conv mapping needs wpack and bias_dense");
  if (d->mapping == 0 && !w_stacked) return fail(CGAT_EINVAL, "linear mapping needs w_stacked");
  PrepArgs A{};
  A.d = *d;
  for (int k = 0; k < d->heads; ++k) {
    A.w.p[k] = w[k];
    A.bias.p[k] = bias ? bias[k] : nullptr;
    A.a.p[k] = a[k];
    A.B.p[k] = B[k];
  }
  A.wpack = (__nv_bfloat16*)wpack;
  A.wpack_dgrad = (__nv_bfloat16*)wpack_dgrad;
  A.w_stacked = w_stacked;
  A.bias_dense = bias_dense;
  A.a_stacked = a_stacked;
  A.adj = adj;
  const StreamGeom g = make_geom(*d);
  long long work = d->mapping == 1 ? (long long)g.npairs * 2 * g.npad * 8 : (long long)g.heads * g.ci * g.co;
  if (d->mapping == 1 && d->wgrad_cols) work = (long long)3 * (3 * g.nchunk + 1) * g.npad * 8;
  int blocks = (int)((work + ADJ_THREADS - 1) / ADJ_THREADS);
  if (blocks < 1) blocks = 1;
  if (blocks > 148) blocks = 148;
  stream_prepare_kernel<<<d->heads + blocks, ADJ_THREADS, 0, (cudaStream_t)stream>>>(A);
  return check_launch("stream_prepare_kernel");
}

extern "C" int cgat_stream_param_grads(const cgat_stream_desc* d, const float* wg_partial, int ncta, int nt,
                                       const float* gW_lin, const float* ga, const float* gadj, const float* const* B,
                                       float* const* g_w, float* const* g_bias, float* const* g_a, float* const* g_B,
                                       int accumulate, void* stream) {
  if (int rc = check_desc(d)) return rc;
  if (!ga || !gadj || !B || !g_w || !g_a || !g_B) return fail(CGAT_EINVAL, "null argument");
  if (d->mapping == 1 && (!wg_partial || ncta < 1 || nt < 1)) return fail(CGAT_EINVAL, "conv mapping needs wgrad partials");
  if (d->mapping == 0 && !gW_lin) return fail(CGAT_EINVAL, "linear mapping needs gW");
  GradArgs A{};
  A.d = *d;
  A.wg_partial = wg_partial; A.ncta = ncta; A.nt = nt; A.gW_lin = gW_lin; A.ga = ga; A.gadj = gadj;
  for (int k = 0; k < d->heads; ++k) {
    A.B.p[k] = B[k];
    A.g_w.p[k] = g_w[k];
    A.g_bias.p[k] = g_bias ? g_bias[k] : nullptr;
    A.g_a.p[k] = g_a[k];
    A.g_B.p[k] = g_B[k];
  }
  A.accumulate = accumulate;
  const StreamGeom g = make_geom(*d);
  const long long work = d->mapping == 1 ? (long long)g.heads * (g.co * g.ci * g.taps + g.co) : (long long)g.heads * g.ci * g.co;
  // conv mapping: one block per output (see the kernel); linear: one thread per element
  int blocks = d->mapping == 1 ? (int)work : (int)((work + ADJ_THREADS - 1) / ADJ_THREADS);
  if (blocks < 1) blocks = 1;
  if (blocks > 148 * 8) blocks = 148 * 8;  // 8 resident blocks of 256 threads per SM: one wave
  stream_param_grads_kernel<<<d->heads + blocks, ADJ_THREADS, 0, (cudaStream_t)stream>>>(A);
  return check_launch("stream_param_grads_kernel");
}
