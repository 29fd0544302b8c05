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
  int sph, ext;           // fused layer kernels: score rows per head / in total behind the cout feature rows (0: none)
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
  g.sph = lf_score_rows_per_head(d.nodes);
  g.ext = d.wgrad_cols ? lf_score_rows(d.nodes, d.co, d.heads) : 0;
  g.npad = (g.cout + g.ext + 15) & ~15;
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

// score rows behind the feature rows (fused layer kernels): row = cout + k*sph + j, j < nodes: s1 of node j, else s2 of
// node j - nodes;  value = sum_u a[k][which*co + u] * w[k][u][c][tap] on the node's own input channels
__device__ __forceinline__ bool score_row(const StreamGeom& g, int row, int& k, int& which, int& node) {
  const int rs = row - g.cout;
  if (rs < 0 || rs >= g.ext) return false;
  k = rs / g.sph;
  const int j = rs - k * g.sph;
  if (j >= 2 * g.nodes) return false;
  which = j / g.nodes;
  node = j - which * g.nodes;
  return true;
}
__device__ __forceinline__ float dense_ws(const StreamGeom& g, const PtrArr& w, const PtrArr& a, int row, int tap, int cin_idx) {
  int k, which, node, node2, c;
  if (cin_idx >= g.cin || !score_row(g, row, k, which, node)) return 0.f;
  rec_inv(g.spatial, g.nodes, g.ci, cin_idx, node2, c);
  if (node != node2) return 0.f;
  float v = 0.f;
  for (int u = 0; u < g.co; ++u) v = fmaf(a.p[k][which * g.co + u], w.p[k][(u * g.ci + c) * g.taps + tap], v);
  return v;
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
  uint4* zero;                 // optional: a buffer this launch clears (the step's accumulators: no separate memset node)
  long long zero_n16;          // in 16-byte units
};

__global__ void __launch_bounds__(ADJ_THREADS) stream_prepare_kernel(const PrepArgs A) {
  griddep_launch();  // the layer kernel behind this one starts fetching its input tiles right away (common.cuh)
  griddep_wait();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < A.zero_n16; i += (long long)gridDim.x * blockDim.x)
    A.zero[i] = make_uint4(0, 0, 0, 0);
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
    // fused layer kernels (layer_fused.cu): K-chunk kk = (r * nchunk + c) * 3 + s for vertical tap r, channel chunk c,
    // horizontal tap s (the order of the interleaved stage planes); then one zero chunk when 9 * nchunk is odd, the bias
    // chunk (hi + lo bf16 parts in elements 0, 1: it multiplies a plane of ones) and one more zero chunk
    const int nj = 9 * g.nchunk, kbias = nj + (nj & 1);
    const long long n_f = (long long)lf_weight_chunks(g.cin) * g.npad * 8;
    for (long long i = tid; i < n_f; i += nt) {
      const int e = (int)(i & 7);
      const long long qq = i >> 3;
      const int row = (int)(qq % g.npad);
      const int kk = (int)(qq / g.npad);
      float v = 0.f;
      if (kk < nj) {
        const int r = kk / (3 * g.nchunk), rem = kk - r * 3 * g.nchunk, c = rem / 3, sh = rem - c * 3;
        v = row < g.cout ? dense_w(g, A.w, row, r * 3 + sh, c * 8 + e) : dense_ws(g, A.w, A.a, row, r * 3 + sh, c * 8 + e);
      } else if (kk == kbias && e < 2) {
        float b = 0.f;
        int k, which, node, u;
        if (row < g.cout) {
          const int per_head = g.nodes * g.co;
          k = row / per_head;
          rec_inv(g.spatial, g.nodes, g.co, row - k * per_head, node, u);
          b = A.bias.p[k] ? A.bias.p[k][u] : 0.f;
        } else if (score_row(g, row, k, which, node) && A.bias.p[k]) {
          for (u = 0; u < g.co; ++u) b = fmaf(A.a.p[k][which * g.co + u], A.bias.p[k][u], b);
        }
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
  const float* gbias;       // fused layer kernels: [heads][co + 2]  sum d(Wh) per channel | sum ds1 | sum ds2
  PtrArr B, w, bias, a;     // parameters (w, bias, a: only read when the partials carry score rows)
  MutPtrArr g_w, g_bias, g_a, g_B;
  int accumulate;
  // guarded train step (cgat_layer_train + cgat_layer_train_fp32): which accumulator set is valid is decided on the device
  const float* select;      // NULL, or: select[0] != 0 -> ga, gadj, gbias are read alt_offset floats further on
  long long alt_offset;
  float* loss_mse;          // optional: [0..1] := [alt_offset .. alt_offset + 1] when the alternative set is selected
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
  const bool ext = A.d.mapping == 1 && g.ext > 0;  // score rows present: d(a) is finished below, from the partials
  if (!ext) {
    for (long long i = tid; i < g.heads * 2 * g.co; i += nt) {
      float* dst = A.g_a.p[i / (2 * g.co)] + i % (2 * g.co);
      *dst = A.accumulate ? *dst + A.ga[i] : A.ga[i];
    }
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
  //
  // Score rows (fused layer kernels, g.ext > 0): the kernel's d(Wh) rows do NOT contain the terms through the scores
  // s = Wh.a; instead the partials carry  G_which[k][node][tap][c] = d(W.a_which)  in the rows behind the features:
  //   dW[k][u][c][tap] += a1[k][u] G1 + a2[k][u] G2        (same for the bias with the ones column)
  //   d(a_which)[k][u]  = sum_{node,tap,c} w[k][u][c][tap] G_which + bias[k][u] G_which(ones)     (+ ga from the kernel)
  __shared__ float wsum[ADJ_THREADS / 32];
  const int nwe = g.co * g.ci * g.taps;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int NW = ADJ_THREADS / 32;
  const int per_k = nwe + g.co + (ext ? 2 * g.co : 0);
  const size_t cta_stride = (size_t)128 * A.nt;
  const int nq = 3 * g.nchunk + 1;
  auto column = [&](int tap, int ci_idx) {  // column of (tap, dense input channel); tap < 0: the ones (bias) column
    (void)nq;  // (the fused layer kernels' partials, wgrad_cols = 1, never come here: stream_param_grads_reduced_kernel)
    return tap < 0 ? g.taps * g.cin : tap * g.cin + ci_idx;
  };
  // this thread's CTA (warp w takes the CTAs w, w+8, ...: at most one per lane with <= 256 partial tiles)
  const int my_cta = warp + NW * lane;
  const float* my_part = A.wg_partial + (size_t)(my_cta < A.ncta ? my_cta : 0) * cta_stride;
  const bool have = my_cta < A.ncta;
  for (long long i = b; i < (long long)g.heads * per_k; i += nb) {
    const int k = (int)(i / per_k);
    const int r = (int)(i - (long long)k * per_k);
    float* dst;
    float acc = 0.f;
    if (r < nwe + g.co) {
      int u, c = 0, tap = -1;
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
      const float a1 = ext ? A.a.p[k][u] : 0.f, a2 = ext ? A.a.p[k][g.co + u] : 0.f;
      // all loads of a batch of 4 nodes are issued before anything is added: a block pays the L2 latency once
      for (int node0 = 0; node0 < g.nodes; node0 += 4) {
        float v[4][3];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          v[j][0] = v[j][1] = v[j][2] = 0.f;
          const int node = node0 + j;
          if (node >= g.nodes || !have) continue;
          const int col = column(tap, rec_of(g.spatial, g.nodes, g.ci, node, c));
          v[j][0] = __ldcg(my_part + (size_t)(k * g.nodes * g.co + rec_of(g.spatial, g.nodes, g.co, node, u)) * A.nt + col);
          if (ext) {
            const int rs = g.cout + k * g.sph + node;
            v[j][1] = __ldcg(my_part + (size_t)rs * A.nt + col);
            v[j][2] = __ldcg(my_part + (size_t)(rs + g.nodes) * A.nt + col);
          }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) acc += fmaf(a2, v[j][2], fmaf(a1, v[j][1], v[j][0]));
      }
      for (int cta = my_cta + NW * 32; cta < A.ncta; cta += NW * 32) {  // more than 256 partial tiles (not on B200)
        for (int node = 0; node < g.nodes; ++node) {
          const float* part = A.wg_partial + (size_t)cta * cta_stride;
          const int col = column(tap, rec_of(g.spatial, g.nodes, g.ci, node, c));
          acc += __ldcg(part + (size_t)(k * g.nodes * g.co + rec_of(g.spatial, g.nodes, g.co, node, u)) * A.nt + col);
          if (ext) {
            const int rs = g.cout + k * g.sph + node;
            acc = fmaf(a1, __ldcg(part + (size_t)rs * A.nt + col), acc);
            acc = fmaf(a2, __ldcg(part + (size_t)(rs + g.nodes) * A.nt + col), acc);
          }
        }
      }
    } else {
      // d(a)[which*co + u] = sum_{node, tap, c} w[k][u][c][tap] G_which[k][node][tap][c] + bias[k][u] G_which(ones):
      // taps*ci + 1 terms per node and CTA, loaded in batches of 8 before they are used
      const int j0 = r - nwe - g.co, which = j0 / g.co, u = j0 - which * g.co;
      dst = A.g_a.p[k] + j0;
      const int nterm = g.taps * g.ci + 1;
      for (int cta = my_cta; cta < A.ncta; cta += NW * 32) {
        const float* part = A.wg_partial + (size_t)cta * cta_stride;
        for (int node = 0; node < g.nodes; ++node) {
          const float* prow = part + (size_t)(g.cout + k * g.sph + which * g.nodes + node) * A.nt;
          for (int t0 = 0; t0 < nterm; t0 += 8) {
            float v[8], wv[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const int t = t0 + j;
              v[j] = 0.f;
              wv[j] = 0.f;
              if (t < nterm - 1) {
                const int tap = t % g.taps, c = t / g.taps;
                v[j] = __ldcg(prow + column(tap, rec_of(g.spatial, g.nodes, g.ci, node, c)));
                wv[j] = A.w.p[k][(u * g.ci + c) * g.taps + tap];
              } else if (t == nterm - 1 && A.bias.p[k]) {
                v[j] = __ldcg(prow + column(-1, 0));
                wv[j] = A.bias.p[k][u];
              }
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) acc = fmaf(wv[j], v[j], acc);
          }
        }
      }
    }
    acc = warp_sum(acc);
    if (lane == 0) wsum[warp] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < NW; ++w) t += wsum[w];
      if (r >= nwe + g.co) t += A.ga[k * 2 * g.co + (r - nwe - g.co)];  // whatever the kernel accumulated itself
      *dst = A.accumulate ? *dst + t : t;
    }
    __syncthreads();
  }
}

// ---- fused layer kernels (wgrad_cols = 1): ONE launch finishes the step's gradients ---------------------------------------
// The layer kernel leaves compact per-CTA slots [row][9 * ci] (layer_fused.cu).  Phase A: block `row` adds that row over the
// CTAs (coalesced, fixed order: deterministic) into R (rows x 9ci floats, ~20 KB, L2-resident).  A grid barrier (all
// blocks are co-resident: <= 128 blocks of 256 threads), then phase B spread over the whole grid: the conv weight / bias
// gradients over the block-diagonal (a thread per value), d(a) through the score rows (a warp per value), the
// adjacency-normalisation backward (a block per head); a second barrier and -- on a single GPU, where no gradient
// exchange sits in between -- torch.optim.Adam on the flat parameter buffer.  The reduce / parameter-gradient / Adam
// launches of the earlier chain and their gaps become one launch.  (A first version let the LAST block to finish phase A
// do all of phase B alone: 30 us of dependent L2 round trips in one block; the earlier three launches took 12.)
//
// Score rows: the kernel's d(Wh) rows do NOT contain the terms through the scores s = (Wh + b).a; instead the slots carry
//   G_which[k][node][tap][c] = d(W.a_which)  in the rows behind the features:
//   dW[k][u][c][tap] += a1[k][u] G1 + a2[k][u] G2
//   d(a_which)[k][u]  = sum_{node,tap,c} w[k][u][c][tap] G_which + bias[k][u] sum(ds_which)     (+ ga from the kernel)
constexpr int FIN_THREADS = 256, FIN_MAX_R = 128 * 64;  // (1024 threads: the block then owns an SM and cannot overlap the loader gather or start under the train kernel's tail: step +3 us, e2e +11 us)
struct AdamArgs {
  float* p;                 // NULL: no optimiser step in this launch
  const float* g;
  float* m;
  float* v;
  long long n;
  long long* step_dev;      // optimiser steps taken so far (device counter: the launch is replayed from a CUDA graph)
  const float* hyper;       // device: lr, beta1, beta2, eps, weight_decay, grad_scale (schedulers change lr between replays)
};
struct FinishArgs {
  GradArgs G;
  unsigned int* counter;    // zeroed by the caller (per step)
  float* R;                 // [rows][nc] global scratch
  int rows, nc;
  AdamArgs adam;
  // optional: the step's scalar loss goes straight to host memory (a ring in mapped pinned memory) -- no memcpy node, no
  // copy-engine round trip between two steps
  const float* mirror_src;       // the accumulator set's base: loss at [0] (or at [alt_offset] when the fp32 re-run is valid)
  float* mirror;                 // ring of mirror_n floats, device-accessible host memory
  unsigned int mirror_n;
  unsigned int* mirror_cursor;   // device counter of values written so far
};

__device__ __forceinline__ void fin_grid_barrier(unsigned int* counter) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(counter, 1u);
    unsigned int seen;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(counter) : "memory");
    } while (seen < gridDim.x);
  }
  __syncthreads();
}

__global__ void __launch_bounds__(FIN_THREADS) stream_finish_kernel(const FinishArgs F) {
  griddep_launch();
  griddep_wait();
  constexpr int FIN_GROUPS = FIN_THREADS / 64, FIN_PER = (148 + FIN_GROUPS - 1) / FIN_GROUPS;
  __shared__ float s_part[FIN_GROUPS][64];
  const GradArgs& A = F.G;
  {  // ---- phase A: this block's row, summed over the CTAs: 64 columns x FIN_GROUPS groups of slots; a thread has ALL its
     //      loads (<= FIN_PER for up to 148 slots) in flight at once -- with 4 groups the 37 loads per thread went out in
     //      five dependent batches of L2 latency ----
    const int row = blockIdx.x, col = threadIdx.x & 63, grp = threadIdx.x >> 6;
    for (int c0 = 0; c0 < F.nc; c0 += 64) {
      float acc = 0.f;
      if (c0 + col < F.nc) {
        const float* p = A.wg_partial + (size_t)row * F.nc + c0 + col;
        const size_t cs = (size_t)F.rows * F.nc;
        for (int cb = grp; cb < A.ncta; cb += FIN_GROUPS * FIN_PER) {
          float v[FIN_PER];
#pragma unroll
          for (int j = 0; j < FIN_PER; ++j) {
            const int c = cb + j * FIN_GROUPS;
            v[j] = c < A.ncta ? __ldcg(p + (size_t)c * cs) : 0.f;
          }
#pragma unroll
          for (int j = 0; j < FIN_PER; ++j) acc += v[j];
        }
      }
      s_part[grp][col] = acc;
      __syncthreads();
      if (grp == 0 && c0 + col < F.nc) {
        float t = 0.f;
#pragma unroll
        for (int q = 0; q < FIN_GROUPS; ++q) t += s_part[q][col];
        F.R[(size_t)row * F.nc + c0 + col] = t;
      }
      __syncthreads();
    }
  }
  fin_grid_barrier(F.counter);
  // ---- phase B: everything from the reduced matrix, spread over the grid ----
  const StreamGeom g = make_geom(A.d);
  const long long sel = (A.select != nullptr && *A.select != 0.f) ? A.alt_offset : 0;  // the fp32 re-run's accumulators
  if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0 && sel != 0 && A.loss_mse != nullptr) {
    A.loss_mse[0] = A.loss_mse[sel];
    A.loss_mse[1] = A.loss_mse[sel + 1];
  }
  if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0 && F.mirror != nullptr) {
    const unsigned int cur = *F.mirror_cursor;
    F.mirror[cur % F.mirror_n] = __ldcg(F.mirror_src + sel);  // a posted write over PCIe: the kernel does not wait for it
    *F.mirror_cursor = cur + 1;
  }
  const bool ext = g.ext > 0;
  const int nwe = g.co * g.ci * g.taps;
  const float* gb = A.gbias + sel;  // [heads][co + 2]
  auto R = [&](int row, int tap, int c) { return __ldcg(F.R + (size_t)row * F.nc + tap * g.ci + c); };
  if ((int)blockIdx.x < g.heads) {
    const int k = blockIdx.x;
    adj_norm_bwd_block(A.B.p[k], A.gadj + sel + (size_t)k * g.nodes * g.nodes, A.g_B.p[k], g.nodes, A.d.transpose_adj, 0,
                       A.accumulate);
  } else {
    const int nb = gridDim.x - g.heads, b = blockIdx.x - g.heads;
    // conv weights and biases: a thread per value
    const int per_wb = nwe + g.co;
    for (int i = b * FIN_THREADS + threadIdx.x; i < g.heads * per_wb; i += nb * FIN_THREADS) {
      const int k = i / per_wb, r = i - k * per_wb;
      float* dst;
      float acc = 0.f;
      if (r < nwe) {
        const int tap = r % g.taps, c = (r / g.taps) % g.ci, u = r / (g.taps * g.ci);
        dst = A.g_w.p[k] + r;
        const float a1 = ext ? A.a.p[k][u] : 0.f, a2 = ext ? A.a.p[k][g.co + u] : 0.f;
        float v[ADJ_MAX_NODES > 8 ? 8 : ADJ_MAX_NODES][3];
        for (int node0 = 0; node0 < g.nodes; node0 += 8) {  // all loads of a batch of nodes before they are used
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            v[j][0] = v[j][1] = v[j][2] = 0.f;
            const int node = node0 + j;
            if (node >= g.nodes) continue;
            v[j][0] = R(k * g.nodes * g.co + rec_of(g.spatial, g.nodes, g.co, node, u), tap, c);
            if (ext) {
              const int rs = g.cout + k * g.sph + node;
              v[j][1] = R(rs, tap, c);
              v[j][2] = R(rs + g.nodes, tap, c);
            }
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) acc += fmaf(a2, v[j][2], fmaf(a1, v[j][1], v[j][0]));
        }
      } else {
        // bias gradient: the kernel's own sum of d(Wh) over pixels and nodes, plus the terms through the scores
        const int u = r - nwe;
        dst = A.g_bias.p[k] ? A.g_bias.p[k] + u : nullptr;
        if (dst == nullptr) continue;
        acc = gb[k * (g.co + 2) + u];
        if (ext) acc += A.a.p[k][u] * gb[k * (g.co + 2) + g.co] + A.a.p[k][g.co + u] * gb[k * (g.co + 2) + g.co + 1];
      }
      *dst = A.accumulate ? *dst + acc : acc;
    }
    // d(a): a warp per value (nodes * ci * taps terms)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = FIN_THREADS / 32;
    const int nterm = g.nodes * g.ci * g.taps;
    for (int i = (nb - 1 - b) * nw + warp; i < g.heads * 2 * g.co; i += nb * nw) {  // (from the far end: other blocks than above)
      const int k = i / (2 * g.co), j0 = i - k * 2 * g.co, which = j0 / g.co, u = j0 - which * g.co;
      float acc = 0.f;
      if (ext) {
        for (int t = lane; t < nterm; t += 32) {
          const int node = t / (g.ci * g.taps), rem = t - node * g.ci * g.taps, c = rem / g.taps, tap = rem - c * g.taps;
          acc = fmaf(A.w.p[k][(u * g.ci + c) * g.taps + tap], R(g.cout + k * g.sph + which * g.nodes + node, tap, c), acc);
        }
      }
      acc = warp_sum(acc);
      if (lane == 0) {
        if (ext && A.bias.p[k]) acc = fmaf(A.bias.p[k][u], gb[k * (g.co + 2) + g.co + which], acc);  // s = (Wh + b).a
        acc += A.ga[sel + k * 2 * g.co + j0];  // whatever the kernel accumulated itself (no score rows: all of it)
        float* dst = A.g_a.p[k] + j0;
        *dst = A.accumulate ? *dst + acc : acc;
      }
    }
  }
  if (F.adam.p == nullptr) return;
  // ---- torch.optim.Adam (convolutional_gat/train.py:212) on the flat buffers, once every gradient is in place ----
  fin_grid_barrier(F.counter + 1);
  const AdamArgs& O = F.adam;
  const long long step_i = *O.step_dev + 1;
  const float lr = O.hyper[0], b1 = O.hyper[1], b2 = O.hyper[2], eps = O.hyper[3], wd = O.hyper[4], gscale = O.hyper[5];
  const float step = (float)step_i;
  const float bc1 = 1.f - powf(b1, step), bc2 = 1.f - powf(b2, step);
  const float step_size = lr / bc1, inv_sqrt_bc2 = rsqrtf(bc2);
  for (long long i = (long long)blockIdx.x * FIN_THREADS + threadIdx.x; i < O.n; i += (long long)gridDim.x * FIN_THREADS) {
    const float pi = O.p[i];
    const float gi = fmaf(wd, pi, __ldcg(O.g + i) * gscale);
    const float mi = fmaf(b1, O.m[i], (1.f - b1) * gi);
    const float vi = fmaf(b2, O.v[i], (1.f - b2) * gi * gi);
    O.m[i] = mi;
    O.v[i] = vi;
    O.p[i] = pi - step_size * (mi / (sqrtf(vi) * inv_sqrt_bc2 + eps));
  }
  // the step counter: bumped by the last block to get here (everyone has read it above)
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(F.counter + 2, 1u) == gridDim.x - 1) *O.step_dev = step_i;
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
  if (!dgrad && d->wgrad_cols) return (int64_t)lf_weight_chunks(g.cin) * g.npad * 16;
  return dgrad ? (int64_t)g.d_npairs * 2 * g.d_npad * 16 : (int64_t)g.npairs * 2 * g.npad * 16;
}

static int prepare_impl(const cgat_stream_desc* d, const float* const* w, const float* const* bias,
                        const float* const* a, const float* const* B, void* wpack, void* wpack_dgrad,
                        float* w_stacked, float* bias_dense, float* a_stacked, float* adj, void* zero, int64_t zero_bytes,
                        void* stream) {
  if (int rc = check_desc(d)) return rc;
  if (zero_bytes < 0 || (zero_bytes > 0 && (!zero || !aligned16(zero) || zero_bytes % 16)))
    return fail(CGAT_EALIGN, "the buffer to clear must be 16-byte aligned and a multiple of 16 bytes");
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
  A.zero = (uint4*)zero;
  A.zero_n16 = zero_bytes / 16;
  const StreamGeom g = make_geom(*d);
  long long work = d->mapping == 1 ? (long long)g.npairs * 2 * g.npad * 8 : (long long)g.heads * g.ci * g.co;
  if (d->mapping == 1 && d->wgrad_cols) work = (long long)lf_weight_chunks(g.cin) * g.npad * 8;
  int blocks = (int)((work + ADJ_THREADS - 1) / ADJ_THREADS);
  if (blocks < 1) blocks = 1;
  if (blocks > 148) blocks = 148;
  cudaError_t e = launch_pdl(stream_prepare_kernel, dim3(d->heads + blocks), dim3(ADJ_THREADS), 0, (cudaStream_t)stream, A);
  if (e != cudaSuccess) return fail((int)e, "stream_prepare_kernel: %s", cudaGetErrorString(e));
  return check_launch("stream_prepare_kernel");
}

static int param_grads_impl(const cgat_stream_desc* d, const float* wg_partial, int ncta, int nt,
                            const float* gW_lin, const float* ga, const float* gadj, const float* gbias,
                            const float* const* B,
                            const float* const* w, const float* const* bias, const float* const* a,
                            float* const* g_w, float* const* g_bias, float* const* g_a, float* const* g_B,
                            int accumulate, const float* select, long long alt_offset, float* loss_mse,
                            unsigned int* counter, const AdamArgs* adam, void* stream, const float* mirror_src = nullptr,
                            float* mirror = nullptr, int mirror_n = 0, unsigned int* mirror_cursor = nullptr) {
  if (int rc = check_desc(d)) return rc;
  if (!ga || !gadj || !B || !g_w || !g_a || !g_B) return fail(CGAT_EINVAL, "null argument");
  const bool ext = d->mapping == 1 && make_geom(*d).ext > 0;
  if (ext && (!w || !a)) return fail(CGAT_EINVAL, "the fused layer kernels' partials carry score rows: w and a are needed");
  if (d->mapping == 1 && (!wg_partial || ncta < 1 || nt < 1)) return fail(CGAT_EINVAL, "conv mapping needs wgrad partials");
  if (d->mapping == 0 && !gW_lin) return fail(CGAT_EINVAL, "linear mapping needs gW");
  GradArgs A{};
  A.d = *d;
  A.wg_partial = wg_partial; A.ncta = ncta; A.nt = nt; A.gW_lin = gW_lin; A.ga = ga; A.gadj = gadj; A.gbias = gbias;
  if (d->mapping == 1 && d->wgrad_cols && !gbias) return fail(CGAT_EINVAL, "the fused layer kernels' partials need gbias");
  for (int k = 0; k < d->heads; ++k) {
    A.B.p[k] = B[k];
    A.w.p[k] = w ? w[k] : nullptr;
    A.bias.p[k] = bias ? bias[k] : nullptr;
    A.a.p[k] = a ? a[k] : nullptr;
    A.g_w.p[k] = g_w[k];
    A.g_bias.p[k] = g_bias ? g_bias[k] : nullptr;
    A.g_a.p[k] = g_a[k];
    A.g_B.p[k] = g_B[k];
  }
  A.accumulate = accumulate;
  A.select = select; A.alt_offset = alt_offset; A.loss_mse = loss_mse;
  const StreamGeom g = make_geom(*d);
  if (select != nullptr && !(d->mapping == 1 && d->wgrad_cols))
    return fail(CGAT_EINVAL, "the accumulator selector belongs to the fused layer kernels' train step (wgrad_cols = 1)");
  if (d->mapping == 1 && d->wgrad_cols) {
    // fused layer kernels: compact slots [ncta][rows][nt = 9 ci]; the reduced matrix goes to the slot behind them (the
    // workspace of cgat_layer_* has it)
    FinishArgs F{};
    F.G = A;
    F.rows = lf_partial_rows(g.nodes, g.co, g.heads);
    F.nc = nt;
    if (nt != g.taps * g.ci) return fail(CGAT_EINVAL, "partial columns %d, expected 9 * ci = %d", nt, g.taps * g.ci);
    if (F.rows > 128) return fail(CGAT_EUNSUPPORTED, "%d partial rows: the finishing grid must be co-resident", F.rows);
    if (!counter) return fail(CGAT_EINVAL, "the fused layer kernels' gradient launch needs three zeroed counter words");
    F.counter = counter;
    F.R = const_cast<float*>(wg_partial) + (size_t)ncta * F.rows * F.nc;
    if (adam) F.adam = *adam;
    if (mirror != nullptr) {
      if (!mirror_src || mirror_n < 1 || !mirror_cursor) return fail(CGAT_EINVAL, "loss mirror: source, ring size and cursor are needed");
      F.mirror_src = mirror_src; F.mirror = mirror; F.mirror_n = (unsigned)mirror_n; F.mirror_cursor = mirror_cursor;
    }
    cudaError_t e = launch_pdl(stream_finish_kernel, dim3(F.rows), dim3(FIN_THREADS), 0, (cudaStream_t)stream, F);
    if (e != cudaSuccess) return fail((int)e, "stream_finish_kernel: %s", cudaGetErrorString(e));
    return check_launch("stream_finish_kernel");
  }
  const long long work = d->mapping == 1 ? (long long)g.heads * (g.co * g.ci * g.taps + g.co + (ext ? 2 * g.co : 0))
                                         : (long long)g.heads * g.ci * g.co;
  // conv mapping: one block per output (see the kernel); linear: one thread per element
  int blocks = d->mapping == 1 ? (int)work : (int)((work + ADJ_THREADS - 1) / ADJ_THREADS);
  if (blocks < 1) blocks = 1;
  if (blocks > 148 * 8) blocks = 148 * 8;  // 8 resident blocks of 256 threads per SM: one wave
  stream_param_grads_kernel<<<d->heads + blocks, ADJ_THREADS, 0, (cudaStream_t)stream>>>(A);
  return check_launch("stream_param_grads_kernel");
}

extern "C" int cgat_stream_param_grads(const cgat_stream_desc* d, const float* wg_partial, int ncta, int nt,
                                       const float* gW_lin, const float* ga, const float* gadj, const float* gbias,
                                       const float* const* B, const float* const* w, const float* const* bias,
                                       const float* const* a, float* const* g_w, float* const* g_bias, float* const* g_a,
                                       float* const* g_B, int accumulate, void* stream) {
  if (d && d->mapping == 1 && d->wgrad_cols)
    return fail(CGAT_EINVAL, "the fused layer kernels' partials are finished by cgat_stream_finish");
  return param_grads_impl(d, wg_partial, ncta, nt, gW_lin, ga, gadj, gbias, B, w, bias, a, g_w, g_bias, g_a, g_B, accumulate,
                          nullptr, 0, nullptr, nullptr, nullptr, stream);
}

extern "C" int cgat_stream_finish_mirror(const cgat_stream_desc* d, const float* wg_partial, int ncta, int nt, const float* ga,
                                         const float* gadj, const float* gbias, const float* const* B, const float* const* w,
                                         const float* const* bias, const float* const* a, float* const* g_w,
                                         float* const* g_bias, float* const* g_a, float* const* g_B, int accumulate,
                                         const float* select, int64_t alt_offset, float* loss_mse, uint32_t* counter,
                                         float* adam_param, const float* adam_grad, float* adam_m, float* adam_v,
                                         int64_t adam_n, int64_t* adam_step_dev, const float* adam_hyper,
                                         const float* mirror_src, float* mirror, int32_t mirror_n, uint32_t* mirror_cursor,
                                         void* stream) {
  if (!d || d->mapping != 1 || !d->wgrad_cols) return fail(CGAT_EINVAL, "cgat_stream_finish serves the fused layer kernels (wgrad_cols = 1)");
  AdamArgs O{};
  if (adam_param != nullptr) {
    if (!adam_grad || !adam_m || !adam_v || adam_n < 1 || !adam_step_dev || !adam_hyper) return fail(CGAT_EINVAL, "bad Adam arguments");
    O.p = adam_param; O.g = adam_grad; O.m = adam_m; O.v = adam_v; O.n = adam_n; O.step_dev = (long long*)adam_step_dev;
    O.hyper = adam_hyper;
  }
  return param_grads_impl(d, wg_partial, ncta, nt, nullptr, ga, gadj, gbias, B, w, bias, a, g_w, g_bias, g_a, g_B, accumulate,
                          select, (long long)alt_offset, loss_mse, counter, adam_param ? &O : nullptr, stream, mirror_src, mirror,
                          mirror_n, mirror_cursor);
}

extern "C" int cgat_stream_finish(const cgat_stream_desc* d, const float* wg_partial, int ncta, int nt, const float* ga,
                                  const float* gadj, const float* gbias, const float* const* B, const float* const* w,
                                  const float* const* bias, const float* const* a, float* const* g_w, float* const* g_bias,
                                  float* const* g_a, float* const* g_B, int accumulate, const float* select,
                                  int64_t alt_offset, float* loss_mse, uint32_t* counter, float* adam_param,
                                  const float* adam_grad, float* adam_m, float* adam_v, int64_t adam_n,
                                  int64_t* adam_step_dev, const float* adam_hyper, void* stream) {
  return cgat_stream_finish_mirror(d, wg_partial, ncta, nt, ga, gadj, gbias, B, w, bias, a, g_w, g_bias, g_a, g_B, accumulate,
                                   select, alt_offset, loss_mse, counter, adam_param, adam_grad, adam_m, adam_v, adam_n,
                                   adam_step_dev, adam_hyper, nullptr, nullptr, 0, nullptr, stream);
}

extern "C" int cgat_stream_prepare(const cgat_stream_desc* d, const float* const* w, const float* const* bias,
                                   const float* const* a, const float* const* B, void* wpack, void* wpack_dgrad,
                                   float* w_stacked, float* bias_dense, float* a_stacked, float* adj, void* stream) {
  return prepare_impl(d, w, bias, a, B, wpack, wpack_dgrad, w_stacked, bias_dense, a_stacked, adj, nullptr, 0, stream);
}

extern "C" int cgat_stream_prepare_clear(const cgat_stream_desc* d, const float* const* w, const float* const* bias,
                                         const float* const* a, const float* const* B, void* wpack, void* wpack_dgrad,
                                         float* w_stacked, float* bias_dense, float* a_stacked, float* adj, void* clear,
                                         int64_t clear_bytes, void* stream) {
  return prepare_impl(d, w, bias, a, B, wpack, wpack_dgrad, w_stacked, bias_dense, a_stacked, adj, clear, clear_bytes, stream);
}
