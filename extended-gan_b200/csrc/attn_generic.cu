// K4g / K5g: fused graph attention for ANY node count up to 64 (BASELINE.json config 4: V in {32, 64} nodes at
// 128 x 128, the attention-kernel stress shape).  Same arithmetic and C ABI as attn_kernels.cu (reference
// convolutional_gat/baseline_model.py:127-160, neighbour soft-max), different work split: with 32-64 nodes the V x V
// attention of one pixel no longer fits one thread's registers, so a pixel is owned by a ROW OF THREADS -- thread
// (pixel p, node i) computes row i of the attention (logits, LeakyReLU, mask, soft-max over the neighbours j and the
// aggregation h'_i = sum_j att_ij Wh_j, two passes over j with Wh and s2 of the pixel staged in shared memory),
// then thread (p, v) mixes the rows with the normalised adjacency, z_v = sum_i h'_i adj[i][v], and applies ELU.
// The [V][V] attention is never stored.  The backward recomputes the forward; the sums over the row index i that
// d(Wh_j) and ds2_j need are formed by column j's own thread from per-row statistics (s1, max, 1/sum, soft-max dot)
// published in shared memory -- one extra exp per (i, j) instead of V*C shared-memory atomics.  fp32 math; fp32 or
// bf16 tensors.
#include "attn_common.cuh"

namespace cgat {

constexpr int GEN_THREADS = 512;  // V = 64: 8 pixels share the CTA's [heads][V^2] adjacency accumulators (49 KB): 2 CTAs = 32 warps per SM (at 128 threads: 8 warps, ncu warps_active 12.5 %)
constexpr int GEN_MAX_NODES = 64;
constexpr int GEN_MAX_C = 8;

struct GenGeom {
  int nodes, ci, co, heads, spatial, pre, concat, np, ppb;  // np = threads per pixel (power of two >= nodes)
  int in_rec, out_rec;
};

__host__ __device__ inline GenGeom gen_geom(int nodes, int ci, int co, int heads, int layout, int proj, int merge) {
  GenGeom g;
  g.nodes = nodes; g.ci = ci; g.co = co; g.heads = heads;
  g.spatial = layout == CGAT_LAYOUT_SPATIAL;
  g.pre = proj == CGAT_PROJ_PRE;
  g.concat = merge == CGAT_MERGE_CONCAT;
  int np = 1;
  while (np < nodes) np <<= 1;
  g.np = np;
  g.ppb = GEN_THREADS / np;
  g.in_rec = g.pre ? heads * nodes * co : nodes * ci;
  g.out_rec = g.concat ? heads * nodes * co : nodes * co;
  return g;
}

// element offset of (node, c) inside a record of `nodes` nodes with C channels each
__device__ __forceinline__ int gen_off(int spatial, int nodes, int C, int node, int c) {
  return spatial ? c * nodes + node : node * C + c;
}
// element offset inside the OUTPUT record: head k, node v, channel u
__device__ __forceinline__ int gen_out_off(const GenGeom& g, int k, int v, int u) {
  if (!g.concat) return gen_off(g.spatial, g.nodes, g.co, v, u);
  if (g.spatial) return k * g.nodes * g.co + u * g.nodes + v;
  return v * (g.heads * g.co) + k * g.co + u;
}

template <typename T> __device__ __forceinline__ float gen_ld(const T* p) { return DT<T>::to_f(*p); }

// shared memory carve-up (floats): per-CTA parameter block, then per-local-pixel blocks
struct GenSmem {
  float* a;      // [2co]
  float* W;      // [ci*co]
  float* adj;    // [nodes*nodes]
  float* gadj;   // [heads][nodes*nodes]   (backward)
  float* ga;     // [heads][2co]           (backward)
  float* gW;     // [heads][ci*co]         (backward, linear)
  uint64_t* mask;  // [nodes]
  float* Wh;     // [ppb][nodes*co]
  float* s2;     // [ppb][nodes]
  float* hp;     // [ppb][nodes*co]
  float* dz;     // [ppb][nodes*co]        (backward)
  float* dhp;    // [ppb][nodes*co]        (backward: d h'_i)
  float* st;     // [ppb][4][nodes]        (backward: s1, row max, 1/row sum, soft-max dot of every row)
};

__host__ __device__ inline size_t gen_carve(const GenGeom& g, bool bwd, unsigned char* base, GenSmem* s) {
  size_t off = 0;
  auto take = [&](size_t n_floats) {
    float* p = reinterpret_cast<float*>(base + off);
    off += ((n_floats * 4 + 15) / 16) * 16;
    return p;
  };
  const int nn = g.nodes * g.nodes, nc = g.nodes * g.co;
  float* a = take(2 * g.co);
  float* W = take(g.ci * g.co);
  float* adj = take(nn);
  float* gadj = bwd ? take((size_t)g.heads * nn) : nullptr;
  float* ga = bwd ? take((size_t)g.heads * 2 * g.co) : nullptr;
  float* gW = bwd ? take((size_t)g.heads * g.ci * g.co) : nullptr;
  uint64_t* mask = reinterpret_cast<uint64_t*>(take(2 * g.nodes));
  float* Wh = take((size_t)g.ppb * nc);
  float* s2 = take((size_t)g.ppb * g.nodes);
  float* hp = take((size_t)g.ppb * nc);
  float* dz = bwd ? take((size_t)g.ppb * nc) : nullptr;
  float* dhp = bwd ? take((size_t)g.ppb * nc) : nullptr;
  float* st = bwd ? take((size_t)g.ppb * 4 * g.nodes) : nullptr;
  if (s) *s = GenSmem{a, W, adj, gadj, ga, gW, mask, Wh, s2, hp, dz, dhp, st};
  return off;
}

// COT = channels per node at compile time (per-thread channel arrays stay in registers), 0 = run-time
template <typename T, bool BWD, int COT>
__global__ void __launch_bounds__(GEN_THREADS) attn_generic_kernel(const AttnArgs A, const GenGeom g) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  GenSmem S;
  gen_carve(g, BWD, smem_raw, &S);
  const int tid = threadIdx.x;
  const int lp = tid / g.np, i = tid % g.np;  // local pixel, node (row) of this thread
  const bool row = lp < g.ppb && i < g.nodes;
  const int nodes = g.nodes, co = COT ? COT : g.co, ci = (COT && !g.pre) ? COT : g.ci, nc = nodes * co, nn = nodes * nodes;
  const float alpha = A.alpha;
  const float inv_heads = 1.f / (float)g.heads;
  const T* in = reinterpret_cast<const T*>(A.in);

  if (tid < nodes) {
    uint64_t m = 0;
    for (int j = 0; j < nodes; ++j)
      if (A.mask == nullptr || A.mask[tid * nodes + j] != 0) m |= (1ull << j);
    S.mask[tid] = m;
  }
  if (BWD) {
    for (int q = tid; q < g.heads * nn; q += GEN_THREADS) S.gadj[q] = 0.f;
    for (int q = tid; q < g.heads * 2 * co; q += GEN_THREADS) S.ga[q] = 0.f;
    for (int q = tid; q < g.heads * ci * co; q += GEN_THREADS) S.gW[q] = 0.f;
  }
  __syncthreads();

  const long long ngroups = (A.n_pix + g.ppb - 1) / g.ppb;
  for (long long grp = blockIdx.x; grp < ngroups; grp += gridDim.x) {
    const long long pix = grp * g.ppb + lp;
    const bool act = row && pix < A.n_pix;
    float* Wh = S.Wh + (size_t)lp * nc;
    float* s2 = S.s2 + (size_t)lp * nodes;
    float* hp = S.hp + (size_t)lp * nc;
    float X[GEN_MAX_C], dX[GEN_MAX_C], oacc[GEN_MAX_C];
#pragma unroll
    for (int t = 0; t < GEN_MAX_C; ++t) { X[t] = 0.f; dX[t] = 0.f; oacc[t] = 0.f; }
    if (act && !g.pre)
      _Pragma("unroll") for (int t = 0; t < ci; ++t) X[t] = gen_ld(in + pix * g.in_rec + gen_off(g.spatial, nodes, ci, i, t));

    for (int k = 0; k < g.heads; ++k) {
      __syncthreads();  // previous head / group is done with the shared blocks
      for (int q = tid; q < 2 * co; q += GEN_THREADS) S.a[q] = A.a[k * 2 * co + q];
      for (int q = tid; q < nn; q += GEN_THREADS) S.adj[q] = A.adj[(size_t)k * nn + q];
      if (!g.pre)
        for (int q = tid; q < ci * co; q += GEN_THREADS) S.W[q] = A.W[(size_t)k * ci * co + q];
      __syncthreads();
      // ---- 1. projected features of node i, s1_i, s2_i ----
      float wh[GEN_MAX_C];
      float s1 = 0.f, s2i = 0.f;
      if (act) {
        _Pragma("unroll") for (int u = 0; u < co; ++u) {
          float v;
          if (g.pre) {
            v = gen_ld(in + pix * g.in_rec + (size_t)k * nc + gen_off(g.spatial, nodes, co, i, u));
          } else {
            v = 0.f;
            _Pragma("unroll") for (int t = 0; t < ci; ++t) v = fmaf(X[t], S.W[t * co + u], v);
          }
          wh[u] = v;
          Wh[i * co + u] = v;
          s1 = fmaf(v, S.a[u], s1);
          s2i = fmaf(v, S.a[co + u], s2i);
        }
        s2[i] = s2i;
      }
      __syncthreads();
      // ---- 2. row i: soft-max over the neighbours j and aggregation (two passes over j) ----
      float mx = -INFINITY, rinv = 0.f;
      float hpi[GEN_MAX_C];
#pragma unroll
      for (int u = 0; u < GEN_MAX_C; ++u) hpi[u] = 0.f;
      const uint64_t mrow = act ? S.mask[i] : 0;
      if (act) {
#pragma unroll 4
        for (int j = 0; j < nodes; ++j) {
          const float pre = s1 + s2[j];
          float e = pre > 0.f ? pre : alpha * pre;
          if (!((mrow >> j) & 1ull)) e = kMaskFill;
          mx = fmaxf(mx, e);
        }
        float sum = 0.f;
#pragma unroll 4
        for (int j = 0; j < nodes; ++j) {
          const float pre = s1 + s2[j];
          float e = pre > 0.f ? pre : alpha * pre;
          if (!((mrow >> j) & 1ull)) e = kMaskFill;
          const float p = fast_exp(e - mx);
          sum += p;
          _Pragma("unroll") for (int u = 0; u < co; ++u) hpi[u] = fmaf(p, Wh[j * co + u], hpi[u]);
        }
        rinv = 1.f / sum;
        _Pragma("unroll") for (int u = 0; u < co; ++u) {
          hpi[u] *= rinv;
          hp[i * co + u] = hpi[u];
        }
      }
      __syncthreads();
      // ---- 3. adjacency mix for node v = i:  z_v[u] = sum_i' hp_i'[u] adj[i'][v] ----
      float z[GEN_MAX_C];
#pragma unroll
      for (int u = 0; u < GEN_MAX_C; ++u) z[u] = 0.f;
      if (act) {
#pragma unroll 4
        for (int ii = 0; ii < nodes; ++ii) {
          const float w = S.adj[ii * nodes + i];
          _Pragma("unroll") for (int u = 0; u < co; ++u) z[u] = fmaf(hp[ii * co + u], w, z[u]);
        }
      }
      if constexpr (!BWD) {
        if (act) {
          T* out = reinterpret_cast<T*>(A.out) + pix * g.out_rec;
          _Pragma("unroll") for (int u = 0; u < co; ++u) {
            const float o = A.apply_elu ? elu_fwd<F32>(z[u]) : z[u];
            if (g.concat) out[gen_out_off(g, k, i, u)] = DT<T>::from_f(o);
            else oacc[u] += o;
          }
        }
      } else {
        // ---- 4. dz_v = dout * ELU'(z_v) ----
        float* dzs = S.dz + (size_t)lp * nc;
        float* dhps = S.dhp + (size_t)lp * nc;
        float* sts = S.st + (size_t)lp * 4 * nodes;
        if (act) {
          const T* dout = reinterpret_cast<const T*>(A.dout) + pix * g.out_rec;
          const float gs = g.concat ? 1.f : inv_heads;
          _Pragma("unroll") for (int u = 0; u < co; ++u) {
            const float d = gen_ld(dout + gen_out_off(g, k, i, u)) * gs;
            dzs[i * co + u] = A.apply_elu ? d * elu_grad<F32>(z[u]) : d;
          }
        }
        __syncthreads();
        // ---- 5a. row i: d h'_i, adjacency gradient, soft-max dot; publish the row statistics ----
        float ds1 = 0.f;
        if (act) {
          float dhp[GEN_MAX_C];
#pragma unroll
          for (int u = 0; u < GEN_MAX_C; ++u) dhp[u] = 0.f;
          float* gadj = S.gadj + (size_t)k * nn + (size_t)i * nodes;
#pragma unroll 4
          for (int t = 0; t < nodes; ++t) {
            int v = i + t;  // staggered: the pixels of a CTA that share the accumulator row hit different columns
            if (v >= nodes) v -= nodes;
            const float w = S.adj[i * nodes + v];
            float ga_iv = 0.f;
            _Pragma("unroll") for (int u = 0; u < co; ++u) {
              const float d = dzs[v * co + u];
              dhp[u] = fmaf(d, w, dhp[u]);
              ga_iv = fmaf(hpi[u], d, ga_iv);
            }
            atomicAdd(&gadj[v], ga_iv);
          }
          // one pass over the row: dot = sum_j att_ij datt_ij and, with l_ij = att_ij * LeakyReLU'(pre_ij) on the unmasked
          // entries, ds1 = sum_j l_ij (datt_ij - dot) = sum_j l_ij datt_ij - dot * sum_j l_ij  (the column sums are formed
          // by the column's thread in 5b)
          float dot = 0.f, la = 0.f, lb = 0.f;
#pragma unroll 4
          for (int j = 0; j < nodes; ++j) {
            const float pre = s1 + s2[j];
            const bool on = (mrow >> j) & 1ull;
            float e = pre > 0.f ? pre : alpha * pre;
            if (!on) e = kMaskFill;
            const float att = fast_exp(e - mx) * rinv;
            float datt = 0.f;
            _Pragma("unroll") for (int u = 0; u < co; ++u) datt = fmaf(dhp[u], Wh[j * co + u], datt);
            dot = fmaf(att, datt, dot);
            const float l = on ? att * (pre > 0.f ? 1.f : alpha) : 0.f;
            la = fmaf(l, datt, la);
            lb += l;
          }
          ds1 = fmaf(-dot, lb, la);
          _Pragma("unroll") for (int u = 0; u < co; ++u) dhps[i * co + u] = dhp[u];
          sts[i] = s1;
          sts[nodes + i] = mx;
          sts[2 * nodes + i] = rinv;
          sts[3 * nodes + i] = dot;
        }
        __syncthreads();
        // ---- 5b. column j = this thread's node: d(Wh_j) = sum_i att_ij d h'_i and ds2_j = sum_i dp_ij, with att_ij
        //      recomputed from the published row statistics (no atomics) ----
        float dwh[GEN_MAX_C];
#pragma unroll
        for (int u = 0; u < GEN_MAX_C; ++u) dwh[u] = 0.f;
        float d2 = 0.f;
        if (act) {
          const float s2j = s2[i];
#pragma unroll 4
          for (int r = 0; r < nodes; ++r) {
            const float pre = sts[r] + s2j;
            const bool on = (S.mask[r] >> i) & 1ull;
            float e = pre > 0.f ? pre : alpha * pre;
            if (!on) e = kMaskFill;
            const float att = fast_exp(e - sts[nodes + r]) * sts[2 * nodes + r];
            float datt = 0.f;
            _Pragma("unroll") for (int u = 0; u < co; ++u) {
              const float dh = dhps[r * co + u];
              datt = fmaf(dh, wh[u], datt);
              dwh[u] = fmaf(att, dh, dwh[u]);
            }
            d2 += on ? att * (datt - sts[3 * nodes + r]) * (pre > 0.f ? 1.f : alpha) : 0.f;
          }
        }
        // ---- 6. d(Wh_i), parameter-gradient partials ----
        float ga_part[2 * GEN_MAX_C];
#pragma unroll
        for (int u = 0; u < 2 * GEN_MAX_C; ++u) ga_part[u] = 0.f;
        float gw_part[GEN_MAX_C * GEN_MAX_C];  // linear projection: this thread's X^T d(Wh) terms, reduced per warp below
#pragma unroll
        for (int u = 0; u < GEN_MAX_C * GEN_MAX_C; ++u) gw_part[u] = 0.f;
        if (act) {
          T* din = reinterpret_cast<T*>(A.out) + pix * g.in_rec;
          _Pragma("unroll") for (int u = 0; u < co; ++u) {
            const float dw = dwh[u] + ds1 * S.a[u] + d2 * S.a[co + u];
            ga_part[u] = ds1 * wh[u];
            ga_part[co + u] = d2 * wh[u];
            if (g.pre) {
              din[(size_t)k * nc + gen_off(g.spatial, nodes, co, i, u)] = DT<T>::from_f(dw);
            } else {
              _Pragma("unroll") for (int t = 0; t < ci; ++t) {
                dX[t] = fmaf(dw, S.W[t * co + u], dX[t]);
                gw_part[t * GEN_MAX_C + u] = X[t] * dw;
              }
            }
          }
        }
        if (!g.pre) {
          _Pragma("unroll") for (int t = 0; t < ci; ++t)
            _Pragma("unroll") for (int u = 0; u < co; ++u) {
              const float s = warp_sum(gw_part[t * GEN_MAX_C + u]);
              if ((tid & 31) == 0) atomicAdd(&S.gW[(size_t)k * ci * co + t * co + u], s);
            }
        }
        for (int u = 0; u < 2 * co; ++u) {
          const float s = warp_sum(ga_part[u]);
          if ((tid & 31) == 0) atomicAdd(&S.ga[k * 2 * co + u], s);
        }
      }
    }
    if constexpr (!BWD) {
      if (act && !g.concat) {
        T* out = reinterpret_cast<T*>(A.out) + pix * g.out_rec;
        _Pragma("unroll") for (int u = 0; u < co; ++u) out[gen_out_off(g, 0, i, u)] = DT<T>::from_f(oacc[u] * inv_heads);
      }
    } else if (act && !g.pre) {
      T* din = reinterpret_cast<T*>(A.out) + pix * g.in_rec;
      _Pragma("unroll") for (int t = 0; t < ci; ++t) din[gen_off(g.spatial, nodes, ci, i, t)] = DT<T>::from_f(dX[t]);
    }
  }
  if constexpr (BWD) {
    __syncthreads();
    for (int q = tid; q < g.heads * nn; q += GEN_THREADS) atomicAdd(A.gadj + q, S.gadj[q]);
    for (int q = tid; q < g.heads * 2 * co; q += GEN_THREADS) atomicAdd(A.ga + q, S.ga[q]);
    if (!g.pre)
      for (int q = tid; q < g.heads * ci * co; q += GEN_THREADS) atomicAdd(A.gW + q, S.gW[q]);
  }
}

int attn_generic_supported(const cgat_attn_desc* d) {
  const int ci = d->proj == CGAT_PROJ_PRE ? d->co : d->ci;
  return d->nodes >= 1 && d->nodes <= GEN_MAX_NODES && d->co >= 1 && d->co <= GEN_MAX_C && ci >= 1 && ci <= GEN_MAX_C;
}

template <typename T>
static int gen_launch(AttnOp op, const cgat_attn_desc* d, const AttnArgs& A, const GenGeom& g, cudaStream_t st) {
  const bool bwd = op == OP_BWD;
  const size_t smem = gen_carve(g, bwd, nullptr, nullptr);
  if (smem > 227 * 1024) return fail(CGAT_EUNSUPPORTED, "generic attention tile needs %zu B of shared memory", smem);
  const long long ngroups = (d->n_pix + g.ppb - 1) / g.ppb;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  long long grid = (long long)sms * 4;  // persistent grid-stride CTAs: the per-CTA gradient accumulators are flushed once
  if (grid > ngroups) grid = ngroups;
  auto go = [&](auto kern) -> int {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return fail((int)e, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    kern<<<(unsigned)grid, GEN_THREADS, smem, st>>>(A, g);
    return 0;
  };
  const bool c4 = g.co == 4 && (g.pre || g.ci == 4);  // the reference's T = T' = 4 frames per node
  int rc;
  if (bwd) rc = c4 ? go(attn_generic_kernel<T, true, 4>) : go(attn_generic_kernel<T, true, 0>);
  else rc = c4 ? go(attn_generic_kernel<T, false, 4>) : go(attn_generic_kernel<T, false, 0>);
  if (rc) return rc;
  return check_launch(bwd ? "attn_generic_kernel<bwd>" : "attn_generic_kernel<fwd>");
}

// neighbour soft-max only (stats == NULL); called by the dispatcher of attn_kernels.cu for shapes that have no
// register-resident instantiation
int attn_generic_launch(AttnOp op, const cgat_attn_desc* d, const AttnArgs& A, cudaStream_t st) {
  if ((op != OP_FWD && op != OP_BWD) || A.stats != nullptr || !attn_generic_supported(d)) return CGAT_EUNSUPPORTED;
  const GenGeom g = gen_geom(d->nodes, d->proj == CGAT_PROJ_PRE ? d->co : d->ci, d->co, d->heads, d->layout, d->proj, d->merge);
  return d->dtype == CGAT_F32 ? gen_launch<float>(op, d, A, g, st) : gen_launch<__nv_bfloat16>(op, d, A, g, st);
}

}  // namespace cgat
