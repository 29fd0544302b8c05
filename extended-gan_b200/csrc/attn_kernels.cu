// K4 / K5: fused graph-attention forward and backward for the conv-GAT layer (sm_100a).
//
// One thread owns one pixel record and runs the whole per-pixel attention in registers
// (attn_math.cuh).  A CTA handles a tile of TILE consecutive pixel records, which are CONTIGUOUS in
// HBM for every layout the layer uses ([N,H,W,T,V] has the T*V record innermost), so the tile moves
// with one TMA bulk copy (cp.async.bulk, global->shared, mbarrier completion) and leaves with one
// bulk store; the threads only touch shared memory with 16-byte LDS/STS.  HBM traffic is exactly the
// algorithmic bytes: read `in` once, write `out` once (forward); read `in`, `dout`, write `din`
// (backward).  Nothing of the reference's [N,V,V,P,2T'] concat or [N,V,V,P,P] diag_embed
// (convolutional_gat/baseline_model.py:146,162-169) is ever materialised.
//
// Parameter gradients (W, a, adjacency) are reduced per CTA through shared memory and added to the
// fp32 accumulators with one atomicAdd per value per CTA.
#include "attn_common.cuh"
#include <cstdlib>

namespace cgat {

// ===================================================================================================
// K4 forward
// ===================================================================================================
template <int NODES, int CI, int CO, bool SPATIAL, bool PRE, typename T>
__global__ void __launch_bounds__(TILE) attn_fwd_kernel(const AttnArgs A) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  constexpr int IN_SUB = PRE ? NODES * CO : NODES * CI;  // elements per (pixel, head) input sub-record
  constexpr int OUT_SUB = NODES * CO;
  const int heads = A.heads;
  const int in_rec = PRE ? heads * IN_SUB : IN_SUB;
  const int out_rec = (A.merge == CGAT_MERGE_MEAN) ? OUT_SUB : heads * OUT_SUB;

  using SP = SmemParams<NODES, CI, CO>;
  SP& sp = *reinterpret_cast<SP*>(smem_raw);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw + ((sizeof(SP) + 15) & ~15));
  T* s_in = reinterpret_cast<T*>(smem_raw + ((sizeof(SP) + 15) & ~15) + 128);
  T* s_out = s_in + (size_t)TILE * in_rec;

  const int tid = threadIdx.x;
  const long long pix0 = (long long)blockIdx.x * TILE;
  const int npix = (int)min((long long)TILE, A.n_pix - pix0);

  if (tid == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  __syncthreads();
  if (tid == 0) {
    const uint32_t bytes = (uint32_t)npix * in_rec * sizeof(T);
    mbar_arrive_expect_tx(bar, bytes);
    bulk_g2s(s_in, reinterpret_cast<const T*>(A.in) + pix0 * in_rec, bytes, bar);
  }
  load_params(sp, A, !PRE, true);
  __syncthreads();
  mbar_wait(bar, 0);

  if (tid < npix) {
    const bool pixel_mode = A.stats != nullptr;
    const long long sample = (pix0 + tid) / A.pix_per_sample;
    float X[NODES][PRE ? 1 : CI];
    if constexpr (!PRE) {
      float r[NODES * CI];
      load_rec<NODES * CI, T>(s_in + (size_t)tid * in_rec, r);
      rec_to_mat<NODES, CI, SPATIAL>(r, X);
    }
    float acc[NODES][CO];
#pragma unroll
    for (int v = 0; v < NODES; ++v)
#pragma unroll
      for (int u = 0; u < CO; ++u) acc[v][u] = 0.f;
    const float inv_heads = 1.f / (float)heads;

    for (int k = 0; k < heads; ++k) {
      float Wh[NODES][CO];
      if constexpr (PRE) {
        float r[NODES * CO];
        load_rec<NODES * CO, T>(s_in + (size_t)tid * in_rec + k * IN_SUB, r);
        rec_to_mat<NODES, CO, SPATIAL>(r, Wh);
      } else {
        project_linear<F32, NODES, CI, CO>(X, sp.W[k], Wh);
      }
      float z[NODES][CO];
#pragma unroll
      for (int v = 0; v < NODES; ++v)
#pragma unroll
        for (int u = 0; u < CO; ++u) z[v][u] = 0.f;
      if (pixel_mode) {
        const float* st = A.stats + ((sample * heads + k) * 2) * (NODES * NODES);
        attn_forward_pixel<F32, NODES, CO, true>(Wh, sp.a[k], sp.adj[k], sp.maskrow, A.alpha, st,
                                            st + NODES * NODES, z);
      } else {
        attn_forward_pixel<F32, NODES, CO, false>(Wh, sp.a[k], sp.adj[k], sp.maskrow, A.alpha, nullptr, nullptr, z);
      }
      if (A.apply_elu) {
#pragma unroll
        for (int v = 0; v < NODES; ++v)
#pragma unroll
          for (int u = 0; u < CO; ++u) z[v][u] = elu_fwd<F32>(z[v][u]);
      }
      if (A.merge == CGAT_MERGE_MEAN) {
#pragma unroll
        for (int v = 0; v < NODES; ++v)
#pragma unroll
          for (int u = 0; u < CO; ++u) acc[v][u] = fmaf(z[v][u], inv_heads, acc[v][u]);
      } else if (SPATIAL) {
        // concat, spatial: element (k,u,v) at (k*CO+u)*NODES+v  -> head sub-record is contiguous
        float r[NODES * CO];
        mat_to_rec<NODES, CO, true>(z, r);
        store_rec<NODES * CO, T>(s_out + (size_t)tid * out_rec + k * OUT_SUB, r);
      } else {
        // concat, temporal: element (t,k,u) at t*(heads*CO) + k*CO + u
        T* o = s_out + (size_t)tid * out_rec;
#pragma unroll
        for (int v = 0; v < NODES; ++v)
#pragma unroll
          for (int u = 0; u < CO; ++u) o[v * (heads * CO) + k * CO + u] = DT<T>::from_f(z[v][u]);
      }
    }
    if (A.merge == CGAT_MERGE_MEAN) {
      float r[NODES * CO];
      mat_to_rec<NODES, CO, SPATIAL>(acc, r);
      store_rec<NODES * CO, T>(s_out + (size_t)tid * out_rec, r);
    }
  }
  fence_proxy_async_smem();
  __syncthreads();
  if (tid == 0) {
    bulk_s2g(reinterpret_cast<T*>(A.out) + pix0 * out_rec, s_out, (uint32_t)npix * out_rec * sizeof(T));
    bulk_commit();
    bulk_wait_read0();
  }
}

// ===================================================================================================
// K5 backward.  MODE 0: full backward.  MODE 1: pixel-mode pre-pass (bstats += att*dAtt).
// ===================================================================================================
template <int NODES, int CI, int CO, bool PRE>
struct RedLayout {
  static constexpr int ADJ = 0;
  static constexpr int AV = NODES * NODES;
  static constexpr int WV = AV + 2 * CO;
  static constexpr int R = WV + (PRE ? 0 : CI * CO);
};

template <int NODES, int CI, int CO, bool SPATIAL, bool PRE, typename T, int MODE>
__global__ void __launch_bounds__(TILE) attn_bwd_kernel(const AttnArgs A) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  constexpr int IN_SUB = PRE ? NODES * CO : NODES * CI;
  constexpr int OUT_SUB = NODES * CO;
  using RL = RedLayout<NODES, CI, CO, PRE>;
  constexpr int R = (MODE == 0) ? RL::R : NODES * NODES;
  const int heads = A.heads;
  const int in_rec = PRE ? heads * IN_SUB : IN_SUB;
  const int out_rec = (A.merge == CGAT_MERGE_MEAN) ? OUT_SUB : heads * OUT_SUB;

  using SP = SmemParams<NODES, CI, CO>;
  SP& sp = *reinterpret_cast<SP*>(smem_raw);
  size_t off = (sizeof(SP) + 15) & ~size_t(15);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw + off);
  off += 128;
  float* red = reinterpret_cast<float*>(smem_raw + off);
  off += sizeof(float) * R * TILE;
  T* s_in = reinterpret_cast<T*>(smem_raw + off);
  T* s_dout = s_in + (size_t)TILE * in_rec;
  T* s_din = s_dout + (size_t)TILE * out_rec;  // MODE 0 only

  const int tid = threadIdx.x;
  const long long pix0 = (long long)blockIdx.x * TILE;
  const int npix = (int)min((long long)TILE, A.n_pix - pix0);

  if (tid == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  __syncthreads();
  if (tid == 0) {
    const uint32_t b_in = (uint32_t)npix * in_rec * sizeof(T);
    const uint32_t b_do = (uint32_t)npix * out_rec * sizeof(T);
    mbar_arrive_expect_tx(bar, b_in + b_do);
    bulk_g2s(s_in, reinterpret_cast<const T*>(A.in) + pix0 * in_rec, b_in, bar);
    bulk_g2s(s_dout, reinterpret_cast<const T*>(A.dout) + pix0 * out_rec, b_do, bar);
  }
  load_params(sp, A, !PRE, true);
  __syncthreads();
  mbar_wait(bar, 0);

  const bool active = tid < npix;
  const bool pixel_mode = A.stats != nullptr;
  const long long sample = active ? (pix0 + tid) / A.pix_per_sample : 0;
  // sample range covered by this tile (for the MODE 1 / pixel-mode reductions a tile may straddle samples)
  const long long s_first = pix0 / A.pix_per_sample;
  const long long s_last = (pix0 + npix - 1) / A.pix_per_sample;

  float X[NODES][PRE ? 1 : CI];
  float dX[NODES][PRE ? 1 : CI];
  if constexpr (!PRE) {
    if (active) {
      float r[NODES * CI];
      load_rec<NODES * CI, T>(s_in + (size_t)tid * in_rec, r);
      rec_to_mat<NODES, CI, SPATIAL>(r, X);
    } else {
#pragma unroll
      for (int n = 0; n < NODES; ++n)
#pragma unroll
        for (int c = 0; c < CI; ++c) X[n][c] = 0.f;
    }
#pragma unroll
    for (int n = 0; n < NODES; ++n)
#pragma unroll
      for (int c = 0; c < CI; ++c) dX[n][c] = 0.f;
  }
  const float gscale = (A.merge == CGAT_MERGE_MEAN) ? 1.f / (float)heads : 1.f;

  for (int k = 0; k < heads; ++k) {
    float* mycol = red + tid;  // red[r*TILE + tid]
    if (active) {
      float Wh[NODES][CO];
      if constexpr (PRE) {
        float r[NODES * CO];
        load_rec<NODES * CO, T>(s_in + (size_t)tid * in_rec + k * IN_SUB, r);
        rec_to_mat<NODES, CO, SPATIAL>(r, Wh);
      } else {
        project_linear<F32, NODES, CI, CO>(X, sp.W[k], Wh);
      }
      const float* st = pixel_mode ? A.stats + ((sample * heads + k) * 2) * (NODES * NODES) : nullptr;
      // ---- recompute z, then dz = dout * scale * ELU'(z) ----
      float z[NODES][CO];
#pragma unroll
      for (int v = 0; v < NODES; ++v)
#pragma unroll
        for (int u = 0; u < CO; ++u) z[v][u] = 0.f;
      if (pixel_mode)
        attn_forward_pixel<F32, NODES, CO, true>(Wh, sp.a[k], sp.adj[k], sp.maskrow, A.alpha, st, st + NODES * NODES, z);
      else
        attn_forward_pixel<F32, NODES, CO, false>(Wh, sp.a[k], sp.adj[k], sp.maskrow, A.alpha, nullptr, nullptr, z);
      {
        float dz[NODES][CO];
        if (A.merge == CGAT_MERGE_MEAN || SPATIAL) {
          float r[NODES * CO];
          const T* src = s_dout + (size_t)tid * out_rec + (A.merge == CGAT_MERGE_MEAN ? 0 : k * OUT_SUB);
          load_rec<NODES * CO, T>(src, r);
          rec_to_mat<NODES, CO, SPATIAL>(r, dz);
        } else {
          const T* o = s_dout + (size_t)tid * out_rec;
#pragma unroll
          for (int v = 0; v < NODES; ++v)
#pragma unroll
            for (int u = 0; u < CO; ++u) dz[v][u] = DT<T>::to_f(o[v * (heads * CO) + k * CO + u]);
        }
#pragma unroll
        for (int v = 0; v < NODES; ++v)
#pragma unroll
          for (int u = 0; u < CO; ++u)
            z[v][u] = dz[v][u] * gscale * (A.apply_elu ? elu_grad<F32>(z[v][u]) : 1.f);
      }
      // z now holds dz
      if constexpr (MODE == 1) {
        float dot[NODES * NODES];
#pragma unroll
        for (int i = 0; i < NODES * NODES; ++i) dot[i] = 0.f;
        float dWh_unused[NODES][CO];
        attn_backward_pixel<F32, NODES, CO, true, 1>(Wh, z, sp.a[k], sp.adj[k], sp.maskrow, A.alpha, st,
                                                st + NODES * NODES, nullptr, dWh_unused, nullptr, nullptr, dot);
#pragma unroll
        for (int i = 0; i < NODES * NODES; ++i) mycol[i * TILE] = dot[i];
      } else {
        float dWh[NODES][CO];
#pragma unroll
        for (int v = 0; v < NODES; ++v)
#pragma unroll
          for (int u = 0; u < CO; ++u) dWh[v][u] = 0.f;
        float g_a[2 * CO];
#pragma unroll
        for (int u = 0; u < 2 * CO; ++u) g_a[u] = 0.f;
        float g_adj[NODES * NODES];
#pragma unroll
        for (int i = 0; i < NODES * NODES; ++i) g_adj[i] = 0.f;
        if (pixel_mode) {
          const float* bs = A.bstats + (sample * heads + k) * (NODES * NODES);
          attn_backward_pixel<F32, NODES, CO, true, 0>(Wh, z, sp.a[k], sp.adj[k], sp.maskrow, A.alpha, st,
                                                  st + NODES * NODES, bs, dWh, g_a, g_adj, nullptr);
        } else {
          attn_backward_pixel<F32, NODES, CO, false, 0>(Wh, z, sp.a[k], sp.adj[k], sp.maskrow, A.alpha, nullptr,
                                                   nullptr, nullptr, dWh, g_a, g_adj, nullptr);
        }
#pragma unroll
        for (int i = 0; i < NODES * NODES; ++i) mycol[(RL::ADJ + i) * TILE] = g_adj[i];
#pragma unroll
        for (int u = 0; u < 2 * CO; ++u) mycol[(RL::AV + u) * TILE] = g_a[u];
        if constexpr (PRE) {
          float r[NODES * CO];
          mat_to_rec<NODES, CO, SPATIAL>(dWh, r);
          store_rec<NODES * CO, T>(s_din + (size_t)tid * in_rec + k * IN_SUB, r);
        } else {
          float g_W[CI * CO];
#pragma unroll
          for (int i = 0; i < CI * CO; ++i) g_W[i] = 0.f;
          project_linear_bwd<F32, NODES, CI, CO>(X, dWh, sp.W[k], dX, g_W);
#pragma unroll
          for (int i = 0; i < CI * CO; ++i) mycol[(RL::WV + i) * TILE] = g_W[i];
        }
      }
    } else {
#pragma unroll 4
      for (int r = 0; r < R; ++r) mycol[r * TILE] = 0.f;
    }
    __syncthreads();
    // ---- CTA reduction of the R partial sums, one atomicAdd per value ----
    {
      const int warp = tid >> 5, lane = tid & 31;
      if (MODE == 1 && s_first != s_last) {
        // tile straddles samples: per-thread atomics (rare: only when P % TILE != 0)
        if (active)
          for (int r = 0; r < R; ++r)
            atomicAdd(A.stats_out + (sample * heads + k) * (NODES * NODES) + r, red[r * TILE + tid]);
      } else {
        for (int r = warp; r < R; r += TILE / 32) {
          float s = 0.f;
#pragma unroll
          for (int q = 0; q < TILE / 32; ++q) s += red[r * TILE + lane + 32 * q];
          s = warp_sum(s);
          if (lane == 0) {
            if constexpr (MODE == 1) {
              atomicAdd(A.stats_out + (s_first * heads + k) * (NODES * NODES) + r, s);
            } else {
              if (r < RL::AV)
                atomicAdd(A.gadj + (size_t)k * NODES * NODES + r, s);
              else if (r < RL::WV)
                atomicAdd(A.ga + (size_t)k * 2 * CO + (r - RL::AV), s);
              else
                atomicAdd(A.gW + (size_t)k * CI * CO + (r - RL::WV), s);
            }
          }
        }
      }
    }
    __syncthreads();
  }

  if constexpr (MODE == 0) {
    if constexpr (!PRE) {
      if (active) {
        float r[NODES * CI];
        mat_to_rec<NODES, CI, SPATIAL>(dX, r);
        store_rec<NODES * CI, T>(s_din + (size_t)tid * in_rec, r);
      }
    }
    fence_proxy_async_smem();
    __syncthreads();
    if (tid == 0) {
      bulk_s2g(reinterpret_cast<T*>(A.out) + pix0 * in_rec, s_din, (uint32_t)npix * in_rec * sizeof(T));
      bulk_commit();
      bulk_wait_read0();
    }
  }
}

// ===================================================================================================
// Pixel-axis soft-max statistics (compat mode, baseline_model.py:131).  One CTA per (sample, head);
// every thread keeps an online (max, sum) pair per (i,j) over its pixels, then the CTA merges.
// stats [N][heads][2][NODES*NODES] = (max, 1/sum).
// ===================================================================================================
template <int NODES, int CI, int CO, bool SPATIAL, bool PRE, typename T>
__global__ void __launch_bounds__(256) attn_pixstats_kernel(const AttnArgs A) {
  constexpr int NN = NODES * NODES;
  constexpr int IN_SUB = PRE ? NODES * CO : NODES * CI;
  __shared__ SmemParams<NODES, CI, CO> sp;
  __shared__ float s_m[8][NN];
  __shared__ float s_s[8][NN];
  const int heads = A.heads;
  const int in_rec = PRE ? heads * IN_SUB : IN_SUB;
  const long long sample = blockIdx.x / heads;
  const int k = blockIdx.x % heads;
  load_params(sp, A, !PRE, false);
  __syncthreads();

  float m[NN], s[NN];
#pragma unroll
  for (int i = 0; i < NN; ++i) { m[i] = -INFINITY; s[i] = 0.f; }
  const T* base = reinterpret_cast<const T*>(A.in) + sample * A.pix_per_sample * in_rec;
  for (long long p = threadIdx.x; p < A.pix_per_sample; p += blockDim.x) {
    float Wh[NODES][CO];
    if constexpr (PRE) {
      float r[NODES * CO];
      load_rec<NODES * CO, T>(base + p * in_rec + k * IN_SUB, r);
      rec_to_mat<NODES, CO, SPATIAL>(r, Wh);
    } else {
      float r[NODES * CI];
      float X[NODES][CI];
      load_rec<NODES * CI, T>(base + p * in_rec, r);
      rec_to_mat<NODES, CI, SPATIAL>(r, X);
      project_linear<F32, NODES, CI, CO>(X, sp.W[k], Wh);
    }
    float e[NODES][NODES];
    attn_logits_pixel<NODES, CO>(Wh, sp.a[k], sp.maskrow, A.alpha, e);
#pragma unroll
    for (int i = 0; i < NODES; ++i)
#pragma unroll
      for (int j = 0; j < NODES; ++j) {
        const int q = i * NODES + j;
        const float v = e[i][j];
        const float nm = fmaxf(m[q], v);
        s[q] = s[q] * fast_exp(m[q] - nm) + fast_exp(v - nm);
        m[q] = nm;
      }
  }
  // merge within warp
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int q = 0; q < NN; ++q) {
    float mm = m[q], ss = s[q];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float om = __shfl_xor_sync(0xffffffffu, mm, o);
      const float os = __shfl_xor_sync(0xffffffffu, ss, o);
      const float nm = fmaxf(mm, om);
      const float sa = (mm == -INFINITY) ? 0.f : ss * fast_exp(mm - nm);
      const float sb = (om == -INFINITY) ? 0.f : os * fast_exp(om - nm);
      ss = sa + sb;
      mm = nm;
    }
    if (lane == 0) { s_m[warp][q] = mm; s_s[warp][q] = ss; }
  }
  __syncthreads();
  if (threadIdx.x < NN) {
    const int q = threadIdx.x;
    float mm = -INFINITY, ss = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) {
      const float om = s_m[w][q], os = s_s[w][q];
      const float nm = fmaxf(mm, om);
      const float sa = (mm == -INFINITY) ? 0.f : ss * fast_exp(mm - nm);
      const float sb = (om == -INFINITY) ? 0.f : os * fast_exp(om - nm);
      ss = sa + sb;
      mm = nm;
    }
    float* st = A.stats_out + ((sample * heads + k) * 2) * NN;
    st[q] = mm;
    st[NN + q] = 1.f / ss;
  }
}

// ===================================================================================================
// host-side dispatch
// ===================================================================================================
template <int NODES, int CI, int CO>
static size_t fwd_smem(int heads, bool pre, int merge, size_t esz) {
  const int in_rec = pre ? heads * NODES * CO : NODES * CI;
  const int out_rec = merge == CGAT_MERGE_MEAN ? NODES * CO : heads * NODES * CO;
  return ((sizeof(SmemParams<NODES, CI, CO>) + 15) & ~size_t(15)) + 128 + (size_t)TILE * (in_rec + out_rec) * esz;
}
template <int NODES, int CI, int CO>
static size_t bwd_smem(int heads, bool pre, int merge, size_t esz, int mode) {
  const int in_rec = pre ? heads * NODES * CO : NODES * CI;
  const int out_rec = merge == CGAT_MERGE_MEAN ? NODES * CO : heads * NODES * CO;
  const int R = mode == 0 ? (NODES * NODES + 2 * CO + (pre ? 0 : CI * CO)) : NODES * NODES;
  return ((sizeof(SmemParams<NODES, CI, CO>) + 15) & ~size_t(15)) + 128 + sizeof(float) * R * TILE +
         (size_t)TILE * (in_rec + out_rec + (mode == 0 ? in_rec : 0)) * esz;
}

static int validate(const cgat_attn_desc* d, const void* in, const float* a) {
  if (!d || !in || !a) return fail(CGAT_EINVAL, "null argument");
  if (d->n_pix <= 0 || d->pix_per_sample <= 0 || d->n_pix % d->pix_per_sample)
    return fail(CGAT_EINVAL, "n_pix=%lld must be a positive multiple of pix_per_sample=%lld", (long long)d->n_pix,
                (long long)d->pix_per_sample);
  if (d->heads < 1 || d->heads > MAX_HEADS) return fail(CGAT_EINVAL, "heads=%d out of range 1..%d", d->heads, MAX_HEADS);
  if (d->dtype != CGAT_F32 && d->dtype != CGAT_BF16) return fail(CGAT_EINVAL, "bad dtype %d", d->dtype);
  if (d->layout != CGAT_LAYOUT_SPATIAL && d->layout != CGAT_LAYOUT_TEMPORAL) return fail(CGAT_EINVAL, "bad layout");
  if (d->proj != CGAT_PROJ_LINEAR && d->proj != CGAT_PROJ_PRE) return fail(CGAT_EINVAL, "bad proj");
  if (d->merge != CGAT_MERGE_CONCAT && d->merge != CGAT_MERGE_MEAN) return fail(CGAT_EINVAL, "bad merge");
  if (!aligned16(in)) return fail(CGAT_EALIGN, "input pointer not 16-byte aligned");
  return 0;
}

using Op = AttnOp;

template <int NODES, int CI, int CO, bool SPATIAL, bool PRE, typename T>
static int launch(Op op, const cgat_attn_desc* d, const AttnArgs& A, cudaStream_t st) {
  const unsigned grid = (unsigned)((d->n_pix + TILE - 1) / TILE);
  if (op == OP_FWD) {
    auto kern = attn_fwd_kernel<NODES, CI, CO, SPATIAL, PRE, T>;
    const size_t smem = fwd_smem<NODES, CI, CO>(d->heads, PRE, d->merge, sizeof(T));
    if (smem > 227 * 1024) return fail(CGAT_EUNSUPPORTED, "forward tile needs %zu B of shared memory", smem);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return fail((int)e, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    kern<<<grid, TILE, smem, st>>>(A);
    return check_launch("attn_fwd_kernel");
  } else if (op == OP_BWD) {
    auto kern = attn_bwd_kernel<NODES, CI, CO, SPATIAL, PRE, T, 0>;
    const size_t smem = bwd_smem<NODES, CI, CO>(d->heads, PRE, d->merge, sizeof(T), 0);
    if (smem > 227 * 1024) return fail(CGAT_EUNSUPPORTED, "backward tile needs %zu B of shared memory", smem);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return fail((int)e, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    kern<<<grid, TILE, smem, st>>>(A);
    return check_launch("attn_bwd_kernel");
  } else if (op == OP_BSTATS) {
    auto kern = attn_bwd_kernel<NODES, CI, CO, SPATIAL, PRE, T, 1>;
    const size_t smem = bwd_smem<NODES, CI, CO>(d->heads, PRE, d->merge, sizeof(T), 1);
    if (smem > 227 * 1024) return fail(CGAT_EUNSUPPORTED, "backward tile needs %zu B of shared memory", smem);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return fail((int)e, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    kern<<<grid, TILE, smem, st>>>(A);
    return check_launch("attn_bwd_kernel<stats>");
  } else {
    const unsigned g = (unsigned)((d->n_pix / d->pix_per_sample) * d->heads);
    attn_pixstats_kernel<NODES, CI, CO, SPATIAL, PRE, T><<<g, 256, 0, st>>>(A);
    return check_launch("attn_pixstats_kernel");
  }
}

template <int NODES, int CI, int CO, bool SPATIAL>
static int dispatch2(Op op, const cgat_attn_desc* d, const AttnArgs& A, cudaStream_t st) {
  const bool pre = d->proj == CGAT_PROJ_PRE;
  if (d->dtype == CGAT_F32) {
    return pre ? launch<NODES, CI, CO, SPATIAL, true, float>(op, d, A, st)
               : launch<NODES, CI, CO, SPATIAL, false, float>(op, d, A, st);
  }
  return pre ? launch<NODES, CI, CO, SPATIAL, true, __nv_bfloat16>(op, d, A, st)
             : launch<NODES, CI, CO, SPATIAL, false, __nv_bfloat16>(op, d, A, st);
}

// Instantiated (nodes, ci, co, layout) combinations.  (6,4,4,spatial) and (4,6,6,temporal) are the
// reference's KNMI graph (V=6 regions, T=4 frames: kmni_dataset/__main__.py:49-56, kmni_data_loader.py:91-93);
// (8,4,4) / (4,8,8) cover an 8-node variant.  Other shapes are rejected with CGAT_EUNSUPPORTED.
static bool h2_enabled() {
  static int v = -1;
  if (v < 0) v = std::getenv("CGAT_NO_H2") ? 0 : 1;
  return v == 1;
}

static int dispatch(Op op, const cgat_attn_desc* d, const AttnArgs& A, cudaStream_t st) {
  if (d->dtype == CGAT_BF16 && A.stats == nullptr && (op == OP_FWD || op == OP_BWD) && h2_enabled()) {
    const int rc = attn_h2_launch(op, d, A, st);
    if (rc != CGAT_EUNSUPPORTED) return rc;
  }
  const bool sp = d->layout == CGAT_LAYOUT_SPATIAL;
  const int ci = d->proj == CGAT_PROJ_PRE ? d->co : d->ci;
  if (sp && d->nodes == 6 && ci == 4 && d->co == 4) return dispatch2<6, 4, 4, true>(op, d, A, st);
  if (!sp && d->nodes == 4 && ci == 6 && d->co == 6) return dispatch2<4, 6, 6, false>(op, d, A, st);
  if (sp && d->nodes == 8 && ci == 4 && d->co == 4) return dispatch2<8, 4, 4, true>(op, d, A, st);
  if (!sp && d->nodes == 4 && ci == 8 && d->co == 8) return dispatch2<4, 8, 8, false>(op, d, A, st);
  {
    const int rc = attn_generic_launch(op, d, A, st);  // any node count <= 64, neighbour soft-max
    if (rc != CGAT_EUNSUPPORTED) return rc;
  }
  return fail(CGAT_EUNSUPPORTED,
              "attention kernel not instantiated for nodes=%d ci=%d co=%d layout=%d (register-resident kernels: spatial "
              "6/4/4, 8/4/4; temporal 4/6/6, 4/8/8; generic kernel: nodes <= 64, channels <= 8, neighbour soft-max)",
              d->nodes, ci, d->co, d->layout);
}

}  // namespace cgat

using namespace cgat;

extern "C" int cgat_attn_fwd(const cgat_attn_desc* d, const void* in, void* out, const float* W, const float* a,
                             const float* adj, const uint8_t* mask, const float* stats, void* stream) {
  if (int rc = validate(d, in, a)) return rc;
  if (!out || !adj) return fail(CGAT_EINVAL, "null out/adj");
  if (d->proj == CGAT_PROJ_LINEAR && !W) return fail(CGAT_EINVAL, "W required for linear projection");
  if (!aligned16(out)) return fail(CGAT_EALIGN, "output pointer not 16-byte aligned");
  AttnArgs A{};
  A.in = in; A.out = out; A.W = W; A.a = a; A.adj = adj; A.mask = mask; A.stats = stats;
  A.n_pix = d->n_pix; A.pix_per_sample = d->pix_per_sample; A.heads = d->heads; A.merge = d->merge;
  A.apply_elu = d->apply_elu; A.alpha = d->alpha;
  return dispatch(OP_FWD, d, A, (cudaStream_t)stream);
}

extern "C" int cgat_attn_pixstats(const cgat_attn_desc* d, const void* in, const float* W, const float* a,
                                  const uint8_t* mask, float* stats, void* stream) {
  if (int rc = validate(d, in, a)) return rc;
  if (!stats) return fail(CGAT_EINVAL, "null stats");
  if (d->proj == CGAT_PROJ_LINEAR && !W) return fail(CGAT_EINVAL, "W required for linear projection");
  AttnArgs A{};
  A.in = in; A.W = W; A.a = a; A.mask = mask; A.stats_out = stats;
  A.n_pix = d->n_pix; A.pix_per_sample = d->pix_per_sample; A.heads = d->heads; A.merge = d->merge;
  A.apply_elu = d->apply_elu; A.alpha = d->alpha;
  return dispatch(OP_STATS, d, A, (cudaStream_t)stream);
}

extern "C" int cgat_attn_bwd(const cgat_attn_desc* d, const void* in, const void* dout, void* din, const float* W,
                             const float* a, const float* adj, const uint8_t* mask, const float* stats,
                             const float* bstats, float* gW, float* ga, float* gadj, void* stream) {
  if (int rc = validate(d, in, a)) return rc;
  if (!dout || !din || !adj || !ga || !gadj) return fail(CGAT_EINVAL, "null dout/din/adj/ga/gadj");
  if (d->proj == CGAT_PROJ_LINEAR && (!W || !gW)) return fail(CGAT_EINVAL, "W and gW required for linear projection");
  if (!aligned16(dout) || !aligned16(din)) return fail(CGAT_EALIGN, "dout/din pointer not 16-byte aligned");
  if ((stats == nullptr) != (bstats == nullptr)) return fail(CGAT_EINVAL, "stats and bstats must be given together");
  AttnArgs A{};
  A.in = in; A.out = din; A.dout = dout; A.W = W; A.a = a; A.adj = adj; A.mask = mask; A.stats = stats;
  A.bstats = bstats; A.gW = gW; A.ga = ga; A.gadj = gadj;
  A.n_pix = d->n_pix; A.pix_per_sample = d->pix_per_sample; A.heads = d->heads; A.merge = d->merge;
  A.apply_elu = d->apply_elu; A.alpha = d->alpha;
  return dispatch(OP_BWD, d, A, (cudaStream_t)stream);
}

extern "C" int cgat_attn_pixstats_bwd(const cgat_attn_desc* d, const void* in, const void* dout, const float* W,
                                      const float* a, const float* adj, const uint8_t* mask, const float* stats,
                                      float* bstats, void* stream) {
  if (int rc = validate(d, in, a)) return rc;
  if (!dout || !adj || !stats || !bstats) return fail(CGAT_EINVAL, "null dout/adj/stats/bstats");
  if (d->proj == CGAT_PROJ_LINEAR && !W) return fail(CGAT_EINVAL, "W required for linear projection");
  if (!aligned16(dout)) return fail(CGAT_EALIGN, "dout pointer not 16-byte aligned");
  AttnArgs A{};
  A.in = in; A.dout = dout; A.W = W; A.a = a; A.adj = adj; A.mask = mask; A.stats = stats; A.stats_out = bstats;
  A.n_pix = d->n_pix; A.pix_per_sample = d->pix_per_sample; A.heads = d->heads; A.merge = d->merge;
  A.apply_elu = d->apply_elu; A.alpha = d->alpha;
  return dispatch(OP_BSTATS, d, A, (cudaStream_t)stream);
}
