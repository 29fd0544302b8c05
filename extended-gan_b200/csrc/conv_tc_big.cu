// tcgen05 implicit-GEMM convolution for LARGE channel counts (sm_100a): operands streamed by TMA.
//
// conv_tc.cu keeps the whole packed weight matrix resident in shared memory, which serves the conv-GAT node
// conv (24 -> 72 channels) but not the dense convs of the stacks around the layer: the DCGAN discriminators
// (dcgan/model.py:150-165: 64->128->256->512, k=4 s=2, taken here as stride-1 2x2 convs over 2x2-regrouped
// input, i.e. 256/512/1024 K-channels per tap) and the SmaAt-UNet pointwise convs (unet_model.py:20).  Those
// are plain tensor-bound contractions, so this file is a classic streamed GEMM pipeline:
//
//   fprop / dgrad   D[128 pixels][BN couts] += A[128 px][64 ch] . B[BN][64 ch]^T   per (tap, 64-channel block)
//     A: one 4-D TMA box (64 ch, tw, th, tn) of the NHWC input at the tap-shifted coordinates (out-of-bounds =
//        zero = the conv padding, so there is no im2col and no halo logic), 128-byte swizzle, K-major.
//     B: one 3-D TMA box (64 ch, 1 tap, BN) of the [cout][tap][cin] weights (fprop: the KRSC tensor as it is;
//        dgrad: a rotated + transposed copy in the workspace), 128-byte swizzle, K-major.
//   wgrad           dW[128 couts][BN cins] += dY[64 px][128 co]^T . X_tap[64 px][BN ci]  per pixel block
//     both operands MN-major (the contraction runs over pixels, memory is channel-contiguous), the same
//     4-D boxes; split-K over pixel blocks across CTAs, fp32 partial sums reduced in a fixed order.
//
// 4-stage mbarrier pipeline, one TMA warp, one MMA-issuing thread, four epilogue warps; two TMEM accumulators
// so the epilogue of tile i overlaps the MMAs of tile i+1; persistent CTAs, one per SM.
#include <cstdlib>
#include "tc_common.cuh"

namespace cgat {

constexpr int BIG_THREADS = 192;  // warp 0 TMA, warp 1 MMA (+ TMEM alloc), warps 2-5 epilogue
constexpr int BIG_STAGES = 4;
constexpr int BIG_KC = 64;        // bf16 elements of one 128-byte swizzle row

// shared-memory matrix descriptor, SWIZZLE_128B canonical layouts (tiles are 1024-byte aligned):
//   K-major : rows of 128 B (64 elements of K), 8-row groups SBO = 1024 B apart; LBO unused; a K=16 step inside
//             the swizzle row advances the start address by 32 B.
//   MN-major: rows of 128 B (64 elements of M/N) per K index, 8 K-rows per 1024-byte atom, atoms SBO = 1024 B
//             apart along K, 64-element M/N blocks LBO apart.
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return make_smem_desc(smem_addr, lbo_bytes, sbo_bytes) | ((uint64_t)2 << 61);
}

struct BigArgs {
  int n, ho, wo;                 // pixel grid the tiles cover (fprop/dgrad: output; wgrad: dY)
  int gk, gn;                    // fprop/dgrad: K channels per tap, N channels.  wgrad: gk = cout (M), gn = cin (N)
  int kh, kw, pad_t, pad_l;
  int tw, th, tn;                // pixel tile (tw*th*tn = 128 fprop, 64 wgrad)
  int tiles_w, tiles_h, tiles_n;
  int bn, n_tiles, m_tiles, kchunks;
  int splits;                    // wgrad: split-K factor
  int act;
  uint32_t tmem_cols;
  const float* bias;
  void* out;                     // fprop/dgrad: bf16 [n][ho][wo][gn]; wgrad: fp32 [splits][cout][taps][cin]
};

struct BigSmem {
  uint64_t full[BIG_STAGES], empty[BIG_STAGES], tfull[2], tempty[2];
  uint32_t tmem_slot;
};

__device__ __forceinline__ uint8_t* big_align_smem(uint8_t* raw) {
  const uint32_t a = smem_u32(raw);
  return raw + ((1024u - (a & 1023u)) & 1023u);
}

__device__ __forceinline__ float big_act(float v, int act) {
  if (act == 1) return fmaxf(v, 0.f);
  if (act == 2) return v > 0.f ? v : 0.2f * v;
  if (act == 3) return 1.f / (1.f + __expf(-v));
  return v;
}

// ---------------------------------------------------------------------------------------------------------
// fprop (and dgrad through rotated/transposed weights)
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(BIG_THREADS, 1)
conv_big_fprop_kernel(const __grid_constant__ CUtensorMap amap, const __grid_constant__ CUtensorMap bmap,
                      const BigArgs A) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = big_align_smem(smem_raw);
  const uint32_t a_bytes = 128u * 128u, b_bytes = (uint32_t)A.bn * 128u, stage_bytes = a_bytes + b_bytes;
  BigSmem* S = reinterpret_cast<BigSmem*>(smem + BIG_STAGES * stage_bytes);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < BIG_STAGES; ++s) {
      mbar_init(&S->full[s], 1);
      mbar_init(&S->empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&S->tfull[b], 1);
      mbar_init(&S->tempty[b], 128);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(&S->tmem_slot, A.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = S->tmem_slot;

  const int total = A.m_tiles * A.n_tiles;
  const int taps = A.kh * A.kw;
  const int ktotal = taps * A.kchunks;

  if (warp == 0) {
    if (lane == 0) {
      tma_prefetch_desc(&amap);
      tma_prefetch_desc(&bmap);
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
        const int mt = tile / A.n_tiles, nt = tile - mt * A.n_tiles;
        const int iw = mt % A.tiles_w, ih = (mt / A.tiles_w) % A.tiles_h, in = mt / (A.tiles_w * A.tiles_h);
        for (int tap = 0; tap < taps; ++tap) {
          const int r = tap / A.kw, s = tap - r * A.kw;
          for (int kc = 0; kc < A.kchunks; ++kc, ++it) {
            const uint32_t st = it % BIG_STAGES, ph = (it / BIG_STAGES) & 1u;
            mbar_wait(&S->empty[st], ph ^ 1u);
            mbar_arrive_expect_tx(&S->full[st], stage_bytes);
            uint8_t* sa = smem + st * stage_bytes;
            tma_load_4d(sa, &amap, kc * BIG_KC, iw * A.tw - A.pad_l + s, ih * A.th - A.pad_t + r, in * A.tn,
                        &S->full[st]);
            tma_load_3d(sa + a_bytes, &bmap, kc * BIG_KC, tap, nt * A.bn, &S->full[st]);
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = make_idesc_bf16(128, A.bn, 0, 0);
      uint32_t it = 0, lt = 0;
      for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++lt) {
        const uint32_t buf = lt & 1u;
        mbar_wait(&S->tempty[buf], ((lt >> 1) & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + buf * (uint32_t)A.bn;
        for (int kb = 0; kb < ktotal; ++kb, ++it) {
          const uint32_t st = it % BIG_STAGES, ph = (it / BIG_STAGES) & 1u;
          mbar_wait(&S->full[st], ph);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + st * stage_bytes), b_addr = a_addr + a_bytes;
#pragma unroll
          for (int k = 0; k < BIG_KC / 16; ++k)
            umma_bf16(d_tmem, make_desc_sw128(a_addr + k * 32, 16, 1024), make_desc_sw128(b_addr + k * 32, 16, 1024),
                      idesc, (kb | k) != 0);
          umma_commit(&S->empty[st]);
        }
        umma_commit(&S->tfull[buf]);
      }
    }
  } else {
    const int q = warp & 3;  // TMEM lane quarter this warp may read
    const int row = q * 32 + lane;
    __nv_bfloat16* y = reinterpret_cast<__nv_bfloat16*>(A.out);
    uint32_t lt = 0;
    for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++lt) {
      const int mt = tile / A.n_tiles, nt = tile - mt * A.n_tiles;
      const int iw = mt % A.tiles_w, ih = (mt / A.tiles_w) % A.tiles_h, in = mt / (A.tiles_w * A.tiles_h);
      const int pw = iw * A.tw + row % A.tw, ph_ = ih * A.th + (row / A.tw) % A.th, pn = in * A.tn + row / (A.tw * A.th);
      const bool valid = pn < A.n && ph_ < A.ho && pw < A.wo;
      const uint32_t buf = lt & 1u;
      mbar_wait(&S->tfull[buf], (lt >> 1) & 1u);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + buf * (uint32_t)A.bn;
      __nv_bfloat16* yrow = y + ((size_t)(pn * A.ho + ph_) * A.wo + pw) * A.gn;
      for (int c0 = 0; c0 < A.bn; c0 += 16) {
        float v[16];
        tmem_ld16(taddr + c0, v);
        const int col = nt * A.bn + c0;
        if (valid && col < A.gn) {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            if (col + 8 * h < A.gn) {  // gn % 8 == 0
              uint32_t pk[4];
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const int c = col + 8 * h + 2 * j;
                float f0 = v[8 * h + 2 * j], f1 = v[8 * h + 2 * j + 1];
                if (A.bias) {
                  f0 += __ldg(A.bias + c);
                  f1 += __ldg(A.bias + c + 1);
                }
                __nv_bfloat162 p = __floats2bfloat162_rn(big_act(f0, A.act), big_act(f1, A.act));
                pk[j] = *reinterpret_cast<uint32_t*>(&p);
              }
              *reinterpret_cast<uint4*>(yrow + col + 8 * h) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&S->tempty[buf]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, A.tmem_cols);
}

// ---------------------------------------------------------------------------------------------------------
// wgrad: work item = (split, tap, cout tile of 128, cin tile of bn); K = the split's pixel blocks of 64
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(BIG_THREADS, 1)
conv_big_wgrad_kernel(const __grid_constant__ CUtensorMap ymap, const __grid_constant__ CUtensorMap xmap,
                      const BigArgs A) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = big_align_smem(smem_raw);
  const int nb = A.bn / 64;  // 64-channel boxes of the X operand
  const uint32_t box_bytes = 64u * 128u, a_bytes = 2u * box_bytes, stage_bytes = a_bytes + (uint32_t)nb * box_bytes;
  BigSmem* S = reinterpret_cast<BigSmem*>(smem + BIG_STAGES * stage_bytes);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < BIG_STAGES; ++s) {
      mbar_init(&S->full[s], 1);
      mbar_init(&S->empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&S->tfull[b], 1);
      mbar_init(&S->tempty[b], 128);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(&S->tmem_slot, A.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = S->tmem_slot;

  const int taps = A.kh * A.kw;
  const int ot = taps * A.m_tiles * A.n_tiles;
  const int total = ot * A.splits;
  const int kblocks = A.tiles_w * A.tiles_h * A.tiles_n;

  if (warp == 0) {
    if (lane == 0) {
      tma_prefetch_desc(&ymap);
      tma_prefetch_desc(&xmap);
      uint32_t it = 0;
      for (int item = blockIdx.x; item < total; item += gridDim.x) {
        const int split = item / ot, o = item - split * ot;
        const int nt = o % A.n_tiles, mt = (o / A.n_tiles) % A.m_tiles, tap = o / (A.n_tiles * A.m_tiles);
        const int r = tap / A.kw, s = tap - r * A.kw;
        const int kb0 = (int)((long long)kblocks * split / A.splits), kb1 = (int)((long long)kblocks * (split + 1) / A.splits);
        for (int kb = kb0; kb < kb1; ++kb, ++it) {
          const int iw = kb % A.tiles_w, ih = (kb / A.tiles_w) % A.tiles_h, in = kb / (A.tiles_w * A.tiles_h);
          const uint32_t st = it % BIG_STAGES, ph = (it / BIG_STAGES) & 1u;
          mbar_wait(&S->empty[st], ph ^ 1u);
          // rows 64..127 of the accumulator are never stored when the cout tile ends before them: their dY box is
          // not fetched (rows of D are independent, whatever that part of shared memory holds)
          const int na = A.gk - mt * 128 > 64 ? 2 : 1;
          mbar_arrive_expect_tx(&S->full[st], (uint32_t)(na + nb) * box_bytes);
          uint8_t* sa = smem + st * stage_bytes;
          for (int b = 0; b < na; ++b)
            tma_load_4d(sa + b * box_bytes, &ymap, mt * 128 + b * 64, iw * A.tw, ih * A.th, in * A.tn, &S->full[st]);
          for (int b = 0; b < nb; ++b)
            tma_load_4d(sa + a_bytes + b * box_bytes, &xmap, nt * A.bn + b * 64, iw * A.tw - A.pad_l + s,
                        ih * A.th - A.pad_t + r, in * A.tn, &S->full[st]);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = make_idesc_bf16(128, A.bn, 1, 1);
      uint32_t it = 0, lt = 0;
      for (int item = blockIdx.x; item < total; item += gridDim.x, ++lt) {
        const int split = item / ot;
        const int kb0 = (int)((long long)kblocks * split / A.splits), kb1 = (int)((long long)kblocks * (split + 1) / A.splits);
        const uint32_t buf = lt & 1u;
        mbar_wait(&S->tempty[buf], ((lt >> 1) & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + buf * (uint32_t)A.bn;
        for (int kb = kb0; kb < kb1; ++kb, ++it) {
          const uint32_t st = it % BIG_STAGES, ph = (it / BIG_STAGES) & 1u;
          mbar_wait(&S->full[st], ph);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + st * stage_bytes), b_addr = a_addr + a_bytes;
#pragma unroll
          for (int k = 0; k < 4; ++k)  // 16 pixels = two 8-row swizzle atoms per instruction
            umma_bf16(d_tmem, make_desc_sw128(a_addr + k * 2048, box_bytes, 1024),
                      make_desc_sw128(b_addr + k * 2048, box_bytes, 1024), idesc, (kb > kb0) || k != 0);
          umma_commit(&S->empty[st]);
        }
        umma_commit(&S->tfull[buf]);
      }
    }
  } else {
    const int q = warp & 3;
    const int row = q * 32 + lane;
    float* out = reinterpret_cast<float*>(A.out);
    uint32_t lt = 0;
    for (int item = blockIdx.x; item < total; item += gridDim.x, ++lt) {
      const int split = item / ot, o = item - split * ot;
      const int nt = o % A.n_tiles, mt = (o / A.n_tiles) % A.m_tiles, tap = o / (A.n_tiles * A.m_tiles);
      const int kb0 = (int)((long long)kblocks * split / A.splits), kb1 = (int)((long long)kblocks * (split + 1) / A.splits);
      const int co = mt * 128 + row;
      const uint32_t buf = lt & 1u;
      mbar_wait(&S->tfull[buf], (lt >> 1) & 1u);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + buf * (uint32_t)A.bn;
      float* orow = out + (((size_t)split * A.gk + co) * taps + tap) * A.gn;
      for (int c0 = 0; c0 < A.bn; c0 += 16) {
        float v[16];
        tmem_ld16(taddr + c0, v);
        if (kb1 == kb0) {  // empty split (more splits than pixel blocks): no MMA ran, the accumulator is stale
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = 0.f;
        }
        const int ci = nt * A.bn + c0;
        if (co < A.gk) {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (ci + 4 * j < A.gn)  // gn % 8 == 0
              *reinterpret_cast<float4*>(orow + ci + 4 * j) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        }
      }
      tc_fence_before();
      mbar_arrive(&S->tempty[buf]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, A.tmem_cols);
}

// sums the split-K partials in a fixed order: dw[i] = sum_s partial[s][i]
__global__ void __launch_bounds__(256) big_wgrad_reduce_kernel(const float4* __restrict__ partial, float4* __restrict__ dw,
                                                               long long n4, int splits) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 a = partial[i];
    for (int s = 1; s < splits; ++s) {
      const float4 b = partial[(long long)s * n4 + i];
      a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
    }
    dw[i] = a;
  }
}

// dgrad weights: wp[ci][r*kw+s][co] = w[co][kh-1-r][kw-1-s][ci]   (rotated taps, channels transposed: the same
// [N][tap][K] order the fprop map reads)
__global__ void __launch_bounds__(256) big_pack_dgrad_kernel(const __nv_bfloat16* __restrict__ w,
                                                             __nv_bfloat16* __restrict__ wp, int cout, int cin, int kh,
                                                             int kw) {
  __shared__ __nv_bfloat16 t[32][33];
  // one (tap, 32x32 channel block) per block: coalesced read along ci, coalesced write along co
  const int taps = kh * kw;
  const int cb = (cin + 31) / 32, ob = (cout + 31) / 32;
  const long long nblk = (long long)taps * cb * ob;
  for (long long b = blockIdx.x; b < nblk; b += gridDim.x) {
    const int tap = (int)(b / (cb * ob)), rem = (int)(b % (cb * ob));
    const int co0 = (rem / cb) * 32, ci0 = (rem % cb) * 32;
    const int src_tap = taps - 1 - tap;  // (kh-1-r)*kw + (kw-1-s)
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int j = ty; j < 32; j += 8) {
      const int co = co0 + j, ci = ci0 + tx;
      t[j][tx] = (co < cout && ci < cin) ? w[((size_t)co * taps + src_tap) * cin + ci] : __float2bfloat16(0.f);
    }
    __syncthreads();
    for (int j = ty; j < 32; j += 8) {
      const int ci = ci0 + j, co = co0 + tx;
      if (ci < cin && co < cout) wp[((size_t)ci * taps + tap) * cout + co] = t[tx][j];
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------------------
// host
// ---------------------------------------------------------------------------------------------------------
static int big_sm_count() {
  static int num_sms = 0;
  if (!num_sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
  }
  return num_sms;
}

// both kernels opt in once (per device) to the largest dynamic shared-memory size any of their shapes uses
constexpr size_t BIG_MAX_SMEM = (size_t)BIG_STAGES * (128 * 128 + 256 * 128) + sizeof(BigSmem) + 1024;
static int big_set_smem(const void* kernel, int slot) {
  static int done[2][64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 64 && done[slot][dev]) return 0;
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BIG_MAX_SMEM);
  if (e != cudaSuccess) return fail((int)e, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
  if (dev < 64) done[slot][dev] = 1;
  return 0;
}

static int pow2_floor(int v) {
  int p = 1;
  while (p * 2 <= v) p *= 2;
  return p;
}

static int make_map(CUtensorMap* map, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides,
                    const cuuint32_t* box) {
  EncodeTiledFn enc = get_encode_tiled();
  if (!enc) return fail(CGAT_EUNSUPPORTED, "cuTensorMapEncodeTiled not available from the driver");
  ensure_context();
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), dims, strides, box,
                   estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(CGAT_EINVAL, "cuTensorMapEncodeTiled (swizzle 128B) failed with CUresult %d", (int)r);
  return 0;
}

// 4-D map over an NHWC bf16 tensor, box (64 channels, tw, th, tn)
static int make_nhwc_map(CUtensorMap* map, const void* base, int n, int h, int w, int c, int tw, int th, int tn) {
  cuuint64_t dims[4] = {(cuuint64_t)c, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)n};
  cuuint64_t strides[3] = {(cuuint64_t)c * 2, (cuuint64_t)w * c * 2, (cuuint64_t)h * w * c * 2};
  cuuint32_t box[4] = {BIG_KC, (cuuint32_t)tw, (cuuint32_t)th, (cuuint32_t)tn};
  return make_map(map, base, 4, dims, strides, box);
}

struct BigGeom {
  int tw, th, tn, tiles_w, tiles_h, tiles_n;
};
static BigGeom pixel_tiles(int n, int ho, int wo, int pixels) {
  BigGeom g;
  g.tw = pow2_floor(wo) < 16 ? pow2_floor(wo) : 16;
  const int th_cap = pixels / g.tw;
  g.th = pow2_floor(ho) < th_cap ? pow2_floor(ho) : th_cap;
  g.tn = pixels / (g.tw * g.th);
  g.tiles_w = (wo + g.tw - 1) / g.tw;
  g.tiles_h = (ho + g.th - 1) / g.th;
  g.tiles_n = (n + g.tn - 1) / g.tn;
  return g;
}

static uint32_t tmem_cols_for(int bn) {
  uint32_t c = 32;
  while ((int)c < 2 * bn) c *= 2;
  return c;
}

static int round_up(int v, int m) { return (v + m - 1) / m * m; }

// which: 0 fprop, 1 dgrad, 2 wgrad.  Served: bf16, stride 1, dense, channels % 8 == 0 (TMA strides), and enough
// channels on the contraction / tile axes for 64-wide blocks to pay.
int conv_big_supported(const cgat_conv_desc* d, int which) {
  if (d->dtype != CGAT_BF16 || d->stride != 1 || d->groups != 1) return 0;
  if (d->cin % 8 || d->cout % 8) return 0;
  if (d->kh > 16 || d->kw > 16) return 0;
  if (which == 0) return d->cin >= 64;
  if (which == 1) return d->cout >= 64 && d->kh - 1 - d->pad_top >= 0 && d->kw - 1 - d->pad_left >= 0;
  return d->cin >= 64 && d->cout >= 64;
}

static int wgrad_splits(const cgat_conv_desc* d, int* bn_out) {
  const int bn = round_up(d->cin, 64) < 256 ? round_up(d->cin, 64) : 256;
  const BigGeom g = pixel_tiles(d->n, d->ho, d->wo, 64);
  const int kblocks = g.tiles_w * g.tiles_h * g.tiles_n;
  const int ot = d->kh * d->kw * ((d->cout + 127) / 128) * ((d->cin + bn - 1) / bn);
  static const int waves = getenv("CGAT_BIG_WAVES") ? atoi(getenv("CGAT_BIG_WAVES")) : 1;  // developer knob
  int splits = (waves * big_sm_count()) / ot;  // one wave of work items (measured: two waves cost 7-35 % more)
  if (splits > kblocks / 2) splits = kblocks / 2;  // at least two pixel blocks per item
  if (splits < 1) splits = 1;
  if (splits > 64) splits = 64;
  if (bn_out) *bn_out = bn;
  return splits;
}

// wgrad shapes below the 64-channel bar that the kernel still serves correctly (zero-filled channel blocks): used
// where the resident-weight wgrad of conv_tc.cu does not take the shape (the DCGAN generator's k=4 convs)
int conv_big_wgrad_small_ok(const cgat_conv_desc* d) {
  if (d->dtype != CGAT_BF16 || d->stride != 1 || d->groups != 1) return 0;
  if (d->cin % 8 || d->cout % 8 || d->kh > 16 || d->kw > 16) return 0;
  return d->cin >= 16 && d->cout >= 8;
}

size_t conv_dbias_workspace(const cgat_conv_desc* d);
size_t conv_big_workspace(const cgat_conv_desc* d, int which) {
  if (!conv_big_supported(d, which) && !(which == 2 && conv_big_wgrad_small_ok(d))) return 0;
  if (which == 0) return 0;
  if (which == 1) return (size_t)d->kh * d->kw * d->cin * d->cout * 2;
  const int splits = wgrad_splits(d, nullptr);
  const size_t part = splits > 1 ? (size_t)splits * d->cout * d->kh * d->kw * d->cin * 4 : 0;
  const size_t db = conv_dbias_workspace(d);  // the bias gradient's partial rows reuse the workspace after the wgrad reduction
  return part > db ? part : db;
}

// input [n][hi][wi][gk] -> output [n][hout][wout][gn]; weights [gn][taps][gk] bf16
static int launch_big_fprop(const void* in, int n, int hi, int wi, int gk, const void* w, int kh, int kw, int pad_t,
                            int pad_l, int hout, int wout, int gn, const float* bias, int act, void* out,
                            cudaStream_t st) {
  if (!aligned16(in) || !aligned16(out) || !aligned16(w)) return fail(CGAT_EALIGN, "conv tensors must be 16-byte aligned");
  const BigGeom g = pixel_tiles(n, hout, wout, 128);
  BigArgs A{};
  A.n = n; A.ho = hout; A.wo = wout; A.gk = gk; A.gn = gn;
  A.kh = kh; A.kw = kw; A.pad_t = pad_t; A.pad_l = pad_l;
  A.tw = g.tw; A.th = g.th; A.tn = g.tn; A.tiles_w = g.tiles_w; A.tiles_h = g.tiles_h; A.tiles_n = g.tiles_n;
  static const int bn_cap = getenv("CGAT_BIG_BN") ? atoi(getenv("CGAT_BIG_BN")) : 256;  // developer knob
  A.bn = round_up(gn, 16) < bn_cap ? round_up(gn, 16) : bn_cap;
  A.m_tiles = g.tiles_w * g.tiles_h * g.tiles_n;
  // small problems: narrower cout tiles until at least half of the SMs have a tile (measured on the DCGAN
  // discriminator convs at N = 64: conv3 21.6 -> 15.4 us, conv4 30 -> 23.7 us)
  while (A.bn > 64 && A.bn % 32 == 0 && A.m_tiles * ((gn + A.bn - 1) / A.bn) < big_sm_count() / 2) A.bn /= 2;
  A.n_tiles = (gn + A.bn - 1) / A.bn;
  A.kchunks = (gk + BIG_KC - 1) / BIG_KC;
  A.splits = 1;
  A.act = act;
  A.tmem_cols = tmem_cols_for(A.bn);
  A.bias = bias;
  A.out = out;
  CUtensorMap amap, bmap;
  if (int rc = make_nhwc_map(&amap, in, n, hi, wi, gk, g.tw, g.th, g.tn)) return rc;
  {
    cuuint64_t dims[3] = {(cuuint64_t)gk, (cuuint64_t)kh * kw, (cuuint64_t)gn};
    cuuint64_t strides[2] = {(cuuint64_t)gk * 2, (cuuint64_t)kh * kw * gk * 2};
    // box order follows the dims: (64 channels, 1 tap, bn couts) -> bn rows of 128 B
    cuuint32_t box[3] = {BIG_KC, 1, (cuuint32_t)A.bn};
    if (int rc = make_map(&bmap, w, 3, dims, strides, box)) return rc;
  }
  const size_t smem = (size_t)BIG_STAGES * (128 * 128 + A.bn * 128) + sizeof(BigSmem) + 1024;
  if (int rc = big_set_smem((const void*)conv_big_fprop_kernel, 0)) return rc;
  const int total = A.m_tiles * A.n_tiles;
  const int grid = total < big_sm_count() ? total : big_sm_count();
  conv_big_fprop_kernel<<<grid, BIG_THREADS, smem, st>>>(amap, bmap, A);
  return check_launch("conv_big_fprop_kernel");
}

int conv_big_fprop_launch(const cgat_conv_desc* d, const void* x, const void* w, const float* bias, void* y,
                          cudaStream_t st) {
  return launch_big_fprop(x, d->n, d->h, d->w, d->cin, w, d->kh, d->kw, d->pad_top, d->pad_left, d->ho, d->wo, d->cout,
                          bias, d->act, y, st);
}

int conv_big_dgrad_launch(const cgat_conv_desc* d, const void* dy, const void* w, void* dx, void* workspace,
                          cudaStream_t st) {
  if (!workspace || !aligned16(workspace)) return fail(CGAT_EINVAL, "dgrad needs a 16-byte aligned workspace");
  const long long nblk = (long long)d->kh * d->kw * ((d->cin + 31) / 32) * ((d->cout + 31) / 32);
  const int blocks = (int)(nblk < 148 * 16 ? nblk : 148 * 16);
  big_pack_dgrad_kernel<<<blocks, 256, 0, st>>>((const __nv_bfloat16*)w, (__nv_bfloat16*)workspace, d->cout, d->cin,
                                                d->kh, d->kw);
  if (int rc = check_launch("big_pack_dgrad_kernel")) return rc;
  // dx = conv(dy, rot180(w)^T) with leading padding k-1-pad
  return launch_big_fprop(dy, d->n, d->ho, d->wo, d->cout, workspace, d->kh, d->kw, d->kh - 1 - d->pad_top,
                          d->kw - 1 - d->pad_left, d->h, d->w, d->cin, nullptr, 0, dx, st);
}

int conv_big_wgrad_launch(const cgat_conv_desc* d, const void* x, const void* dy, float* dw, void* workspace,
                          cudaStream_t st) {
  if (!aligned16(x) || !aligned16(dy) || !aligned16(dw)) return fail(CGAT_EALIGN, "conv tensors must be 16-byte aligned");
  int bn = 0;
  const int splits = wgrad_splits(d, &bn);
  if (splits > 1 && (!workspace || !aligned16(workspace))) return fail(CGAT_EINVAL, "wgrad needs a 16-byte aligned workspace");
  const BigGeom g = pixel_tiles(d->n, d->ho, d->wo, 64);
  BigArgs A{};
  A.n = d->n; A.ho = d->ho; A.wo = d->wo; A.gk = d->cout; A.gn = d->cin;
  A.kh = d->kh; A.kw = d->kw; A.pad_t = d->pad_top; A.pad_l = d->pad_left;
  A.tw = g.tw; A.th = g.th; A.tn = g.tn; A.tiles_w = g.tiles_w; A.tiles_h = g.tiles_h; A.tiles_n = g.tiles_n;
  A.bn = bn;
  A.n_tiles = (d->cin + bn - 1) / bn;
  A.m_tiles = (d->cout + 127) / 128;
  A.kchunks = 0;
  A.splits = splits;
  A.tmem_cols = tmem_cols_for(bn);
  A.out = splits > 1 ? workspace : (void*)dw;
  CUtensorMap ymap, xmap;
  if (int rc = make_nhwc_map(&ymap, dy, d->n, d->ho, d->wo, d->cout, g.tw, g.th, g.tn)) return rc;
  if (int rc = make_nhwc_map(&xmap, x, d->n, d->h, d->w, d->cin, g.tw, g.th, g.tn)) return rc;
  const size_t smem = (size_t)BIG_STAGES * (2 + bn / 64) * 64 * 128 + sizeof(BigSmem) + 1024;
  if (int rc = big_set_smem((const void*)conv_big_wgrad_kernel, 1)) return rc;
  const int total = d->kh * d->kw * A.m_tiles * A.n_tiles * splits;
  const int grid = total < big_sm_count() ? total : big_sm_count();
  conv_big_wgrad_kernel<<<grid, BIG_THREADS, smem, st>>>(ymap, xmap, A);
  if (int rc = check_launch("conv_big_wgrad_kernel")) return rc;
  if (splits > 1) {
    const long long n4 = (long long)d->cout * d->kh * d->kw * d->cin / 4;
    const int blocks = (int)((n4 + 255) / 256 < 148 * 8 ? (n4 + 255) / 256 : 148 * 8);
    big_wgrad_reduce_kernel<<<blocks, 256, 0, st>>>((const float4*)workspace, (float4*)dw, n4, splits);
    return check_launch("big_wgrad_reduce_kernel");
  }
  return 0;
}

}  // namespace cgat
