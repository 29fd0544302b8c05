// K1/K2/K3: implicit-GEMM convolution on the 5th-generation tensor cores (tcgen05 + TMEM + TMA), sm_100a.
//
// Shape class served here: stride-1 convolutions over NHWC bf16 activations with cin % 8 == 0
// (the conv-GAT node conv, 3x3, 24 -> 72 channels as a block-diagonal dense conv; DCGAN generator
// layers; 1x1 pointwise convs).  Other shapes take the direct kernel (conv_direct.cu).
//
// Mapping (fprop):  D[128 pixels x NPAD couts] = sum over taps (r,s) and channel chunks of
//                   A_tap[128 pixels x 16 ch] . B[16 ch x NPAD couts]        (tcgen05.mma M=128, K=16)
//  * an output tile is TH=16 rows x TW=8 columns of one image (M = 128);
//  * the INPUT HALO tile ((TH+kh-1) x (TW+kw-1) pixels) is loaded ONCE per tile by TMA, one
//    cp.async.bulk.tensor box per 8-channel chunk, out-of-image pixels zero-filled by the TMA unit
//    (= the conv zero padding, including PyTorch's asymmetric padding="same" for even kernels);
//  * no im2col is materialised: shared memory holds [chunk][halo row][halo col][8 ch] and every tap's
//    A operand is the SAME buffer addressed through a different UMMA shared-memory descriptor --
//    start address shifted by (r*WP+s) pixels, 8 consecutive pixels of an image row form one core
//    matrix (8 rows x 16 B, SWIZZLE_NONE K-major), core matrices step by one halo row (SBO = WP*16 B),
//    the two K-chunks of an instruction step by one chunk plane (LBO);
//  * weights are pre-packed once per call into the matching K-major core-matrix order and stay resident
//    in shared memory for the whole persistent CTA;
//  * accumulators live in TMEM (2 stages x NPAD columns); the epilogue warps read them with tcgen05.ld,
//    add bias, apply the activation and store bf16 NHWC, overlapping the next tile's MMAs.
// Warp roles: warp 0 = TMA producer, warp 1 = MMA issuer (one thread) + TMEM allocator, warps 2..5 = epilogue.
#include "tc_common.cuh"

namespace cgat {

constexpr int TC_TH = 16, TC_TW = 8;  // output tile (rows x cols) -> M = 128
constexpr int TC_THREADS = 192;
constexpr int TC_STAGES = 4;

// ---- host helpers --------------------------------------------------------------------------------
EncodeTiledFn get_encode_tiled() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

int make_nhwc_map(CUtensorMap* map, const void* base, int n, int h, int w, int c, int bw, int bh) {
  EncodeTiledFn enc = get_encode_tiled();
  if (!enc) return fail(CGAT_EUNSUPPORTED, "cuTensorMapEncodeTiled not available from the driver");
  cuuint64_t dims[4] = {(cuuint64_t)c, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)n};
  cuuint64_t strides[3] = {(cuuint64_t)c * 2, (cuuint64_t)w * c * 2, (cuuint64_t)h * w * c * 2};
  cuuint32_t box[4] = {8, (cuuint32_t)bw, (cuuint32_t)bh, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(CGAT_EINVAL, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return 0;
}

struct TcGeom {
  int nchunk;  // real 8-channel chunks of the GEMM-K side per tap
  int ch2;     // chunks per tap rounded up to even (one tcgen05.mma consumes two)
  int npad;    // GEMM-N padded to a multiple of 16
  int hp, wp;  // halo tile
  int chb;     // bytes of one chunk plane (128-byte aligned)
  int kc;      // total K chunks = taps * ch2
  size_t wbytes, stage_bytes, smem;
  int tiles_h, tiles_w, tiles;
  uint32_t tmem_cols;
};

static TcGeom geom(int n, int ho, int wo, int cin, int cout, int kh, int kw) {
  TcGeom g;
  g.nchunk = cin / 8;
  g.ch2 = (g.nchunk + 1) & ~1;
  g.npad = (cout + 15) & ~15;
  g.hp = TC_TH + kh - 1;
  g.wp = TC_TW + kw - 1;
  g.chb = (g.hp * g.wp * 16 + 127) & ~127;
  g.kc = kh * kw * g.ch2;
  g.wbytes = (size_t)g.kc * g.npad * 16;
  g.stage_bytes = (size_t)g.ch2 * g.chb;
  g.smem = 1024 + g.wbytes + TC_STAGES * g.stage_bytes;
  g.tiles_h = (ho + TC_TH - 1) / TC_TH;
  g.tiles_w = (wo + TC_TW - 1) / TC_TW;
  g.tiles = n * g.tiles_h * g.tiles_w;
  uint32_t c = 32;
  while (c < (uint32_t)(2 * g.npad)) c <<= 1;
  g.tmem_cols = c;
  return g;
}

// ---- weight packing ----------------------------------------------------------------------------------
// out[(tap*ch2 + c) * npad + row][e]  (bf16, 16 bytes per (chunk,row))
//   fprop: row = cout index, value w[row][tap][c*8+e]
//   dgrad: row = cin index,  value w[c*8+e][flipped tap][row]   (roles of cin/cout swapped, kernel rotated 180 deg)
template <typename T>
__global__ void pack_weights_kernel(const T* __restrict__ w, __nv_bfloat16* __restrict__ out, int cout, int cin,
                                    int kh, int kw, int rows, int kdim, int ch2, int npad, int dgrad) {
  const int taps = kh * kw;
  const long long total = (long long)taps * ch2 * npad * 8;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int e = (int)(i & 7);
    long long q = i >> 3;
    const int row = (int)(q % npad);
    q /= npad;
    const int c = (int)(q % ch2);
    const int tap = (int)(q / ch2);
    const int k = c * 8 + e;
    float v = 0.f;
    if (row < rows && k < kdim) {
      if (!dgrad) {
        v = DT<T>::to_f(w[((long long)row * taps + tap) * cin + k]);
      } else {
        const int ftap = taps - 1 - tap;
        v = DT<T>::to_f(w[((long long)k * taps + ftap) * cin + row]);
      }
    }
    out[i] = __float2bfloat16_rn(v);
  }
}

// ---- the kernel ----------------------------------------------------------------------------------------
struct ConvTcArgs {
  const __nv_bfloat16* wpack;
  const float* bias;
  __nv_bfloat16* y;
  int ho, wo, cout, npad;
  int kh, kw, pad_t, pad_l;
  int nchunk, ch2, hp, wp, chb, kc;
  int tiles_h, tiles_w, tiles;
  int act;
  uint32_t tmem_cols;
  uint32_t wbytes, stage_bytes;
};

__device__ __forceinline__ float tc_act(float v, int act) {
  switch (act) {
    case 1: return fmaxf(v, 0.f);
    case 2: return v > 0.f ? v : 0.2f * v;
    case 3: return 1.f / (1.f + __expf(-v));
    default: return v;
  }
}

__global__ void __launch_bounds__(TC_THREADS, 1)
conv_fprop_tc_kernel(const __grid_constant__ CUtensorMap tmap_x, const ConvTcArgs A) {
  extern __shared__ __align__(1024) unsigned char smem[];
  // [0,1024): barriers + tmem pointer;  then packed weights;  then TC_STAGES halo stages
  uint64_t* full = reinterpret_cast<uint64_t*>(smem);         // [TC_STAGES]
  uint64_t* empty = full + TC_STAGES;                         // [TC_STAGES]
  uint64_t* tfull = empty + TC_STAGES;                        // [2]
  uint64_t* tempty = tfull + 2;                               // [2]
  uint64_t* wbar = tempty + 2;                                // [1]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(wbar + 1);
  unsigned char* s_w = smem + 1024;
  unsigned char* s_halo = s_w + A.wbytes;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int i = 0; i < TC_STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 128); }
    mbar_init(wbar, 1);
    fence_mbar_init();
    tma_prefetch_desc(&tmap_x);
  }
  // zero the padding chunk planes (never written by TMA, multiplied by zero weights but must be finite)
  if (A.ch2 != A.nchunk) {
    for (int s = 0; s < TC_STAGES; ++s) {
      uint4* p = reinterpret_cast<uint4*>(s_halo + (size_t)s * A.stage_bytes + (size_t)A.nchunk * A.chb);
      for (int i = threadIdx.x; i < A.chb / 16; i += TC_THREADS) p[i] = make_uint4(0, 0, 0, 0);
    }
    fence_proxy_async_smem();
  }
  if (warp == 1) tmem_alloc(tmem_ptr, A.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      mbar_arrive_expect_tx(wbar, A.wbytes);
      bulk_g2s(s_w, A.wpack, A.wbytes, wbar);
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t tx = (uint32_t)A.nchunk * A.hp * A.wp * 16;
      for (int tile = blockIdx.x; tile < A.tiles; tile += gridDim.x) {
        const int tw = tile % A.tiles_w;
        const int th = (tile / A.tiles_w) % A.tiles_h;
        const int n = tile / (A.tiles_w * A.tiles_h);
        mbar_wait(&empty[stage], phase ^ 1);
        mbar_arrive_expect_tx(&full[stage], tx);
        unsigned char* dst = s_halo + (size_t)stage * A.stage_bytes;
        for (int c = 0; c < A.nchunk; ++c)
          tma_load_4d(dst + (size_t)c * A.chb, &tmap_x, c * 8, tw * TC_TW - A.pad_l, th * TC_TH - A.pad_t, n,
                      &full[stage]);
        if (++stage == TC_STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc = make_idesc_bf16(128, A.npad, 0, 0);
      const uint32_t a_sbo = (uint32_t)A.wp * 16, a_lbo = (uint32_t)A.chb;
      const uint32_t b_sbo = 128, b_lbo = (uint32_t)A.npad * 16;
      const uint32_t w_addr = smem_u32(s_w);
      mbar_wait(wbar, 0);
      int stage = 0, acc = 0;
      uint32_t phase = 0, aphase = 0;
      for (int tile = blockIdx.x; tile < A.tiles; tile += gridDim.x) {
        mbar_wait(&tempty[acc], aphase ^ 1);
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        const uint32_t h_addr = smem_u32(s_halo + (size_t)stage * A.stage_bytes);
        const uint32_t d_addr = tmem_base + (uint32_t)(acc * A.npad);
        uint32_t accum = 0;
        for (int r = 0; r < A.kh; ++r)
          for (int s = 0; s < A.kw; ++s) {
            const int tap = r * A.kw + s;
            for (int kp = 0; kp < A.ch2 / 2; ++kp) {
              const uint64_t ad = make_smem_desc(h_addr + (uint32_t)((r * A.wp + s) * 16 + 2 * kp * A.chb), a_lbo, a_sbo);
              const uint64_t bd = make_smem_desc(w_addr + (uint32_t)((tap * A.ch2 + 2 * kp) * A.npad * 16), b_lbo, b_sbo);
              umma_bf16(d_addr, ad, bd, idesc, accum);
              accum = 1;
            }
          }
        umma_commit(&empty[stage]);  // halo stage reusable once these MMAs retire
        umma_commit(&tfull[acc]);    // accumulator ready for the epilogue
        if (++stage == TC_STAGES) { stage = 0; phase ^= 1; }
        if (++acc == 2) { acc = 0; aphase ^= 1; }
      }
    }
  } else {
    // ===================== epilogue (warps 2..5 -> TMEM lane groups 2,3,0,1) =====================
    const int lg = warp & 3;
    const int m = lg * 32 + lane;  // accumulator row = pixel of the tile
    const int hrow = m >> 3, wcol = m & 7;
    int acc = 0;
    uint32_t aphase = 0;
    for (int tile = blockIdx.x; tile < A.tiles; tile += gridDim.x) {
      const int tw = tile % A.tiles_w;
      const int th = (tile / A.tiles_w) % A.tiles_h;
      const int n = tile / (A.tiles_w * A.tiles_h);
      const int h = th * TC_TH + hrow, w = tw * TC_TW + wcol;
      const bool valid = h < A.ho && w < A.wo;
      mbar_wait(&tfull[acc], aphase);
      tc_fence_after();
      __nv_bfloat16* yp = A.y + (((long long)n * A.ho + h) * A.wo + w) * A.cout;
      const uint32_t t_addr = tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(acc * A.npad);
      for (int c0 = 0; c0 < A.npad; c0 += 16) {
        float v[16];
        tmem_ld16(t_addr + c0, v);
        if (valid) {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int c = c0 + i;
            const float b = (A.bias != nullptr && c < A.cout) ? A.bias[c] : 0.f;
            v[i] = tc_act(v[i] + b, A.act);
          }
          if ((A.cout & 7) == 0 && c0 + 16 <= A.cout) {
            uint4 o0, o1;
            __nv_bfloat162 t;
#define PK(a, b) (t = __floats2bfloat162_rn(a, b), *reinterpret_cast<uint32_t*>(&t))
            o0.x = PK(v[0], v[1]); o0.y = PK(v[2], v[3]); o0.z = PK(v[4], v[5]); o0.w = PK(v[6], v[7]);
            o1.x = PK(v[8], v[9]); o1.y = PK(v[10], v[11]); o1.z = PK(v[12], v[13]); o1.w = PK(v[14], v[15]);
#undef PK
            reinterpret_cast<uint4*>(yp + c0)[0] = o0;
            reinterpret_cast<uint4*>(yp + c0)[1] = o1;
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i)
              if (c0 + i < A.cout) yp[c0 + i] = __float2bfloat16_rn(v[i]);
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&tempty[acc]);
      if (++acc == 2) { acc = 0; aphase ^= 1; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, A.tmem_cols);
}

// ---- support matrix / launchers ----------------------------------------------------------------------------
int conv_tc_supported(const cgat_conv_desc* d, int which) {
  if (d->dtype != CGAT_BF16 || d->stride != 1) return 0;
  if (which == 2) return 0;  // wgrad: see conv_wgrad_tc (not yet routed here)
  // dgrad of a stride-1 conv is a stride-1 conv with cin/cout swapped and the kernel rotated
  const int gk = which == 0 ? d->cin : d->cout;   // GEMM-K channels
  const int gn = which == 0 ? d->cout : d->cin;   // GEMM-N channels
  if (gk % 8 != 0 || gn < 1 || gn > 256) return 0;
  if (which == 1) {
    // padding of the equivalent forward conv must be non-negative
    if (d->kh - 1 - d->pad_top < 0 || d->kw - 1 - d->pad_left < 0) return 0;
  }
  const TcGeom g = geom(d->n, which == 0 ? d->ho : d->h, which == 0 ? d->wo : d->w, gk, gn, d->kh, d->kw);
  if (g.smem > 227 * 1024 || g.tmem_cols > 512) return 0;
  return 1;
}

size_t conv_tc_workspace(const cgat_conv_desc* d, int which) {
  if (!conv_tc_supported(d, which)) return 0;
  const int gk = which == 0 ? d->cin : d->cout;
  const int gn = which == 0 ? d->cout : d->cin;
  return geom(d->n, d->ho, d->wo, gk, gn, d->kh, d->kw).wbytes;
}

// generic stride-1 launch: input [n][hi][wi][gk] -> output [n][hout][wout][gn]
static int launch_tc(const void* in, int n, int hi, int wi, int gk, const void* w_krsc, int w_cout, int w_cin,
                     int dgrad, int kh, int kw, int pad_t, int pad_l, int hout, int wout, int gn, const float* bias,
                     int act, void* out, void* workspace, cudaStream_t st) {
  if (!aligned16(in) || !aligned16(out) || !aligned16(workspace))
    return fail(CGAT_EALIGN, "conv tensors / workspace must be 16-byte aligned");
  const TcGeom g = geom(n, hout, wout, gk, gn, kh, kw);
  CUtensorMap map;
  if (int rc = make_nhwc_map(&map, in, n, hi, wi, gk, g.wp, g.hp)) return rc;
  {
    const long long total = (long long)g.kc * g.npad * 8;
    int blocks = (int)((total + 255) / 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    pack_weights_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>((const __nv_bfloat16*)w_krsc, (__nv_bfloat16*)workspace,
                                                              w_cout, w_cin, kh, kw, gn, gk, g.ch2, g.npad, dgrad);
    if (int rc = check_launch("pack_weights_kernel")) return rc;
  }
  ConvTcArgs A{};
  A.wpack = (const __nv_bfloat16*)workspace;
  A.bias = bias;
  A.y = (__nv_bfloat16*)out;
  A.ho = hout; A.wo = wout; A.cout = gn; A.npad = g.npad;
  A.kh = kh; A.kw = kw; A.pad_t = pad_t; A.pad_l = pad_l;
  A.nchunk = g.nchunk; A.ch2 = g.ch2; A.hp = g.hp; A.wp = g.wp; A.chb = g.chb; A.kc = g.kc;
  A.tiles_h = g.tiles_h; A.tiles_w = g.tiles_w; A.tiles = g.tiles;
  A.act = act;
  A.tmem_cols = g.tmem_cols;
  A.wbytes = (uint32_t)g.wbytes;
  A.stage_bytes = (uint32_t)g.stage_bytes;
  static int num_sms = 0;
  if (!num_sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
  }
  cudaError_t e = cudaFuncSetAttribute(conv_fprop_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem);
  if (e != cudaSuccess) return fail((int)e, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
  const int grid = g.tiles < num_sms ? g.tiles : num_sms;
  conv_fprop_tc_kernel<<<grid, TC_THREADS, g.smem, st>>>(map, A);
  return check_launch("conv_fprop_tc_kernel");
}

int conv_fprop_tc_launch(const cgat_conv_desc* d, const void* x, const void* w, const float* bias, void* y,
                         void* workspace, cudaStream_t st) {
  return launch_tc(x, d->n, d->h, d->w, d->cin, w, d->cout, d->cin, 0, d->kh, d->kw, d->pad_top, d->pad_left, d->ho,
                   d->wo, d->cout, bias, d->act, y, workspace, st);
}

int conv_dgrad_tc_launch(const cgat_conv_desc* d, const void* dy, const void* w, void* dx, void* workspace,
                         cudaStream_t st) {
  // dx = conv(dy, rot180(w)^T) with leading padding k-1-pad
  return launch_tc(dy, d->n, d->ho, d->wo, d->cout, w, d->cout, d->cin, 1, d->kh, d->kw, d->kh - 1 - d->pad_top,
                   d->kw - 1 - d->pad_left, d->h, d->w, d->cin, nullptr, 0, dx, workspace, st);
}

int conv_wgrad_tc_launch(const cgat_conv_desc*, const void*, const void*, float*, float*, void*, cudaStream_t) {
  return fail(CGAT_EUNSUPPORTED, "tcgen05 wgrad not built yet");
}

}  // namespace cgat
