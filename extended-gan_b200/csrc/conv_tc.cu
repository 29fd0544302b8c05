// K1/K2/K3: implicit-GEMM convolution on the 5th-generation tensor cores (tcgen05 + TMEM + TMA), sm_100a.
//
// Shape class served here: stride-1 convolutions over NHWC bf16 activations with cin % 8 == 0
// (the conv-GAT node conv, 3x3, 24 -> 72 channels as a block-diagonal dense conv; DCGAN generator
// layers; 1x1 pointwise convs).  Other shapes take the direct kernel (conv_direct.cu).
//
// Mapping (fprop):  D[128 pixels x NPAD couts] = sum over taps (r,s) and channel chunks of
//                   A_tap[128 pixels x 16 ch] . B[16 ch x NPAD couts]        (tcgen05.mma M=128, K=16)
//  * an output tile is TH=16 rows x TW=8 columns of one image (M = 128);
//  * the INPUT HALO tile ((TH+kh-1) x (TW+kw-1) pixels) is loaded ONCE per tile by TMA as WIDE rows: the
//    tensor is viewed as [n][h][w*c] so one box row is a whole halo row (WP pixels x C channels, contiguous
//    in HBM); out-of-image elements are zero-filled by the TMA unit (= the conv zero padding, including
//    PyTorch's asymmetric padding="same" for even kernels).  (A first version issued one 16-byte-row box
//    per 8-channel chunk; ncu showed the TMA unit, not HBM or the tensor pipe, was the limiter.)
//  * four producer warps re-lay the staged rows into [chunk][halo row][halo col][8 ch] (16-byte units);
//  * no im2col is materialised: every tap's A operand is that SAME buffer addressed through a different
//    UMMA shared-memory descriptor -- start address shifted by (r*WP+s) pixels, 8 consecutive pixels of an
//    image row form one core matrix (8 rows x 16 B, SWIZZLE_NONE K-major), core matrices step by one halo
//    row (SBO = WP*16 B), the two K-chunks of an instruction step by one chunk plane (LBO);
//  * weights are pre-packed once per call into the matching K-major core-matrix order and stay resident
//    in shared memory for the whole persistent CTA;
//  * accumulators live in TMEM (2 stages x NPAD columns); the epilogue warps read them with tcgen05.ld,
//    add bias, apply the activation and store bf16 NHWC, overlapping the next tile's MMAs.
// Warp roles: warps 0-3 = TMA issue + re-layout, warp 4 = MMA issuer (one thread) + TMEM allocator,
// warps 5-8 = epilogue.
#include "tc_common.cuh"

namespace cgat {

constexpr int TC_TH = 16, TC_TW = 8;  // output tile (rows x cols) -> M = 128
constexpr int TC_THREADS = 288;
constexpr int TC_PROD = 128;          // producer threads (warps 0-3)
constexpr int TC_MMA_WARP = 4;
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

// The driver entry point needs a current context on the CALLING thread.  Autograd runs backward on worker
// threads that may not have touched the runtime yet, so bind the primary context of the thread's device.
static void ensure_context() {
  using GetCurFn = CUresult (*)(CUcontext*);
  static GetCurFn get_cur = nullptr;
  if (!get_cur) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuCtxGetCurrent", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      get_cur = reinterpret_cast<GetCurFn>(p);
  }
  CUcontext ctx = nullptr;
  if (get_cur && get_cur(&ctx) == CUDA_SUCCESS && ctx != nullptr) return;
  int dev = 0;
  cudaGetDevice(&dev);
  cudaSetDevice(dev);  // CUDA >= 12: initialises the runtime and makes the primary context current
}

int make_rows_map(CUtensorMap* map, const void* base, int n, int h, int w, int c, int boxe, int rows) {
  EncodeTiledFn enc = get_encode_tiled();
  if (!enc) return fail(CGAT_EUNSUPPORTED, "cuTensorMapEncodeTiled not available from the driver");
  ensure_context();
  cuuint64_t dims[3] = {(cuuint64_t)w * c, (cuuint64_t)h, (cuuint64_t)n};
  cuuint64_t strides[2] = {(cuuint64_t)w * c * 2, (cuuint64_t)h * w * c * 2};
  cuuint32_t box[3] = {(cuuint32_t)boxe, (cuuint32_t)rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(CGAT_EINVAL, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return 0;
}

static int sm_count() {
  static int num_sms = 0;
  if (!num_sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
  }
  return num_sms;
}

// A staged row-tile: `rows` rows of `rl` contiguous elements, fetched as `nbox` boxes of `boxe` elements
struct RowStage {
  int rl, boxe, nbox, rows;
  uint32_t bytes;  // 128-byte aligned
};
static RowStage row_stage(int pixels_per_row, int c, int rows) {
  RowStage s;
  s.rl = pixels_per_row * c;
  s.boxe = s.rl < 256 ? s.rl : 256;
  s.nbox = (s.rl + s.boxe - 1) / s.boxe;
  s.rows = rows;
  s.bytes = (uint32_t)(((size_t)s.nbox * rows * s.boxe * 2 + 127) & ~(size_t)127);
  return s;
}

struct TcGeom {
  int nchunk;  // real 8-channel chunks of the GEMM-K side per tap
  int ch2;     // chunks per tap rounded up to even (one tcgen05.mma consumes two)
  int npad;    // GEMM-N padded to a multiple of 16
  int hp, wp;  // halo tile
  int chb;     // bytes of one chunk plane (128-byte aligned)
  int kc;      // total K chunks = taps * ch2
  RowStage xs;
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
  g.xs = row_stage(g.wp, cin, g.hp);
  g.wbytes = (size_t)g.kc * g.npad * 16;
  g.stage_bytes = (size_t)g.ch2 * g.chb;
  g.smem = 2048 + g.wbytes + 2 * (size_t)g.xs.bytes + TC_STAGES * g.stage_bytes;
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

// ---- producer-side helpers (warps 0-3) -----------------------------------------------------------------
struct StageDesc {  // device copy of RowStage
  int rl, boxe, nbox, rows;
  uint32_t bytes;
};

__device__ __forceinline__ void issue_rows(const CUtensorMap* map, const StageDesc& s, unsigned char* dst, int c0,
                                           int row0, int n, uint64_t* bar, uint32_t extra_tx = 0) {
  mbar_arrive_expect_tx(bar, (uint32_t)s.nbox * s.rows * s.boxe * 2 + extra_tx);
  for (int b = 0; b < s.nbox; ++b)
    tma_load_3d(dst + (size_t)b * s.rows * s.boxe * 2, map, c0 + b * s.boxe, row0, n, bar);
}

// staged rows [box][row][boxe] -> planes [chunk][row][col][8 ch]   (one 16-byte unit per (chunk,row,col))
__device__ __forceinline__ void relayout(const unsigned char* __restrict__ stag, unsigned char* __restrict__ planes,
                                         const StageDesc& s, int cols, int c_total, int nchunk, int plane_bytes,
                                         int pwarp, int lane) {
  const int nq = nchunk * cols;
  for (int row = pwarp; row < s.rows; row += TC_PROD / 32)
    for (int q = lane; q < nq; q += 32) {
      const int c = q / cols;
      const int col = q - c * cols;
      const int e = col * c_total + c * 8;
      const int b = e / s.boxe;
      const int off = e - b * s.boxe;
      const uint4 v = *reinterpret_cast<const uint4*>(stag + ((size_t)(b * s.rows + row) * s.boxe + off) * 2);
      *reinterpret_cast<uint4*>(planes + (size_t)c * plane_bytes + (size_t)(row * cols + col) * 16) = v;
    }
}

// Optional timeline instrumentation (developer aid): when a buffer is registered, CTA 0 records clock64()
// at the pipeline hand-off points of its first DBG_TILES tiles: [tile][event] with events
// 0 P:loop top, 1 P:empty ok, 2 P:rows landed, 3 P:relayout done, 4 M:tempty ok, 5 M:full ok, 6 M:issued,
// 7 E:tfull ok, 8 E:done
constexpr int DBG_TILES = 16, DBG_EVENTS = 9;
static long long* g_dbg = nullptr;
void set_debug_buffer(long long* p) { g_dbg = p; }
#define DBG(ev)                                                                                   \
  do {                                                                                            \
    if (A.dbg != nullptr && blockIdx.x == 0 && it < DBG_TILES) A.dbg[it * DBG_EVENTS + (ev)] = clock64(); \
  } while (0)

// ---- fprop / dgrad kernel ------------------------------------------------------------------------------------
struct ConvTcArgs {
  long long* dbg;
  const __nv_bfloat16* wpack;
  const float* bias;
  __nv_bfloat16* y;
  int ho, wo, cin, cout, npad;
  int kh, kw, pad_t, pad_l;
  int nchunk, ch2, hp, wp, chb, kc;
  int tiles_h, tiles_w, tiles;
  int act;
  uint32_t tmem_cols;
  uint32_t wbytes, stage_bytes;
  StageDesc xs;
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
  // [0,1024): barriers + tmem pointer; packed weights; 2 row-staging buffers; TC_STAGES halo stages
  uint64_t* full = reinterpret_cast<uint64_t*>(smem);  // [TC_STAGES]  producers -> MMA
  uint64_t* empty = full + TC_STAGES;                  // [TC_STAGES]  MMA -> producers
  uint64_t* tfull = empty + TC_STAGES;                 // [2]          MMA -> epilogue
  uint64_t* tempty = tfull + 2;                        // [2]          epilogue -> MMA
  uint64_t* wbar = tempty + 2;                         // [1]          packed weights landed
  uint64_t* sbar = wbar + 1;                           // [2]          TMA rows landed in staging
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(sbar + 2);
  float* s_bias = reinterpret_cast<float*>(smem + 1024);  // [npad <= 256]
  unsigned char* s_w = smem + 2048;
  unsigned char* s_stag = s_w + A.wbytes;
  unsigned char* s_halo = s_stag + 2 * (size_t)A.xs.bytes;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int i = 0; i < TC_STAGES; ++i) { mbar_init(&full[i], TC_PROD); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 128); mbar_init(&sbar[i], 1); }
    mbar_init(wbar, 1);
    fence_mbar_init();
    tma_prefetch_desc(&tmap_x);
  }
  for (int i = threadIdx.x; i < A.npad; i += TC_THREADS)
    s_bias[i] = (A.bias != nullptr && i < A.cout) ? A.bias[i] : 0.f;
  // zero the padding chunk planes (never written, multiplied by zero weights but must be finite)
  if (A.ch2 != A.nchunk) {
    for (int s = 0; s < TC_STAGES; ++s) {
      uint4* p = reinterpret_cast<uint4*>(s_halo + (size_t)s * A.stage_bytes + (size_t)A.nchunk * A.chb);
      for (int i = threadIdx.x; i < A.chb / 16; i += TC_THREADS) p[i] = make_uint4(0, 0, 0, 0);
    }
    fence_proxy_async_smem();
  }
  if (warp == TC_MMA_WARP) tmem_alloc(tmem_ptr, A.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp < TC_PROD / 32) {
    // ===================== producers: TMA rows -> staging -> chunk planes =====================
    const int ptid = threadIdx.x;
    auto origin = [&](int tile, int& c0, int& row0, int& n) {
      const int tw = tile % A.tiles_w;
      const int th = (tile / A.tiles_w) % A.tiles_h;
      n = tile / (A.tiles_w * A.tiles_h);
      c0 = (tw * TC_TW - A.pad_l) * A.cin;
      row0 = th * TC_TH - A.pad_t;
    };
    if (ptid == 0) {
      mbar_arrive_expect_tx(wbar, A.wbytes);
      bulk_g2s(s_w, A.wpack, A.wbytes, wbar);
      if ((int)blockIdx.x < A.tiles) {
        int c0, row0, n;
        origin(blockIdx.x, c0, row0, n);
        issue_rows(&tmap_x, A.xs, s_stag, c0, row0, n, &sbar[0]);
      }
    }
    int stage = 0, it = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < A.tiles; tile += gridDim.x, ++it) {
      named_bar_sync(1, TC_PROD);  // everyone is done reading staging[(it+1)&1] (tile it-1)
      if (ptid == 0) {
        DBG(0);
        const int next = tile + gridDim.x;
        if (next < A.tiles) {
          int c0, row0, n;
          origin(next, c0, row0, n);
          issue_rows(&tmap_x, A.xs, s_stag + (size_t)((it + 1) & 1) * A.xs.bytes, c0, row0, n, &sbar[(it + 1) & 1]);
        }
      }
      mbar_wait(&empty[stage], phase ^ 1);
      if (ptid == 0) DBG(1);
      mbar_wait(&sbar[it & 1], (uint32_t)(it >> 1) & 1u);
      if (ptid == 0) DBG(2);
      relayout(s_stag + (size_t)(it & 1) * A.xs.bytes, s_halo + (size_t)stage * A.stage_bytes, A.xs, A.wp, A.cin,
               A.nchunk, A.chb, warp, lane);
      fence_proxy_async_smem();
      if (ptid == 0) DBG(3);
      mbar_arrive(&full[stage]);
      if (++stage == TC_STAGES) { stage = 0; phase ^= 1; }
    }
  } else if (warp == TC_MMA_WARP) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc = make_idesc_bf16(128, A.npad, 0, 0);
      const uint32_t a_sbo = (uint32_t)A.wp * 16, a_lbo = (uint32_t)A.chb;
      const uint32_t b_sbo = 128, b_lbo = (uint32_t)A.npad * 16;
      const uint32_t w_addr = smem_u32(s_w);
      mbar_wait(wbar, 0);
      int stage = 0, acc = 0, it = 0;
      uint32_t phase = 0, aphase = 0;
      for (int tile = blockIdx.x; tile < A.tiles; tile += gridDim.x, ++it) {
        mbar_wait(&tempty[acc], aphase ^ 1);
        DBG(4);
        mbar_wait(&full[stage], phase);
        DBG(5);
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
        DBG(6);
        if (++stage == TC_STAGES) { stage = 0; phase ^= 1; }
        if (++acc == 2) { acc = 0; aphase ^= 1; }
      }
    }
  } else {
    // ===================== epilogue (warps 5..8 -> TMEM lane groups 1,2,3,0) =====================
    const int lg = warp & 3;
    const int m = lg * 32 + lane;  // accumulator row = pixel of the tile
    const int hrow = m >> 3, wcol = m & 7;
    int acc = 0, it = 0;
    uint32_t aphase = 0;
    for (int tile = blockIdx.x; tile < A.tiles; tile += gridDim.x, ++it) {
      const int tw = tile % A.tiles_w;
      const int th = (tile / A.tiles_w) % A.tiles_h;
      const int n = tile / (A.tiles_w * A.tiles_h);
      const int h = th * TC_TH + hrow, w = tw * TC_TW + wcol;
      const bool valid = h < A.ho && w < A.wo;
      mbar_wait(&tfull[acc], aphase);
      if (m == 0) DBG(7);
      tc_fence_after();
      __nv_bfloat16* yp = A.y + (((long long)n * A.ho + h) * A.wo + w) * A.cout;
      const uint32_t t_addr = tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(acc * A.npad);
      for (int c0 = 0; c0 < A.npad; c0 += 32) {
        // two 16-column TMEM loads in flight before the single wait
        float v[32];
        tmem_ld16_nowait(t_addr + c0, *reinterpret_cast<float(*)[16]>(&v[0]));
        if (c0 + 16 < A.npad) tmem_ld16_nowait(t_addr + c0 + 16, *reinterpret_cast<float(*)[16]>(&v[16]));
        tmem_ld_wait();
        if (valid) {
#pragma unroll
          for (int hseg = 0; hseg < 2; ++hseg) {
            const int cb = c0 + 16 * hseg;
            if (cb >= A.npad) break;
            float* u = &v[16 * hseg];
#pragma unroll
            for (int i = 0; i < 16; ++i) u[i] = tc_act(u[i] + s_bias[cb + i], A.act);
            if ((A.cout & 7) == 0 && cb + 16 <= A.cout) {
              uint4 o0, o1;
              __nv_bfloat162 t;
#define PK(a, b) (t = __floats2bfloat162_rn(a, b), *reinterpret_cast<uint32_t*>(&t))
              o0.x = PK(u[0], u[1]); o0.y = PK(u[2], u[3]); o0.z = PK(u[4], u[5]); o0.w = PK(u[6], u[7]);
              o1.x = PK(u[8], u[9]); o1.y = PK(u[10], u[11]); o1.z = PK(u[12], u[13]); o1.w = PK(u[14], u[15]);
#undef PK
              reinterpret_cast<uint4*>(yp + cb)[0] = o0;
              reinterpret_cast<uint4*>(yp + cb)[1] = o1;
            } else {
#pragma unroll
              for (int i = 0; i < 16; ++i)
                if (cb + i < A.cout) yp[cb + i] = __float2bfloat16_rn(u[i]);
            }
          }
        }
      }
      tc_fence_before();
      if (m == 0) DBG(8);
      mbar_arrive(&tempty[acc]);
      if (++acc == 2) { acc = 0; aphase ^= 1; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == TC_MMA_WARP) {
    __syncwarp();
    tmem_dealloc(tmem_base, A.tmem_cols);
  }
}

// ---- support matrix / launchers ----------------------------------------------------------------------------
int conv_wgrad_tc_supported(const cgat_conv_desc* d);
size_t conv_wgrad_tc_workspace(const cgat_conv_desc* d);

int conv_tc_supported(const cgat_conv_desc* d, int which) {
  if (d->dtype != CGAT_BF16 || d->stride != 1) return 0;
  if (which == 2) return conv_wgrad_tc_supported(d);
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
  if (which == 2) return conv_wgrad_tc_workspace(d);
  const int gk = which == 0 ? d->cin : d->cout;
  const int gn = which == 0 ? d->cout : d->cin;
  return geom(d->n, d->ho, d->wo, gk, gn, d->kh, d->kw).wbytes;
}

static StageDesc to_dev(const RowStage& s) { return StageDesc{s.rl, s.boxe, s.nbox, s.rows, s.bytes}; }

// generic stride-1 launch: input [n][hi][wi][gk] -> output [n][hout][wout][gn]
static int launch_tc(const void* in, int n, int hi, int wi, int gk, const void* w_krsc, int w_cout, int w_cin,
                     int dgrad, int kh, int kw, int pad_t, int pad_l, int hout, int wout, int gn, const float* bias,
                     int act, void* out, void* workspace, cudaStream_t st) {
  if (!aligned16(in) || !aligned16(out) || !aligned16(workspace))
    return fail(CGAT_EALIGN, "conv tensors / workspace must be 16-byte aligned");
  const TcGeom g = geom(n, hout, wout, gk, gn, kh, kw);
  CUtensorMap map;
  if (int rc = make_rows_map(&map, in, n, hi, wi, gk, g.xs.boxe, g.xs.rows)) return rc;
  {
    const long long total = (long long)g.kc * g.npad * 8;
    int blocks = (int)((total + 255) / 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    pack_weights_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>((const __nv_bfloat16*)w_krsc, (__nv_bfloat16*)workspace,
                                                              w_cout, w_cin, kh, kw, gn, gk, g.ch2, g.npad, dgrad);
    if (int rc = check_launch("pack_weights_kernel")) return rc;
  }
  ConvTcArgs A{};
  A.dbg = g_dbg;
  A.wpack = (const __nv_bfloat16*)workspace;
  A.bias = bias;
  A.y = (__nv_bfloat16*)out;
  A.ho = hout; A.wo = wout; A.cin = gk; A.cout = gn; A.npad = g.npad;
  A.kh = kh; A.kw = kw; A.pad_t = pad_t; A.pad_l = pad_l;
  A.nchunk = g.nchunk; A.ch2 = g.ch2; A.hp = g.hp; A.wp = g.wp; A.chb = g.chb; A.kc = g.kc;
  A.tiles_h = g.tiles_h; A.tiles_w = g.tiles_w; A.tiles = g.tiles;
  A.act = act;
  A.tmem_cols = g.tmem_cols;
  A.wbytes = (uint32_t)g.wbytes;
  A.stage_bytes = (uint32_t)g.stage_bytes;
  A.xs = to_dev(g.xs);
  cudaError_t e = cudaFuncSetAttribute(conv_fprop_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem);
  if (e != cudaSuccess) return fail((int)e, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
  const int grid = g.tiles < sm_count() ? g.tiles : sm_count();
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

// =====================================================================================================
// K3 wgrad:  dW[cout][tap][cin] = sum over pixels  dY[pix][cout] * X[pix + tap][cin]
//
// GEMM view: M = cout (<= 128, the TMEM lanes), N = cin per tap (padded to 16), K = pixels.  Both operands
// are MN-major in shared memory (pixels are the contraction): dY tiles are re-laid as
// [cout chunk][row][8 px][8 co] and the X halo tile is the very same [cin chunk][halo row][halo col][8 ci]
// buffer the forward uses; tap (r,s) is again only a shifted descriptor start address.  One tcgen05.mma
// (K=16) contracts two image rows of 8 pixels.  A persistent CTA keeps ONE accumulator set in TMEM
// (taps x NT columns, + 16 columns fed by a plane of ones that yield dbias) across all its tiles, then
// writes its partial sums once; a small kernel reduces the per-CTA partials in a fixed order
// (deterministic, no atomics).
// =====================================================================================================
constexpr int WG_STAGES = 3;

struct WgGeom {
  int nchunk, ch2, nt;    // cin chunks, even-padded, N per tap = ch2*8
  int mchunk;             // cout / 8
  int hp, wp, chb;        // halo tile of x
  int cols;               // TMEM columns used = taps*nt + 16
  uint32_t tmem_cols;
  RowStage xs, ys;
  size_t x_stage, dy_stage, stage_bytes, smem;
  int tiles_h, tiles_w, tiles;
};

static WgGeom wgeom(const cgat_conv_desc* d) {
  WgGeom g;
  g.nchunk = d->cin / 8;
  g.ch2 = (g.nchunk + 1) & ~1;
  g.nt = g.ch2 * 8;
  g.mchunk = d->cout / 8;
  g.hp = TC_TH + d->kh - 1;
  g.wp = TC_TW + d->kw - 1;
  g.chb = (g.hp * g.wp * 16 + 127) & ~127;
  g.cols = d->kh * d->kw * g.nt + 16;
  uint32_t c = 32;
  while (c < (uint32_t)g.cols) c <<= 1;
  g.tmem_cols = c;
  g.xs = row_stage(g.wp, d->cin, g.hp);
  g.ys = row_stage(TC_TW, d->cout, TC_TH);
  g.x_stage = (size_t)g.ch2 * g.chb;
  g.dy_stage = (size_t)16 * 2048;  // 16 cout-chunk planes (M = 128) of 16 rows x 8 px x 16 B
  g.stage_bytes = g.x_stage + g.dy_stage;
  g.smem = 1024 + 4096 + 2 * ((size_t)g.xs.bytes + g.ys.bytes) + WG_STAGES * g.stage_bytes;
  g.tiles_h = (d->ho + TC_TH - 1) / TC_TH;
  g.tiles_w = (d->wo + TC_TW - 1) / TC_TW;
  g.tiles = d->n * g.tiles_h * g.tiles_w;
  return g;
}

struct WgArgs {
  float* partial;  // [grid][128][cols]
  int cin, cout;
  int kh, kw, pad_t, pad_l;
  int nchunk, ch2, nt, mchunk, hp, wp, chb, cols;
  int tiles_h, tiles_w, tiles;
  uint32_t tmem_cols, x_stage, stage_bytes;
  StageDesc xs, ys;
};

__global__ void __launch_bounds__(TC_THREADS, 1)
conv_wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_dy,
                     const WgArgs A) {
  extern __shared__ __align__(1024) unsigned char smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem);  // [WG_STAGES]
  uint64_t* empty = full + WG_STAGES;                  // [WG_STAGES]
  uint64_t* done = empty + WG_STAGES;                  // [1]
  uint64_t* sbar = done + 1;                           // [2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(sbar + 2);
  unsigned char* s_ones = smem + 1024;  // [2 chunk planes][16 rows][8 px][8 x bf16(1.0)]
  unsigned char* s_stag = s_ones + 4096;
  const size_t stag_bytes = (size_t)A.xs.bytes + A.ys.bytes;
  unsigned char* s_stage = s_stag + 2 * stag_bytes;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < WG_STAGES; ++i) { mbar_init(&full[i], TC_PROD); mbar_init(&empty[i], 1); }
    mbar_init(done, 1);
    mbar_init(&sbar[0], 1);
    mbar_init(&sbar[1], 1);
    fence_mbar_init();
    tma_prefetch_desc(&tmap_x);
    tma_prefetch_desc(&tmap_dy);
  }
  {
    uint32_t* o = reinterpret_cast<uint32_t*>(s_ones);
    for (int i = threadIdx.x; i < 1024; i += TC_THREADS) o[i] = 0x3f803f80u;  // two bf16 1.0
    // planes that are never written must hold finite numbers: x padding chunk, dY planes >= mchunk
    for (int s = 0; s < WG_STAGES; ++s) {
      unsigned char* st = s_stage + (size_t)s * A.stage_bytes;
      if (A.ch2 != A.nchunk) {
        uint4* p = reinterpret_cast<uint4*>(st + (size_t)A.nchunk * A.chb);
        for (int i = threadIdx.x; i < A.chb / 16; i += TC_THREADS) p[i] = make_uint4(0, 0, 0, 0);
      }
      uint4* q = reinterpret_cast<uint4*>(st + A.x_stage + (size_t)A.mchunk * 2048);
      for (int i = threadIdx.x; i < (16 - A.mchunk) * 128; i += TC_THREADS) q[i] = make_uint4(0, 0, 0, 0);
    }
    fence_proxy_async_smem();
  }
  if (warp == TC_MMA_WARP) tmem_alloc(tmem_ptr, A.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp < TC_PROD / 32) {
    const int ptid = threadIdx.x;
    auto issue = [&](int tile, int buf) {
      const int tw = tile % A.tiles_w;
      const int th = (tile / A.tiles_w) % A.tiles_h;
      const int n = tile / (A.tiles_w * A.tiles_h);
      unsigned char* dst = s_stag + (size_t)buf * stag_bytes;
      // one mbarrier phase covers both operands: arm it once with the sum of the bytes
      const uint32_t ybytes = (uint32_t)A.ys.nbox * A.ys.rows * A.ys.boxe * 2;
      issue_rows(&tmap_x, A.xs, dst, (tw * TC_TW - A.pad_l) * A.cin, th * TC_TH - A.pad_t, n, &sbar[buf], ybytes);
      for (int b = 0; b < A.ys.nbox; ++b)
        tma_load_3d(dst + A.xs.bytes + (size_t)b * A.ys.rows * A.ys.boxe * 2, &tmap_dy,
                    tw * TC_TW * A.cout + b * A.ys.boxe, th * TC_TH, n, &sbar[buf]);
    };
    if (ptid == 0 && (int)blockIdx.x < A.tiles) issue(blockIdx.x, 0);
    int stage = 0, it = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < A.tiles; tile += gridDim.x, ++it) {
      named_bar_sync(1, TC_PROD);
      if (ptid == 0) {
        const int next = tile + gridDim.x;
        if (next < A.tiles) issue(next, (it + 1) & 1);
      }
      mbar_wait(&empty[stage], phase ^ 1);
      mbar_wait(&sbar[it & 1], (uint32_t)(it >> 1) & 1u);
      const unsigned char* sg = s_stag + (size_t)(it & 1) * stag_bytes;
      unsigned char* st = s_stage + (size_t)stage * A.stage_bytes;
      relayout(sg, st, A.xs, A.wp, A.cin, A.nchunk, A.chb, warp, lane);
      relayout(sg + A.xs.bytes, st + A.x_stage, A.ys, TC_TW, A.cout, A.mchunk, 2048, warp, lane);
      fence_proxy_async_smem();
      mbar_arrive(&full[stage]);
      if (++stage == WG_STAGES) { stage = 0; phase ^= 1; }
    }
  } else if (warp == TC_MMA_WARP) {
    if (lane == 0) {
      const uint32_t idesc_w = make_idesc_bf16(128, A.nt, 1, 1);
      const uint32_t idesc_b = make_idesc_bf16(128, 16, 1, 1);
      const uint32_t ones_addr = smem_u32(s_ones);
      int stage = 0;
      uint32_t phase = 0;
      uint32_t accum = 0;
      for (int tile = blockIdx.x; tile < A.tiles; tile += gridDim.x) {
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        const uint32_t x_addr = smem_u32(s_stage + (size_t)stage * A.stage_bytes);
        const uint32_t dy_addr = x_addr + A.x_stage;
        for (int j = 0; j < TC_TH / 2; ++j) {  // K step: image rows 2j, 2j+1 of the tile (16 pixels)
          const uint64_t ad = make_smem_desc(dy_addr + j * 256, /*LBO: next 8 pixels*/ 128, /*SBO: next 8 couts*/ 2048);
          const uint32_t acc_j = accum | (j > 0);
          for (int r = 0; r < A.kh; ++r)
            for (int s = 0; s < A.kw; ++s) {
              const uint64_t bd = make_smem_desc(x_addr + (uint32_t)(((2 * j + r) * A.wp + s) * 16),
                                                 /*LBO: next halo row*/ (uint32_t)A.wp * 16, /*SBO: next 8 cin*/ (uint32_t)A.chb);
              umma_bf16(tmem_base + (uint32_t)((r * A.kw + s) * A.nt), ad, bd, idesc_w, acc_j);
            }
          const uint64_t od = make_smem_desc(ones_addr + j * 256, 128, 2048);
          umma_bf16(tmem_base + (uint32_t)(A.kh * A.kw * A.nt), ad, od, idesc_b, acc_j);
        }
        accum = 1;
        umma_commit(&empty[stage]);
        if (++stage == WG_STAGES) { stage = 0; phase ^= 1; }
      }
      umma_commit(done);
    }
  } else {
    // epilogue: once, after every MMA of this CTA has retired
    const int lg = warp & 3;
    const int m = lg * 32 + lane;  // cout index
    mbar_wait(done, 0);
    tc_fence_after();
    float* out = A.partial + ((size_t)blockIdx.x * 128 + m) * A.cols;
    for (int c0 = 0; c0 < A.cols; c0 += 16) {
      float v[16];
      tmem_ld16(tmem_base + ((uint32_t)(lg * 32) << 16) + c0, v);
      if (m < A.cout) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
          reinterpret_cast<float4*>(out + c0)[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == TC_MMA_WARP) {
    __syncwarp();
    tmem_dealloc(tmem_base, A.tmem_cols);
  }
}

// dw[co][tap][ci] = sum_cta partial[cta][co][tap*nt + ci];  dbias[co] = sum_cta partial[cta][co][taps*nt]
// 64 outputs x 4 CTA-slices per block; fixed summation order (deterministic).
__global__ void __launch_bounds__(256) wgrad_reduce_kernel(const float* __restrict__ partial, float* __restrict__ dw,
                                                           float* __restrict__ dbias, int ncta, int cout, int taps,
                                                           int cin, int nt, int cols) {
  __shared__ float s[4][64];
  const int nw = cout * taps * cin;
  const int total = nw + (dbias ? cout : 0);
  const int o = blockIdx.x * 64 + (threadIdx.x & 63);
  const int slice = threadIdx.x >> 6;
  float acc = 0.f;
  int co = 0, col = 0;
  if (o < total) {
    if (o < nw) {
      const int ci = o % cin;
      const int tap = (o / cin) % taps;
      co = o / (cin * taps);
      col = tap * nt + ci;
    } else {
      co = o - nw;
      col = taps * nt;
    }
    const float* p = partial + (size_t)co * cols + col;
#pragma unroll 4
    for (int b = slice; b < ncta; b += 4) acc += p[(size_t)b * 128 * cols];
  }
  s[slice][threadIdx.x & 63] = acc;
  __syncthreads();
  if (slice == 0 && o < total) {
    const float v = (s[0][threadIdx.x] + s[1][threadIdx.x]) + (s[2][threadIdx.x] + s[3][threadIdx.x]);
    if (o < nw) dw[o] = v; else dbias[co] = v;
  }
}

int conv_wgrad_tc_supported(const cgat_conv_desc* d) {
  if (d->dtype != CGAT_BF16 || d->stride != 1) return 0;
  if (d->cin % 8 != 0 || d->cout % 8 != 0 || d->cout > 128) return 0;
  const WgGeom g = wgeom(d);
  if (g.nt > 256 || g.cols > 512 || g.smem > 227 * 1024) return 0;
  return 1;
}

size_t conv_wgrad_tc_workspace(const cgat_conv_desc* d) {
  const WgGeom g = wgeom(d);
  return (size_t)148 * 128 * g.cols * sizeof(float);
}

int conv_wgrad_tc_launch(const cgat_conv_desc* d, const void* x, const void* dy, float* dw, float* dbias,
                         void* workspace, cudaStream_t st) {
  if (!aligned16(x) || !aligned16(dy) || !aligned16(workspace))
    return fail(CGAT_EALIGN, "conv tensors / workspace must be 16-byte aligned");
  const WgGeom g = wgeom(d);
  CUtensorMap mx, mdy;
  if (int rc = make_rows_map(&mx, x, d->n, d->h, d->w, d->cin, g.xs.boxe, g.xs.rows)) return rc;
  if (int rc = make_rows_map(&mdy, dy, d->n, d->ho, d->wo, d->cout, g.ys.boxe, g.ys.rows)) return rc;
  WgArgs A{};
  A.partial = (float*)workspace;
  A.cin = d->cin; A.cout = d->cout;
  A.kh = d->kh; A.kw = d->kw; A.pad_t = d->pad_top; A.pad_l = d->pad_left;
  A.nchunk = g.nchunk; A.ch2 = g.ch2; A.nt = g.nt; A.mchunk = g.mchunk; A.hp = g.hp; A.wp = g.wp; A.chb = g.chb;
  A.cols = g.cols;
  A.tiles_h = g.tiles_h; A.tiles_w = g.tiles_w; A.tiles = g.tiles;
  A.tmem_cols = g.tmem_cols; A.x_stage = (uint32_t)g.x_stage; A.stage_bytes = (uint32_t)g.stage_bytes;
  A.xs = to_dev(g.xs); A.ys = to_dev(g.ys);
  const int grid = g.tiles < sm_count() ? g.tiles : sm_count();
  if (grid > 148) return fail(CGAT_EUNSUPPORTED, "wgrad workspace sized for <= 148 CTAs");
  cudaError_t e = cudaFuncSetAttribute(conv_wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem);
  if (e != cudaSuccess) return fail((int)e, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
  conv_wgrad_tc_kernel<<<grid, TC_THREADS, g.smem, st>>>(mx, mdy, A);
  if (int rc = check_launch("conv_wgrad_tc_kernel")) return rc;
  const int taps = d->kh * d->kw;
  const int total = d->cout * taps * d->cin + (dbias ? d->cout : 0);
  wgrad_reduce_kernel<<<(total + 63) / 64, 256, 0, st>>>((const float*)workspace, dw, dbias, grid, d->cout, taps,
                                                        d->cin, g.nt, g.cols);
  return check_launch("wgrad_reduce_kernel");
}

}  // namespace cgat
