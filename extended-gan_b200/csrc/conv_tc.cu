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
//  * measured on B200 (tools/microbench/mma_rate.cu): one tcgen05.mma costs >= 71 cycles to issue whatever its
//    N (<= 128) or layout, and ~350 cycles when consecutive instructions switch accumulator, so the kernels
//    issue FEW, LARGE instructions: fprop packs K densely (chunk-major order, ceil(taps*cin/16) instructions
//    per tile), wgrad builds an im2col operand in shared memory so one instruction covers all taps;
//  * no im2col is materialised: every tap's A operand is that SAME buffer addressed through a different
//    UMMA shared-memory descriptor -- start address shifted by (r*WP+s) pixels, 8 consecutive pixels of an
//    image row form one core matrix (8 rows x 16 B, SWIZZLE_NONE K-major), core matrices step by one halo
//    row (SBO = WP*16 B), the two K-chunks of an instruction step by one chunk plane (LBO);
//  * weights are pre-packed once per call into the matching K-major core-matrix order and stay resident
//    in shared memory for the whole persistent CTA;
//  * accumulators live in TMEM (2 stages x NPAD columns); the epilogue warps read them with tcgen05.ld,
//    add bias, apply the activation and store bf16 NHWC, overlapping the next tile's MMAs.
// Warp roles: warps 0-3 = re-layout, warp 4 = MMA issuer (one thread) + TMEM allocator, warps 5-8 and 10-13 = two
// epilogue groups (one per TMEM accumulator stage), warp 9 = TMA issuer (one thread).
#include "tc_common.cuh"

namespace cgat {

constexpr int TC_TH = 16, TC_TW = 8;  // output tile (rows x cols) -> M = 128
constexpr int TC_THREADS = 448;
constexpr int TC_PROD = 128;          // re-layout threads (warps 0-3)
constexpr int TC_MMA_WARP = 4;
constexpr int TC_TMA_WARP = 9;        // issues the TMA row loads (kept off the re-layout warps' critical path)
// epilogue group 0 = warps 5-8, group 1 = warps 10-13; group g drains accumulator stage g (tiles it % 2 == g)
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
void ensure_context() {
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
  int nchunk;   // 8-channel chunks of the GEMM-K side per tap
  int taps;
  int npairs;   // tcgen05.mma instructions per tile = ceil(nchunk*taps / 2)
  int npad;     // GEMM-N padded to a multiple of 16
  int hp, wp;   // halo tile
  int chb;      // bytes of one chunk plane (128-byte aligned)
  RowStage xs;
  size_t wbytes, stage_bytes, out_bytes, smem;
  int tiles_h, tiles_w, tiles;
  int stages;  // halo stages that fit shared memory (2..TC_STAGES)
  uint32_t tmem_cols;
};

static TcGeom geom(int n, int ho, int wo, int cin, int cout, int kh, int kw) {
  TcGeom g;
  g.nchunk = cin / 8;
  g.taps = kh * kw;
  g.npairs = (g.nchunk * g.taps + 1) / 2;
  g.npad = (cout + 15) & ~15;
  g.hp = TC_TH + kh - 1;
  g.wp = TC_TW + kw - 1;
  g.chb = (g.hp * g.wp * 16 + 127) & ~127;
  g.xs = row_stage(g.wp, cin, g.hp);
  g.wbytes = (size_t)g.npairs * 2 * g.npad * 16;
  g.stage_bytes = (size_t)g.nchunk * g.chb + 128;  // +128: the odd last chunk's partner reads one unit further
  g.out_bytes = ((size_t)128 * cout * 2 + 127) & ~(size_t)127;
  g.stages = TC_STAGES;
  do {
    g.smem = 2048 + 1024 + g.wbytes + 2 * (size_t)g.xs.bytes + g.stages * g.stage_bytes + 2 * g.out_bytes;
  } while (g.smem > 227 * 1024 && --g.stages >= 2);
  if (g.stages < 2) g.stages = 2;
  g.tiles_h = (ho + TC_TH - 1) / TC_TH;
  g.tiles_w = (wo + TC_TW - 1) / TC_TW;
  g.tiles = n * g.tiles_h * g.tiles_w;
  uint32_t c = 32;
  while (c < (uint32_t)(2 * g.npad)) c <<= 1;
  g.tmem_cols = c;
  return g;
}

// ---- weight packing ----------------------------------------------------------------------------------
// K-chunks are ordered CHUNK-MAJOR: i = c*taps + tap.   out[i*npad + row][e]  (bf16, 16 bytes per (i,row))
//   fprop: row = cout index, value w[row][tap][c*8+e]
//   dgrad: row = cin index,  value w[c*8+e][flipped tap][row]   (roles of cin/cout swapped, kernel rotated 180 deg)
// chunks beyond nchunk*taps (odd count) are zero.
template <typename T>
__global__ void pack_weights_kernel(const T* __restrict__ w, __nv_bfloat16* __restrict__ out, int cin, int taps,
                                    int rows, int kdim, int nchunk, int kchunks2, int npad, int dgrad) {
  const long long total = (long long)kchunks2 * npad * 8;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int e = (int)(i & 7);
    long long q = i >> 3;
    const int row = (int)(q % npad);
    const int ki = (int)(q / npad);
    const int c = ki / taps;
    const int tap = ki - c * taps;
    const int k = c * 8 + e;
    float v = 0.f;
    if (row < rows && c < nchunk && k < kdim) {
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

__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, const uint4& v) {
  asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// Copy the `nchunk` 16-byte channel chunks of ONE staged pixel (row, col) to `dst + c*plane_bytes` (c = chunk).
// staging layout [box][row][boxe elements];  element offset of the pixel inside its row = col*c_total.
// All addresses are 32-bit shared-space addresses.
__device__ __forceinline__ void copy_pixel_chunks(uint32_t stag, const StageDesc& s, int row, int col, int c_total,
                                                  int nchunk, uint32_t dst, uint32_t plane_bytes) {
  int e = col * c_total;
  const int b = e / s.boxe;
  int off = e - b * s.boxe;
  const uint32_t box_bytes = (uint32_t)s.rows * s.boxe * 2;
  uint32_t src = stag + (uint32_t)b * box_bytes + (uint32_t)(row * s.boxe + off) * 2;
  int left = (s.boxe - off) >> 3;  // chunks left in this box row
  for (int c = 0; c < nchunk; ++c) {
    sts128(dst, lds128(src));
    dst += plane_bytes;
    src += 16;
    if (--left == 0) {  // continue at the same row of the next box
      src += box_bytes - (uint32_t)s.boxe * 2;
      left = s.boxe >> 3;
    }
  }
}

// Optional timeline instrumentation (developer aid): when a buffer is registered, CTA 0 records clock64()
// at the pipeline hand-off points of its first DBG_TILES tiles: [tile][event] with events
// 0 P:loop top, 1 P:empty ok, 2 P:rows landed, 3 P:relayout done, 4 M:tempty ok, 5 M:full ok, 6 M:issued,
// 7 E:tfull ok, 8 E:done
constexpr int DBG_TILES = 16, DBG_EVENTS = 9;
static long long* g_dbg = nullptr;
void set_debug_buffer(long long* p) { g_dbg = p; }
long long* get_debug_buffer() { return g_dbg; }
#define DBG(ev)                                                                                           \
  do {                                                                                                    \
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
  int nchunk, taps, npairs, hp, wp, chb;
  int tiles_h, tiles_w, tiles;
  int act, stages;
  uint32_t tmem_cols;
  uint32_t wbytes, stage_bytes, out_bytes;
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

constexpr int TC_MAX_PAIRS = 120;  // instructions per tile (descriptor table in shared memory)
constexpr int TC_MAX_OWN = 3;      // halo pixels a producer thread re-lays per tile

__global__ void __launch_bounds__(TC_THREADS, 1)
conv_fprop_tc_kernel(const __grid_constant__ CUtensorMap tmap_x, const ConvTcArgs A) {
  extern __shared__ __align__(1024) unsigned char smem[];
  // [0,1024): barriers + tmem pointer; [1024,2048): bias; [2048,3072): descriptor table; packed weights;
  // 2 row-staging buffers; TC_STAGES halo stages; 2 output staging tiles
  uint64_t* full = reinterpret_cast<uint64_t*>(smem);  // [TC_STAGES]  producers -> MMA
  uint64_t* empty = full + TC_STAGES;                  // [TC_STAGES]  MMA -> producers
  uint64_t* tfull = empty + TC_STAGES;                 // [2]          MMA -> epilogue
  uint64_t* tempty = tfull + 2;                        // [2]          epilogue -> MMA
  uint64_t* wbar = tempty + 2;                         // [1]          packed weights landed
  uint64_t* sbar = wbar + 1;                           // [2]          TMA rows landed in staging
  uint64_t* sfree = sbar + 2;                          // [2]          staging buffer re-laid, may be overwritten
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(sfree + 2);
  float* s_bias = reinterpret_cast<float*>(smem + 1024);       // [npad <= 256]
  uint32_t* s_tab = reinterpret_cast<uint32_t*>(smem + 2048);  // [npairs][2]: A start offset, A LBO (bytes)
  unsigned char* s_w = smem + 3072;
  unsigned char* s_stag = s_w + A.wbytes;
  unsigned char* s_halo = s_stag + 2 * (size_t)A.xs.bytes;
  unsigned char* s_out = s_halo + (size_t)A.stages * A.stage_bytes;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int i = 0; i < TC_STAGES; ++i) { mbar_init(&full[i], TC_PROD); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 128); mbar_init(&sbar[i], 1); mbar_init(&sfree[i], TC_PROD);
    }
    mbar_init(wbar, 1);
    fence_mbar_init();
    tma_prefetch_desc(&tmap_x);
  }
  for (int i = threadIdx.x; i < A.npad; i += TC_THREADS)
    s_bias[i] = (A.bias != nullptr && i < A.cout) ? A.bias[i] : 0.f;
  // descriptor table: K-chunk i = c*taps + tap lives at plane c, shifted by the tap's pixel offset
  for (int p = threadIdx.x; p < A.npairs; p += TC_THREADS) {
    auto addr = [&](int i) {
      const int c = i / A.taps, tap = i - c * A.taps;
      const int r = tap / A.kw, s = tap - r * A.kw;
      return (uint32_t)(c * A.chb + (r * A.wp + s) * 16);
    };
    const uint32_t a0 = addr(2 * p);
    const bool has2 = 2 * p + 1 < A.nchunk * A.taps;
    s_tab[2 * p] = a0;
    s_tab[2 * p + 1] = has2 ? addr(2 * p + 1) - a0 : 16u;  // odd tail: partner weights are zero, data must be finite
  }
  // halo stages start zeroed: alignment padding / the odd tail's partner unit must hold finite numbers
  {
    uint4* p = reinterpret_cast<uint4*>(s_halo);
    const int n16 = (int)((size_t)A.stages * A.stage_bytes / 16);
    for (int i = threadIdx.x; i < n16; i += TC_THREADS) p[i] = make_uint4(0, 0, 0, 0);
    fence_proxy_async_smem();
  }
  if (warp == TC_MMA_WARP) tmem_alloc(tmem_ptr, A.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == TC_TMA_WARP) {
    // ===================== TMA issuer: staged rows of tile it -> staging[it & 1] =====================
    if (lane == 0) {
      mbar_arrive_expect_tx(wbar, A.wbytes);
      bulk_g2s(s_w, A.wpack, A.wbytes, wbar);
      int it = 0;
      for (int tile = blockIdx.x; tile < A.tiles; tile += gridDim.x, ++it) {
        const int tw = tile % A.tiles_w;
        const int th = (tile / A.tiles_w) % A.tiles_h;
        const int n = tile / (A.tiles_w * A.tiles_h);
        mbar_wait(&sfree[it & 1], ((uint32_t)(it >> 1) & 1u) ^ 1u);
        issue_rows(&tmap_x, A.xs, s_stag + (size_t)(it & 1) * A.xs.bytes, (tw * TC_TW - A.pad_l) * A.cin,
                   th * TC_TH - A.pad_t, n, &sbar[it & 1]);
      }
    }
  } else if (warp < TC_PROD / 32) {
    // ===================== re-layout: staging -> chunk planes =====================
    const int ptid = threadIdx.x;
    // the halo pixels this thread re-lays every tile (fixed): q = ptid + k*128
    int own_row[TC_MAX_OWN], own_col[TC_MAX_OWN];
    const int npix = A.hp * A.wp;
#pragma unroll
    for (int k = 0; k < TC_MAX_OWN; ++k) {
      const int q = ptid + k * TC_PROD;
      own_row[k] = q < npix ? q / A.wp : -1;
      own_col[k] = q < npix ? q % A.wp : 0;
    }
    int stage = 0, it = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < A.tiles; tile += gridDim.x, ++it) {
      if (ptid == 0) DBG(0);
      mbar_wait(&empty[stage], phase ^ 1);
      if (ptid == 0) DBG(1);
      mbar_wait(&sbar[it & 1], (uint32_t)(it >> 1) & 1u);
      if (ptid == 0) DBG(2);
      const uint32_t sg = smem_u32(s_stag) + (uint32_t)(it & 1) * A.xs.bytes;
      const uint32_t hs = smem_u32(s_halo) + (uint32_t)stage * A.stage_bytes;
#pragma unroll
      for (int k = 0; k < TC_MAX_OWN; ++k)
        if (own_row[k] >= 0)
          copy_pixel_chunks(sg, A.xs, own_row[k], own_col[k], A.cin, A.nchunk, hs + (uint32_t)(ptid + k * TC_PROD) * 16,
                            (uint32_t)A.chb);
      mbar_arrive(&sfree[it & 1]);  // staging buffer consumed (generic-proxy reads done)
      fence_proxy_async_smem();
      if (ptid == 0) DBG(3);
      mbar_arrive(&full[stage]);
      if (++stage == A.stages) { stage = 0; phase ^= 1; }
    }
  } else if (warp == TC_MMA_WARP) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc = make_idesc_bf16(128, A.npad, 0, 0);
      const uint32_t a_sbo = (uint32_t)A.wp * 16;
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
        uint64_t bd = make_smem_desc(w_addr, b_lbo, b_sbo);
        const uint64_t bstep = (uint64_t)((2 * A.npad * 16) >> 4);  // two K-chunks per instruction
        for (int p = 0; p < A.npairs; ++p) {
          const uint2 t = *reinterpret_cast<const uint2*>(&s_tab[2 * p]);
          const uint64_t ad = make_smem_desc(h_addr + t.x, t.y, a_sbo);
          umma_bf16(d_addr, ad, bd, idesc, p > 0);
          bd += bstep;
        }
        umma_commit(&empty[stage]);  // halo stage reusable once these MMAs retire
        umma_commit(&tfull[acc]);    // accumulator ready for the epilogue
        DBG(6);
        if (++stage == A.stages) { stage = 0; phase ^= 1; }
        if (++acc == 2) { acc = 0; aphase ^= 1; }
      }
    }
  } else {
    // ===================== epilogue: group 0 = warps 5..8, group 1 = warps 10..13 =====================
    // (any 4 consecutive warps cover the 4 TMEM lane quarters: warp % 4)
    const int grp = warp > TC_TMA_WARP ? 1 : 0;
    const int lg = warp & 3;
    const int m = lg * 32 + lane;  // accumulator row = pixel of the tile
    const int etid = threadIdx.x - (grp ? (TC_TMA_WARP + 1) : (TC_MMA_WARP + 1)) * 32;
    const int hrow = m >> 3, wcol = m & 7;
    const bool staged = (A.cout & 7) == 0;  // coalesced path: tile -> shared memory -> bulk stores per image row
    const int acc = grp;  // this group's accumulator stage
    int it = grp;
    uint32_t aphase = 0;
    for (int tile = blockIdx.x + grp * gridDim.x; tile < A.tiles; tile += 2 * gridDim.x, it += 2) {
      const int tw = tile % A.tiles_w;
      const int th = (tile / A.tiles_w) % A.tiles_h;
      const int n = tile / (A.tiles_w * A.tiles_h);
      const int h = th * TC_TH + hrow, w = tw * TC_TW + wcol;
      const bool valid = h < A.ho && w < A.wo;
      unsigned char* ob = s_out + (size_t)grp * A.out_bytes;
      mbar_wait(&tfull[acc], aphase);
      if (m == 0) DBG(7);
      tc_fence_after();
      __nv_bfloat16* yp = A.y + (((long long)n * A.ho + h) * A.wo + w) * A.cout;
      __nv_bfloat16* op = reinterpret_cast<__nv_bfloat16*>(ob) + (size_t)m * A.cout;
      const uint32_t t_addr = tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(acc * A.npad);
      for (int c0 = 0; c0 < A.npad; c0 += 32) {
        float v[32];
        tmem_ld16_nowait(t_addr + c0, *reinterpret_cast<float(*)[16]>(&v[0]));
        if (c0 + 16 < A.npad) tmem_ld16_nowait(t_addr + c0 + 16, *reinterpret_cast<float(*)[16]>(&v[16]));
        tmem_ld_wait();
#pragma unroll
        for (int hseg = 0; hseg < 2; ++hseg) {
          const int cb = c0 + 16 * hseg;
          if (cb >= A.npad) break;
          float* u = &v[16 * hseg];
          {
            const float4* b4 = reinterpret_cast<const float4*>(s_bias + cb);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float4 b = b4[i];
              u[4 * i] += b.x; u[4 * i + 1] += b.y; u[4 * i + 2] += b.z; u[4 * i + 3] += b.w;
            }
            if (A.act != 0) {
#pragma unroll
              for (int i = 0; i < 16; ++i) u[i] = tc_act(u[i], A.act);
            }
          }
          if (staged) {
            if (cb < A.cout) {  // cout % 8 == 0: segments of 8 are all-or-nothing
              uint4 o0, o1;
              __nv_bfloat162 t;
#define PK(a, b) (t = __floats2bfloat162_rn(a, b), *reinterpret_cast<uint32_t*>(&t))
              o0.x = PK(u[0], u[1]); o0.y = PK(u[2], u[3]); o0.z = PK(u[4], u[5]); o0.w = PK(u[6], u[7]);
              o1.x = PK(u[8], u[9]); o1.y = PK(u[10], u[11]); o1.z = PK(u[12], u[13]); o1.w = PK(u[14], u[15]);
#undef PK
              reinterpret_cast<uint4*>(op + cb)[0] = o0;
              if (cb + 8 < A.cout) reinterpret_cast<uint4*>(op + cb)[1] = o1;
            }
          } else if (valid) {
#pragma unroll
            for (int i = 0; i < 16; ++i)
              if (cb + i < A.cout) yp[cb + i] = __float2bfloat16_rn(u[i]);
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&tempty[acc]);  // accumulator drained: the MMA warp may start the tile after next
      if (staged) {
        // tile is complete in shared memory: all 128 threads stream it out, one contiguous image row segment
        // (wvalid * cout bf16) at a time, consecutive threads -> consecutive 16-byte chunks
        named_bar_sync(2 + grp, 128);
        const int w0 = tw * TC_TW;
        const int wvalid = min(TC_TW, A.wo - w0);
        const int row_u4 = wvalid * A.cout / 8;        // uint4 per valid row segment
        const int rows = min(TC_TH, A.ho - th * TC_TH);
        const int tile_row_u4 = TC_TW * A.cout / 8;    // uint4 pitch of a tile row in shared memory
        int r = 0, q = etid;
        while (q >= row_u4) { q -= row_u4; ++r; }
        while (r < rows) {
          const uint4 v = reinterpret_cast<const uint4*>(ob)[r * tile_row_u4 + q];
          reinterpret_cast<uint4*>(A.y + (((long long)n * A.ho + th * TC_TH + r) * A.wo + w0) * A.cout)[q] = v;
          q += 128;
          while (q >= row_u4) { q -= row_u4; ++r; }
        }
        // the buffer is reused two tiles later; the barrier at the top of that iteration orders these reads
      }
      if (m == 0) DBG(8);
      aphase ^= 1;
      if (staged) named_bar_sync(2 + grp, 128);  // everyone has read the staging tile before it is rewritten
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
  if (d->dtype != CGAT_BF16 || d->stride != 1 || d->groups != 1) return 0;
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
  if (g.smem > 227 * 1024 || g.tmem_cols > 512 || g.npairs > TC_MAX_PAIRS) return 0;
  if (g.hp * g.wp > TC_MAX_OWN * TC_PROD) return 0;
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
static int launch_tc(const void* in, int n, int hi, int wi, int gk, const void* w_krsc, int w_cin, int dgrad, int kh,
                     int kw, int pad_t, int pad_l, int hout, int wout, int gn, const float* bias, int act, void* out,
                     void* workspace, cudaStream_t st, bool prepacked = false) {
  if (!aligned16(in) || !aligned16(out) || !aligned16(workspace))
    return fail(CGAT_EALIGN, "conv tensors / workspace must be 16-byte aligned");
  const TcGeom g = geom(n, hout, wout, gk, gn, kh, kw);
  CUtensorMap map;
  if (int rc = make_rows_map(&map, in, n, hi, wi, gk, g.xs.boxe, g.xs.rows)) return rc;
  if (!prepacked) {
    const long long total = (long long)g.npairs * 2 * g.npad * 8;
    int blocks = (int)((total + 255) / 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    pack_weights_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>((const __nv_bfloat16*)w_krsc, (__nv_bfloat16*)workspace,
                                                              w_cin, g.taps, gn, gk, g.nchunk, g.npairs * 2, g.npad,
                                                              dgrad);
    if (int rc = check_launch("pack_weights_kernel")) return rc;
  }
  ConvTcArgs A{};
  A.dbg = g_dbg;
  A.wpack = (const __nv_bfloat16*)workspace;
  A.bias = bias;
  A.y = (__nv_bfloat16*)out;
  A.ho = hout; A.wo = wout; A.cin = gk; A.cout = gn; A.npad = g.npad;
  A.kh = kh; A.kw = kw; A.pad_t = pad_t; A.pad_l = pad_l;
  A.nchunk = g.nchunk; A.taps = g.taps; A.npairs = g.npairs; A.hp = g.hp; A.wp = g.wp; A.chb = g.chb;
  A.tiles_h = g.tiles_h; A.tiles_w = g.tiles_w; A.tiles = g.tiles;
  A.act = act;
  A.stages = g.stages;
  A.tmem_cols = g.tmem_cols;
  A.wbytes = (uint32_t)g.wbytes;
  A.stage_bytes = (uint32_t)g.stage_bytes;
  A.out_bytes = (uint32_t)g.out_bytes;
  A.xs = to_dev(g.xs);
  cudaError_t e = cudaFuncSetAttribute(conv_fprop_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem);
  if (e != cudaSuccess) return fail((int)e, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
  const int grid = g.tiles < sm_count() ? g.tiles : sm_count();
  conv_fprop_tc_kernel<<<grid, TC_THREADS, g.smem, st>>>(map, A);
  return check_launch("conv_fprop_tc_kernel");
}

int conv_fprop_tc_launch(const cgat_conv_desc* d, const void* x, const void* w, const float* bias, void* y,
                         void* workspace, cudaStream_t st) {
  return launch_tc(x, d->n, d->h, d->w, d->cin, w, d->cin, 0, d->kh, d->kw, d->pad_top, d->pad_left, d->ho, d->wo,
                   d->cout, bias, d->act, y, workspace, st);
}

// weights already in the packed chunk-major UMMA order (cgat_stream_prepare); `dgrad` selects the geometry
int conv_tc_packed_launch(const cgat_conv_desc* d, int dgrad, const void* in, const void* wpack, const float* bias,
                          void* out, cudaStream_t st) {
  if (!dgrad)
    return launch_tc(in, d->n, d->h, d->w, d->cin, nullptr, d->cin, 0, d->kh, d->kw, d->pad_top, d->pad_left, d->ho,
                     d->wo, d->cout, bias, d->act, out, const_cast<void*>(wpack), st, true);
  return launch_tc(in, d->n, d->ho, d->wo, d->cout, nullptr, d->cin, 1, d->kh, d->kw, d->kh - 1 - d->pad_top,
                   d->kw - 1 - d->pad_left, d->h, d->w, d->cin, nullptr, 0, out, const_cast<void*>(wpack), st, true);
}

int conv_dgrad_tc_launch(const cgat_conv_desc* d, const void* dy, const void* w, void* dx, void* workspace,
                         cudaStream_t st) {
  // dx = conv(dy, rot180(w)^T) with leading padding k-1-pad
  return launch_tc(dy, d->n, d->ho, d->wo, d->cout, w, d->cin, 1, d->kh, d->kw, d->kh - 1 - d->pad_top,
                   d->kw - 1 - d->pad_left, d->h, d->w, d->cin, nullptr, 0, dx, workspace, st);
}

// =====================================================================================================
// K3 wgrad:  dW[cout][tap][cin] = sum over pixels  dY[pix][cout] * X[pix + tap][cin]   (+ dbias)
//
// GEMM view: M = cout (<= 128, the TMEM lanes), N = taps*cin (+8 columns of ones that yield dbias), K = pixels.
// Both operands are MN-major in shared memory (pixels are the contraction) as [8-wide chunk][128 px][8 elems]
// planes: dY is re-laid from its staged rows, and the X side is an IM2COL tile built by the producer warps
// from the staged halo rows (plane index = tap*nchunk + c, i.e. exactly dW's [tap][cin] column order).  One
// tcgen05.mma (K = 16 pixels = two image rows of the tile) therefore covers ALL taps: 8 instructions per tile,
// all accumulating into the same TMEM columns.  A persistent CTA keeps that accumulator across all its tiles,
// then writes its partial sums once; a small kernel reduces the per-CTA partials in a fixed order
// (deterministic, no atomics).
// =====================================================================================================
constexpr int WG_STAGES = 2;
constexpr int WG_THREADS = 448;   // warps 0-7 re-layout, 8 MMA, 9-12 epilogue, 13 TMA
constexpr int WG_PROD = 256;
constexpr int WG_MMA_WARP = 8;
constexpr int WG_TMA_WARP = 13;
constexpr int WG_MAXU = 24;       // 16-byte units one re-layout thread moves per tile

struct WgGeom {
  int nchunk, taps, mchunk;
  int nt;       // N = round16(taps*cin + 8)
  int planes;   // planes per stage: mchunk dY planes followed by nt/8 im2col planes (>= 16 in total)
  int hp, wp;
  uint32_t tmem_cols;
  RowStage xs, ys;
  size_t stage_bytes, smem;
  int tiles_h, tiles_w, tiles;
};

static WgGeom wgeom(const cgat_conv_desc* d) {
  WgGeom g;
  g.nchunk = d->cin / 8;
  g.taps = d->kh * d->kw;
  g.mchunk = d->cout / 8;
  g.nt = (g.taps * d->cin + 8 + 15) & ~15;
  g.planes = g.mchunk + g.nt / 8;
  if (g.planes < 16) g.planes = 16;
  g.hp = TC_TH + d->kh - 1;
  g.wp = TC_TW + d->kw - 1;
  uint32_t c = 32;
  while (c < (uint32_t)g.nt) c <<= 1;
  g.tmem_cols = c;
  g.xs = row_stage(g.wp, d->cin, g.hp);
  g.ys = row_stage(TC_TW, d->cout, TC_TH);
  g.stage_bytes = (size_t)g.planes * 2048;
  g.smem = 1024 + 2 * ((size_t)g.xs.bytes + g.ys.bytes) + WG_STAGES * g.stage_bytes;
  g.tiles_h = (d->ho + TC_TH - 1) / TC_TH;
  g.tiles_w = (d->wo + TC_TW - 1) / TC_TW;
  g.tiles = d->n * g.tiles_h * g.tiles_w;
  return g;
}

struct WgArgs {
  long long* dbg;
  float* partial;  // [grid][128][nt]
  int cin, cout;
  int kh, kw, pad_t, pad_l;
  int nchunk, taps, mchunk, nt, planes;
  int tiles_h, tiles_w, tiles;
  uint32_t tmem_cols, stage_bytes;
  StageDesc xs, ys;
};

__global__ void __launch_bounds__(WG_THREADS, 1)
conv_wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_dy,
                     const WgArgs A) {
  extern __shared__ __align__(1024) unsigned char smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem);  // [WG_STAGES]
  uint64_t* empty = full + WG_STAGES;                  // [WG_STAGES]
  uint64_t* done = empty + WG_STAGES;                  // [1]
  uint64_t* sbar = done + 1;                           // [2]
  uint64_t* sfree = sbar + 2;                          // [2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(sfree + 2);
  unsigned char* s_stag = smem + 1024;
  const size_t stag_bytes = (size_t)A.xs.bytes + A.ys.bytes;
  unsigned char* s_stage = s_stag + 2 * stag_bytes;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < WG_STAGES; ++i) { mbar_init(&full[i], WG_PROD); mbar_init(&empty[i], 1); }
    mbar_init(done, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&sbar[i], 1); mbar_init(&sfree[i], WG_PROD); }
    fence_mbar_init();
    tma_prefetch_desc(&tmap_x);
    tma_prefetch_desc(&tmap_dy);
  }
  {
    // every plane starts zeroed (padding planes must be finite), then the plane of ones behind the im2col planes
    uint4* p = reinterpret_cast<uint4*>(s_stage);
    const int n16 = (int)((size_t)WG_STAGES * A.stage_bytes / 16);
    for (int i = threadIdx.x; i < n16; i += WG_THREADS) p[i] = make_uint4(0, 0, 0, 0);
    __syncthreads();
    for (int s = 0; s < WG_STAGES; ++s) {
      uint32_t* o = reinterpret_cast<uint32_t*>(s_stage + (size_t)s * A.stage_bytes +
                                                (size_t)(A.mchunk + A.taps * A.nchunk) * 2048);
      for (int i = threadIdx.x; i < 512; i += WG_THREADS) o[i] = 0x3f803f80u;  // bf16 1.0 x2
    }
    fence_proxy_async_smem();
  }
  if (warp == WG_MMA_WARP) tmem_alloc(tmem_ptr, A.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == WG_TMA_WARP) {
    if (lane == 0) {
      int it = 0;
      for (int tile = blockIdx.x; tile < A.tiles; tile += gridDim.x, ++it) {
        const int tw = tile % A.tiles_w;
        const int th = (tile / A.tiles_w) % A.tiles_h;
        const int n = tile / (A.tiles_w * A.tiles_h);
        const int buf = it & 1;
        mbar_wait(&sfree[buf], ((uint32_t)(it >> 1) & 1u) ^ 1u);
        unsigned char* dst = s_stag + (size_t)buf * stag_bytes;
        // one mbarrier phase covers both operands: arm it once with the sum of the bytes
        const uint32_t ybytes = (uint32_t)A.ys.nbox * A.ys.rows * A.ys.boxe * 2;
        issue_rows(&tmap_x, A.xs, dst, (tw * TC_TW - A.pad_l) * A.cin, th * TC_TH - A.pad_t, n, &sbar[buf], ybytes);
        for (int b = 0; b < A.ys.nbox; ++b)
          tma_load_3d(dst + A.xs.bytes + (size_t)b * A.ys.rows * A.ys.boxe * 2, &tmap_dy,
                      tw * TC_TW * A.cout + b * A.ys.boxe, th * TC_TH, n, &sbar[buf]);
      }
    }
  } else if (warp < WG_PROD / 32) {
    // Two threads per tile pixel.  The 16-byte units a thread moves are the same for every tile, so their
    // (source, destination) offsets are computed once and kept in registers (packed, 16-byte granularity);
    // per tile the thread then issues its loads in batches of 8 before the matching stores (ILP: the
    // re-layout is latency bound at 2 warps per scheduler otherwise).
    const int ptid = threadIdx.x & 127;
    const int half = threadIdx.x >> 7;
    const int hr = ptid >> 3, wc = ptid & 7;
    uint32_t unit[WG_MAXU];
    int nunit = 0;
    {
      auto add_pixel = [&](uint32_t stag_off, const StageDesc& sd, int row, int col, int c_total, int nch,
                           uint32_t dst_off) {
        int e = col * c_total;
        const int b = e / sd.boxe;
        int off = e - b * sd.boxe;
        const uint32_t box_bytes = (uint32_t)sd.rows * sd.boxe * 2;
        uint32_t src = stag_off + (uint32_t)b * box_bytes + (uint32_t)(row * sd.boxe + off) * 2;
        int left = (sd.boxe - off) >> 3;
        for (int c = 0; c < nch; ++c) {
#pragma unroll
          for (int k = 0; k < WG_MAXU; ++k)
            if (k == nunit) unit[k] = ((src >> 4) << 16) | (dst_off >> 4);
          ++nunit;
          dst_off += 2048;
          src += 16;
          if (--left == 0) { src += box_bytes - (uint32_t)sd.boxe * 2; left = sd.boxe >> 3; }
        }
      };
      const uint32_t st0 = (uint32_t)ptid * 16;
      if (half == 0) add_pixel(A.xs.bytes, A.ys, hr, wc, A.cout, A.mchunk, st0);
      for (int tap = 1 - half; tap < A.taps; tap += 2) {
        const int r = tap / A.kw, s2 = tap - r * A.kw;
        add_pixel(0, A.xs, hr + r, wc + s2, A.cin, A.nchunk, st0 + (uint32_t)(A.mchunk + tap * A.nchunk) * 2048);
      }
    }
    int stage = 0, it = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < A.tiles; tile += gridDim.x, ++it) {
      if (threadIdx.x == 0) DBG(0);
      mbar_wait(&empty[stage], phase ^ 1);
      if (threadIdx.x == 0) DBG(1);
      mbar_wait(&sbar[it & 1], (uint32_t)(it >> 1) & 1u);
      if (threadIdx.x == 0) DBG(2);
      const uint32_t sg = smem_u32(s_stag) + (uint32_t)((it & 1) * stag_bytes);
      const uint32_t st = smem_u32(s_stage) + (uint32_t)stage * A.stage_bytes;
#pragma unroll
      for (int k0 = 0; k0 < WG_MAXU; k0 += 8) {
        uint4 v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k)
          if (k0 + k < nunit) v[k] = lds128(sg + ((unit[k0 + k] >> 16) << 4));
#pragma unroll
        for (int k = 0; k < 8; ++k)
          if (k0 + k < nunit) sts128(st + ((unit[k0 + k] & 0xffffu) << 4), v[k]);
      }
      mbar_arrive(&sfree[it & 1]);
      fence_proxy_async_smem();
      if (threadIdx.x == 0) DBG(3);
      mbar_arrive(&full[stage]);
      if (++stage == WG_STAGES) { stage = 0; phase ^= 1; }
    }
  } else if (warp == WG_MMA_WARP) {
    if (lane == 0) {
      const uint32_t idesc = make_idesc_bf16(128, A.nt, 1, 1);
      int stage = 0, it = 0;
      uint32_t phase = 0;
      uint32_t accum = 0;
      for (int tile = blockIdx.x; tile < A.tiles; tile += gridDim.x, ++it) {
        DBG(4);
        mbar_wait(&full[stage], phase);
        DBG(5);
        tc_fence_after();
        const uint32_t dy_addr = smem_u32(s_stage + (size_t)stage * A.stage_bytes);
        const uint32_t im_addr = dy_addr + (uint32_t)A.mchunk * 2048;
#pragma unroll
        for (int j = 0; j < TC_TH / 2; ++j) {  // K step: image rows 2j, 2j+1 of the tile (16 pixels)
          // LBO: next 8 pixels (128 B);  SBO: next 8 couts / next 8 im2col columns (one 2048 B plane)
          const uint64_t ad = make_smem_desc(dy_addr + j * 256, 128, 2048);
          const uint64_t bd = make_smem_desc(im_addr + j * 256, 128, 2048);
          umma_bf16(tmem_base, ad, bd, idesc, accum | (j > 0));
        }
        accum = 1;
        umma_commit(&empty[stage]);
        DBG(6);
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
    float* out = A.partial + ((size_t)blockIdx.x * 128 + m) * A.nt;
    for (int c0 = 0; c0 < A.nt; c0 += 16) {
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
  if (warp == WG_MMA_WARP) {
    __syncwarp();
    tmem_dealloc(tmem_base, A.tmem_cols);
  }
}

// dw[co][k] = sum_cta partial[cta][co][k]  (k < taps*cin);  dbias[co] = sum_cta partial[cta][co][taps*cin]
// 64 outputs x 4 CTA-slices per block; fixed summation order (deterministic).
__global__ void __launch_bounds__(256) wgrad_reduce_kernel(const float* __restrict__ partial, float* __restrict__ dw,
                                                           float* __restrict__ dbias, int ncta, int cout, int kdim,
                                                           int nt) {
  __shared__ float s[4][64];
  const int nw = cout * kdim;
  const int total = nw + (dbias ? cout : 0);
  const int o = blockIdx.x * 64 + (threadIdx.x & 63);
  const int slice = threadIdx.x >> 6;
  float acc = 0.f;
  int co = 0;
  if (o < total) {
    int col;
    if (o < nw) { co = o / kdim; col = o - co * kdim; } else { co = o - nw; col = kdim; }
    const float* p = partial + (size_t)co * nt + col;
#pragma unroll 4
    for (int b = slice; b < ncta; b += 4) acc += p[(size_t)b * 128 * nt];
  }
  s[slice][threadIdx.x & 63] = acc;
  __syncthreads();
  if (slice == 0 && o < total) {
    const float v = (s[0][threadIdx.x] + s[1][threadIdx.x]) + (s[2][threadIdx.x] + s[3][threadIdx.x]);
    if (o < nw) dw[o] = v; else dbias[co] = v;
  }
}

int conv_wgrad_tc_supported(const cgat_conv_desc* d) {
  if (d->dtype != CGAT_BF16 || d->stride != 1 || d->groups != 1) return 0;
  if (d->cin % 8 != 0 || d->cout % 8 != 0 || d->cout > 128) return 0;
  const WgGeom g = wgeom(d);
  if (g.nt > 256 || g.smem > 227 * 1024) return 0;
  // units per re-layout thread: the dY pixel plus every other tap (first half) must fit the register table
  if (g.mchunk + ((g.taps + 1) / 2) * g.nchunk > WG_MAXU) return 0;
  return 1;
}

size_t conv_wgrad_tc_workspace(const cgat_conv_desc* d) {
  const WgGeom g = wgeom(d);
  return (size_t)148 * 128 * g.nt * sizeof(float);
}

int conv_wgrad_tc_launch(const cgat_conv_desc* d, const void* x, const void* dy, float* dw, float* dbias,
                         void* workspace, cudaStream_t st, int* ncta_out, int* nt_out) {
  if (!aligned16(x) || !aligned16(dy) || !aligned16(workspace))
    return fail(CGAT_EALIGN, "conv tensors / workspace must be 16-byte aligned");
  const WgGeom g = wgeom(d);
  CUtensorMap mx, mdy;
  if (int rc = make_rows_map(&mx, x, d->n, d->h, d->w, d->cin, g.xs.boxe, g.xs.rows)) return rc;
  if (int rc = make_rows_map(&mdy, dy, d->n, d->ho, d->wo, d->cout, g.ys.boxe, g.ys.rows)) return rc;
  WgArgs A{};
  A.dbg = g_dbg;
  A.partial = (float*)workspace;
  A.cin = d->cin; A.cout = d->cout;
  A.kh = d->kh; A.kw = d->kw; A.pad_t = d->pad_top; A.pad_l = d->pad_left;
  A.nchunk = g.nchunk; A.taps = g.taps; A.mchunk = g.mchunk; A.nt = g.nt; A.planes = g.planes;
  A.tiles_h = g.tiles_h; A.tiles_w = g.tiles_w; A.tiles = g.tiles;
  A.tmem_cols = g.tmem_cols; A.stage_bytes = (uint32_t)g.stage_bytes;
  A.xs = to_dev(g.xs); A.ys = to_dev(g.ys);
  const int grid = g.tiles < sm_count() ? g.tiles : sm_count();
  if (grid > 148) return fail(CGAT_EUNSUPPORTED, "wgrad workspace sized for <= 148 CTAs");
  cudaError_t e = cudaFuncSetAttribute(conv_wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem);
  if (e != cudaSuccess) return fail((int)e, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
  conv_wgrad_tc_kernel<<<grid, WG_THREADS, g.smem, st>>>(mx, mdy, A);
  if (int rc = check_launch("conv_wgrad_tc_kernel")) return rc;
  if (ncta_out != nullptr) {  // caller reduces the partial sums itself
    *ncta_out = grid;
    *nt_out = g.nt;
    return 0;
  }
  const int kdim = g.taps * d->cin;
  const int total = d->cout * kdim + (dbias ? d->cout : 0);
  wgrad_reduce_kernel<<<(total + 63) / 64, 256, 0, st>>>((const float*)workspace, dw, dbias, grid, d->cout, kdim, g.nt);
  return check_launch("wgrad_reduce_kernel");
}

}  // namespace cgat
