// K6 / K7: one conv-mapped conv-GAT stream -- node conv AND graph attention -- as ONE kernel per direction.
//
// Why: the unfused path (conv_tc.cu -> attn_h2.cu) writes the projected features Wh[N,H,W,heads*nodes*co] to HBM
// and reads them back (twice in the backward, plus d(Wh) once more for wgrad): 3/4 of the step's HBM traffic is
// that intermediate.  Here Wh never leaves the SM:
//
//   x --TMA--> "row planes" in shared memory
//        --tcgen05.mma (M = 128 pixels, N = heads*nodes*co, K = 9 taps * cin)--> Wh in TMEM (fp32)
//        --tcgen05.ld: TMEM lane = pixel, so each thread receives exactly ITS pixel's record-->
//        per-pixel attention in registers (attn_math.cuh, fp32): logits, LeakyReLU, mask, soft-max over the
//        neighbours, aggregation, adjacency mix, ELU (reference baseline_model.py:127-160)
//   forward : head merge through shared memory, coalesced bf16 store of out
//   backward: the forward is recomputed, then d(Wh) (bf16) goes straight into the shared-memory planes that
//             are the A operand of the wgrad MMA (M = couts, K = pixels); the wgrad accumulators stay in TMEM
//             across all tiles of the persistent CTA.  Only x and d(out) (train mode: x and y) are read from
//             HBM; only per-CTA partial sums of the parameter gradients are written.
//
// Interleaved planes (the im2col that is NOT materialised).  x arrives PADDED CHUNK-PLANAR, [n][cin/8][h][wp][8] with
// wp = tiles_w * 8 + 2 (one zero pixel left, zeros right; cgat_loader_gather_planar / cgat_records_to_planar write it).
// ONE 5-D TMA box per tile -- (8 pixels x 8 channels = 128 B; horizontal tap s: 3, stride ONE pixel, i.e. the dimension
// overlaps the innermost one; chunk c; 18 rows; 1 image) -- lands as  stage[row 18][plane q = c*3 + s][128 B]:  every
// 128-byte unit is one SWIZZLE_NONE core matrix (8 pixels of a tile row x 8 channels, 128-byte aligned), a row of the
// stage holds the 9 planes side by side (1152 B), rows outside the image are zero-filled by the TMA unit and the
// horizontal padding is the zero columns of the global layout.  Flattening  j = r * 9 + q  (r = vertical tap = one row
// down) gives address j * 128: BOTH MMAs see uniform strides --
//   fprop  A (K-major): K-chunk pair (j, j+1) at LBO = 128, pixel rows at SBO = 1152: 14 instructions for the 27 chunks
//          (chunk 27 meets zero weights) + 1 for the bias (a constant plane of ones against the bias K-chunk);
//   wgrad  B (MN-major): N = 27 chunks x 8 channels = 216 columns at SBO = 128, K = 16 pixels = two tile rows at
//          LBO = 1152: ONE accumulator and 8 instructions per tile.
// The first layout of this kernel kept one plane per (s, c) with 2304-byte plane stride: wgrad then needed one
// accumulator per vertical tap (24 instructions, N = 80 each, and a ~280-cycle penalty whenever consecutive tcgen05.mma
// switch accumulators) and the TMA unit 9 boxes per tile; with a ~105-cycle floor per tcgen05.mma whatever N <= 128
// (tools/microbench/mma_rate.cu) the tensor pipe was the kernel's bottleneck (11 300 of 11 400 cycles per tile pair).
// dbias no longer comes out of the wgrad MMA (no plane of ones among its columns): the attention threads sum d(Wh)
// over their pixels themselves.
//
// Shared memory: [header | packed weights | d(Wh) buffers (ring of 2-3) | x stages (ring of 4)].  The x planes of a
// tile live from its TMA load until its wgrad MMAs have read them; a d(Wh) buffer lives from the head exchange of
// its tile until the same wgrad.  The two rings are independent, which is what lets 4 x stages + 3 d(Wh) buffers fit.
//
// Warp roles (512 threads, 1 CTA / SM): warp 0 TMA issuer, warp 1 MMA issuer + TMEM allocator, warps 4-15 three
// attention groups (group g owns heads g, g+3, ...; warp % 4 selects the TMEM lane quarter).  The kernel launches
// with 128 registers per thread; setmaxnreg moves warpgroup 0's budget to the attention warpgroups
// (128 * 32 + 384 * 160 = 65 536 = the CTA's pool, exactly).
#include "tc_common.cuh"
#include "attn_common.cuh"
#include <cstdlib>
#include <type_traits>

namespace cgat {

constexpr int LF_TH = 16, LF_TW = 8;  // output tile -> M = 128 pixels
constexpr int LF_THREADS = 512;
constexpr int LF_ATT_WARP0 = 4;
constexpr int LF_GROUPS = 3;
constexpr int LF_TMA_WARP = 0, LF_MMA_WARP = 1;
constexpr int LF_MAXSTG = 4;        // x-plane stages (TMA prefetch depth)
constexpr int LF_MAXDW = 3;         // d(Wh) buffers (backward kernels)
constexpr int LF_FP_COL0 = 256;     // TMEM: wgrad accumulators at columns [0,256), fprop accumulators at 256 + 128*acc
constexpr int LF_PR = LF_TH + 2;    // stage rows (vertical halo)
constexpr int LF_ROW = LF_TW * 16;  // 128 B: 8 pixels x 8 channels = one core matrix
constexpr int LF_ONES = 4096;       // constant A operand of the bias MMA: 16 core matrices of ones, then 16 of zeros
// Range guard of the paired-half train kernel: largest |s1|, |s2| (attention score halves) it accepts.  The logits
// s1_i + s2_j then stay below 16, where fp16 resolves 2^-7: e^(logit - max) is good to < 1 %.  Beyond it -- or when any
// sum the kernel produces is not finite (fp16 overflows at 65 504) -- the kernel raises LfArgs::guard and the step is
// re-run by the fp32 instantiation (cgat_layer_train_fp32; tests/test_gpu_config2_pinned.py drives it with inputs x 100).
constexpr float LF_SMAX = 8.f;
constexpr int LF_HDR = 8704;        // barriers + parameters (fp32 and packed-half2 copies)
constexpr int LF_SLOT = 64;         // floats per attention warp in the end-of-kernel reduction scratch (>= RG + 3)

struct LfArgs {
  long long* dbg;              // developer aid: clock64() timeline of CTA 0 (cgat_layer_debug_timeline)
  const __nv_bfloat16* x;      // input records [n][h][w][cin] (read through the tensor map; the pointer serves L2 prefetches)
  const __nv_bfloat16* wpack;  // [nj + 3][npad][8] chunk-major packed dense weights (cgat_stream_prepare)
  const float* bias;           // [cout] dense bias
  const float* a;              // [heads][2co]
  const float* adj;            // [heads][nodes][nodes]
  const uint8_t* mask;         // [nodes][nodes] or NULL
  __nv_bfloat16* out;          // fwd
  const __nv_bfloat16* dout;   // bwd
  __nv_bfloat16* dwh;          // bwd, optional: d(Wh) [n][h][w][cout] for a following dgrad
  float* partial;              // bwd: [grid][128][nt] wgrad partial sums
  float* ga;                   // bwd: [heads][2co]   accumulated into
  float* gadj;                 // bwd: [heads][nodes*nodes] accumulated into
  float* gbias;                // bwd: [heads][co + 2] accumulated into: sum of d(Wh) per output channel, sum of ds1, of ds2
  const __nv_bfloat16* y;      // train mode (bwd kernel): target, same layout as out; d(out) is derived in-kernel
  float* loss_out;             // train mode: scalar loss, accumulated into
  float* mse_out;              // train mode, optional: mean squared error alone (the reference's running train loss)
  float* guard;                // paired-half train kernel, optional: guard[0] := 1 when the step left the fp16 math's range
  const float* run_if;         // optional: the whole launch is a no-op unless run_if[0] != 0 (the fp32 re-run of such a step)
  float out_scale;             // bwd: factor applied to the gradient sums when they leave the kernel (PAIR: 1/numel)
  float lambda, inv_n;         // train mode: loss = mean((out-y)^2) - lambda*mean(out); inv_n = 1/numel(out)
  int h, w, wp, cin, cout, ext, npad, heads, merge, apply_elu;  // wp: padded row of x, lf_padded_width(w);  // ext: score rows behind the cout feature rows (0: none)
  float alpha;
  int nchunk, np, nj, mchunk, nt;  // np = 3*nchunk planes per stage row, nj = 3*np K-chunks (fprop) = N-chunks (wgrad); nt = 9*ci
  int rowp;                        // bytes of a stage row: np * 128
  int tiles_h, tiles_w, tiles, nstg, ndw;
  int rows_pad;                 // rows of a CTA's partial-sum slot: cout + ext rounded up to a lane quarter (32)
  uint32_t wbytes, stage_bytes, dw_bytes;  // bytes of one x stage / of one d(Wh) buffer (0 in the forward kernel)
};

// timeline events of CTA 0's first LF_DBG_TILES tiles: [tile][event]
//  0 P:loop top  1 P:stage empty  2 P:x landed  3 P:im2col done | 4 M:im2col full  5 M:acc free  6 M:fprop issued
//  7 M:dWh planes full  8 M:wgrad issued | 9 A:Wh ready  10 A:Wh in registers  11 A:forward done  12 A:after exchange
//  13 A:backward done  14 A:tile done           (A = attention group 0, warp 0, lane 0)
constexpr int LF_DBG_TILES = 16, LF_DBG_EVENTS = 16;
// The stamps cost registers in the 32-register TMA / MMA warps (their spills go to L2: the CTA's shared memory leaves
// ~8 KB of L1), so they are compiled in only with -DCGAT_LF_TIMELINE (make TIMELINE=1; tools/layer_timeline.py).
#ifdef CGAT_LF_TIMELINE
#define LDBG(ev)                                                                                                 \
  do {                                                                                                           \
    if (A.dbg != nullptr && blockIdx.x == 0 && it < LF_DBG_TILES) A.dbg[it * LF_DBG_EVENTS + (ev)] = clock64();   \
  } while (0)
#else
#define LDBG(ev) do { (void)it; } while (0)
#endif
// whole-kernel stamps of CTA 0 in row 14: 0 entry, 1 init done, 2 A:regs granted, 3 A:tiles done, 4 A:sums flushed,
// 5 A:wgrad complete, 6 A:partials written, 7 all warps joined, 8 exit
#ifdef CGAT_LF_TIMELINE
#define LDBGX(ev)                                                                            \
  do {                                                                                       \
    if (A.dbg != nullptr && blockIdx.x == 0) A.dbg[14 * LF_DBG_EVENTS + (ev)] = clock64();   \
  } while (0)
#else
#define LDBGX(ev) do { } while (0)
#endif
static long long* g_lf_dbg = nullptr;

__device__ __forceinline__ void tmem_ld8_nowait(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
}
template <int N>
__device__ __forceinline__ void setmaxnreg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N>
__device__ __forceinline__ void setmaxnreg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }

// mbarrier wait of a CONVERGED warp: the vote makes the loop condition warp-uniform, so the compiler still knows the warp
// is converged afterwards (a per-lane `while (!try_wait)` ends that knowledge: every tcgen05.mma behind it is then wrapped
// in ELECT / VOTEU / R2UR sequences)
__device__ __forceinline__ void mbar_wait_warp(uint64_t* bar, uint32_t parity) {
  while (!__any_sync(0xffffffffu, mbar_try_wait(bar, parity))) {
  }
}

__device__ __forceinline__ uint4 lf_lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void lf_sts128(uint32_t addr, const uint4& v) {
  asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// bf16 record held in registers (uint4 per 8 elements) -> fp32
template <int N>
__device__ __forceinline__ void unpack_rec(const uint4* __restrict__ p, float (&r)[N]) {
#pragma unroll
  for (int i = 0; i < N / 8; ++i) {
    const uint32_t w[4] = {p[i].x, p[i].y, p[i].z, p[i].w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      r[8 * i + 2 * k] = __uint_as_float(w[k] << 16);
      r[8 * i + 2 * k + 1] = __uint_as_float(w[k] & 0xffff0000u);
    }
  }
}

template <int NODES, int CO, bool SPATIAL, bool BWD, bool MASKED, bool PAIR>
__global__ void __launch_bounds__(LF_THREADS, 1)
layer_kernel(const __grid_constant__ CUtensorMap tmap_x, const LfArgs A) {
  constexpr int REC = NODES * CO;          // elements of one head's pixel record
  constexpr int RGA = 2 * CO + NODES * NODES;  // adjacency-grad + a-grad values per head ...
  constexpr int RG = RGA + CO + 2;             // ... + sum of d(Wh) per output channel (dbias), sum of ds1, sum of ds2
  static_assert(REC % 8 == 0, "record must be a multiple of 16 bytes");
  static_assert(RG + 3 <= LF_SLOT, "reduction slot too small");
  if (A.run_if != nullptr) {  // the fp32 re-run of a guarded train step: a no-op unless the paired-half kernel asked for it
    griddep_wait();           // (launched with programmatic serialisation: that kernel may still be running)
    if (*A.run_if == 0.f) return;  // uniform over the grid; nothing has been set up yet
  }
  griddep_launch();  // whatever follows in the stream may be scheduled as soon as this grid's CTAs leave their SMs
  extern __shared__ __align__(1024) unsigned char smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem);  // [4]  TMA -> MMA         (x planes landed)
  uint64_t* empty = full + LF_MAXSTG;                  // [4]  MMA -> TMA         (x stage reusable)
  uint64_t* dyfull = empty + LF_MAXSTG;                // [3]  attention -> MMA   (d(Wh) buffer written)
  uint64_t* dwfree = dyfull + LF_MAXDW;                // [3]  MMA -> attention   (d(Wh) buffer consumed by its wgrad)
  uint64_t* tfull = dwfree + LF_MAXDW;                 // [2]  MMA -> attention   (Wh accumulator ready)
  uint64_t* tempty = tfull + 2;                        // [2]  attention -> MMA   (accumulator drained)
  uint64_t* wbar = tempty + 2;                         // [1]
  uint64_t* done = wbar + 1;                           // [1]
  uint64_t* dhread = done + 1;                         // [1]  attention -> attention (d(out) read by every group; PAIR)
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(dhread + 1);
  float* s_a = reinterpret_cast<float*>(smem + 768);             // [MAX_HEADS][2*CO]
  float* s_adj = reinterpret_cast<float*>(smem + 1280);          // [MAX_HEADS][NODES*NODES]
  uint64_t* s_mask = reinterpret_cast<uint64_t*>(smem + 3328);   // [NODES]
  float* s_gacc = reinterpret_cast<float*>(smem + 3392);         // [MAX_HEADS][RG]
  __half2* s_a2 = reinterpret_cast<__half2*>(smem + 6144);       // [MAX_HEADS][2*CO]       both lanes = the parameter
  __half2* s_adj2 = reinterpret_cast<__half2*>(smem + 6656);     // [MAX_HEADS][NODES*NODES]
  static_assert(MAX_HEADS * 2 * CO * 4 <= 512 && MAX_HEADS * NODES * NODES * 4 <= 2048 && NODES * 8 <= 64 &&
                    3392 + (MAX_HEADS * RG + 2) * 4 <= 6144 && 6656 + MAX_HEADS * NODES * NODES * 4 <= LF_HDR, "parameter block overflows the header");
  unsigned char* s_w = smem + LF_HDR;
  // d(Wh) buffers come BEFORE the x stages: the wgrad A operand always spans 16 planes (M = 128 rows), so the last
  // buffer's unused rows read on into the x stages (finite data; those accumulator rows are never stored)
  unsigned char* s_dw = s_w + ((A.wbytes + 127u) & ~127u);
  unsigned char* s_stage = s_dw + (size_t)A.ndw * A.dw_bytes;
  // (a stage is 18 rows + 128 zeroed bytes: fprop's K-chunk 27 -- the odd 27th chunk's partner, zero weights -- reads
  // the first core matrix of "row 18" for the tile's last pixel row, and 0 x garbage may be NaN)
  unsigned char* s_one = s_stage + (size_t)A.nstg * A.stage_bytes;
  float4* s_slab = reinterpret_cast<float4*>(s_one + LF_ONES);  // fwd: [group][REC/4][128]

  // (the shuffle makes the warp index provably warp-uniform: role branches and everything computed under them can then
  // use the uniform datapath -- what CUTLASS calls canonical_warp_idx_sync)
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int nact = A.heads < LF_GROUPS ? A.heads : LF_GROUPS;  // attention groups that own at least one head

  if (threadIdx.x == 0) LDBGX(0);
#ifdef CGAT_LF_TIMELINE
  if (threadIdx.x == 0 && A.dbg != nullptr) {  // every CTA: start / end on the global timer (ns), entries [256 + 2 * cta]
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    A.dbg[256 + 2 * blockIdx.x] = (long long)t;
  }
#endif
  if (threadIdx.x == 0) {
    for (int i = 0; i < LF_MAXSTG; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < LF_MAXDW; ++i) { mbar_init(&dyfull[i], 128 * nact); mbar_init(&dwfree[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 128 * nact); }
    mbar_init(wbar, 1);
    mbar_init(done, 1);
    mbar_init(dhread, 128 * nact);
    fence_mbar_init();
    tma_prefetch_desc(&tmap_x);
  }
  __syncthreads();  // barriers are live: the TMA thread starts fetching while everyone else initialises shared memory

  auto issue_tile = [&](int tile, int stage) {  // ONE box per tile: [18 rows][np planes][128 B] (TMA thread only)
    const int tw = tile % A.tiles_w;
    const int th = (tile / A.tiles_w) % A.tiles_h;
    const int n = tile / (A.tiles_w * A.tiles_h);
    mbar_arrive_expect_tx(&full[stage], (uint32_t)(LF_PR * A.rowp));
    // coordinates (element in the padded row, tap s, chunk c, row, image): padded column tw*8 = image column tw*8 - 1
    tma_load_5d(s_stage + (size_t)stage * A.stage_bytes, &tmap_x, tw * LF_TW * 8, 0, 0, th * LF_TH - 1, n, &full[stage]);
  };
  if (warp == LF_TMA_WARP && lane == 0) {
    // the first tile's planes before the (tile-independent) weights: the x planes are written whole by the TMA unit
    // (padding included), so they need no initialisation and their HBM latency overlaps the set-up below
    // (only the first tile: a TMA instruction takes ~100 cycles to issue, and everyone waits for this thread below)
    // x does not depend on the kernel before this one in the stream (cgat_stream_prepare packs the weights): under
    // programmatic dependent launch the first two tiles are on their way before that kernel has finished
    if ((int)blockIdx.x < A.tiles) issue_tile(blockIdx.x, 0);
    if (A.nstg > 1 && (int)(blockIdx.x + gridDim.x) < A.tiles) issue_tile(blockIdx.x + gridDim.x, 1);
    griddep_wait();
    mbar_arrive_expect_tx(wbar, A.wbytes);
    bulk_g2s(s_w, A.wpack, A.wbytes, wbar);
  }
  if (warp != LF_TMA_WARP) {
    // (warp 0 is busy issuing TMA instructions, ~100 cycles each: the other 15 warps initialise shared memory)
    constexpr int NI = LF_THREADS - 32;
    const int ti = threadIdx.x - 32;
    griddep_wait();  // a, adj come from the kernel before this one
    for (int i = ti; i < A.heads * 2 * CO; i += NI) { s_a[i] = A.a[i]; s_a2[i] = __float2half2_rn(A.a[i]); }
    for (int i = ti; i < A.heads * NODES * NODES; i += NI) {
      s_adj[i] = A.adj[i];
      s_adj2[i] = __float2half2_rn(A.adj[i]);
    }
    for (int i = ti; i < MAX_HEADS * RG + 2; i += NI) s_gacc[i] = 0.f;
    if (ti < NODES) {
      uint64_t mrow = 0;
      for (int j = 0; j < NODES; ++j)
        if (A.mask == nullptr || A.mask[ti * NODES + j] != 0) mrow |= (1ull << j);
      s_mask[ti] = mrow;
    }
    // d(Wh) buffers start zeroed (rows the attention groups never write must hold finite numbers); the constant
    // operand of the bias MMA: 2 KB of bf16 ones (its fprop K-chunk holds the bias, hi + lo bf16 parts), 2 KB of zeros
    uint4* p = reinterpret_cast<uint4*>(s_dw);
    const int n16 = (int)((size_t)A.ndw * A.dw_bytes / 16);
    for (int i = ti; i < n16; i += NI) p[i] = make_uint4(0, 0, 0, 0);
    for (int i = ti; i < A.nstg * 8; i += NI)  // the 128 zero bytes behind every stage's 18 rows (never written by the TMA unit)
      reinterpret_cast<uint4*>(s_stage + (size_t)(i >> 3) * A.stage_bytes + (size_t)LF_PR * A.rowp)[i & 7] = make_uint4(0, 0, 0, 0);
    uint4* o = reinterpret_cast<uint4*>(s_one);
    for (int i = ti; i < LF_ONES / 16; i += NI) {
      const uint32_t v = i < LF_ONES / 32 ? 0x3f803f80u : 0u;  // bf16 1.0 x2
      o[i] = make_uint4(v, v, v, v);
    }
    fence_proxy_async_smem();
  }
  if (warp == LF_MMA_WARP) tmem_alloc(tmem_ptr, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  if (threadIdx.x == 0) LDBGX(1);

  if (warp < LF_ATT_WARP0) {
    setmaxnreg_dec<32>();
    if (warp == LF_TMA_WARP && lane == 0) {
      // ===================== TMA issuer: 9 column-plane boxes per tile (the first tile is on its way) ====
      int it = 0, stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < A.tiles; tile += gridDim.x, ++it) {
        LDBG(0);
        if (it > (A.nstg > 1 ? 1 : 0)) {  // (the first two tiles went out during set-up)
          if (it >= A.nstg) {
#ifdef CGAT_LF_SLEEPY_TMA
            while (!mbar_try_wait(&empty[stage], phase ^ 1u)) __nanosleep(64);
#else
            mbar_wait(&empty[stage], phase ^ 1u);
#endif
          }
          LDBG(1);
          issue_tile(tile, stage);
        }
        if (++stage == A.nstg) { stage = 0; phase ^= 1u; }
      }
    } else if (warp == LF_MMA_WARP) {
      // ===================== MMA issuer: the whole warp runs this converged, one elected lane issues (tc_common.cuh) ====
      const uint32_t idesc_f = make_idesc_bf16(128, A.npad, 0, 0);
      // wgrad N: the nj chunks rounded up to a multiple of 16 columns (M = 128 needs that); the odd 28th chunk is the x
      // data one stage row further down -- finite numbers into accumulator columns nobody reads
      // (kind::f16 wants both operands in the same format: an fp16 A = d(Wh) against the bf16 x planes is an illegal
      // instruction, so the train kernel converts its packed-half d(Wh) to bf16)
      const uint32_t idesc_w = make_idesc_bf16(128, (A.nj * 8 + 15) & ~15, 1, 1);
      const uint32_t w_addr = smem_u32(s_w);
      const uint32_t b_lbo = (uint32_t)A.npad * 16;
      uint32_t wg_accum = 0;
      int wstage = 0, wbuf = 0, wj = 0;  // wj: next tile whose wgrad is to be issued; its x stage and d(Wh) buffer
      uint32_t wphase = 0, bphase = 0;
      auto wgrad = [&](int j) {  // tiles are retired in order: (wstage, wphase), (wbuf, bphase) follow tile j
        mbar_wait_warp(&dyfull[wbuf], bphase);
        { const int it = j; LDBG(7); }
        tc_fence_after();
        const uint32_t dy_addr = smem_u32(s_dw) + (uint32_t)wbuf * A.dw_bytes;
        const uint32_t im_addr = smem_u32(s_stage) + (uint32_t)wstage * A.stage_bytes;
        // A = d(Wh) planes (MN-major: 8-cout chunks 2048 B apart, pixel groups = tile rows 128 B apart); B = the stage
        // (MN-major: the nj chunks j = r * np + q 128 B apart, pixel groups = stage rows)
        constexpr int ROWP = 3 * (NODES * CO / 8) * LF_ROW;  // (ci == co: the stage geometry is a compile-time constant)
        const uint64_t ad0 = make_smem_desc(dy_addr, 128, 2048);
        const uint64_t bd0 = make_smem_desc(im_addr, ROWP, LF_ROW);
#pragma unroll
        for (int jj = 0; jj < LF_TH / 2; ++jj)  // K step: image rows 2jj, 2jj+1 of the tile (16 pixels)
          umma_bf16_warp(tmem_base, ad0 + (uint64_t)((jj * 256) >> 4), bd0 + (uint64_t)((2 * jj * ROWP) >> 4), idesc_w,
                         wg_accum | (uint32_t)(jj > 0));
        wg_accum = 1;
        umma_commit_warp(&empty[wstage]);
        umma_commit_warp(&dwfree[wbuf]);
        { const int it = j; LDBG(8); }
#ifdef CGAT_LF_TIMELINE
        if (A.dbg != nullptr && A.dbg[255] != 0) {
          mbar_wait_warp(&dwfree[wbuf], bphase);
          { const int it = j; LDBG(3); }
        }
#endif
        if (++wstage == A.nstg) { wstage = 0; wphase ^= 1u; }
        if (++wbuf == A.ndw) { wbuf = 0; bphase ^= 1u; }
      };
      mbar_wait_warp(wbar, 0);
      LDBGX(9);
      int it = 0, stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < A.tiles; tile += gridDim.x, ++it) {
        const int acc = it & 1;
        if constexpr (PAIR) {
          // tile pairs: fprop(p0) fprop(p1) wgrad(p0) fprop(p2) wgrad(p1) ...  The wgrad of pair p-2 goes BEFORE the fprop
          // of pair p: its d(Wh) buffers are complete by then (pair p-1 is in its math), and the d(Wh) ring (3 buffers)
          // needs pair p-2's first buffer back before pair p-1 publishes its head exchange
          if ((it & 1) == 0)
            while (wj < it - 2) wgrad(wj++);
        }
        mbar_wait_warp(&full[stage], phase);
        LDBG(4);
        if (it == 0) LDBGX(10);
        mbar_wait_warp(&tempty[acc], (((uint32_t)it >> 1) & 1u) ^ 1u);
        LDBG(5);
        tc_fence_after();
        const uint32_t im_addr = smem_u32(s_stage) + (uint32_t)stage * A.stage_bytes;
        const uint32_t d_addr = tmem_base + LF_FP_COL0 + (uint32_t)acc * 128;
        // (the served shapes have ci == co, so the chunk count is a compile-time constant and the descriptor pairs are
        // the first one plus immediates: a handful of uniform-datapath instructions per tcgen05.mma)
        constexpr int NP = 3 * (NODES * CO / 8), NJ = 3 * NP, ROWP = NP * LF_ROW;
        const uint64_t ad0 = make_smem_desc(im_addr, LF_ROW, ROWP);  // K-chunk pair (j, j+1) 128 B apart, tile rows ROWP apart
        const uint64_t bd0 = make_smem_desc(w_addr, b_lbo, 128);
        const uint32_t b_step = (2 * b_lbo) >> 4;
#pragma unroll
        for (int i = 0; i < (NJ + 1) / 2; ++i)  // K = 16: chunks 2i, 2i+1 of j = r * NP + q
          umma_bf16_warp(d_addr, ad0 + (uint64_t)((2 * i * LF_ROW) >> 4), bd0 + (uint64_t)(i * b_step), idesc_f, i > 0);
        // + bias: a plane of ones (then zeros) against the K-chunk pair behind the last x pair
        umma_bf16_warp(d_addr, make_smem_desc(smem_u32(s_one), LF_ONES / 2, 128), bd0 + (uint64_t)(((NJ + 1) / 2) * b_step), idesc_f, 1);
        umma_commit_warp(&tfull[acc]);
        LDBG(6);
#ifdef CGAT_LF_TIMELINE
        if (A.dbg != nullptr && A.dbg[255] != 0) {  // developer probe: execution time of the fprop MMAs in isolation
          mbar_wait_warp(&tfull[acc], ((uint32_t)it >> 1) & 1u);
          LDBG(2);
        }
#endif
        if constexpr (!BWD) {
          umma_commit_warp(&empty[stage]);
        } else if constexpr (!PAIR) {
          while (wj < it) wgrad(wj++);
        }
        if (++stage == A.nstg) { stage = 0; phase ^= 1u; }
      }
      if constexpr (BWD) {
        while (wj < it) wgrad(wj++);
        umma_commit_warp(done);
      }
    }
  } else if (warp >= LF_ATT_WARP0 && warp < LF_ATT_WARP0 + 4 * LF_GROUPS) {
    // ===================== attention groups =====================
    setmaxnreg_inc<160>();
    if (threadIdx.x == LF_ATT_WARP0 * 32) LDBGX(2);
    const int g = (warp - LF_ATT_WARP0) >> 2;
    const int lg = warp & 3;
    const int m = lg * 32 + lane;  // TMEM lane = pixel of the tile
    const int hrow = m >> 3, wcol = m & 7;
    const bool concat = A.merge == CGAT_MERGE_CONCAT;
    const int out_rec = concat ? A.heads * REC : REC;
    const bool vec_io = !concat || SPATIAL;  // a head's record is contiguous in the output record
    if (g < nact) {
      float gacc[BWD ? RG : 1];
      float loss_acc = 0.f, mse_acc = 0.f;
      int cur_head = -1;
#pragma unroll
      for (int i = 0; i < (BWD ? RG : 1); ++i) gacc[i] = 0.f;
      auto flush = [&](int head) {
        if (!BWD || head < 0) return;
#pragma unroll
        for (int i = 0; i < (BWD ? RG : 1); ++i) {
          const float s = warp_sum(gacc[i]);
          if (lane == 0) atomicAdd(&s_gacc[head * RG + i], s);
          gacc[i] = 0.f;
        }
      };
      const float inv_heads = 1.f / (float)A.heads;
      if constexpr (PAIR) {
        // ============ train mode, two tiles per pass: pixel m of tile A in the low half2 lane, of tile B in the
        // high lane (A, B = consecutive tiles of this CTA, in the two fprop accumulators).  One head per group.
        // All quantities are O(1) in fp16: d(out) is carried WITHOUT its 1/numel factor (out_scale re-applies it
        // to everything that leaves the kernel).
        //   * the scores s1, s2 arrive from the fprop MMA (extra output columns W.a); their gradients leave as extra
        //     rows of the wgrad A operand: no a-terms in registers, no d(a) accumulators (stream_ops.cu finishes them);
        //   * head exchange: every group publishes ELU(z) of its head (fp16) in the d(Wh) planes, then forms out, the
        //     loss terms and d(out) for ONE third of the record and publishes that in the (still unused) score-row
        //     planes: two named barriers per pair, no redundant work between the groups. ============
        constexpr int CH = REC / 8;                 // 16-byte chunks of a head's record
        constexpr int SL = (2 * NODES + 7) / 8;     // 8-column loads / 16-byte chunks of a head's scores
        const int k = g;
        cur_head = g;
        const __half2* adj2 = s_adj2 + k * NODES * NODES;
        const __half2 alpha2 = __float2half2_rn(A.alpha);
        const __half2 gs2 = __float2half2_rn(inv_heads);
        __half2 g2[NODES * NODES], gb2[CO + 2];
        __half2 smax2 = H2::zero();  // range guard: largest |s1|, |s2| this thread has seen (see LF_SMAX)
#pragma unroll
        for (int i = 0; i < NODES * NODES; ++i) g2[i] = H2::zero();
#pragma unroll
        for (int i = 0; i < CO + 2; ++i) gb2[i] = H2::zero();
        const uint32_t lane_off = (uint32_t)(lg * 32) << 16;
        const uint32_t xplane = (uint32_t)A.mchunk * 2048;  // the score-row planes (after the feature planes)
        // tile coordinates and the d(Wh) ring position are CARRIED from tile to tile (every runtime division or modulo
        // is ~25 instructions; the loop top used to cost 1 300 cycles per pair)
        const int d_tw = (int)gridDim.x % A.tiles_w, d_q = (int)gridDim.x / A.tiles_w, d_th = d_q % A.tiles_h, d_n = d_q / A.tiles_h;
        int c_tw = (int)blockIdx.x % A.tiles_w, c_th = ((int)blockIdx.x / A.tiles_w) % A.tiles_h,
            c_n = (int)blockIdx.x / (A.tiles_w * A.tiles_h);
        auto advance = [&](int& tw, int& th, int& n) {  // + gridDim.x tiles
          tw += d_tw;
          const int cw = tw >= A.tiles_w ? 1 : 0;
          tw -= cw ? A.tiles_w : 0;
          th += d_th + cw;
          const int ch = th >= A.tiles_h ? 1 : 0;
          th -= ch ? A.tiles_h : 0;
          n += d_n + ch;
        };
        int c_buf = 0, c_use = 0;  // tile t of this CTA uses d(Wh) buffer t % ndw for the (t / ndw + 1)-th time
        int itp = 0;
        for (int tileA = blockIdx.x; tileA < A.tiles; tileA += 2 * gridDim.x, ++itp) {
          const int it = 2 * itp;
          const int tileB = tileA + gridDim.x;
          const bool hasB = tileB < A.tiles;
          const int bufA = c_buf, useA = c_use;
          if (++c_buf == A.ndw) { c_buf = 0; ++c_use; }
          const int bufB = c_buf, useB = c_use;
          if (++c_buf == A.ndw) { c_buf = 0; ++c_use; }
          const uint32_t ph = (uint32_t)itp & 1u;
          long long pixA, pixB = 0;
          bool validA, validB = false;
          {
            const int h = c_th * LF_TH + hrow, w = c_tw * LF_TW + wcol;
            validA = h < A.h && w < A.w;
            pixA = ((long long)c_n * A.h + h) * A.w + w;
          }
          advance(c_tw, c_th, c_n);
          if (hasB) {
            const int h = c_th * LF_TH + hrow, w = c_tw * LF_TW + wcol;
            validB = h < A.h && w < A.w;
            pixB = ((long long)c_n * A.h + h) * A.w + w;
          }
          advance(c_tw, c_th, c_n);
          if (g == 0) {  // the targets of this pair: pull their lines into L2 while the forward runs
            if (validA) asm volatile("prefetch.global.L2 [%0];" ::"l"(A.y + pixA * REC));
            if (validB) asm volatile("prefetch.global.L2 [%0];" ::"l"(A.y + pixB * REC));
          }
          mbar_wait(&tfull[0], ph);
          if (hasB) mbar_wait(&tfull[1], ph);
          tc_fence_after();
          const bool dbg_thread = g == 0 && m == 0;
          if (dbg_thread) LDBG(9);
          __half2 Wh2[NODES][CO];
          NbState<H2, NODES> st;
          {
            float ra[REC], rb[REC], sa[8 * SL], sb[8 * SL];
            const uint32_t t_addr = tmem_base + lane_off + LF_FP_COL0 + k * REC;
            const uint32_t s_addr = tmem_base + lane_off + LF_FP_COL0 + A.cout + k * (8 * SL);
#pragma unroll
            for (int q = 0; q < CH; ++q) tmem_ld8_nowait(t_addr + q * 8, &ra[q * 8]);
#pragma unroll
            for (int q = 0; q < SL; ++q) tmem_ld8_nowait(s_addr + q * 8, &sa[q * 8]);
            if (hasB) {
#pragma unroll
              for (int q = 0; q < CH; ++q) tmem_ld8_nowait(t_addr + 128 + q * 8, &rb[q * 8]);
#pragma unroll
              for (int q = 0; q < SL; ++q) tmem_ld8_nowait(s_addr + 128 + q * 8, &sb[q * 8]);
            }
            tmem_ld_wait();
            if (!hasB) {
#pragma unroll
              for (int i = 0; i < REC; ++i) rb[i] = 0.f;
#pragma unroll
              for (int i = 0; i < 8 * SL; ++i) sb[i] = 0.f;
            }
            tc_fence_before();
            mbar_arrive(&tempty[0]);
            if (hasB) mbar_arrive(&tempty[1]);
#pragma unroll
            for (int v = 0; v < NODES; ++v) {
#pragma unroll
              for (int u = 0; u < CO; ++u)
                Wh2[v][u] = __floats2half2_rn(ra[rec_off<NODES, CO, SPATIAL>(v, u)], rb[rec_off<NODES, CO, SPATIAL>(v, u)]);
              st.s1[v] = __floats2half2_rn(sa[v], sb[v]);
              st.s2[v] = __floats2half2_rn(sa[NODES + v], sb[NODES + v]);
              smax2 = __hmax2(smax2, __hmax2(__habs2(st.s1[v]), __habs2(st.s2[v])));
            }
          }
          if (dbg_thread) LDBG(10);
          __half2 z2[NODES][CO];
          attn_nb_forward<H2, NODES, CO, MASKED, true, true>(Wh2, nullptr, adj2, s_mask, alpha2, st, z2);
          if (dbg_thread) LDBG(11);
          // ---- swap ELU(z) of the three heads through the d(Wh) buffers of the two tiles (fp16) ----
          const uint32_t exA = smem_u32(s_dw) + (uint32_t)bufA * A.dw_bytes + (uint32_t)m * 16;
          const uint32_t exB = smem_u32(s_dw) + (uint32_t)bufB * A.dw_bytes + (uint32_t)m * 16;
          // the wgrad MMAs that read these buffers last time round must have completed
          if (useA > 0) mbar_wait(&dwfree[bufA], (uint32_t)(useA - 1) & 1u);
          if (hasB && useB > 0) mbar_wait(&dwfree[bufB], (uint32_t)(useB - 1) & 1u);
          // ELU(z) = max(z, t - 1) with t = exp(min(z, 0)) = ELU'(z): t replaces z (only the derivative is needed
          // later).  One 16-byte chunk of the record at a time: 8 outputs live, not the whole record.
#pragma unroll
          for (int q = 0; q < CH; ++q) {
            __half2 o2[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              constexpr int dummy = 0; (void)dummy;
              const int idx = 8 * q + e;
              const int v = SPATIAL ? idx % NODES : idx / CO, u = SPATIAL ? idx / NODES : idx % CO;
              if (A.apply_elu) {
                const __half2 t = H2::exp(H2::min(z2[v][u], H2::zero()));
                o2[e] = H2::max(z2[v][u], H2::sub(t, H2::bc(1.f)));
                z2[v][u] = t;
              } else {
                o2[e] = z2[v][u];
                z2[v][u] = H2::bc(1.f);
              }
            }
            uint4 va, vb;
            __half2 t;
#define LO2(i) (t = __lows2half2(o2[(i)], o2[(i) + 1]), *reinterpret_cast<uint32_t*>(&t))
#define HI2(i) (t = __highs2half2(o2[(i)], o2[(i) + 1]), *reinterpret_cast<uint32_t*>(&t))
            va.x = LO2(0); va.y = LO2(2); va.z = LO2(4); va.w = LO2(6);
            vb.x = HI2(0); vb.y = HI2(2); vb.z = HI2(4); vb.w = HI2(6);
#undef LO2
#undef HI2
            lf_sts128(exA + (uint32_t)(k * CH + q) * 2048, va);
            if (hasB) lf_sts128(exB + (uint32_t)(k * CH + q) * 2048, vb);
          }
          if (dbg_thread) LDBG(15);
          // ---- this group's share of the record (chunks c = g, g + nact, ...): targets requested before the barrier
          //      (L2-resident by now), consumed after it;  d(out) * numel = 2 (out - y) - lambda in packed fp16 on the
          //      exchanged words themselves (a word = two consecutive record elements of one tile; every quantity is
          //      O(1)); the result goes to the score-row planes for everybody.  NC = chunks per group, compile-time so
          //      that the usual case (as many heads as chunks: one chunk each) holds 2 target words, not 2 * CH ----
          auto exchange = [&](auto nc_tag) {
            constexpr int NC = decltype(nc_tag)::value;
            uint4 yraw[2][NC];
#pragma unroll
            for (int half = 0; half < 2; ++half)
#pragma unroll
              for (int j = 0; j < NC; ++j) {
                const int c = g + j * nact;
                yraw[half][j] = make_uint4(0, 0, 0, 0);
                if (c < CH && (half ? validB : validA))
                  asm volatile("ld.global.nc.v4.u32 {%0, %1, %2, %3}, [%4];"
                               : "=r"(yraw[half][j].x), "=r"(yraw[half][j].y), "=r"(yraw[half][j].z), "=r"(yraw[half][j].w)
                               : "l"(reinterpret_cast<const uint4*>(A.y + (half ? pixB : pixA) * REC) + c));
              }
            named_bar_sync(1, 128 * nact);
#pragma unroll
            for (int half = 0; half < 2; ++half) {
              const bool valid = half ? validB : validA;
              const uint32_t ex = half ? exB : exA;
              if (half && !hasB) continue;
              // dz = d(out)/heads: the mean's 1/heads is folded into the constants
              const __half2 c2 = __float2half2_rn(valid ? 2.f * inv_heads : 0.f);
              const __half2 cl = __float2half2_rn(valid ? -A.lambda * inv_heads : 0.f);
              __half2 ssq = H2::zero(), so = H2::zero();
#pragma unroll
              for (int j = 0; j < NC; ++j) {
                const int c = g + j * nact;
                if (c >= CH) continue;
                __half2 os[4];
                for (int kk = 0; kk < A.heads; ++kk) {
                  const uint4 v = lf_lds128(ex + (uint32_t)(kk * CH + c) * 2048);
                  const uint32_t w4[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                  for (int e = 0; e < 4; ++e) {
                    const __half2 w = *reinterpret_cast<const __half2*>(&w4[e]);
                    os[e] = kk == 0 ? w : __hadd2(os[e], w);
                  }
                }
                const uint32_t y4[4] = {yraw[half][j].x, yraw[half][j].y, yraw[half][j].z, yraw[half][j].w};
                uint32_t d4[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const __half2 y2 = __floats2half2_rn(__uint_as_float(y4[e] << 16), __uint_as_float(y4[e] & 0xffff0000u));
                  const __half2 df = __hfma2(os[e], gs2, __hneg2(y2));  // out - y,  out = mean over heads
                  const __half2 o = __hadd2(df, y2);
                  const __half2 dh = __hfma2(df, c2, cl);
                  d4[e] = *reinterpret_cast<const uint32_t*>(&dh);
                  ssq = __hfma2(df, df, ssq);
                  so = __hadd2(so, o);
                }
                lf_sts128(ex + xplane + (uint32_t)c * 2048, make_uint4(d4[0], d4[1], d4[2], d4[3]));
              }
              if (valid) {
                const float2 a = __half22float2(ssq), b = __half22float2(so);
                loss_acc += (a.x + a.y) - A.lambda * (b.x + b.y);
                mse_acc += a.x + a.y;
              }
            }
          };
          if (nact == CH) exchange(std::integral_constant<int, 1>{});
          else exchange(std::integral_constant<int, CH>{});
          named_bar_sync(1, 128 * nact);  // d(out) complete; all heads' ELU words read: the feature planes may take d(Wh)
          // ---- dz = d(out)/heads * ELU'(z), in place over z2 (which holds ELU'); one record chunk at a time ----
#pragma unroll
          for (int c = 0; c < CH; ++c) {
            const uint4 wa = lf_lds128(exA + xplane + (uint32_t)c * 2048);
            uint4 wb = make_uint4(0, 0, 0, 0);
            if (hasB) wb = lf_lds128(exB + xplane + (uint32_t)c * 2048);
            const uint32_t a4[4] = {wa.x, wa.y, wa.z, wa.w}, b4[4] = {wb.x, wb.y, wb.z, wb.w};
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const int idx = 8 * c + e;
              const int v = SPATIAL ? idx % NODES : idx / CO, u = SPATIAL ? idx / NODES : idx % CO;
              // lane A = element idx of tile A, lane B = element idx of tile B
              const __half2 da = *reinterpret_cast<const __half2*>(&a4[e >> 1]);
              const __half2 db = *reinterpret_cast<const __half2*>(&b4[e >> 1]);
              const __half2 ab = (e & 1) ? __highs2half2(da, db) : __lows2half2(da, db);
              z2[v][u] = __hmul2(ab, z2[v][u]);
            }
          }
          mbar_arrive(dhread);  // this thread is done with the score-row planes
          if (dbg_thread) LDBG(12);
          // adjacency gradient sums of this thread stay packed across its pairs (O(1) terms, <= 8 pairs per CTA:
          // fp16 accumulation error ~1e-3 of a per-thread partial; the 19K partials are then summed in fp32)
          __half2 ds[2 * NODES];
          attn_nb_backward_inplace<H2, NODES, CO, MASKED>(Wh2, z2, adj2, s_mask, alpha2, st, &g2[0], ds);  // z2: dz -> d(Wh)
          if (dbg_thread) LDBG(13);
          // bias gradient: d(Wh) summed over this thread's nodes and pixels, and the sums of ds1 / ds2 (the bias terms of
          // the score rows); packed like the adjacency sums (O(1) terms, <= 8 pairs)
#pragma unroll
          for (int u = 0; u < CO; ++u) {
            __half2 t = z2[0][u];
#pragma unroll
            for (int v = 1; v < NODES; ++v) t = __hadd2(t, z2[v][u]);
            gb2[u] = __hadd2(gb2[u], t);
          }
#pragma unroll
          for (int w = 0; w < 2; ++w) {
            __half2 t = ds[w * NODES];
#pragma unroll
            for (int v = 1; v < NODES; ++v) t = __hadd2(t, ds[w * NODES + v]);
            gb2[CO + w] = __hadd2(gb2[CO + w], t);
          }
          // ---- d(Wh) (z2) -> bf16 A planes of the wgrad MMA of both tiles ----
          {
            const uint32_t dyA = exA + (uint32_t)(k * CH) * 2048, dyB = exB + (uint32_t)(k * CH) * 2048;
#pragma unroll
            for (int q = 0; q < CH; ++q) {
              float fa[8], fb[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const int idx = 8 * q + j;
                const int v = SPATIAL ? idx % NODES : idx / CO, u = SPATIAL ? idx / NODES : idx % CO;
                const float2 f = __half22float2(z2[v][u]);
                fa[j] = f.x;
                fb[j] = f.y;
              }
              uint4 va, vb;
              va.x = pack_bf16x2(fa[0], fa[1]); va.y = pack_bf16x2(fa[2], fa[3]);
              va.z = pack_bf16x2(fa[4], fa[5]); va.w = pack_bf16x2(fa[6], fa[7]);
              vb.x = pack_bf16x2(fb[0], fb[1]); vb.y = pack_bf16x2(fb[2], fb[3]);
              vb.z = pack_bf16x2(fb[4], fb[5]); vb.w = pack_bf16x2(fb[6], fb[7]);
              lf_sts128(dyA + (uint32_t)q * 2048, va);
              if (hasB) lf_sts128(dyB + (uint32_t)q * 2048, vb);
            }
          }
          // ---- score gradients ds1 | ds2 (| zero padding) -> this head's score-row planes, once every group has read d(out)
          mbar_wait(dhread, ph);
          {
            const uint32_t syA = exA + xplane + (uint32_t)(k * SL) * 2048, syB = exB + xplane + (uint32_t)(k * SL) * 2048;
#pragma unroll
            for (int q = 0; q < SL; ++q) {
              float fa[8], fb[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                fa[j] = 0.f;
                fb[j] = 0.f;
                if (8 * q + j < 2 * NODES) {
                  const float2 f = __half22float2(ds[8 * q + j]);
                  fa[j] = f.x;
                  fb[j] = f.y;
                }
              }
              uint4 va, vb;
              va.x = pack_bf16x2(fa[0], fa[1]); va.y = pack_bf16x2(fa[2], fa[3]);
              va.z = pack_bf16x2(fa[4], fa[5]); va.w = pack_bf16x2(fa[6], fa[7]);
              vb.x = pack_bf16x2(fb[0], fb[1]); vb.y = pack_bf16x2(fb[2], fb[3]);
              vb.z = pack_bf16x2(fb[4], fb[5]); vb.w = pack_bf16x2(fb[6], fb[7]);
              lf_sts128(syA + (uint32_t)q * 2048, va);
              if (hasB) lf_sts128(syB + (uint32_t)q * 2048, vb);
            }
          }
          fence_proxy_async_smem();
          mbar_arrive(&dyfull[bufA]);
          if (hasB) mbar_arrive(&dyfull[bufB]);
          if (dbg_thread) LDBG(14);
        }
        if (A.guard != nullptr) {
          const float2 f = __half22float2(smax2);
          if (!(fmaxf(f.x, f.y) <= LF_SMAX)) *A.guard = 1.f;  // (NaN fails the comparison too)
        }
#pragma unroll
        for (int i = 0; i < NODES * NODES; ++i) {
          const float2 f = __half22float2(g2[i]);
          gacc[i] += f.x + f.y;
        }
#pragma unroll
        for (int i = 0; i < CO + 2; ++i) {
          const float2 f = __half22float2(gb2[i]);
          gacc[RGA + i] += f.x + f.y;
        }
      } else {
      int it = 0;
      for (int tile = blockIdx.x; tile < A.tiles; tile += gridDim.x, ++it) {
        const int acc = it & 1;
        const uint32_t ph = ((uint32_t)it >> 1) & 1u;
        const int buf = BWD ? it % A.ndw : 0, use = BWD ? it / A.ndw : 0;  // this tile's d(Wh) buffer (backward kernels)
        const uint32_t dwb = smem_u32(s_dw) + (uint32_t)buf * A.dw_bytes + (uint32_t)m * 16;
        const int tw = tile % A.tiles_w;
        const int th = (tile / A.tiles_w) % A.tiles_h;
        const int n = tile / (A.tiles_w * A.tiles_h);
        const int h = th * LF_TH + hrow, w = tw * LF_TW + wcol;
        const bool valid = h < A.h && w < A.w;
        const long long pix = ((long long)n * A.h + h) * A.w + w;
        // upstream record of this pixel (d(out), or y in train mode): requested now, consumed after the forward has
        // been recomputed, so its HBM latency hides behind the math; the next tile's lines are pulled into L2
        uint4 pre[BWD ? REC / 8 : 1];
        if constexpr (BWD) {
          const __nv_bfloat16* src = A.y != nullptr ? A.y + pix * REC
                                                    : (vec_io ? A.dout + pix * out_rec + (concat ? g * REC : 0) : nullptr);
#pragma unroll
          for (int q = 0; q < REC / 8; ++q) {
            pre[q] = make_uint4(0, 0, 0, 0);
            if (valid && src != nullptr)  // volatile: the request must be issued HERE, not sunk to its use
              asm volatile("ld.global.nc.v4.u32 {%0, %1, %2, %3}, [%4];"
                           : "=r"(pre[q].x), "=r"(pre[q].y), "=r"(pre[q].z), "=r"(pre[q].w)
                           : "l"(reinterpret_cast<const uint4*>(src) + q));
          }
          const int ntile = tile + gridDim.x;
          if (g == 0 && ntile < A.tiles) {
            const int ntw = ntile % A.tiles_w, nth = (ntile / A.tiles_w) % A.tiles_h, nn = ntile / (A.tiles_w * A.tiles_h);
            const int nh = nth * LF_TH + hrow, nw = ntw * LF_TW + wcol;
            if (nh < A.h && nw < A.w) {
              const long long npix = ((long long)nn * A.h + nh) * A.w + nw;
              const void* pa = A.y != nullptr ? (const void*)(A.y + npix * REC) : (const void*)(A.dout + npix * out_rec);
              asm volatile("prefetch.global.L2 [%0];" ::"l"(pa));
            }
          }
        }
        mbar_wait(&tfull[acc], ph);
        if (BWD && use > 0) mbar_wait(&dwfree[buf], (uint32_t)(use - 1) & 1u);  // the buffer's previous wgrad has completed
        tc_fence_after();
        const bool dbg_thread = g == 0 && m == 0;
        if (dbg_thread) LDBG(9);
        float oacc[(!BWD) ? REC : 1];
        if (!BWD) {
#pragma unroll
          for (int i = 0; i < REC; ++i) oacc[i] = 0.f;
        }
        for (int k = g; k < A.heads; k += LF_GROUPS) {
          float rec[REC];
          const uint32_t t_addr = tmem_base + ((uint32_t)(lg * 32) << 16) + LF_FP_COL0 + acc * 128 + k * REC;
#pragma unroll
          for (int q = 0; q < REC / 8; ++q) tmem_ld8_nowait(t_addr + q * 8, &rec[q * 8]);
          tmem_ld_wait();
          if (dbg_thread) LDBG(10);
          if (k + LF_GROUPS >= A.heads) {  // last head of this group: the accumulator may be overwritten
            tc_fence_before();
            mbar_arrive(&tempty[acc]);
          }
          // (the conv bias arrives through the MMA: the plane of ones times the bias K-chunk of the packed weights)
          float Wh[NODES][CO];
          rec_to_mat<NODES, CO, SPATIAL>(rec, Wh);
          float z[NODES][CO];
          NbState<F32, NODES> st;
          attn_nb_forward<F32, NODES, CO, MASKED>(Wh, s_a + k * 2 * CO, s_adj + k * NODES * NODES, s_mask, A.alpha, st, z);
          if (dbg_thread) LDBG(11);
          if constexpr (!BWD) {
            if (A.apply_elu) {
#pragma unroll
              for (int v = 0; v < NODES; ++v)
#pragma unroll
                for (int u = 0; u < CO; ++u) z[v][u] = elu_fwd<F32>(z[v][u]);
            }
            if (!concat) {
#pragma unroll
              for (int v = 0; v < NODES; ++v)
#pragma unroll
                for (int u = 0; u < CO; ++u) oacc[rec_off<NODES, CO, SPATIAL>(v, u)] += z[v][u];
            } else if (valid) {
              __nv_bfloat16* op = A.out + pix * out_rec;
              if (vec_io) {
                mat_to_rec<NODES, CO, SPATIAL>(z, rec);
                store_rec<REC, __nv_bfloat16>(op + k * REC, rec);
              } else {
#pragma unroll
                for (int v = 0; v < NODES; ++v)
#pragma unroll
                  for (int u = 0; u < CO; ++u) op[v * (A.heads * CO) + k * CO + u] = __float2bfloat16_rn(z[v][u]);
              }
            }
          } else {
            // ---- upstream gradient of this pixel / head, times ELU'(z) ----
            const float gscale = concat ? 1.f : inv_heads;
            float dz[NODES][CO];
            if (A.y != nullptr) {
              // ---- train mode (mean merge, one head per group): out = mean_k ELU(z_k) needs every head of the
              //      pixel, so the groups swap their ELU(z) through the (still unused) d(Wh) planes of this stage
              //      as fp16; d(out) = (2 (out - y) - lambda) / numel   (convolutional_gat/train.py:131) ----
              const uint32_t exb = dwb;
#pragma unroll
              for (int v = 0; v < NODES; ++v)
#pragma unroll
                for (int u = 0; u < CO; ++u)
                  rec[rec_off<NODES, CO, SPATIAL>(v, u)] = A.apply_elu ? elu_fwd<F32>(z[v][u]) : z[v][u];
#pragma unroll
              for (int q = 0; q < REC / 8; ++q) {
                uint4 v;
                __half2 t;
#define PKH(a, b) (t = __floats2half2_rn(a, b), *reinterpret_cast<uint32_t*>(&t))
                v.x = PKH(rec[8 * q + 0], rec[8 * q + 1]); v.y = PKH(rec[8 * q + 2], rec[8 * q + 3]);
                v.z = PKH(rec[8 * q + 4], rec[8 * q + 5]); v.w = PKH(rec[8 * q + 6], rec[8 * q + 7]);
#undef PKH
                lf_sts128(exb + (uint32_t)(k * (REC / 8) + q) * 2048, v);
              }
              if (m == 0 && g == 0) LDBG(15);
              named_bar_sync(1, 128 * nact);
              float yv[REC];
              unpack_rec<REC>(pre, yv);
#pragma unroll
              for (int i = 0; i < REC; ++i) rec[i] = 0.f;
              for (int kk = 0; kk < A.heads; ++kk) {
#pragma unroll
                for (int q = 0; q < REC / 8; ++q) {
                  const uint4 v = lf_lds128(exb + (uint32_t)(kk * (REC / 8) + q) * 2048);
                  const uint32_t w4[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                  for (int e = 0; e < 4; ++e) {
                    const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w4[e]));
                    rec[8 * q + 2 * e] += f.x;
                    rec[8 * q + 2 * e + 1] += f.y;
                  }
                }
              }
              float lsum = 0.f, msum = 0.f;
#pragma unroll
              for (int i = 0; i < REC; ++i) {
                const float o = rec[i] * inv_heads;
                const float dd = o - yv[i];
                lsum += dd * dd - A.lambda * o;
                msum += dd * dd;
                rec[i] = valid ? (2.f * dd - A.lambda) * A.inv_n : 0.f;
              }
              if (g == 0 && valid) loss_acc += lsum;
              named_bar_sync(1, 128 * nact);  // all heads read: the planes may now take d(Wh)
              if (dbg_thread) LDBG(12);
              rec_to_mat<NODES, CO, SPATIAL>(rec, dz);
#pragma unroll
              for (int v = 0; v < NODES; ++v)
#pragma unroll
                for (int u = 0; u < CO; ++u)
                  dz[v][u] = dz[v][u] * gscale * (A.apply_elu ? elu_grad<F32>(z[v][u]) : 1.f);
            } else {
              const __nv_bfloat16* dp = A.dout + pix * out_rec;
              if (vec_io && (!concat || k == g)) {
                unpack_rec<REC>(pre, rec);
              } else if (!valid) {
#pragma unroll
                for (int i = 0; i < REC; ++i) rec[i] = 0.f;
              } else if (vec_io) {
                load_rec<REC, __nv_bfloat16>(dp + k * REC, rec);
              } else {
#pragma unroll
                for (int v = 0; v < NODES; ++v)
#pragma unroll
                  for (int u = 0; u < CO; ++u)
                    rec[rec_off<NODES, CO, SPATIAL>(v, u)] = __bfloat162float(dp[v * (A.heads * CO) + k * CO + u]);
              }
              rec_to_mat<NODES, CO, SPATIAL>(rec, dz);
#pragma unroll
              for (int v = 0; v < NODES; ++v)
#pragma unroll
                for (int u = 0; u < CO; ++u)
                  dz[v][u] = dz[v][u] * gscale * (A.apply_elu ? elu_grad<F32>(z[v][u]) : 1.f);
            }
            if (k != cur_head) { flush(cur_head); cur_head = k; }
            attn_nb_backward<F32, NODES, CO, MASKED>(Wh, dz, s_a + k * 2 * CO, s_adj + k * NODES * NODES, s_mask, A.alpha, st,
                                                     z, &gacc[NODES * NODES], &gacc[0]);  // z now holds d(Wh)
            if (dbg_thread) LDBG(13);
#pragma unroll
            for (int u = 0; u < CO; ++u)  // bias gradient: d(Wh) summed over nodes (and, in gacc, over this thread's pixels)
#pragma unroll
              for (int v = 0; v < NODES; ++v) gacc[RGA + u] += z[v][u];
            mat_to_rec<NODES, CO, SPATIAL>(z, rec);
            // d(Wh) -> the MN-major A operand of the wgrad MMA: plane = dense cout / 8, 16 bytes per pixel
            const uint32_t dy = dwb + (uint32_t)(k * (REC / 8)) * 2048;
#pragma unroll
            for (int q = 0; q < REC / 8; ++q) {
              uint4 v;
              v.x = pack_bf16x2(rec[8 * q + 0], rec[8 * q + 1]);
              v.y = pack_bf16x2(rec[8 * q + 2], rec[8 * q + 3]);
              v.z = pack_bf16x2(rec[8 * q + 4], rec[8 * q + 5]);
              v.w = pack_bf16x2(rec[8 * q + 6], rec[8 * q + 7]);
              lf_sts128(dy + (uint32_t)q * 2048, v);
              if (A.dwh != nullptr && valid) reinterpret_cast<uint4*>(A.dwh + pix * A.cout + k * REC)[q] = v;
            }
          }
        }
        if constexpr (BWD) {
          fence_proxy_async_smem();
          mbar_arrive(&dyfull[buf]);
          if (dbg_thread) LDBG(14);
        } else if (!concat) {
          // ---- head mean: every group leaves its partial sum in its slab, then all active threads combine ----
          float4* slab = s_slab + (size_t)g * (REC / 4) * 128;
#pragma unroll
          for (int i = 0; i < REC / 4; ++i)
            slab[i * 128 + m] = make_float4(oacc[4 * i], oacc[4 * i + 1], oacc[4 * i + 2], oacc[4 * i + 3]);
          named_bar_sync(1, 128 * nact);
          const int at = g * 128 + m;
          for (int q = at; q < 128 * (REC / 8); q += 128 * nact) {
            const int part = q >> 7, p = q & 127;
            float4 lo = make_float4(0.f, 0.f, 0.f, 0.f), hi = lo;
            for (int gg = 0; gg < nact; ++gg) {
              const float4 a0 = s_slab[((size_t)gg * (REC / 4) + 2 * part) * 128 + p];
              const float4 a1 = s_slab[((size_t)gg * (REC / 4) + 2 * part + 1) * 128 + p];
              lo.x += a0.x; lo.y += a0.y; lo.z += a0.z; lo.w += a0.w;
              hi.x += a1.x; hi.y += a1.y; hi.z += a1.z; hi.w += a1.w;
            }
            const int ph2 = th * LF_TH + (p >> 3), pw2 = tw * LF_TW + (p & 7);
            if (ph2 < A.h && pw2 < A.w) {
              uint4 v;
              v.x = pack_bf16x2(lo.x * inv_heads, lo.y * inv_heads);
              v.y = pack_bf16x2(lo.z * inv_heads, lo.w * inv_heads);
              v.z = pack_bf16x2(hi.x * inv_heads, hi.y * inv_heads);
              v.w = pack_bf16x2(hi.z * inv_heads, hi.w * inv_heads);
              reinterpret_cast<uint4*>(A.out + (((long long)n * A.h + ph2) * A.w + pw2) * REC)[part] = v;
            }
          }
          named_bar_sync(1, 128 * nact);  // slabs are rewritten by the next tile
        }
      }
      }  // !PAIR
      if constexpr (BWD) {
        if (threadIdx.x == LF_ATT_WARP0 * 32) LDBGX(3);
        // ---- this warp's sums -> its slot in the (now idle) weight buffer: butterflies interleaved over all values,
        //      plain stores, no shared-memory atomics (fp32 ATOMS is a CAS loop: 28 serialised ones cost 8 600 cycles)
        {
          float* slot = reinterpret_cast<float*>(s_w) + (warp - LF_ATT_WARP0) * LF_SLOT;
          if constexpr (PAIR && NODES * NODES == 16) {
            // the paired kernel carries adjacency sums only (the score gradients left through the wgrad MMA): 16 values by
            // recursive halving -- 8 + 4 + 2 + 1 + 1 shuffles instead of 5 per value; lanes 2i, 2i+1 end with value i
#pragma unroll
            for (int st = 0; st < 4; ++st) {
              const int off = 16 >> st, nk = 8 >> st;
              const bool up = (lane & off) != 0;
#pragma unroll
              for (int i = 0; i < nk; ++i) {
                const float send = up ? gacc[i] : gacc[i + nk], keep = up ? gacc[i + nk] : gacc[i];
                gacc[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
              }
            }
            gacc[0] += __shfl_xor_sync(0xffffffffu, gacc[0], 1);
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
              loss_acc += __shfl_xor_sync(0xffffffffu, loss_acc, off);
              mse_acc += __shfl_xor_sync(0xffffffffu, mse_acc, off);
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1)
#pragma unroll
              for (int i = 0; i < CO + 2; ++i) gacc[RGA + i] += __shfl_xor_sync(0xffffffffu, gacc[RGA + i], off);
            if ((lane & 1) == 0) slot[lane >> 1] = gacc[0];
            if (lane < RGA - NODES * NODES) slot[NODES * NODES + lane] = 0.f;
            if (lane == 0) {
#pragma unroll
              for (int i = 0; i < CO + 2; ++i) slot[RGA + i] = gacc[RGA + i];
              slot[RG] = loss_acc;
              slot[RG + 1] = mse_acc;
              reinterpret_cast<int*>(slot)[RG + 2] = cur_head;
            }
          } else {
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
#pragma unroll
              for (int i = 0; i < RG; ++i) gacc[i] += __shfl_xor_sync(0xffffffffu, gacc[i], off);
              loss_acc += __shfl_xor_sync(0xffffffffu, loss_acc, off);
              mse_acc += __shfl_xor_sync(0xffffffffu, mse_acc, off);
            }
            if (lane == 0) {
#pragma unroll
              for (int i = 0; i < RG; ++i) slot[i] = gacc[i];
              slot[RG] = loss_acc;
              slot[RG + 1] = mse_acc;
              reinterpret_cast<int*>(slot)[RG + 2] = cur_head;
            }
          }
        }
        if (threadIdx.x == LF_ATT_WARP0 * 32) LDBGX(4);
        // ---- wgrad accumulator -> per-CTA partial sums (lane = dense cout) ----
        mbar_wait(done, 0);
        tc_fence_after();
        if (threadIdx.x == LF_ATT_WARP0 * 32) LDBGX(5);
        // COMPACT slots [row][vertical tap r][horizontal tap s][ci]: of the 216 accumulator columns of a row only the
        // 9 * ci that multiply the row's OWN node are gradients of the shared per-node conv (the dense wgrad also holds
        // every cross-node product): a thread picks those with compile-time (column -> node, ci) tables and stores
        // 3 * ci floats per vertical tap -- 21 KB per CTA instead of 83 KB (the 148 CTAs' slots leave and re-enter
        // the L2 at its write bandwidth: 3 300 cycles of this kernel's tail, and as much again in the reduction)
        if (lg * 32 < A.rows_pad) {
          constexpr int CI = CO, NCH = NODES * CO / 8, NC = 9 * CI;
          int node = -1;  // this row's node; -1: a padding row
          if (m < A.cout) {
            const int rr = m % REC;
            node = SPATIAL ? rr % NODES : rr / CO;
          } else if (m < A.cout + A.ext) {
            const int j = (m - A.cout) % lf_score_rows_per_head(NODES);
            if (j < 2 * NODES) node = j % NODES;
          }
          float* prow = A.partial + ((size_t)blockIdx.x * A.rows_pad + m) * NC;
          for (int r = g; r < 3; r += nact) {
            float sel[3 * CI];
#pragma unroll
            for (int i = 0; i < 3 * CI; ++i) sel[i] = 0.f;
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
#pragma unroll
              for (int sh = 0; sh < 3; ++sh) {
                float v[8];
                tmem_ld8_nowait(tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(((r * NCH + c) * 3 + sh) * 8), v);
                tmem_ld_wait();
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                  constexpr int dummy = 0; (void)dummy;
                  const int ch = c * 8 + e;
                  const int nd = SPATIAL ? ch % NODES : ch / CI, ci = SPATIAL ? ch / NODES : ch % CI;
                  if (node == nd) sel[sh * CI + ci] = v[e];
                }
              }
            }
            float2* dst = reinterpret_cast<float2*>(prow + r * 3 * CI);
#pragma unroll
            for (int i = 0; i < 3 * CI / 2; ++i) dst[i] = make_float2(sel[2 * i] * A.out_scale, sel[2 * i + 1] * A.out_scale);
          }
        }
        if (threadIdx.x == LF_ATT_WARP0 * 32) LDBGX(6);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) LDBGX(7);
  if constexpr (BWD) {
    const float* slots = reinterpret_cast<const float*>(s_w);
    for (int i = threadIdx.x; i < A.heads * RG; i += LF_THREADS) {
      const int k = i / RG, r = i - k * RG;
      float v = s_gacc[i];  // heads a group finished before its last one (more than LF_GROUPS heads)
      for (int wa = 0; wa < 4 * nact; ++wa)  // fixed order: the CTA's contribution is deterministic
        if (reinterpret_cast<const int*>(slots + wa * LF_SLOT)[RG + 2] == k) v += slots[wa * LF_SLOT + r];
      v *= A.out_scale;
      if (A.guard != nullptr && !isfinite(v)) *A.guard = 1.f;
      if (r < NODES * NODES) atomicAdd(A.gadj + (size_t)k * NODES * NODES + r, v);
      else if (r < RGA) atomicAdd(A.ga + (size_t)k * 2 * CO + (r - NODES * NODES), v);
      else atomicAdd(A.gbias + (size_t)k * (CO + 2) + (r - RGA), v);
    }
    if (A.y != nullptr && threadIdx.x == 0) {
      float l = 0.f, q = 0.f;
      for (int wa = 0; wa < 4 * nact; ++wa) { l += slots[wa * LF_SLOT + RG]; q += slots[wa * LF_SLOT + RG + 1]; }
      if (A.guard != nullptr && !(isfinite(l) && isfinite(q))) *A.guard = 1.f;
      atomicAdd(A.loss_out, l * A.inv_n);
      if (A.mse_out != nullptr) atomicAdd(A.mse_out, q * A.inv_n);
    }
  }
  if (warp == LF_MMA_WARP) {
    __syncwarp();
    tmem_dealloc(tmem_base, 512);
  }
  if (threadIdx.x == 0) LDBGX(8);
#ifdef CGAT_LF_TIMELINE
  if (threadIdx.x == 0 && A.dbg != nullptr) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    A.dbg[256 + 2 * blockIdx.x + 1] = (long long)t;
  }
#endif
}

// ---- host side ------------------------------------------------------------------------------------------
struct LfGeom {
  int cin, cout, ext, rec, nchunk, np, nj, npad, mchunk, nt, rows_pad, rowp;
  uint32_t wbytes, stage_bytes, dw_bytes;
  size_t smem;
  int tiles_h, tiles_w, tiles, nstg, ndw;
};

static LfGeom lf_geom(const cgat_layer_desc* d, bool bwd) {
  LfGeom g;
  g.rec = d->nodes * d->co;
  g.cin = d->nodes * d->ci;
  g.cout = d->heads * g.rec;
  g.nchunk = g.cin / 8;
  g.np = 3 * g.nchunk;                // planes per stage row: (chunk c, horizontal tap s)
  g.nj = 3 * g.np;                    // K-chunks of fprop = N-chunks of wgrad: j = r * np + q
  g.rowp = g.np * LF_ROW;
  g.ext = lf_score_rows(d->nodes, d->co, d->heads);  // score rows W.a behind the feature rows (common.cuh)
  g.npad = (g.cout + g.ext + 15) & ~15;
  g.mchunk = g.cout / 8;
  g.nt = 9 * d->ci;                   // columns of a COMPACT partial-sum slot row: [r][s][ci] (see the kernel's tail)
  g.rows_pad = lf_partial_rows(d->nodes, d->co, d->heads);
  g.wbytes = (uint32_t)lf_weight_chunks(g.cin) * g.npad * 16;  // nj chunks + a zero one, the bias chunk + a zero one
  g.stage_bytes = (uint32_t)LF_PR * g.rowp + 128;  // + one zeroed core matrix: the odd last K-chunk's partner, see the kernel
  // a d(Wh) buffer: feature planes, then the score-row planes -- at least rec/8 of them: the train kernel passes d(out)
  // between the head groups through them
  const int xplanes = g.ext / 8 > g.rec / 8 ? g.ext / 8 : g.rec / 8;
  g.dw_bytes = bwd ? (uint32_t)(g.mchunk + xplanes) * 2048 : 0;
  g.nstg = LF_MAXSTG;
  g.ndw = bwd ? LF_MAXDW : 0;
  auto total = [&]() {
    return (size_t)LF_HDR + ((g.wbytes + 127u) & ~127u) + (size_t)g.ndw * g.dw_bytes + (size_t)g.nstg * g.stage_bytes +
           LF_ONES + (bwd ? 0 : (size_t)LF_GROUPS * g.rec * 128 * 4);
  };
  // shed a d(Wh) buffer first (two serve a tile pair), then x stages
  while ((g.smem = total()) > 227 * 1024 && g.ndw > 2) --g.ndw;
  while ((g.smem = total()) > 227 * 1024 && g.nstg > 2) --g.nstg;
  g.tiles_h = (d->h + LF_TH - 1) / LF_TH;
  g.tiles_w = (d->w + LF_TW - 1) / LF_TW;
  g.tiles = d->n * g.tiles_h * g.tiles_w;
  return g;
}

static int lf_sm_count() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  }
  return n;
}

int layer_supported(const cgat_layer_desc* d) {
  if (!d || d->heads < 1 || d->heads > MAX_HEADS || d->n < 1 || d->h < 1 || d->w < 1) return 0;
  const bool sp = d->layout == CGAT_LAYOUT_SPATIAL;
  const bool shape_ok = (sp && d->nodes == 6 && d->ci == 4 && d->co == 4) || (!sp && d->nodes == 4 && d->ci == 6 && d->co == 6);
  if (!shape_ok) return 0;
  const LfGeom f = lf_geom(d, false), b = lf_geom(d, true);
  if (f.cin % 8 || f.cout % 8 || f.cout > 128 || f.npad > 128 || ((f.nj * 8 + 15) & ~15) > 256) return 0;
  if (d->ci != d->co) return 0;  // (the kernel derives its plane count from nodes * co)
  if (f.smem > 227 * 1024 || b.smem > 227 * 1024) return 0;
  return 1;
}

size_t layer_partial_bytes(const cgat_layer_desc* d) {
  const LfGeom g = lf_geom(d, true);
  // one slot per CTA plus one: cgat_stream_param_grads sums the slots into the one behind them (stream_ops.cu)
  return (size_t)(148 + 1) * 128 * g.nt * sizeof(float);
}

// Padded chunk-planar x [n][c/8][h][wp][8] (wp = lf_padded_width(w)) as the 5-D tensor
//   (element of the padded row: wp*8 | horizontal tap s: 3, stride ONE pixel = 16 B | chunk | row | image).
// The tap dimension overlaps the innermost one -- the encoder and the TMA unit accept that (tools/microbench/
// tma_overlap.cu checks the landing) -- so ONE box (64, 3, nchunk, hp, 1) lands as [row][chunk][tap][8 px][8 ch]:
// the stage layout of layer_kernel.  Rows outside the image are zero-filled; the horizontal padding is in the data.
static int make_planar_map(CUtensorMap* map, const void* base, int n, int h, int w, int c, int hp) {
  EncodeTiledFn enc = get_encode_tiled();
  if (!enc) return fail(CGAT_EUNSUPPORTED, "cuTensorMapEncodeTiled not available from the driver");
  ensure_context();
  const int nchunk = c / 8, wp = lf_padded_width(w);
  cuuint64_t dims[5] = {(cuuint64_t)wp * 8, 3, (cuuint64_t)nchunk, (cuuint64_t)h, (cuuint64_t)n};
  cuuint64_t strides[4] = {16, (cuuint64_t)h * wp * 16, (cuuint64_t)wp * 16, (cuuint64_t)nchunk * h * wp * 16};
  cuuint32_t box[5] = {64, 3, (cuuint32_t)nchunk, (cuuint32_t)hp, 1};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(CGAT_EINVAL, "cuTensorMapEncodeTiled (planar x) failed with CUresult %d", (int)r);
  return 0;
}

template <int NODES, int CO, bool SPATIAL>
static int lf_launch(bool bwd, const cgat_layer_desc* d, const LfGeom& g, const CUtensorMap& map, const LfArgs& A,
                     cudaStream_t st) {
  const int grid = g.tiles < lf_sm_count() ? g.tiles : lf_sm_count();
  if (grid > 148) return fail(CGAT_EUNSUPPORTED, "partial-sum workspace sized for <= 148 CTAs");
  auto go = [&](auto kern) -> int {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem);
    if (e != cudaSuccess) return fail((int)e, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    e = launch_pdl(kern, dim3(grid), dim3(LF_THREADS), g.smem, st, map, A);
    if (e != cudaSuccess) return fail((int)e, "layer_kernel launch: %s", cudaGetErrorString(e));
    return 0;
  };
  int rc;
  const bool masked = A.mask != nullptr;  // NULL = all ones = the reference's dense attention: no mask arithmetic
  if (bwd && A.y != nullptr && A.out_scale != 1.f)  // train mode: tile pairs in packed half2 (see layer_launch)
    rc = masked ? go(layer_kernel<NODES, CO, SPATIAL, true, true, true>) : go(layer_kernel<NODES, CO, SPATIAL, true, false, true>);
  else if (bwd)
    rc = masked ? go(layer_kernel<NODES, CO, SPATIAL, true, true, false>) : go(layer_kernel<NODES, CO, SPATIAL, true, false, false>);
  else
    rc = masked ? go(layer_kernel<NODES, CO, SPATIAL, false, true, false>) : go(layer_kernel<NODES, CO, SPATIAL, false, false, false>);
  if (rc) return rc;
  return check_launch(bwd ? "layer_kernel<bwd>" : "layer_kernel<fwd>");
}

int layer_launch(bool bwd, const cgat_layer_desc* d, const void* x, const void* wpack, const float* bias, const float* a,
                 const float* adj, const uint8_t* mask, void* out, const void* dout, void* dwh, float* partial,
                 float* ga, float* gadj, float* gbias, int* ncta_out, int* nt_out, cudaStream_t st, const void* y = nullptr,
                 float* loss_out = nullptr, float lambda = 0.f, float* mse_out = nullptr, float* guard = nullptr,
                 const float* run_if = nullptr, bool force_fp32 = false) {
  if (!layer_supported(d)) return fail(CGAT_EUNSUPPORTED, "fused conv-GAT layer kernel does not support this shape");
  if (!aligned16(x) || !aligned16(wpack) || (out && !aligned16(out)) || (dout && !aligned16(dout)) ||
      (dwh && !aligned16(dwh)) || (partial && !aligned16(partial)))
    return fail(CGAT_EALIGN, "layer tensors must be 16-byte aligned");
  const LfGeom g = lf_geom(d, bwd);
  CUtensorMap map;
  if (d->x_layout != CGAT_X_PLANAR)
    return fail(CGAT_EINVAL, "the fused layer kernels read x padded chunk-planar (x_layout = CGAT_X_PLANAR): convert pixel "
                             "records with cgat_records_to_planar, or let cgat_loader_gather_planar write it");
  if (int rc = make_planar_map(&map, x, d->n, d->h, d->w, g.cin, LF_PR)) return rc;
  LfArgs A{};
  A.dbg = g_lf_dbg;
  A.x = (const __nv_bfloat16*)x;
  A.wpack = (const __nv_bfloat16*)wpack; A.bias = bias; A.a = a; A.adj = adj; A.mask = mask;
  A.out = (__nv_bfloat16*)out; A.dout = (const __nv_bfloat16*)dout; A.dwh = (__nv_bfloat16*)dwh;
  A.partial = partial; A.ga = ga; A.gadj = gadj; A.gbias = gbias;
  A.y = (const __nv_bfloat16*)y; A.loss_out = loss_out; A.mse_out = mse_out; A.lambda = lambda;
  A.inv_n = 1.f / ((float)d->n * (float)d->h * (float)d->w * (float)(d->nodes * d->co));
  // the paired half2 kernel carries d(out) without its 1/numel factor; CGAT_NO_PAIR=1 keeps the fp32 one-tile kernel
  static const bool no_pair = std::getenv("CGAT_NO_PAIR") != nullptr;
  A.out_scale = (bwd && y != nullptr && g.nstg == LF_MAXSTG && g.ndw >= 2 && g.ext > 0 && d->heads <= LF_GROUPS && !no_pair &&
                 !force_fp32) ? A.inv_n : 1.f;
  A.guard = A.out_scale != 1.f ? guard : nullptr;
  A.run_if = run_if;
  A.h = d->h; A.w = d->w; A.wp = lf_padded_width(d->w); A.cin = g.cin; A.cout = g.cout; A.ext = g.ext; A.npad = g.npad; A.heads = d->heads; A.merge = d->merge;
  A.apply_elu = d->apply_elu; A.alpha = d->alpha;
  A.nchunk = g.nchunk; A.np = g.np; A.nj = g.nj; A.rowp = g.rowp; A.mchunk = g.mchunk; A.nt = g.nt; A.rows_pad = g.rows_pad;
  A.tiles_h = g.tiles_h; A.tiles_w = g.tiles_w; A.tiles = g.tiles; A.nstg = g.nstg; A.ndw = g.ndw;
  A.wbytes = g.wbytes; A.stage_bytes = g.stage_bytes; A.dw_bytes = g.dw_bytes;
  if (ncta_out) *ncta_out = g.tiles < lf_sm_count() ? g.tiles : lf_sm_count();
  if (nt_out) *nt_out = g.nt;
  if (d->layout == CGAT_LAYOUT_SPATIAL) return lf_launch<6, 4, true>(bwd, d, g, map, A, st);
  return lf_launch<4, 6, false>(bwd, d, g, map, A, st);
}

}  // namespace cgat

using namespace cgat;

// developer aid, not part of the public header: registers a device buffer [16][16] int64 for the kernel timeline
extern "C" void cgat_layer_debug_timeline(long long* device_buffer) { g_lf_dbg = device_buffer; }

extern "C" int cgat_layer_supported(const cgat_layer_desc* d) { return layer_supported(d); }

extern "C" int64_t cgat_layer_workspace_bytes(const cgat_layer_desc* d) {
  return layer_supported(d) ? (int64_t)layer_partial_bytes(d) : 0;
}

extern "C" int cgat_layer_fwd(const cgat_layer_desc* d, const void* x, const void* wpack, const float* bias_dense,
                              const float* a, const float* adj, const uint8_t* mask, void* out, void* stream) {
  if (!d || !x || !wpack || !a || !adj || !out) return fail(CGAT_EINVAL, "null argument");
  return layer_launch(false, d, x, wpack, bias_dense, a, adj, mask, out, nullptr, nullptr, nullptr, nullptr, nullptr,
                      nullptr, nullptr, nullptr, (cudaStream_t)stream);
}

extern "C" int cgat_layer_bwd(const cgat_layer_desc* d, const void* x, const void* dout, const void* wpack,
                              const float* bias_dense, const float* a, const float* adj, const uint8_t* mask,
                              void* dwh, void* workspace, float* ga, float* gadj, float* gbias, int32_t* ncta_out,
                              int32_t* nt_out, void* stream) {
  if (!d || !x || !dout || !wpack || !a || !adj || !workspace || !ga || !gadj || !gbias || !ncta_out || !nt_out)
    return fail(CGAT_EINVAL, "null argument");
  int ncta = 0, nt = 0;
  const int rc = layer_launch(true, d, x, wpack, bias_dense, a, adj, mask, nullptr, dout, dwh, (float*)workspace, ga,
                              gadj, gbias, &ncta, &nt, (cudaStream_t)stream);
  *ncta_out = ncta;
  *nt_out = nt;
  return rc;
}

extern "C" int cgat_layer_train(const cgat_layer_desc* d, const void* x, const void* y, const void* wpack,
                                const float* bias_dense, const float* a, const float* adj, const uint8_t* mask,
                                float lambda, void* workspace, float* ga, float* gadj, float* gbias, float* loss_out,
                                float* mse_out, float* guard, int32_t* ncta_out, int32_t* nt_out, void* stream) {
  if (!d || !x || !y || !wpack || !a || !adj || !workspace || !ga || !gadj || !gbias || !loss_out || !ncta_out || !nt_out)
    return fail(CGAT_EINVAL, "null argument");
  if (d->merge != CGAT_MERGE_MEAN || d->heads > LF_GROUPS)
    return fail(CGAT_EUNSUPPORTED, "cgat_layer_train serves mean-merged streams with at most %d heads", LF_GROUPS);
  if (!aligned16(y)) return fail(CGAT_EALIGN, "y must be 16-byte aligned");
  int ncta = 0, nt = 0;
  const int rc = layer_launch(true, d, x, wpack, bias_dense, a, adj, mask, nullptr, nullptr, nullptr, (float*)workspace,
                              ga, gadj, gbias, &ncta, &nt, (cudaStream_t)stream, y, loss_out, lambda, mse_out, guard);
  *ncta_out = ncta;
  *nt_out = nt;
  return rc;
}

extern "C" int cgat_layer_train_fp32(const cgat_layer_desc* d, const void* x, const void* y, const void* wpack,
                                     const float* bias_dense, const float* a, const float* adj, const uint8_t* mask,
                                     float lambda, void* workspace, float* ga, float* gadj, float* gbias, float* loss_out,
                                     float* mse_out, const float* run_if, int32_t* ncta_out, int32_t* nt_out, void* stream) {
  if (!d || !x || !y || !wpack || !a || !adj || !workspace || !ga || !gadj || !gbias || !loss_out || !ncta_out || !nt_out)
    return fail(CGAT_EINVAL, "null argument");
  if (d->merge != CGAT_MERGE_MEAN || d->heads > LF_GROUPS)
    return fail(CGAT_EUNSUPPORTED, "cgat_layer_train_fp32 serves mean-merged streams with at most %d heads", LF_GROUPS);
  if (!aligned16(y)) return fail(CGAT_EALIGN, "y must be 16-byte aligned");
  int ncta = 0, nt = 0;
  const int rc = layer_launch(true, d, x, wpack, bias_dense, a, adj, mask, nullptr, nullptr, nullptr, (float*)workspace,
                              ga, gadj, gbias, &ncta, &nt, (cudaStream_t)stream, y, loss_out, lambda, mse_out, nullptr, run_if,
                              true);
  *ncta_out = ncta;
  *nt_out = nt;
  return rc;
}
