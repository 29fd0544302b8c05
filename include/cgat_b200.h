/* cgat_b200.h -- C ABI of the B200-native conv-GAT hot path (libcgat_b200.so).
 *
 * Drop-in boundary.  The reference has no native boundary of its own: every
 * op is a stock ATen call made from Python (`convolutional_gat/baseline_model.py`,
 * `dcgan/model.py`).  Each entry point below therefore cites the reference Python lines whose
 * arithmetic it replaces; the Python side (`extended-gan_b200/cgat/_lib.py`) binds them with ctypes
 * exactly as shown in INTEGRATION.md.
 *
 * Conventions
 *   - plain `extern "C"`, raw device pointers, explicit sizes; no torch types.
 *   - every function returns 0 on success, a negative CGAT_E* code for a rejected argument, or a
 *     positive cudaError_t; `cgat_last_error()` returns a message for the calling thread.
 *   - functions never allocate device memory, never synchronise and never throw; all buffers
 *     (inputs, outputs, workspaces) are owned by the caller and must outlive the stream work.
 *   - `stream` is a cudaStream_t passed as void* (torch.cuda.current_stream().cuda_stream).
 *   - dtype tags: CGAT_F32 = 0, CGAT_BF16 = 1.  Parameters and their gradients are always fp32.
 */
#ifndef CGAT_B200_H_
#define CGAT_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CGAT_F32 0
#define CGAT_BF16 1

#define CGAT_LAYOUT_SPATIAL 0  /* pixel record element (node, c) at  c*nodes + node   (x[N,H,W,T,V], nodes = V) */
#define CGAT_LAYOUT_TEMPORAL 1 /* pixel record element (node, c) at  node*C + c       (x[N,H,W,T,V], nodes = T) */

#define CGAT_PROJ_LINEAR 0 /* input = raw node features; Wh = X . W computed in-kernel (baseline_model.py:127) */
#define CGAT_PROJ_PRE 1    /* input = projected features Wh for all heads, head-major records (conv mapping)  */

#define CGAT_MERGE_CONCAT 0 /* heads concatenated on the channel axis (baseline_model.py:196) */
#define CGAT_MERGE_MEAN 1   /* heads averaged after ELU (output shape == input shape)          */

#define CGAT_EINVAL (-1)       /* bad argument                              */
#define CGAT_EUNSUPPORTED (-2) /* shape / dtype combination not implemented */
#define CGAT_EALIGN (-3)       /* pointer or size not 16-byte aligned       */

typedef struct cgat_attn_desc {
  int64_t n_pix;          /* number of pixel records = N*H*W                                   */
  int64_t pix_per_sample; /* H*W (indexes the per-sample pixel-softmax statistics)             */
  int32_t nodes;          /* graph nodes per pixel (V for spatial, T for temporal)             */
  int32_t ci;             /* input channels per node  (ignored for CGAT_PROJ_PRE)              */
  int32_t co;             /* output channels per node                                          */
  int32_t heads;          /* attention heads (1..8)                                            */
  int32_t layout;         /* CGAT_LAYOUT_*                                                     */
  int32_t proj;           /* CGAT_PROJ_*                                                       */
  int32_t merge;          /* CGAT_MERGE_*                                                      */
  int32_t dtype;          /* CGAT_F32 / CGAT_BF16 for in/out/grad tensors                      */
  int32_t apply_elu;      /* 1: out = ELU(z) (baseline_model.py:160); 0: out = z               */
  float alpha;            /* LeakyReLU slope (0.2, baseline_model.py:111,130)                  */
} cgat_attn_desc;

const char* cgat_version(void);
const char* cgat_last_error(void);

/* K4  fused attention forward.
 * Replaces GraphAttentionLayer2D.forward lines 127-160 of convolutional_gat/baseline_model.py
 * (projection, all-pairs concat :162-169, LeakyReLU :130, soft-max :131, V x V aggregation loop
 * :148-152, adjacency mix :158, ELU :160) and the head loop + cat of GATMultiHead2D.forward :194-197.
 *   in    [n_pix][nodes*ci]            (PROJ_LINEAR)  or  [n_pix][heads][nodes*co]  (PROJ_PRE)
 *   out   [n_pix][nodes*co]            (MERGE_MEAN)   or  concat layout, see DESIGN.md
 *   W     [heads][ci][co]  fp32 (PROJ_LINEAR only, else NULL)
 *   a     [heads][2*co]    fp32
 *   adj   [heads][nodes][nodes] fp32, normalised adjacency in kernel orientation adj[i][v]
 *   mask  [nodes][nodes] uint8 or NULL (NULL = all ones = reference behaviour)
 *   stats NULL -> soft-max over neighbours j;  else [N][heads][2][nodes*nodes] (max plane, 1/sum plane) from
 *         cgat_attn_pixstats -> soft-max over the pixel axis (baseline_model.py:131)           */
int cgat_attn_fwd(const cgat_attn_desc* d, const void* in, void* out, const float* W, const float* a,
                  const float* adj, const uint8_t* mask, const float* stats, void* stream);

/* Pixel-axis soft-max statistics (compat mode of baseline_model.py:131: softmax(e, dim=-1) with the
 * pixel axis last).  stats [N][heads][2][nodes*nodes].                                          */
int cgat_attn_pixstats(const cgat_attn_desc* d, const void* in, const float* W, const float* a,
                       const uint8_t* mask, float* stats, void* stream);

/* K5  fused attention backward (forward recomputed in registers; autograd of the lines above).
 *   dout  gradient of `out`, same layout as `out`
 *   din   gradient of `in`,  same layout as `in`
 *   gW [heads][ci][co], ga [heads][2co], gadj [heads][nodes][nodes]: fp32, ACCUMULATED INTO
 *   bstats: pixel mode only, [N][heads][nodes*nodes] from cgat_attn_pixstats_bwd, else NULL     */
int cgat_attn_bwd(const cgat_attn_desc* d, const void* in, const void* dout, void* din, const float* W,
                  const float* a, const float* adj, const uint8_t* mask, const float* stats,
                  const float* bstats, float* gW, float* ga, float* gadj, void* stream);

/* Pixel mode: bstats[n][h][i][j] = sum_p att * dAtt (ACCUMULATED INTO; caller zeroes).           */
int cgat_attn_pixstats_bwd(const cgat_attn_desc* d, const void* in, const void* dout, const float* W,
                           const float* a, const float* adj, const uint8_t* mask, const float* stats,
                           float* bstats, void* stream);

/* a5  learnable adjacency normalisation, baseline_model.py:41-50 / :133-142:
 *   adj = B + I; min-max normalise; A_hat = D^-1/2 adj D^-1/2 with D detached.
 *   B, adj_hat, gB, gadj: [heads][nodes][nodes] fp32.  transpose=1 writes A_hat^T (the 1-D layer's
 *   left-multiplication, baseline_model.py:53).  One launch for all heads.                        */
int cgat_adj_norm_fwd(const float* B, float* adj_hat, int heads, int nodes, int transpose, void* stream);
int cgat_adj_norm_bwd(const float* B, const float* gadj, float* gB, int heads, int nodes, int transpose,
                      void* stream);

/* Conv descriptors: NHWC activations, weights [cout][kh][kw][cin] ("KRSC"), fp32 accumulate.
 * Replaces nn.Conv2d calls of dcgan/model.py:35-43,150-169 and the node conv of the conv mapping. */
typedef struct cgat_conv_desc {
  int32_t n, h, w, cin;   /* input  [n][h][w][cin]      */
  int32_t cout, kh, kw;   /* weight [cout][kh][kw][cin] */
  int32_t stride;         /* same in both directions    */
  int32_t pad_top, pad_left; /* leading zero padding (PyTorch padding="same" with even k pads 1 before / 2 after) */
  int32_t ho, wo;         /* output [n][ho][wo][cout]   */
  int32_t dtype;          /* activations / weights dtype (CGAT_F32 or CGAT_BF16); bias fp32 */
  int32_t act;            /* fused epilogue: 0 none, 1 ReLU, 2 LeakyReLU(0.2), 3 sigmoid     */
  int32_t groups;         /* 1 = dense; cin = cout-groups for depthwise (SmaAt-UNet). weight [cout][kh][kw][cin/groups] */
} cgat_conv_desc;

/* K1 fprop / K2 dgrad / K3 wgrad.  `impl`: 0 = CUDA-core kernels (any shape, no workspace: vectorised depthwise 3x3,
 * full-window dot, pointwise / implicit-GEMM tiles, generic direct),
 * 1 = tcgen05 implicit GEMM (bf16, stride 1; cgat_conv_tc_supported tells whether the shape is served): operands
 *     streamed by TMA for >= 64 channels (csrc/conv_tc_big.cu), weights resident in shared memory below that
 *     (csrc/conv_tc.cu).
 * `workspace`: device scratch of at least cgat_conv_workspace_bytes(d, which) bytes (packed weights /
 * partial sums), 16-byte aligned, owned by the caller; may be NULL when that size is 0.
 * wgrad WRITES dw [cout][kh][kw][cin] fp32 and, if dbias != NULL, dbias [cout] fp32.                */
/* Optional scratch of the impl = 0 wgrad: with a workspace of this many bytes the bias gradient is reduced by coalesced
 * per-CTA partial rows + a fixed-order final sum (deterministic); with workspace == NULL by one block per channel.     */
int64_t cgat_conv_dbias_workspace_bytes(const cgat_conv_desc* d);
int cgat_conv2d_fprop(const cgat_conv_desc* d, const void* x, const void* w, const float* bias, void* y,
                      int impl, void* workspace, void* stream);
int cgat_conv2d_dgrad(const cgat_conv_desc* d, const void* dy, const void* w, void* dx, int impl,
                      void* workspace, void* stream);
int cgat_conv2d_wgrad(const cgat_conv_desc* d, const void* x, const void* dy, float* dw, float* dbias,
                      int impl, void* workspace, void* stream);
int cgat_conv_tc_supported(const cgat_conv_desc* d, int which /*0 fprop,1 dgrad,2 wgrad*/);
int64_t cgat_conv_workspace_bytes(const cgat_conv_desc* d, int which);
/* Whether the packed-weight entry points below (the one-launch conv-GAT stream path; resident-weight kernels only) serve
 * the shape -- fprop and wgrad, plus dgrad when need_dx -- and the size of cgat_conv2d_wgrad_partial's workspace. */
int cgat_conv_stream_supported(const cgat_conv_desc* d, int need_dx);
int64_t cgat_conv_stream_workspace_bytes(const cgat_conv_desc* d);

/* Variants used by the fused conv-GAT stream: fprop with weights already packed by cgat_stream_prepare (no
 * per-call packing launch), and wgrad that leaves its per-CTA partial sums [ncta][128][nt] in `workspace`
 * (reduced by cgat_stream_param_grads together with the block-diagonal contraction).                         */
int cgat_conv2d_fprop_packed(const cgat_conv_desc* d, const void* x, const void* wpack, const float* bias, void* y,
                             void* stream);
int cgat_conv2d_dgrad_packed(const cgat_conv_desc* d, const void* dy, const void* wpack, void* dx, void* stream);
int cgat_conv2d_wgrad_partial(const cgat_conv_desc* d, const void* x, const void* dy, void* workspace,
                              int32_t* ncta_out, int32_t* nt_out, void* stream);

/* One conv-GAT stream's parameter plumbing (heads of baseline_model.py:191-192 style sub-modules `attention_{k}`).
 * mapping 0: linear (w[k] = W [ci][co]); mapping 1: conv (w[k] = conv.weight [co][ci][3][3], bias[k] = conv.bias).
 * The arrays w/bias/a/B/g_* are HOST arrays of `heads` device pointers.                                        */
typedef struct cgat_stream_desc {
  int32_t nodes, ci, co, heads;
  int32_t layout;        /* CGAT_LAYOUT_*                                  */
  int32_t mapping;       /* 0 linear, 1 conv 3x3 pad 1                      */
  int32_t transpose_adj; /* 1: the 1-D layer's A_hat^T (baseline_model.py:53) */
  int32_t wgrad_cols;    /* 0: conv_tc packing, wgrad partial columns [tap][cin] (cgat_conv2d_wgrad_partial);
                            1: fused layer kernels (cgat_layer_*): K-chunks j = (r*cin/8 + c)*3 + s (vertical tap r, channel
                               chunk c, horizontal tap s), partial columns [j][8], the bias as a separate K-chunk, and -- while
                               heads*(nodes*co + 2*nodes rounded up to 8) <= 128 -- score rows W.a behind the heads*nodes*co
                               feature rows (see cgat_stream_param_grads)                                           */
} cgat_stream_desc;

/* bytes of the packed bf16 weight buffer of the block-diagonal dense conv (dgrad != 0: the dgrad packing) */
int64_t cgat_stream_wpack_bytes(const cgat_stream_desc* d, int dgrad);
/* ONE launch: stack a, normalise the adjacency of every head (baseline_model.py:41-50), and either stack W
 * (linear) or expand + pack the conv weights and bias (conv).  wpack_dgrad may be NULL.                       */
int cgat_stream_prepare(const cgat_stream_desc* d, const float* const* w, const float* const* bias,
                        const float* const* a, const float* const* B, void* wpack, void* wpack_dgrad,
                        float* w_stacked, float* bias_dense, float* a_stacked, float* adj, void* stream);
/* The same launch also clears `clear_bytes` bytes at `clear` (16-byte aligned, a multiple of 16): the train step's
 * accumulators and gradient buffer, so that the step needs no separate memset node in front of it.             */
int cgat_stream_prepare_clear(const cgat_stream_desc* d, const float* const* w, const float* const* bias,
                              const float* const* a, const float* const* B, void* wpack, void* wpack_dgrad,
                              float* w_stacked, float* bias_dense, float* a_stacked, float* adj, void* clear,
                              int64_t clear_bytes, void* stream);
/* ONE launch: per-head parameter gradients from the kernels' accumulators (wgrad partials or gW, ga, gadj);
 * accumulate != 0 adds into g_* (e.g. the parameters' .grad buffers) instead of overwriting.  w, bias, a: the same
 * per-head parameter pointers cgat_stream_prepare took; read only when wgrad_cols = 1 and the partials carry the score
 * rows d(W.a) of the fused layer kernels (then  dW += a (x) d(W.a)  and  d(a) = <W, d(W.a)>  are formed here); may be
 * NULL otherwise.  gbias [heads][co + 2]: the fused layer kernels' sums of d(Wh) per output channel and of ds1, ds2
 * (wgrad_cols = 1 only: the bias gradient does not come out of their wgrad MMA); NULL otherwise.                 */
int cgat_stream_param_grads(const cgat_stream_desc* d, const float* wg_partial, int ncta, int nt, const float* gW_lin,
                            const float* ga, const float* gadj, const float* gbias, const float* const* B, const float* const* w,
                            const float* const* bias, const float* const* a, float* const* g_w,
                            float* const* g_bias, float* const* g_a, float* const* g_B, int accumulate, void* stream);
/* The fused layer kernels' (cgat_layer_bwd / cgat_layer_train, wgrad_cols = 1) side of it, as ONE launch: reduces their
 * compact per-CTA partial sums (ncta slots of [rows][nt = 9 ci] floats; slot `ncta` of the workspace is scratch), forms every
 * parameter gradient (conv weight / bias over the block-diagonal, d(a) through the score rows, adjacency backward) and
 * optionally applies torch.optim.Adam to the flat parameter buffer in the same launch (single GPU: nothing sits between
 * the gradients and the optimiser).
 *   counter            THREE uint32 the caller zeroes before every call (grid barriers between the launch's phases)
 *   select, alt_offset, loss_mse   the guarded train step (see cgat_layer_train): select may be NULL; when select[0] != 0
 *                      the accumulators are read alt_offset floats further on (the set cgat_layer_train_fp32 filled) and, if
 *                      loss_mse != NULL, loss_mse[0..1] := loss_mse[alt_offset .. alt_offset + 1]
 *   adam_param         NULL: gradients only.  Else flat fp32 param / grad / m / v of adam_n elements (every g_* pointer
 *                      must point into adam_grad), *adam_step_dev = steps taken so far (incremented by the launch),
 *                      adam_hyper = device floats {lr, beta1, beta2, eps, weight_decay, grad_scale}: both live on the device
 *                      so that the launch can be replayed from a CUDA graph while schedulers change lr.              */
int cgat_stream_finish(const cgat_stream_desc* d, const float* wg_partial, int ncta, int nt, const float* ga,
                       const float* gadj, const float* gbias, const float* const* B, const float* const* w,
                       const float* const* bias, const float* const* a, float* const* g_w, float* const* g_bias,
                       float* const* g_a, float* const* g_B, int accumulate, const float* select, int64_t alt_offset,
                       float* loss_mse, uint32_t* counter, float* adam_param, const float* adam_grad, float* adam_m,
                       float* adam_v, int64_t adam_n, int64_t* adam_step_dev, const float* adam_hyper, void* stream);
/* The same launch, and the step's scalar loss written straight to HOST memory: mirror = a ring of mirror_n floats in mapped
 * pinned memory (device-accessible), *mirror_cursor = values written so far (device counter, advanced by the launch),
 * mirror_src = the base of the accumulator set that holds the loss at [0] (at [alt_offset] when the fp32 re-run is
 * selected).  Replaces a 4-byte device-to-host memcpy between two steps (train.py:135: loss.item()) by a posted write. */
int cgat_stream_finish_mirror(const cgat_stream_desc* d, const float* wg_partial, int ncta, int nt, const float* ga,
                              const float* gadj, const float* gbias, const float* const* B, const float* const* w,
                              const float* const* bias, const float* const* a, float* const* g_w, float* const* g_bias,
                              float* const* g_a, float* const* g_B, int accumulate, const float* select, int64_t alt_offset,
                              float* loss_mse, uint32_t* counter, float* adam_param, const float* adam_grad, float* adam_m,
                              float* adam_v, int64_t adam_n, int64_t* adam_step_dev, const float* adam_hyper,
                              const float* mirror_src, float* mirror, int32_t mirror_n, uint32_t* mirror_cursor, void* stream);

/* K6 / K7  one conv-mapped stream of the conv-GAT layer (shared 3x3 node conv, pad 1, + graph attention) as ONE
 * kernel per direction: the projected features never touch HBM (tcgen05 accumulators in TMEM are read by the
 * attention threads directly).  Replaces, for mapping_type="conv", the node conv at the call sites
 * convolutional_gat/model.py:21-42 plus baseline_model.py:127-160 (see cgat_attn_fwd).  bf16 activations.
 *   x      [n][nodes*ci/8][h][wp][8], wp = 8*ceil(w/8) + 2: PADDED CHUNK-PLANAR (x_layout = CGAT_X_PLANAR): 16-byte
 *          chunk c of the pixel record [T*V] of the loaders' [N,H,W,T,V] layout in plane c, image column i at padded
 *          column i + 1, every other column ZERO.  cgat_loader_gather_planar writes it directly,
 *          cgat_records_to_planar converts pixel records; the padding columns are never written by either (allocate
 *          the buffer zeroed once).  One 5-D TMA box per 16x8-pixel tile then fills a whole pipeline stage.
 *   wpack, bias_dense, a, adj                       as produced by cgat_stream_prepare (fprop packing)
 *   out    [n][h][w][nodes*co] (MERGE_MEAN) or the concat layout of cgat_attn_fwd
 *   dwh    optional (may be NULL): d(Wh) [n][h][w][heads*nodes*co] for cgat_conv2d_dgrad_packed
 *   workspace  cgat_layer_workspace_bytes(d) bytes: per-CTA wgrad partial sums (compact slots of [rows][nt = 9 ci] floats:
 *              of a dense row only the columns of its own node, order [tap][ci]; rows = feature + score rows rounded up to
 *              32), plus one more slot that cgat_stream_finish uses as its reduction scratch
 *   ga [heads][2co], gadj [heads][nodes][nodes], gbias [heads][co+2]     fp32, ACCUMULATED INTO                 */
typedef struct cgat_layer_desc {
  int32_t n, h, w;
  int32_t nodes, ci, co, heads;
  int32_t layout;    /* CGAT_LAYOUT_* */
  int32_t merge;     /* CGAT_MERGE_*  */
  int32_t apply_elu;
  float alpha;
  int32_t x_layout;  /* must be CGAT_X_PLANAR (1): x is padded chunk-planar, see above.  CGAT_X_RECORDS (0), the pixel
                        records themselves, is rejected: convert with cgat_records_to_planar                        */
} cgat_layer_desc;
#define CGAT_X_RECORDS 0
#define CGAT_X_PLANAR 1
int cgat_layer_supported(const cgat_layer_desc* d);
int64_t cgat_layer_workspace_bytes(const cgat_layer_desc* d);
int cgat_layer_fwd(const cgat_layer_desc* d, const void* x, const void* wpack, const float* bias_dense, const float* a,
                   const float* adj, const uint8_t* mask, void* out, void* stream);
int cgat_layer_bwd(const cgat_layer_desc* d, const void* x, const void* dout, const void* wpack,
                   const float* bias_dense, const float* a, const float* adj, const uint8_t* mask, void* dwh,
                   void* workspace, float* ga, float* gadj, float* gbias, int32_t* ncta_out, int32_t* nt_out,
                   void* stream);

/* The same backward kernel in TRAIN mode: the reference step  loss = MSE(model(x), y) - lambda*mean(model(x));
 * loss.backward()  (convolutional_gat/train.py:130-132) for a model that is one mean-merged conv stream (the
 * Spatial/Temporal models of convolutional_gat/model.py:8-88 use only their hidden layer).  The forward is
 * recomputed in-kernel anyway, so out and d(out) never exist in HBM: reads x and y, writes the parameter-gradient
 * partial sums and ACCUMULATES the scalar loss into loss_out[0] and, if mse_out != NULL, the plain mean squared error
 * (the reference's running train loss, train.py:135-139) into mse_out[0].  heads <= 3, CGAT_MERGE_MEAN.        */
int cgat_layer_train(const cgat_layer_desc* d, const void* x, const void* y, const void* wpack,
                     const float* bias_dense, const float* a, const float* adj, const uint8_t* mask, float lambda,
                     void* workspace, float* ga, float* gadj, float* gbias, float* loss_out, float* mse_out, float* guard,
                     int32_t* ncta_out, int32_t* nt_out, void* stream);
/* Arithmetic of cgat_layer_train: bf16 operands on the tensor cores (fp32 accumulation), the per-pixel attention math in
 * PACKED fp16 (two pixels per instruction), every sum that leaves the kernel in fp32.  fp16 resolves the attention logits
 * only while they are O(10) and overflows at 65 504, so the kernel checks itself: guard (optional, may be NULL) is ONE
 * float the caller zeroes; the kernel stores 1 there when a score half |s1|, |s2| exceeded 8 or any of its sums is not
 * finite.  Such a step must be recomputed by cgat_layer_train_fp32 -- the same kernel in its fp32 instantiation (one
 * tile per pass, ~1.6x the time): same arguments; with run_if != NULL the launch is a no-op unless run_if[0] != 0, so
 * the pair (train with guard = g; train_fp32 into a second set of accumulators with run_if = g) can sit in one captured
 * graph, and cgat_stream_finish takes the selector (select / alt_offset) to read the set that is valid.           */
int cgat_layer_train_fp32(const cgat_layer_desc* d, const void* x, const void* y, const void* wpack,
                          const float* bias_dense, const float* a, const float* adj, const uint8_t* mask, float lambda,
                          void* workspace, float* ga, float* gadj, float* gbias, float* loss_out, float* mse_out,
                          const float* run_if, int32_t* ncta_out, int32_t* nt_out, void* stream);

/* a8  the 1-D layer after its GEMM: GraphAttentionLayer.forward lines 36-56 of convolutional_gat/baseline_model.py
 * (scores :36-38 / :58-65, soft-max over neighbours :39, attention <- A_hat . attention :53, aggregation :54,
 * ELU :56).  fp32.  Wh [n][v][f]; a [2f]; adj = A_hat [v][v]; s1, s2 [n][v]; att, M [n][v][v]; out [n][v][f].
 * The backward needs dadj [v][v] and dM [n][v][v] ZEROED by the caller; dWh, da, ds1, ds2 are written.            */
int cgat_gat1d_fwd(const float* Wh, const float* a, const float* adj, const uint8_t* mask, float* s1, float* s2,
                   float* att, float* M, float* out, int n, int v, int f, float alpha, void* stream);
int cgat_gat1d_bwd(const float* Wh, const float* a, const float* adj, const uint8_t* mask, const float* s1,
                   const float* s2, const float* att, const float* M, const float* out, const float* dout, float* dWh,
                   float* da, float* dadj, float* dM, float* ds1, float* ds2, int n, int v, int f, float alpha,
                   void* stream);

/* a12  train-step pieces, convolutional_gat/train.py:131 and :212.
 * loss = mean((yhat-y)^2) - lambda*mean(yhat); writes dloss/dyhat (same dtype as yhat) and
 * ACCUMULATES the scalar loss into loss_out[0] and, if mse_out != NULL, mean((yhat-y)^2) into mse_out[0]
 * (fp32; caller zeroes).                                                                           */
int cgat_loss_fwd_bwd(const void* yhat, const void* y, void* dyhat, float* loss_out, float* mse_out, int64_t n,
                      float lambda, float grad_scale, int dtype, void* stream);
/* torch.optim.Adam(lr, weight_decay) on flat fp32 buffers; `step` is the 1-based step count read
 * from device memory (so the launch is CUDA-graph replayable); grad is multiplied by grad_scale
 * (1/world after a sum all-reduce).                                                                 */
int cgat_adam_step(float* param, const float* grad, float* m, float* v, const int64_t* step_dev, int64_t n,
                   float lr, float beta1, float beta2, float eps, float weight_decay, float grad_scale,
                   void* stream);
/* 8(b)  gradient exchange for hosts that are not PyTorch: NCCL (ncclAllReduce over NVLink / NVSwitch) bound at run time
 * with dlopen("libnccl.so.2") -- no link-time dependency.  rank 0 obtains the 128-byte id (cgat_comm_unique_id) and ships
 * it to the other ranks by its own means; every rank creates ONE communicator (cgat_comm_init) and calls
 * cgat_flat_allreduce (in-place SUM over the flat fp32 gradient buffer, on `stream`; the 1/world scale is the grad_scale of
 * the Adam entry points) once per step.  The Python package keeps torch.distributed's communicator; models of <= 2^20
 * parameters use cgat_p2p_allreduce_adam instead.  cgat_comm_available: 1 when libnccl could be loaded.            */
int cgat_comm_available(void);
int cgat_comm_unique_id(void* id128);
int cgat_comm_init(int32_t rank, int32_t world, const void* id128, void** comm_out);
int cgat_flat_allreduce(void* comm, float* buf, int64_t n, void* stream);
int cgat_comm_destroy(void* comm);

/* a11 / f2  BatchNorm2d + activation + Dropout2d around the convs, NHWC (csrc/norm_act_kernels.cu): the rest of the
 * reference's ConvBlock (dcgan/model.py:35-52: BatchNorm2d -> Dropout2d(0.01) -> activation) and of the SmaAt-UNet double
 * convs (BatchNorm2d -> ReLU).  x, y, dy, dx [n][hw][c] of `dtype` (fp32 / bf16); statistics and parameters fp32 [c].
 * act: 0 none, 1 ReLU, 2 LeakyReLU(slope), 3 sigmoid.  mask: optional [n][c] floats (0 or 1/(1-p), Dropout2d: whole
 * channels of a sample), applied AFTER the activation (identical to the block's order for ReLU / LeakyReLU).
 *   cgat_bn_stats      train-mode batch statistics: mean, rstd = 1/sqrt(biased var + eps); if running_mean != NULL also
 *                      running = (1-momentum)*running + momentum*batch (UNBIASED variance, as torch) and
 *                      ++*num_batches_tracked.  workspace: cgat_bn_workspace_bytes(c) bytes (cleared by the call).
 *   cgat_bn_act_fwd    y = act((x - mean) * rstd * gamma + beta) * mask;  mean == NULL: no normalisation (z = x).
 *   cgat_bn_act_bwd    dz = dy * act'(z) * mask (z recomputed from x);  dbeta = sum dz, dgamma = sum dz * xhat (skipped when
 *                      dgamma == NULL and the statistics are constants);  dx = gamma*rstd*(dz - dbeta/M - xhat*dgamma/M)
 *                      when training != 0 (batch statistics), gamma*rstd*dz otherwise.  Two launches.
 *   cgat_dropout2d_mask  mask[i] = Bernoulli(1-p) / (1-p), i < n, Philox4x32-10 keyed by (seed, *counter, i); ++*counter.  */
int64_t cgat_bn_workspace_bytes(int32_t c);
int cgat_bn_stats(const void* x, int32_t dtype, int64_t n, int64_t hw, int32_t c, void* workspace, float* mean, float* rstd,
                  float* running_mean, float* running_var, int64_t* num_batches_tracked, float momentum, float eps,
                  void* stream);
int cgat_bn_act_fwd(const void* x, void* y, int32_t dtype, int64_t n, int64_t hw, int32_t c, const float* mean,
                    const float* rstd, const float* gamma, const float* beta, const float* mask, int32_t act, float slope,
                    void* stream);
int cgat_bn_act_bwd(const void* x, const void* dy, void* dx, int32_t dtype, int64_t n, int64_t hw, int32_t c,
                    const float* mean, const float* rstd, const float* gamma, const float* beta, const float* mask,
                    int32_t act, float slope, int32_t training, void* workspace, float* dgamma, float* dbeta,
                    int32_t accumulate, void* stream);
/* The same three passes over `sets` statistic sets: images [s*n/sets, (s+1)*n/sets) are normalised with THEIR OWN batch
 * statistics and the running statistics take one momentum update per set, set 0 first -- the reference's UnetModel
 * pushing vertex after vertex through one shared UNet (convolutional_gat/unet_model.py:25-26), as one launch per pass.
 * mean, rstd [sets][c]; workspace cgat_bn_workspace_bytes_sets(c, sets); num_batches_tracked += sets; backward:
 * set_sums [2][sets][c] receives the per-set sum dz | sum dz*xhat (read by the apply pass), dgamma / dbeta [c] their
 * totals over the sets (added to when accumulate != 0).                                                              */
int64_t cgat_bn_workspace_bytes_sets(int32_t c, int32_t sets);
int cgat_bn_stats_sets(const void* x, int32_t dtype, int64_t n, int64_t hw, int32_t c, int32_t sets, void* workspace,
                       float* mean, float* rstd, float* running_mean, float* running_var, int64_t* num_batches_tracked,
                       float momentum, float eps, void* stream);
int cgat_bn_act_fwd_sets(const void* x, void* y, int32_t dtype, int64_t n, int64_t hw, int32_t c, int32_t sets,
                         const float* mean, const float* rstd, const float* gamma, const float* beta, const float* mask,
                         int32_t act, float slope, void* stream);
int cgat_bn_act_bwd_sets(const void* x, const void* dy, void* dx, int32_t dtype, int64_t n, int64_t hw, int32_t c,
                         int32_t sets, const float* mean, const float* rstd, const float* gamma, const float* beta,
                         const float* mask, int32_t act, float slope, int32_t training, void* workspace, float* set_sums,
                         float* dgamma, float* dbeta, int32_t accumulate, void* stream);
int cgat_dropout2d_mask(float* mask, int64_t n, float p, uint64_t seed, uint64_t* counter, void* stream);

/* f2  the non-conv ops of the SmaAt-UNet applied per vertex by convolutional_gat/unet_model.py:20-29 (public
 * architecture), NHWC, dtype fp32 / bf16 (csrc/unet_glue_kernels.cu).  Tensors [n][h][w][c] unless stated.
 *   cgat_maxpool2_*      nn.MaxPool2d(2): y [n][h/2][w/2][c], idx one byte per output (position of the first maximum in the
 *                        window, row-major); backward writes all of dx.
 *   cgat_upcat_*         out [n][h][w][c2+c1] = cat(x2 [n][h][w][c2], zero_pad(bilinear_x2(x1 [n][h1][w1][c1],
 *                        align_corners=True))) with pad offsets ((h-2h1)/2, (w-2w1)/2) as UpDS.forward; backward writes
 *                        dx1 and dx2.
 *   cgat_pool_hw         avg, mx [n][c] fp32 and the first arg-max pixel (int32) over the hw pixels of every (image,
 *                        channel): adaptive_avg_pool2d / adaptive_max_pool2d to 1x1.  workspace:
 *                        cgat_pool_hw_workspace_bytes.  cgat_dot_hw: out[n][c] = sum over pixels of x * dy.
 *   cgat_cbam_mlp_*      CBAM channel gate: scale = sigmoid(MLP(avg) + MLP(mx)), MLP = Linear(c, hid) - ReLU - Linear(hid,
 *                        c) with nn.Linear weight layout; pre [n][2][hid] keeps the hidden pre-activations.  The backward
 *                        takes d(scale) and writes d(pre-sigmoid) ds [n][c], dpre, d(avg), d(mx) and the four parameter
 *                        gradients (overwritten).
 *   cgat_gate_channels_* y = x * scale[n][c];  backward: dx = dy * scale + davg/hw + (pixel == argmax ? dmax : 0), i.e. the
 *                        gate's and both pooling branches' input gradients in one pass (d(scale) itself = cgat_dot_hw).
 *   cgat_chan_pool_*     pooled [npix][2] = (mean, max over the c channels of a pixel), argmax [npix]; backward:
 *                        dx = d(mean)/c + (channel == argmax ? d(max) : 0).
 *   cgat_gate_pixels     y = x * s[pixel] (s of the tensor's dtype; also the backward dx = dy * s); cgat_chan_dot:
 *                        out[pixel] = sum over channels of x * dy (= d(s)).                                             */
int cgat_maxpool2_fwd(const void* x, void* y, uint8_t* idx, int32_t dtype, int64_t n, int32_t h, int32_t w, int32_t c,
                      void* stream);
int cgat_maxpool2_bwd(const void* dy, const uint8_t* idx, void* dx, int32_t dtype, int64_t n, int32_t h, int32_t w, int32_t c,
                      void* stream);
int cgat_upcat_fwd(const void* x1, const void* x2, void* out, int32_t dtype, int64_t n, int32_t h1, int32_t w1, int32_t c1,
                   int32_t h, int32_t w, int32_t c2, void* stream);
int cgat_upcat_bwd(const void* dout, void* dx1, void* dx2, int32_t dtype, int64_t n, int32_t h1, int32_t w1, int32_t c1,
                   int32_t h, int32_t w, int32_t c2, void* stream);
int64_t cgat_pool_hw_workspace_bytes(int64_t n, int64_t hw, int32_t c);
int cgat_pool_hw(const void* x, int32_t dtype, int64_t n, int64_t hw, int32_t c, void* workspace, float* avg, float* mx,
                 int32_t* argmax, void* stream);
int cgat_dot_hw(const void* x, const void* dy, int32_t dtype, int64_t n, int64_t hw, int32_t c, void* workspace, float* out,
                void* stream);
int cgat_cbam_mlp_fwd(const float* avg, const float* mx, const float* w1, const float* b1, const float* w2, const float* b2,
                      int32_t n, int32_t c, int32_t hid, float* pre, float* scale, void* stream);
int cgat_cbam_mlp_bwd(const float* dscale, const float* scale, const float* pre, const float* avg, const float* mx,
                      const float* w1, const float* w2, int32_t n, int32_t c, int32_t hid, float* ds, float* dpre, float* davg,
                      float* dmax, float* dw1, float* db1, float* dw2, float* db2, void* stream);
int cgat_gate_channels_fwd(const void* x, const float* scale, void* y, int32_t dtype, int64_t n, int64_t hw, int32_t c,
                           void* stream);
int cgat_gate_channels_bwd(const void* dy, const float* scale, const float* davg, const float* dmax, const int32_t* argmax,
                           void* dx, int32_t dtype, int64_t n, int64_t hw, int32_t c, void* stream);
int cgat_gate_pixels(const void* x, const void* s, void* y, int32_t dtype, int64_t n, int64_t hw, int32_t c, void* stream);
int cgat_chan_pool_fwd(const void* x, void* pooled, int32_t* argmax, int32_t dtype, int64_t npix, int32_t c, void* stream);
int cgat_chan_pool_bwd(const void* dpooled, const int32_t* argmax, void* dx, int32_t dtype, int64_t npix, int32_t c,
                       void* stream);
int cgat_chan_dot(const void* x, const void* dy, void* out, int32_t dtype, int64_t npix, int32_t c, void* stream);

/* a11  zero-pad (1 pixel) + 2x2 space-to-depth of an NHWC activation in one pass -- the regrouping that serves the DCGAN
 * discriminators' k=4, stride-2, padding-1 convs (dcgan/model.py:150-165) as stride-1 2x2 convs over 4c channels:
 *   inverse == 0:  src x [n][h][w][c] -> dst xs [n][h/2+1][w/2+1][2][2][c],  xs[n][i][j][a][b][:] = x[n][2i+a-1][2j+b-1][:] or 0
 *   inverse != 0:  src dxs (that layout) -> dst dx [n][h][w][c]  (the exact inverse gather = the backward).  h, w even.   */
int cgat_s2d_pad(const void* src, void* dst, int32_t dtype, int64_t n, int32_t h, int32_t w, int32_t c, int32_t inverse,
                 void* stream);

/* f3  the KNMI loader's windowing + normalisation + layout change on the device, replacing
 * convolutional_gat/data_loaders/kmni_data_loader.py:72-127 (__segmentify and the permute of __next__):
 *   x[s, h, w, t, v] = pow(frames[start[s] + t, v, h, w] / normalizing_max, power)            t < steps
 *   y[s, h, w, t, v] = pow(frames[start[s] + steps + t, v, h, w] / normalizing_max, power)
 * frames [n_frames][vertices][h][w] uint8 (the raw integer frames, 0..254), start [n] int32 window starts (every
 * start + 2*steps <= n_frames: the caller's responsibility), x, y [n][crop_h][crop_w][steps][vertices] of `dtype`.  */
int cgat_loader_gather(const uint8_t* frames, int64_t n_frames, const int32_t* start, void* x, void* y, int32_t n,
                       int32_t vertices, int32_t h, int32_t w, int32_t crop_h, int32_t crop_w, int32_t steps,
                       float normalizing_max, float power, int32_t dtype, void* stream);
/* The same gather with x written PADDED CHUNK-PLANAR (CGAT_X_PLANAR, see cgat_layer_fwd): x_planar
 * [n][steps*vertices/8][crop_h][wp][8] bf16, wp = 8*ceil(crop_w/8) + 2, image column i at padded column i + 1 (the
 * padding columns are NOT written: zero the buffer once); y stays [n][crop_h][crop_w][steps][vertices] bf16.        */
int cgat_loader_gather_planar(const uint8_t* frames, int64_t n_frames, const int32_t* start, void* x_planar, void* y,
                              int32_t n, int32_t vertices, int32_t h, int32_t w, int32_t crop_h, int32_t crop_w,
                              int32_t steps, float normalizing_max, float power, void* stream);
/* f3  the ARAI loader's windowing + layout change on the device, replacing
 * convolutional_gat/data_loaders/arai_data_loader.py:57-93 (__batchify + fix_sizes): FLOAT frames
 * [n_frames][vertices][h][w] (the files' [L][regions][1][H][W]), no normalisation:
 *   x[s, h, w, t, v] = frames[start[s] + t, v, h, w]     y[s, h, w, t, v] = frames[start[s] + steps + t, v, h, w]
 * x, y [n][crop_h][crop_w][steps][vertices] of `dtype` (fp32: bit-exact copy; bf16: rounded once).                    */
int cgat_loader_gather_f32(const float* frames, int64_t n_frames, const int32_t* start, void* x, void* y, int32_t n,
                           int32_t vertices, int32_t h, int32_t w, int32_t crop_h, int32_t crop_w, int32_t steps,
                           int32_t dtype, void* stream);
/* pixel records [n][h][w][rec] bf16 -> padded chunk-planar [n][rec/8][h][wp][8] for an x tensor that did not come from
 * the loader kernel (the reference's own DataLoader, convolutional_gat/train.py:128); padding columns not written.   */
int cgat_records_to_planar(const void* x, void* x_planar, int64_t n, int32_t h, int32_t w, int32_t rec, void* stream);
/* f4  validation metrics of convolutional_gat/train.py:53-75 + utils.py:135-167 in one pass over y, y_hat (n elements of
 * `dtype`): ACCUMULATES into out6 (double, caller zeroes)  [0] sum (y'-yh')^2  [1] sum ((y'-yh')*normalizing_max)^2
 * [2] TP [3] FP [4] FN [5] #(bin(y') == bin(yh')),  y' = y^(1/power), bin = the threshold binarisation of utils.py:138-141. */
int cgat_val_metrics(const void* y, const void* y_hat, int64_t n, float power, float threshold, float normalizing_max,
                     int32_t dtype, double* out6, void* stream);
/* Data-parallel gradient exchange + Adam as ONE kernel over NVLink peer memory (small models).  Every
 * rank owns a mailbox of cgat_p2p_mailbox_bytes(n, world) bytes in peer-mapped (symmetric) memory, ZEROED once before
 * the first step; peer_mailboxes is a HOST array of the `world` device addresses (index = rank).  The kernel pushes
 * `grad` into every peer's mailbox as {value, epoch} words (the flag travels with the data: no fence, no flag hop),
 * waits per element for all sources, sums them in rank order (bit-identical on every
 * rank), and applies torch.optim.Adam with the 1/world mean folded in (train.py:212).  The 1-based step count (also the
 * exchange epoch) is read from `step_dev`, or taken from `step_host` when step_dev is NULL.  n <= 2^20 floats; larger models use an NCCL all-reduce + cgat_adam_step.        */
int64_t cgat_p2p_mailbox_bytes(int64_t n, int32_t world);
int cgat_p2p_allreduce_adam(const uint64_t* peer_mailboxes, int32_t rank, int32_t world, const float* grad, float* param,
                            float* m, float* v, const int64_t* step_dev, int64_t step_host, int64_t n, float lr,
                            float beta1, float beta2, float eps, float weight_decay, void* stream);
/* The same exchange + Adam as a CUDA-graph node: the step is *step_counter + 1 (steps taken so far, device memory; the launch
 * stores the new count when it has applied the update, and leaves it on a time-out), hyper = device {lr, beta1, beta2, eps,
 * weight_decay} (a scheduler rewrites lr between replays).  Launched with programmatic dependent launch: it may be
 * scheduled while the kernel that produces `grad` still runs and waits for it on the device.                         */
int cgat_p2p_allreduce_adam_graph(const uint64_t* peer_mailboxes, int32_t rank, int32_t world, const float* grad, float* param,
                                  float* m, float* v, int64_t* step_counter, const float* hyper, int64_t n, void* stream);
/* the same update with the 1-based step count passed by value (no device counter, no increment kernel) */
int cgat_adam_step_at(float* param, const float* grad, float* m, float* v, int64_t step, int64_t n, float lr, float beta1,
                      float beta2, float eps, float weight_decay, float grad_scale, void* stream);
/* dtype conversion of contiguous buffers (fp32 <-> bf16) */
int cgat_cast(const void* src, int src_dtype, void* dst, int dst_dtype, int64_t n, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CGAT_B200_H_ */
