// Shared helpers for the sm_100a kernels: error plumbing, dtype traits, mbarrier / bulk-copy PTX.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdint>
#include <cstdio>
#include <cstdarg>
#include <cstdlib>
#include <utility>

#include "../../include/cgat_b200.h"

namespace cgat {

// ---- error plumbing (thread-local message, C ABI never throws) -------------------------------
char* last_error_buf();
inline int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(last_error_buf(), 512, fmt, ap);
  va_end(ap);
  return code;
}
inline int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail((int)e, "%s: %s", what, cudaGetErrorString(e));
  return 0;
}
inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// ---- dtype traits ------------------------------------------------------------------------------
template <typename T> struct DT;
template <> struct DT<float> {
  static __device__ __forceinline__ float to_f(float v) { return v; }
  static __device__ __forceinline__ float from_f(float v) { return v; }
};
template <> struct DT<__nv_bfloat16> {
  static __device__ __forceinline__ float to_f(__nv_bfloat16 v) { return __bfloat162float(v); }
  static __device__ __forceinline__ __nv_bfloat16 from_f(float v) { return __float2bfloat16_rn(v); }
};

#if defined(__CUDACC__)
// ---- programmatic dependent launch: the kernels of one train step form a chain of small launches; a kernel launched
// with launch_pdl may START (CTA scheduling, set-up, loads that do not depend on its predecessor) while the previous
// kernel of the stream still runs, and calls griddep_wait() before it touches anything that kernel produces.  A
// predecessor that calls griddep_launch() early lets it start even earlier.  Captured into CUDA graphs as programmatic
// edges.  CGAT_NO_PDL=1 falls back to plain stream order (griddep_wait is then a no-op).
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void griddep_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  static const bool no_pdl = std::getenv("CGAT_NO_PDL") != nullptr;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = no_pdl ? 0 : 1;
  return cudaLaunchKernelEx(&cfg, kern, KArgs(std::forward<Args>(args))...);
}

// ---- shared-memory address / mbarrier / bulk async copy (TMA 1-D) ------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// try_wait is potentially blocking (the hardware suspends the thread for a bounded time), so the loop in
// mbar_wait does not spin hot.  (Measured on B200: test_wait polling and try_wait with a long suspend-time hint
// give the same kernel times; the plain form is kept.)
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
// global -> shared bulk copy (bytes multiple of 16, both addresses 16-byte aligned), completes on mbarrier
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(smem_dst)),
      "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
// shared -> global bulk copy, tracked by the bulk async-group of the issuing thread
__device__ __forceinline__ void bulk_s2g(void* gdst, const void* smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst),
               "r"(smem_u32(smem_src)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
#endif

// Fused conv-GAT layer (layer_fused.cu / stream_ops.cu): the attention scores s1 = Wh.a[:co], s2 = Wh.a[co:] are linear
// in the layer input, so the dense conv gets extra output rows  W.a  per (head, node) -- `sph` rows per head (2*nodes
// padded to a 16-byte chunk) behind the heads*nodes*co feature rows -- whenever everything still fits one M = 128 /
// N <= 128 tcgen05 tile.  Both translation units must agree on this rule.
__host__ __device__ inline int lf_score_rows_per_head(int nodes) { return (2 * nodes + 7) & ~7; }
__host__ __device__ inline int lf_score_rows(int nodes, int co, int heads) {
  const int ext = heads * lf_score_rows_per_head(nodes);
  return heads * nodes * co + ext <= 128 ? ext : 0;
}
// rows of a CTA's wgrad partial-sum slot (layer_fused.cu writes, stream_ops.cu reads): feature + score rows, rounded up
// to a TMEM lane quarter.  Slot layout: [nt / 8 column blocks][rows][8 columns]  (coalesced writes, see layer_fused.cu)
__host__ __device__ inline int lf_partial_rows(int nodes, int co, int heads) {
  return (heads * nodes * co + lf_score_rows(nodes, co, heads) + 31) & ~31;
}
__host__ __device__ inline size_t lf_partial_index(int row, int col, int rows) {
  return ((size_t)(col >> 3) * rows + row) * 8 + (col & 7);
}
// padded chunk-planar x (CGAT_X_PLANAR): [n][cin/8][h][wp][8]; one zero pixel left of the image, zeros up to the tile grid
// and one more right of it.  Loader, converter and the layer kernels' tensor map must agree on this.
__host__ __device__ inline int lf_padded_width(int w) { return ((w + 7) & ~7) + 2; }
// K-chunks of the packed fprop weights: 9 taps x cin/8, one zero chunk (27 is odd), the bias chunk, one zero chunk
__host__ __device__ inline int lf_weight_chunks(int cin) { return 9 * (cin / 8) + (9 * (cin / 8)) % 2 + 2; }

}  // namespace cgat
