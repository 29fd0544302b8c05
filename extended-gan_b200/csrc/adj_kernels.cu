// a5: learnable adjacency normalisation and its backward (one CTA per head).
//
// Reference: convolutional_gat/baseline_model.py:41-50 and :133-142
//   adj = B + I;  adj = (adj - min adj) / (max adj - min adj);
//   D = diag(rowsum(adj)) detached (legacy Variable(..., requires_grad=False), :48/:140);
//   A_hat = sqrt(inverse(D)) . adj . sqrt(inverse(D)).
// The reference spends ~10 tiny ATen launches (incl. an LU inverse of a diagonal matrix) per layer
// per step on this; here it is one launch for all heads.  Backward follows torch autograd: the
// gradient of a full min()/max() reduction is distributed evenly over tied elements.
#include "adj_math.cuh"

namespace cgat {

__global__ void __launch_bounds__(ADJ_THREADS) adj_norm_fwd_kernel(const float* __restrict__ B,
                                                                   float* __restrict__ out, int nodes,
                                                                   int transpose) {
  adj_norm_fwd_block(B, out, nodes, transpose, blockIdx.x);
}

__global__ void __launch_bounds__(ADJ_THREADS) adj_norm_bwd_kernel(const float* __restrict__ B,
                                                                   const float* __restrict__ g,
                                                                   float* __restrict__ gB, int nodes,
                                                                   int transpose) {
  adj_norm_bwd_block(B, g, gB, nodes, transpose, blockIdx.x, 0);
}

}  // namespace cgat

using namespace cgat;

extern "C" int cgat_adj_norm_fwd(const float* B, float* adj_hat, int heads, int nodes, int transpose, void* stream) {
  if (!B || !adj_hat) return fail(CGAT_EINVAL, "null argument");
  if (heads < 1 || nodes < 1 || nodes > ADJ_MAX_NODES)
    return fail(CGAT_EINVAL, "heads=%d nodes=%d out of range (nodes <= %d)", heads, nodes, ADJ_MAX_NODES);
  adj_norm_fwd_kernel<<<heads, ADJ_THREADS, 0, (cudaStream_t)stream>>>(B, adj_hat, nodes, transpose);
  return check_launch("adj_norm_fwd_kernel");
}

extern "C" int cgat_adj_norm_bwd(const float* B, const float* gadj, float* gB, int heads, int nodes, int transpose,
                                 void* stream) {
  if (!B || !gadj || !gB) return fail(CGAT_EINVAL, "null argument");
  if (heads < 1 || nodes < 1 || nodes > ADJ_MAX_NODES)
    return fail(CGAT_EINVAL, "heads=%d nodes=%d out of range (nodes <= %d)", heads, nodes, ADJ_MAX_NODES);
  adj_norm_bwd_kernel<<<heads, ADJ_THREADS, 0, (cudaStream_t)stream>>>(B, gadj, gB, nodes, transpose);
  return check_launch("adj_norm_bwd_kernel");
}
