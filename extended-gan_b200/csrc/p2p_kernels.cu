// Gradient exchange + optimiser of the data-parallel train step as ONE kernel over NVLink peer memory.
//
// The conv-GAT models have 6.6K-44K parameters: their gradient all-reduce is pure latency (NCCL: ~12 us at 2 GPUs on
// a 0.13 ms step), and it sits on the step's critical path between the backward and Adam.  Every rank owns a
// MAILBOX in symmetric (peer-mapped) memory, [2 epochs parities][world source ranks][n floats] plus one flag per
// (parity, source).  The kernel
//   1. PUSHES the local gradient into slot `rank` of every peer's mailbox (coalesced 16-byte stores over NVLink),
//   2. fences (system scope) and releases flag[parity][rank] = epoch on every peer,
//   3. acquires its own flags from all sources, sums the `world` slots IN RANK ORDER (every rank computes the same
//      bits: replicas cannot drift) and
//   4. applies torch.optim.Adam (convolutional_gat/train.py:212) with the 1/world mean folded in.
// Epoch = the 1-based step counter read from device memory; slots alternate by its parity, which is enough because
// a rank cannot be two steps ahead of a peer (step e+1 needs that peer's flag of step e+1, sent after it finished e).
// One CTA (the vector is small); larger models keep the NCCL all-reduce.
#include "common.cuh"

namespace cgat {

constexpr int P2P_THREADS = 1024;
constexpr int P2P_MAX_WORLD = 8;

struct P2pPeers {
  float* mailbox[P2P_MAX_WORLD];     // peer p's mailbox base
  uint32_t* flags[P2P_MAX_WORLD];    // peer p's flags [2][world]
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

__global__ void __launch_bounds__(P2P_THREADS)
p2p_allreduce_adam_kernel(const P2pPeers P, int rank, int world, long long n, long long n_pad, const float* __restrict__ g,
                          float* __restrict__ p, float* __restrict__ m, float* __restrict__ v,
                          const long long* __restrict__ step_dev, long long step_host, float lr, float b1, float b2,
                          float eps, float wd) {
  const long long epoch = step_dev != nullptr ? *step_dev : step_host;
  const int par = (int)(epoch & 1);
  const int tid = threadIdx.x;
  // 1. push
  for (int q = 0; q < world; ++q) {
    float* dst = P.mailbox[q] + ((size_t)par * world + rank) * n_pad;
    for (long long i = (long long)tid * 4; i < n; i += (long long)P2P_THREADS * 4) {
      if (i + 4 <= n) {
        *reinterpret_cast<float4*>(dst + i) = *reinterpret_cast<const float4*>(g + i);
      } else {
        for (long long j = i; j < n; ++j) dst[j] = g[j];
      }
    }
  }
  // 2. publish
  __threadfence_system();
  __syncthreads();
  if (tid < world) st_release_sys(P.flags[tid] + par * world + rank, (uint32_t)epoch);
  // 3. wait for every source
  if (tid < world) {
    const uint32_t* f = P.flags[rank] + par * world + tid;
    const long long t0 = clock64();
    while (ld_acquire_sys(f) != (uint32_t)epoch) {
      if (clock64() - t0 > 6000000000ll) {  // ~3 s: a peer never arrived; record it instead of hanging the GPU
        atomicExch(P.flags[rank] + 2 * world, 1u);
        break;
      }
    }
  }
  __syncthreads();
  // 4. rank-ordered sum + Adam
  const float step = (float)epoch;
  const float bc1 = 1.f - powf(b1, step);
  const float bc2 = 1.f - powf(b2, step);
  const float step_size = lr / bc1;
  const float inv_sqrt_bc2 = rsqrtf(bc2);
  const float gscale = 1.f / (float)world;
  const float* mine = P.mailbox[rank] + (size_t)par * world * n_pad;
  for (long long i = tid; i < n; i += P2P_THREADS) {
    float gs = 0.f;
    for (int q = 0; q < world; ++q) gs += __ldcg(mine + (size_t)q * n_pad + i);  // L2: the peers' stores landed there
    const float pi = p[i];
    const float gi = fmaf(wd, pi, gs * gscale);
    const float mi = fmaf(b1, m[i], (1.f - b1) * gi);
    const float vi = fmaf(b2, v[i], (1.f - b2) * gi * gi);
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) * inv_sqrt_bc2 + eps;
    p[i] = pi - step_size * (mi / denom);
  }
}

}  // namespace cgat

using namespace cgat;

extern "C" int64_t cgat_p2p_mailbox_bytes(int64_t n, int32_t world) {
  if (n <= 0 || world < 1 || world > P2P_MAX_WORLD) return 0;
  const int64_t n_pad = (n + 31) & ~(int64_t)31;
  return 2 * (int64_t)world * n_pad * 4 + 256;  // + flags [2][world] u32, then one u32 time-out marker (all zeroed by the caller)
}

extern "C" int cgat_p2p_allreduce_adam(const uint64_t* peer_mailboxes, int32_t rank, int32_t world, const float* grad,
                                       float* param, float* m, float* v, const int64_t* step_dev, int64_t step_host,
                                       int64_t n, float lr, float beta1, float beta2, float eps, float weight_decay,
                                       void* stream) {
  if (!peer_mailboxes || !grad || !param || !m || !v) return fail(CGAT_EINVAL, "null argument");
  if (!step_dev && step_host < 1) return fail(CGAT_EINVAL, "step_dev is NULL and step_host < 1");
  if (world < 1 || world > P2P_MAX_WORLD || rank < 0 || rank >= world) return fail(CGAT_EINVAL, "bad rank/world %d/%d", rank, world);
  if (n <= 0 || n > (1 << 20)) return fail(CGAT_EUNSUPPORTED, "p2p exchange serves vectors of at most 2^20 floats (n=%lld)", (long long)n);
  if (!aligned16(grad)) return fail(CGAT_EALIGN, "grad must be 16-byte aligned");
  const long long n_pad = (n + 31) & ~(long long)31;
  P2pPeers P{};
  for (int q = 0; q < world; ++q) {
    if (!peer_mailboxes[q] || (peer_mailboxes[q] & 15)) return fail(CGAT_EALIGN, "peer mailbox %d null or misaligned", q);
    P.mailbox[q] = reinterpret_cast<float*>(peer_mailboxes[q]);
    P.flags[q] = reinterpret_cast<uint32_t*>(peer_mailboxes[q] + (uint64_t)2 * world * n_pad * 4);
  }
  p2p_allreduce_adam_kernel<<<1, P2P_THREADS, 0, (cudaStream_t)stream>>>(P, rank, world, n, n_pad, grad, param, m, v,
                                                                         (const long long*)step_dev, (long long)step_host, lr,
                                                                         beta1, beta2, eps, weight_decay);
  return check_launch("p2p_allreduce_adam_kernel");
}
