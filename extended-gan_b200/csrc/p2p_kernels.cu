// Gradient exchange + optimiser of the data-parallel train step as ONE kernel over NVLink peer memory.
//
// The conv-GAT models have 1K-44K parameters: their gradient all-reduce is pure latency (NCCL: ~12 us at 2 GPUs on
// a 0.13 ms step), and it sits on the step's critical path between the backward and Adam.  Every rank owns a
// MAILBOX in symmetric (peer-mapped) memory, [2 epoch parities][world source ranks][n] 8-byte words.  A word is
// {gradient bits, epoch}: the flag travels WITH the data (the "LL" idea of NCCL's low-latency protocol), so there
// is no fence and no separate flag hop on the critical path -- one NVLink one-way latency instead of three.
// The kernel
//   1. PUSHES {g[i], epoch} into slot `rank` of every peer's mailbox with single 8-byte stores,
//   2. per element spins on its own mailbox until the word of every source carries this epoch, sums the `world`
//      values IN RANK ORDER (every rank computes the same bits: replicas cannot drift) and
//   3. applies torch.optim.Adam (convolutional_gat/train.py:212) with the 1/world mean folded in.
// Epoch = the 1-based step counter; slots alternate by its parity, which is enough because a rank cannot be two
// steps ahead of a peer (step e+1 needs that peer's words of step e+1, sent after it finished e), and a stale word in
// the same parity slot carries epoch e-2.  The mailbox is zeroed once (epoch 0 never matches).
// One CTA (the vector is small); larger models keep the NCCL all-reduce.
//
// Contract: every rank takes EXACTLY the same number of steps (the epoch is the step counter).  A rank that waits
// ~3 s for a peer gives up: the whole CTA then SKIPS the parameter / m / v update of that step (all-or-nothing, so a
// replica is never updated from a partial or stale sum) and sets the time-out marker behind the mailbox, which
// cgat.train_step.TrainStep polls and turns into a RuntimeError.
#include "common.cuh"

namespace cgat {

constexpr int P2P_THREADS = 1024;
constexpr int P2P_MAX_WORLD = 8;

long long* get_debug_buffer();  // developer timeline (cgat_debug_timeline): 4 globaltimer stamps per step when set

__device__ __forceinline__ long long p2p_now() {
  long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

struct P2pPeers {
  uint2* mailbox[P2P_MAX_WORLD];   // peer p's mailbox base
};

// {value bits, epoch} as ONE 64-bit scalar access: single-copy atomic at 8-byte alignment (a .v2.u32 access is two
// scalar accesses in unspecified order to the PTX memory model -- a reader could see the new epoch with old bits)
__device__ __forceinline__ void st_word_sys(uint2* p, uint32_t bits, uint32_t epoch) {
  const unsigned long long w = (unsigned long long)bits | ((unsigned long long)epoch << 32);
  asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(w) : "memory");
}
__device__ __forceinline__ uint2 ld_word_sys(const uint2* p) {
  unsigned long long w;
  asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(w) : "l"(p) : "memory");
  return make_uint2((uint32_t)w, (uint32_t)(w >> 32));
}

__global__ void __launch_bounds__(P2P_THREADS)
p2p_allreduce_adam_kernel(const P2pPeers P, int rank, int world, long long n, long long n_pad, const float* __restrict__ g,
                          float* __restrict__ p, float* __restrict__ m, float* __restrict__ v,
                          const long long* __restrict__ step_dev, long long step_host, float lr, float b1, float b2,
                          float eps, float wd, uint32_t* timeout_marker, long long* dbg, long long* step_counter,
                          const float* __restrict__ hyper) {
  // graph-resident form (cgat_p2p_allreduce_adam_graph): the step comes from a device counter of steps taken so far, which
  // this launch advances at its end, and the hyper-parameters from device memory (lr, beta1, beta2, eps, weight_decay:
  // a scheduler changes lr between replays); launched with programmatic dependent launch behind the gradient kernel
  griddep_wait();
  if (hyper != nullptr) { lr = hyper[0]; b1 = hyper[1]; b2 = hyper[2]; eps = hyper[3]; wd = hyper[4]; }
  const long long epoch = step_counter != nullptr ? *step_counter + 1 : (step_dev != nullptr ? *step_dev : step_host);
  const bool stamp = dbg != nullptr && threadIdx.x == 0 && epoch < 4096;
  if (stamp) dbg[epoch * 4 + 0] = p2p_now();
  const uint32_t ep = (uint32_t)epoch;
  const int par = (int)(epoch & 1);
  const int tid = threadIdx.x;
  // 1. push {value, epoch} words: 8-byte stores, coalesced per peer
  for (long long i = tid; i < n; i += P2P_THREADS) {
    const uint32_t bits = __float_as_uint(g[i]);
    for (int q = 0; q < world; ++q) st_word_sys(P.mailbox[q] + ((size_t)par * world + rank) * n_pad + i, bits, ep);
  }
  if (stamp) dbg[epoch * 4 + 1] = p2p_now();
  // 2. + 3. per element: wait for every source, rank-ordered sum, Adam
  const float step = (float)epoch;
  const float bc1 = 1.f - powf(b1, step);
  const float bc2 = 1.f - powf(b2, step);
  const float step_size = lr / bc1;
  const float inv_sqrt_bc2 = rsqrtf(bc2);
  const float gscale = 1.f / (float)world;
  const uint2* mine = P.mailbox[rank] + (size_t)par * world * n_pad;
  const long long t0 = clock64();
  // 2a. wait for every word of this thread's elements; the rank-ordered sums of its first P2P_CACHE elements stay in
  //     registers (every conv-GAT model: n <= 4096), later ones are re-read from the (local) mailbox in 2b
  constexpr int P2P_CACHE = 4;
  float cache[P2P_CACHE];
  int timed_out = 0;
  auto wait_sum = [&](long long i) {
    float gs = 0.f;
    for (int q = 0; q < world; ++q) {
      const uint2* src = mine + (size_t)q * n_pad + i;
      uint2 w = ld_word_sys(src);
      while (w.y != ep && !timed_out) {
        if (clock64() - t0 > 6000000000ll) timed_out = 1;  // ~3 s: a peer never arrived
        else w = ld_word_sys(src);
      }
      gs += __uint_as_float(w.x);
    }
    return gs;
  };
#pragma unroll
  for (int s = 0; s < P2P_CACHE; ++s) {
    const long long i = tid + (long long)s * P2P_THREADS;
    cache[s] = i < n ? wait_sum(i) : 0.f;
  }
  for (long long i = tid + (long long)P2P_CACHE * P2P_THREADS; i < n; i += P2P_THREADS) wait_sum(i);
  // all-or-nothing: one late word anywhere in the vector and NO element of p / m / v is touched this step
  if (__syncthreads_or(timed_out)) {
    if (tid == 0) atomicExch(timeout_marker, 1u);
    return;  // (the step counter stays: the host raises on the marker; nothing was updated)
  }
  if (step_counter != nullptr && tid == 0) *step_counter = epoch;  // every thread has read it (the barrier above)
  if (stamp) dbg[epoch * 4 + 2] = p2p_now();
  auto adam = [&](long long i, float gs) {
    const float pi = p[i];
    const float gi = fmaf(wd, pi, gs * gscale);
    const float mi = fmaf(b1, m[i], (1.f - b1) * gi);
    const float vi = fmaf(b2, v[i], (1.f - b2) * gi * gi);
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) * inv_sqrt_bc2 + eps;
    p[i] = pi - step_size * (mi / denom);
  };
#pragma unroll
  for (int s = 0; s < P2P_CACHE; ++s) {
    const long long i = tid + (long long)s * P2P_THREADS;
    if (i < n) adam(i, cache[s]);
  }
  for (long long i = tid + (long long)P2P_CACHE * P2P_THREADS; i < n; i += P2P_THREADS) {
    float gs = 0.f;
    for (int q = 0; q < world; ++q) gs += __uint_as_float(ld_word_sys(mine + (size_t)q * n_pad + i).x);
    adam(i, gs);
  }
  if (stamp) dbg[epoch * 4 + 3] = p2p_now();
}

}  // namespace cgat

using namespace cgat;

extern "C" int64_t cgat_p2p_mailbox_bytes(int64_t n, int32_t world) {
  if (n <= 0 || world < 1 || world > P2P_MAX_WORLD) return 0;
  const int64_t n_pad = (n + 31) & ~(int64_t)31;
  return 2 * (int64_t)world * n_pad * 8 + 256;  // {value, epoch} words, then one u32 time-out marker (all zeroed by the caller)
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
    P.mailbox[q] = reinterpret_cast<uint2*>(peer_mailboxes[q]);
  }
  uint32_t* marker = reinterpret_cast<uint32_t*>(peer_mailboxes[rank] + (uint64_t)2 * world * n_pad * 8);
  p2p_allreduce_adam_kernel<<<1, P2P_THREADS, 0, (cudaStream_t)stream>>>(P, rank, world, n, n_pad, grad, param, m, v,
                                                                         (const long long*)step_dev, (long long)step_host, lr,
                                                                         beta1, beta2, eps, weight_decay, marker, get_debug_buffer(),
                                                                         nullptr, nullptr);
  return check_launch("p2p_allreduce_adam_kernel");
}

extern "C" int cgat_p2p_allreduce_adam_graph(const uint64_t* peer_mailboxes, int32_t rank, int32_t world, const float* grad,
                                             float* param, float* m, float* v, int64_t* step_counter, const float* hyper,
                                             int64_t n, void* stream) {
  if (!peer_mailboxes || !grad || !param || !m || !v || !step_counter || !hyper) return fail(CGAT_EINVAL, "null argument");
  if (world < 1 || world > P2P_MAX_WORLD || rank < 0 || rank >= world) return fail(CGAT_EINVAL, "bad rank/world %d/%d", rank, world);
  if (n <= 0 || n > (1 << 20)) return fail(CGAT_EUNSUPPORTED, "p2p exchange serves vectors of at most 2^20 floats (n=%lld)", (long long)n);
  if (!aligned16(grad)) return fail(CGAT_EALIGN, "grad must be 16-byte aligned");
  const long long n_pad = (n + 31) & ~(long long)31;
  P2pPeers P{};
  for (int q = 0; q < world; ++q) {
    if (!peer_mailboxes[q] || (peer_mailboxes[q] & 15)) return fail(CGAT_EALIGN, "peer mailbox %d null or misaligned", q);
    P.mailbox[q] = reinterpret_cast<uint2*>(peer_mailboxes[q]);
  }
  uint32_t* marker = reinterpret_cast<uint32_t*>(peer_mailboxes[rank] + (uint64_t)2 * world * n_pad * 8);
  cudaError_t e = launch_pdl(p2p_allreduce_adam_kernel, dim3(1), dim3(P2P_THREADS), 0, (cudaStream_t)stream, P, (int)rank,
                             (int)world, (long long)n, n_pad, grad, param, m, v, (const long long*)nullptr, (long long)0, 0.f, 0.f,
                             0.f, 0.f, 0.f, marker, get_debug_buffer(), (long long*)step_counter, hyper);
  if (e != cudaSuccess) return fail((int)e, "p2p_allreduce_adam_kernel: %s", cudaGetErrorString(e));
  return check_launch("p2p_allreduce_adam_kernel");
}
