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
// One cluster of `world` CTAs (the vector is small); larger models keep the NCCL all-reduce.
//
// Contract: every rank takes EXACTLY the same number of steps (the epoch is the step counter).  A rank that waits
// ~3 s for a peer gives up: the whole cluster then SKIPS the parameter / m / v update of that step (all-or-nothing, so a
// replica is never updated from a partial or stale sum) and sets the time-out marker behind the mailbox, which
// cgat.train_step.TrainStep polls and turns into a RuntimeError.
#include "common.cuh"
#include <cstdlib>
#include <utility>

namespace cgat {

constexpr int P2P_THREADS = 256;  // per CTA; the kernel runs as one cluster of `world` CTAs
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
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// the int at the same shared-memory address in CTA `cta` of this cluster (distributed shared memory)
__device__ __forceinline__ int ld_dsmem_s32(const int* local, unsigned cta) {
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(local);
  uint32_t remote;
  int v;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(a), "r"(cta));
  asm volatile("ld.shared::cluster.s32 %0, [%1];" : "=r"(v) : "r"(remote) : "memory");
  return v;
}
__device__ __forceinline__ uint2 ld_word_sys(const uint2* p) {
  unsigned long long w;
  asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(w) : "l"(p) : "memory");
  return make_uint2((uint32_t)w, (uint32_t)(w >> 32));
}

// One CLUSTER of `world` CTAs (<= 8: the portable cluster size).  CTA c pushes the whole vector to peer c -- a single SM
// sustains ~50 GB/s of remote 8-byte stores, so one CTA pushing to 8 peers took 1.8 us and the peer served last saw its
// words that much later (8-GPU timeline: push 1.8, wait 5.2 us); eight SMs push in parallel -- and owns slice c of the
// elements for the wait + rank-ordered sum + Adam.  All-or-nothing across the cluster: every CTA publishes whether it
// timed out in its shared memory, one cluster barrier, every CTA reads all flags through distributed shared memory.
__global__ void __launch_bounds__(P2P_THREADS)
p2p_allreduce_adam_kernel(const P2pPeers P, int rank, int world, long long n, long long n_pad, const float* __restrict__ g,
                          float* __restrict__ p, float* __restrict__ m, float* __restrict__ v,
                          const long long* __restrict__ step_dev, long long step_host, float lr, float b1, float b2,
                          float eps, float wd, uint32_t* timeout_marker, long long* dbg, long long* step_counter,
                          const float* __restrict__ hyper) {
  // graph-resident form (cgat_p2p_allreduce_adam_graph): the step comes from a device counter of steps taken so far, which
  // this launch advances at its end, and the hyper-parameters from device memory (lr, beta1, beta2, eps, weight_decay:
  // a scheduler changes lr between replays); launched with programmatic dependent launch behind the gradient kernel
  __shared__ int s_timed_out;
  const int c = (int)blockIdx.x;  // = rank of this CTA in the cluster (grid = one cluster of `world` CTAs)
  griddep_wait();
  if (hyper != nullptr) { lr = hyper[0]; b1 = hyper[1]; b2 = hyper[2]; eps = hyper[3]; wd = hyper[4]; }
  const long long epoch = step_counter != nullptr ? *step_counter + 1 : (step_dev != nullptr ? *step_dev : step_host);
  const bool stamp = dbg != nullptr && c == 0 && threadIdx.x == 0 && epoch < 4096;
  if (stamp) dbg[epoch * 4 + 0] = p2p_now();
  const uint32_t ep = (uint32_t)epoch;
  const int par = (int)(epoch & 1);
  const int tid = threadIdx.x;
  // 1. push {value, epoch} words to peer c: 8-byte stores, coalesced
  {
    uint2* dst = P.mailbox[c] + ((size_t)par * world + rank) * n_pad;
    for (long long i = tid; i < n; i += P2P_THREADS) st_word_sys(dst + i, __float_as_uint(g[i]), ep);
  }
  if (stamp) dbg[epoch * 4 + 1] = p2p_now();
  // 2. + 3. per element of this CTA's slice: wait for every source, rank-ordered sum, Adam
  const float step = (float)epoch;
  const float bc1 = 1.f - powf(b1, step);
  const float bc2 = 1.f - powf(b2, step);
  const float step_size = lr / bc1;
  const float inv_sqrt_bc2 = rsqrtf(bc2);
  const float gscale = 1.f / (float)world;
  const long long chunk = ((n + world - 1) / world + 31) & ~31ll;
  const long long lo = min(n, (long long)c * chunk), hi = min(n, lo + chunk);
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
    const long long i = lo + tid + (long long)s * P2P_THREADS;
    cache[s] = i < hi ? wait_sum(i) : 0.f;
  }
  for (long long i = lo + tid + (long long)P2P_CACHE * P2P_THREADS; i < hi; i += P2P_THREADS) wait_sum(i);
  // all-or-nothing over the whole cluster: one late word anywhere in the vector and NO element of p / m / v is touched
  const int cta_timed_out = __syncthreads_or(timed_out);
  if (tid == 0) s_timed_out = cta_timed_out;
  cluster_sync_all();  // flags written (and: every CTA has read the step counter)
  int any = 0;
  for (int r = 0; r < world; ++r) any |= ld_dsmem_s32(&s_timed_out, (unsigned)r);
  if (any) {
    if (c == 0 && tid == 0) atomicExch(timeout_marker, 1u);
    cluster_sync_all();  // (nobody leaves while a peer CTA may still read its flag)
    return;  // (the step counter stays: the host raises on the marker; nothing was updated)
  }
  if (step_counter != nullptr && c == 0 && tid == 0) *step_counter = epoch;
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
    const long long i = lo + tid + (long long)s * P2P_THREADS;
    if (i < hi) adam(i, cache[s]);
  }
  for (long long i = lo + tid + (long long)P2P_CACHE * P2P_THREADS; i < hi; i += P2P_THREADS) {
    float gs = 0.f;
    for (int q = 0; q < world; ++q) gs += __uint_as_float(ld_word_sys(mine + (size_t)q * n_pad + i).x);
    adam(i, gs);
  }
  if (stamp) dbg[epoch * 4 + 3] = p2p_now();
  cluster_sync_all();  // (nobody leaves while a peer CTA may still read its flag)
}

// launch as ONE cluster of `world` CTAs, optionally with programmatic dependent launch
template <typename... Args>
static cudaError_t p2p_launch(int world, bool pdl, cudaStream_t st, Args&&... args) {
  static const bool no_pdl = std::getenv("CGAT_NO_PDL") != nullptr;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)world);
  cfg.blockDim = dim3(P2P_THREADS);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = st;
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = (unsigned)world;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = (pdl && !no_pdl) ? 2 : 1;
  return cudaLaunchKernelEx(&cfg, p2p_allreduce_adam_kernel, std::forward<Args>(args)...);
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
  cudaError_t e = p2p_launch((int)world, false, (cudaStream_t)stream, P, (int)rank, (int)world, (long long)n, n_pad, grad, param, m,
                             v, (const long long*)step_dev, (long long)step_host, lr, beta1, beta2, eps, weight_decay, marker,
                             get_debug_buffer(), (long long*)nullptr, (const float*)nullptr);
  if (e != cudaSuccess) return fail((int)e, "p2p_allreduce_adam_kernel: %s", cudaGetErrorString(e));
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
  cudaError_t e = p2p_launch((int)world, true, (cudaStream_t)stream, P, (int)rank, (int)world, (long long)n, n_pad, grad, param, m,
                             v, (const long long*)nullptr, (long long)0, 0.f, 0.f, 0.f, 0.f, 0.f, marker, get_debug_buffer(),
                             (long long*)step_counter, hyper);
  if (e != cudaSuccess) return fail((int)e, "p2p_allreduce_adam_kernel: %s", cudaGetErrorString(e));
  return check_launch("p2p_allreduce_adam_kernel");
}
