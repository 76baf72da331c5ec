// One-shot all-reduce of small fp64 vectors over NVLink peer memory, for the synchronised-BatchNorm
// statistics exchange: the per-channel (sum, sum of squares) of the forward pass and (sum dy, sum dy*xhat)
// of the backward pass (modeling/sync_batchnorm/batchnorm.py:55-78,90-111 of the reference: reduce to the
// master + broadcast through Python queues; here: every rank is its own master).
//
// A step has ~240 such exchanges (60 BatchNorm layers x 2 passes x forward/backward), each a few kilobytes and
// each on the critical path, so what matters is latency, not bandwidth.  One kernel, one CTA:
//   1. every rank stores its vector into slot [seq % DEPTH][rank] of every peer's inbox (plain stores to the
//      peer's mapped memory: NVLink writes through NVSwitch) in the LL wire format described below -- payload and
//      sequence number in the same 8-byte word, so there is no fence and no separate flag;
//   2. it polls the words of its own inbox until they carry seq and sums the contributions in rank order
//      (bitwise identical on every rank) into the caller's buffer.
// seq lives in device memory and is advanced by the kernel itself, so captured CUDA graphs replay correctly.
// A slot is reused only DEPTH exchanges later; a rank can be at most one exchange ahead of any peer (it cannot
// finish exchange k before every peer has contributed to k), so DEPTH = 4 is ample.  Waits are bounded: after
// ~20 s without progress the kernel records an error code (in mapped host memory: steps.check_peer_exchange polls it
// once per training step and raises) and returns instead of hanging the GPU.
// Co-residency: the two exchange channels are used by the two streams of the training step, and a rank may reach
// them in either order.  Each exchange kernel is ONE CTA that only waits for peers' stores, so two of them in
// flight need two free CTA slots out of 148 SMs x several -- but CUDA does not GUARANTEE that independent graph
// branches run concurrently: a rank that serialised channel 0 before channel 1 while a peer serialised them the
// other way round would cross-wait.  That case ends in the bounded wait + the error above (never a hang or silently
// partial sums); S2R_OVERLAP=0 runs the step on one stream and one channel.
//
// Set-up: each process allocates its inbox with cudaMalloc, exports a CUDA IPC handle, and opens the handles
// of its peers (exchanged by the host through torch.distributed -- plumbing, not the data path).
#include <string.h>

#include "bn_tail.cuh"

namespace {

constexpr int CM_THREADS = 1024;

struct CommHost {
  bool ready = false;
  int rank = 0, world = 1, slot = 0;
  void* local = nullptr;                       // cudaMalloc'ed region: inbox | flags | seq
  int* err_host = nullptr;                     // cudaHostAlloc'ed (mapped) sticky error flag
  void* peers[CM_MAX_WORLD] = {nullptr};
  size_t inbox_bytes = 0, flags_bytes = 0;
  CommDev dev;
  CommDev* dev_copy = nullptr;               // the descriptor in device memory (kernels that take a pointer)
};

CommHost g_comm;

// WORLD > 0: the rank count is a compile-time constant, so the polls of all peers are issued together (their L2
// latencies overlap) instead of one peer after the other; WORLD == 0: generic loop.
template <int WORLD>
__global__ void __launch_bounds__(CM_THREADS)
allreduce_small_kernel(double* __restrict__ buf, int n, CommDev c, int ch) {
  // channel ch has its own sequence counter and its own region of every inbox: two streams can run their exchange
  // sequences concurrently as long as every rank issues the same sequence PER CHANNEL
  // Programmatic dependent launch (common.cuh): this one-CTA kernel may become resident early, but it must NOT release
  // its own dependents before the exchange is over -- they would become resident on every SM (a bn_apply grid takes
  // all thread slots) and sit there while this kernel waits for the peers, starving the other stream's kernels that
  // the PEER's exchange on the other channel is waiting for: a cross-rank cycle (seen as the bounded-wait error on
  // 2 GPUs).  No launch_dependents here: the dependents start when this kernel completes.
  pdl_wait();
  unsigned long long* seqp = c.seq + ch;
  const unsigned long long seq = *seqp + 1;       // every thread reads it; thread 0 advances it at the end
  const unsigned long long tag = (seq & 0xffffffffull) << 32;
  const int d = (int)(seq % CM_DEPTH);
  const size_t slot_words = (size_t)c.slot * 2;
  const size_t ch_words = (size_t)ch * CM_DEPTH * c.world * slot_words;
  // 1. contribute to every peer's inbox
  for (int i = threadIdx.x; i < n; i += CM_THREADS) {
    const unsigned long long bits = (unsigned long long)__double_as_longlong(buf[i]);
    const unsigned long long w0 = tag | (bits & 0xffffffffull), w1 = tag | (bits >> 32);
    for (int r = 0; r < c.world; ++r) {
      if (r == c.rank) continue;
      unsigned long long* dst = reinterpret_cast<unsigned long long*>(c.inbox[r]) + ch_words + ((size_t)d * c.world + c.rank) * slot_words;
      cm_st_word(dst + 2 * i, w0);
      cm_st_word(dst + 2 * i + 1, w1);
    }
  }
  // 2. + 3. poll every rank's words and sum in rank order (own value taken from buf)
  const unsigned long long* in = reinterpret_cast<const unsigned long long*>(c.inbox[c.rank]) + ch_words + (size_t)d * c.world * slot_words;
  bool timed_out = false;
  for (int i = threadIdx.x; WORLD > 0 && i < n; i += CM_THREADS) {
    const double mine = buf[i];
    unsigned long long w0[WORLD > 0 ? WORLD : 1], w1[WORLD > 0 ? WORLD : 1];
    const long long t0 = clock64();
    bool all_ok;
    do {
      all_ok = true;
#pragma unroll
      for (int r = 0; r < WORLD; ++r) {
        const unsigned long long* src = in + (size_t)r * slot_words + 2 * i;
        w0[r] = cm_ld_word(src);
        w1[r] = cm_ld_word(src + 1);
      }
#pragma unroll
      for (int r = 0; r < WORLD; ++r)
        if (r != c.rank && ((((w0[r] ^ tag) | (w1[r] ^ tag)) >> 32) != 0)) all_ok = false;
      if (!all_ok && clock64() - t0 > 40000000000ll) {   // ~20 s: a peer is gone; do not hang the GPU
        timed_out = true;
        break;
      }
    } while (!all_ok);
    double acc = 0.0;
#pragma unroll
    for (int r = 0; r < WORLD; ++r)
      acc += (r == c.rank) ? mine : __longlong_as_double((long long)((w0[r] & 0xffffffffull) | (w1[r] << 32)));
    buf[i] = acc;
  }
  for (int i = threadIdx.x; WORLD == 0 && i < n; i += CM_THREADS) {
    double acc = 0.0;
    const double mine = buf[i];
    for (int r = 0; r < c.world; ++r) {
      if (r == c.rank) {
        acc += mine;
        continue;
      }
      const unsigned long long* src = in + (size_t)r * slot_words + 2 * i;
      unsigned long long w0 = cm_ld_word(src), w1 = cm_ld_word(src + 1);
      if (((w0 ^ tag) >> 32) != 0 || ((w1 ^ tag) >> 32) != 0) {
        const long long t0 = clock64();
        while (true) {
          w0 = cm_ld_word(src);
          w1 = cm_ld_word(src + 1);
          if (((w0 ^ tag) >> 32) == 0 && ((w1 ^ tag) >> 32) == 0) break;
          if (timed_out || clock64() - t0 > 40000000000ll) {   // ~20 s: a peer is gone; do not hang the GPU
            timed_out = true;
            break;
          }
        }
      }
      acc += __longlong_as_double((long long)((w0 & 0xffffffffull) | (w1 << 32)));
    }
    buf[i] = acc;
  }
  if (timed_out) {
    *reinterpret_cast<volatile int*>(c.err) = 1;
    __threadfence_system();
  }
  __syncthreads();
  if (threadIdx.x == 0) *seqp = seq;
}

// [exchange] + finalize of one BatchNorm layer in ONE launch of one CTA (bn_tail.cuh): what used to be the exchange
// kernel followed by bn_finalize.  Launched with programmatic dependent launch like every other kernel of the step.
__global__ void __launch_bounds__(CM_THREADS)
bn_tail_kernel(int C, BnTail t) {
  pdl_wait();
  if (!t.comm) pdl_trigger();   // with an exchange: dependents only after it (see allreduce_small_kernel)
  double* stats = const_cast<double*>(t.sums);
  if (t.comm) cm_exchange_cta(*t.comm, t.channel, stats, 2 * C);
  for (int c = threadIdx.x; c < C; c += CM_THREADS) {
    float sc, sh;
    bn_fin1(t, C, c, cm_ld_f64(stats + c), cm_ld_f64(stats + C + c), t.gamma ? t.gamma[c] : 1.f, t.beta ? t.beta[c] : 0.f,
            true, sc, sh);
  }
}

}  // namespace

const CommDev* s2r_comm_dev_ptr() { return g_comm.ready ? g_comm.dev_copy : nullptr; }

int s2r_bn_tail_launch(const s2r_bn_tail* t, int C, cudaStream_t stream) {
  S2R_REQUIRE(t && t->sums && C >= 1, S2R_ERR_SHAPE, "bn_tail: null statistics / tail or C=%d", C);
  // batchnorm.py:116 -- the statistics need more than one value per channel
  S2R_REQUIRE(t->count > 1, S2R_ERR_SHAPE, "BatchNorm computes unbiased standard-deviation, which requires size > 1.");
  S2R_REQUIRE(t->mean_invstd && t->scale_shift, S2R_ERR_SHAPE, "bn_tail: null outputs");
  BnTail b = bn_tail_from(t);
  if (t->channel >= 0) {
    const CommHost& h = g_comm;
    S2R_REQUIRE(h.ready && h.world > 1, S2R_ERR_SHAPE, "bn_tail: exchange channel %d without an initialised peer exchange", t->channel);
    S2R_REQUIRE(t->channel < CM_CHANNELS && 2 * C <= h.slot, S2R_ERR_SHAPE, "bn_tail: channel %d / %d doubles exceed the slot of %d", t->channel, 2 * C, h.slot);
    b.comm = h.dev_copy;
  }
  S2R_CUDA_OK(s2r_launch(bn_tail_kernel, dim3(1), dim3(CM_THREADS), 0, stream, C, b));
  S2R_LAUNCH_OK();
  return S2R_OK;
}

extern "C" int s2r_bn_tail_run(const s2r_bn_tail* tail, int C, s2r_stream_t stream) {
  return s2r_bn_tail_launch(tail, C, (cudaStream_t)stream);
}

/* Allocates this rank's inbox (slot_doubles per contribution) and writes its 64-byte CUDA IPC handle. */
extern "C" int s2r_comm_create(int rank, int world, int slot_doubles, void* handle_out) {
  S2R_REQUIRE(!g_comm.ready && g_comm.local == nullptr, S2R_ERR_SHAPE, "comm: already created");
  S2R_REQUIRE(world >= 1 && world <= CM_MAX_WORLD && rank >= 0 && rank < world && slot_doubles >= 2 && handle_out,
              S2R_ERR_SHAPE, "comm: bad rank/world/slot");
  CommHost& h = g_comm;
  h.rank = rank; h.world = world; h.slot = (slot_doubles + 1) & ~1;
  h.inbox_bytes = (size_t)CM_CHANNELS * CM_DEPTH * world * h.slot * 2 * sizeof(unsigned long long);   // LL words: 2 per double
  h.flags_bytes = (size_t)CM_DEPTH * world * sizeof(unsigned long long);
  const size_t total = h.inbox_bytes + h.flags_bytes + 256;
  S2R_CUDA_OK(cudaMalloc(&h.local, total));
  S2R_CUDA_OK(cudaMemset(h.local, 0, total));
  S2R_CUDA_OK(cudaHostAlloc((void**)&h.err_host, 64, cudaHostAllocMapped | cudaHostAllocPortable));
  *h.err_host = 0;
  S2R_CUDA_OK(cudaDeviceSynchronize());
  cudaIpcMemHandle_t ih;
  S2R_CUDA_OK(cudaIpcGetMemHandle(&ih, h.local));
  static_assert(sizeof(ih) == 64, "CUDA IPC handle size");
  memcpy(handle_out, &ih, sizeof(ih));
  return S2R_OK;
}

/* handles: world x 64 bytes, indexed by rank (this rank's own entry is ignored). */
extern "C" int s2r_comm_open(const void* handles) {
  CommHost& h = g_comm;
  S2R_REQUIRE(h.local != nullptr && !h.ready && handles, S2R_ERR_SHAPE, "comm: create first");
  for (int r = 0; r < h.world; ++r) {
    if (r == h.rank) {
      h.peers[r] = h.local;
    } else {
      cudaIpcMemHandle_t ih;
      memcpy(&ih, (const char*)handles + (size_t)r * 64, 64);
      S2R_CUDA_OK(cudaIpcOpenMemHandle(&h.peers[r], ih, cudaIpcMemLazyEnablePeerAccess));
    }
  }
  for (int r = 0; r < h.world; ++r) {
    h.dev.inbox[r] = (double*)h.peers[r];
    h.dev.flags[r] = (unsigned long long*)((char*)h.peers[r] + h.inbox_bytes);
  }
  h.dev.seq = (unsigned long long*)((char*)h.local + h.inbox_bytes + h.flags_bytes);
  S2R_CUDA_OK(cudaHostGetDevicePointer((void**)&h.dev.err, h.err_host, 0));
  h.dev.rank = h.rank; h.dev.world = h.world; h.dev.slot = h.slot;
  S2R_CUDA_OK(cudaMalloc((void**)&h.dev_copy, sizeof(CommDev)));
  S2R_CUDA_OK(cudaMemcpy(h.dev_copy, &h.dev, sizeof(CommDev), cudaMemcpyHostToDevice));
  h.ready = true;
  return S2R_OK;
}

extern "C" int s2r_comm_ready() { return g_comm.ready ? g_comm.world : 0; }

/* In-place sum over all ranks of buf[0..n) (fp64) on exchange channel `channel` (0 or 1).  Every rank must issue the
 * same sequence of calls per channel; calls on different channels may be in flight together (two streams). */
extern "C" int s2r_allreduce_small_f64_ch(double* buf, int n, int channel, s2r_stream_t stream) {
  const CommHost& h = g_comm;
  S2R_REQUIRE(h.ready, S2R_ERR_SHAPE, "comm: not initialised");
  S2R_REQUIRE(buf && n >= 1 && n <= h.slot, S2R_ERR_SHAPE, "allreduce_small: n=%d exceeds the slot of %d doubles", n, h.slot);
  S2R_REQUIRE(channel >= 0 && channel < CM_CHANNELS, S2R_ERR_SHAPE, "allreduce_small: channel %d", channel);
  if (h.world == 1) return S2R_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (h.world == 2) S2R_CUDA_OK(s2r_launch(allreduce_small_kernel<2>, dim3(1), dim3(CM_THREADS), 0, st, buf, n, h.dev, channel));
  else if (h.world == 4) S2R_CUDA_OK(s2r_launch(allreduce_small_kernel<4>, dim3(1), dim3(CM_THREADS), 0, st, buf, n, h.dev, channel));
  else if (h.world == 8) S2R_CUDA_OK(s2r_launch(allreduce_small_kernel<8>, dim3(1), dim3(CM_THREADS), 0, st, buf, n, h.dev, channel));
  else S2R_CUDA_OK(s2r_launch(allreduce_small_kernel<0>, dim3(1), dim3(CM_THREADS), 0, st, buf, n, h.dev, channel));
  S2R_LAUNCH_OK();
  return S2R_OK;
}

extern "C" int s2r_allreduce_small_f64(double* buf, int n, s2r_stream_t stream) {
  return s2r_allreduce_small_f64_ch(buf, n, 0, stream);
}

/* Non-zero after a bounded wait expired (a peer died, or two exchange kernels that needed to be co-resident were
 * serialised on some rank and cross-waited).  The flag is in mapped host memory: a plain read, no synchronisation --
 * cheap enough to poll once per training step. */
extern "C" int s2r_comm_error() {
  const CommHost& h = g_comm;
  if (!h.ready || h.err_host == nullptr) return 0;
  return *reinterpret_cast<volatile int*>(h.err_host);
}

extern "C" int s2r_comm_destroy() {
  CommHost& h = g_comm;
  if (h.local == nullptr) return S2R_OK;
  cudaDeviceSynchronize();
  for (int r = 0; r < h.world; ++r)
    if (r != h.rank && h.peers[r]) cudaIpcCloseMemHandle(h.peers[r]);
  cudaFree(h.local);
  if (h.dev_copy) cudaFree(h.dev_copy);
  if (h.err_host) cudaFreeHost(h.err_host);
  h = CommHost();
  return S2R_OK;
}
