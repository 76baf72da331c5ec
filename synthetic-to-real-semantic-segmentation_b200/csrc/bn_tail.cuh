// Pending BatchNorm (s2r_bn_tail, include/s2r_b200.h): a producer kernel has left per-channel (sum, sum of squares) in
// an fp64 buffer; the kernel that CONSUMES the normalised tensor turns them into scale / shift in its own prologue
// (bn_fin4: a dozen instructions per channel, after griddepcontrol.wait -- the kernel boundary orders it after the
// producer's atomics) and its first CTA publishes mean / inv-std / scale / shift for the backward pass and updates the
// running statistics.  That removes the bn_finalize launch of every BatchNorm layer (120 per adaptation step, each on
// the critical path).  [A first version ran the finalize in the LAST CTA of the producer (ticket counter): every CTA
// then pays a __threadfence over its outstanding fp64 atomics plus an atomic round trip, and the step got 0.4 ms
// SLOWER than with the separate launch; profiles/r2_notes.md.]
// For a synchronised BatchNorm on several ranks the sums are first exchanged over NVLink peer memory (the one-shot "LL"
// exchange below).  That exchange WAITS for the peers, and a CTA that spins inside a persistent, machine-filling
// kernel can close a cycle across two streams and two ranks (rank 1: stream-A kernel needs the SM held by the
// spinning stream-B CTA, rank 2 the other way round), so with an exchange the tail runs as ONE small kernel of its own
// (bn_tail_kernel in comm.cu: exchange + finalize, one CTA that can always co-reside) instead of two.
// Reference: modeling/sync_batchnorm/batchnorm.py:55-78,90-125 (reduce to master, _compute_mean_std, broadcast) and
// the F.batch_norm fallback at :50-53.
#pragma once
#include "common.cuh"

#ifdef __CUDACC__

constexpr int CM_DEPTH = 4;
constexpr int CM_CHANNELS = 2;   // independent exchange sequences (one per stream of the two-stream training step)
constexpr int CM_MAX_WORLD = 16;

struct CommDev {
  double* inbox[CM_MAX_WORLD];                 // base of every rank's inbox region (peer-mapped)
  unsigned long long* flags[CM_MAX_WORLD];     // base of every rank's flag array [DEPTH][world]
  unsigned long long* seq;                     // local: exchange counters (one per channel)
  int* err;                                    // sticky error flag in MAPPED HOST memory (the host polls it without a sync)
  int rank, world, slot;                       // slot: doubles per (depth, rank) entry
};

// device copy of the exchange descriptor (comm.cu owns it); nullptr until s2r_comm_open
const CommDev* s2r_comm_dev_ptr();

// device-side form of s2r_bn_tail (passed to kernels by value)
struct BnTail {
  double inv_count;      // 1 / count
  double unbias;         // count / (count - 1): running_var takes the unbiased variance (batchnorm.py:121-123)
  const double* sums;    // [2][C]
  const float* gamma;
  const float* beta;
  float* running_mean;
  float* running_var;
  float* mean_invstd;
  float* scale_shift;
  const CommDev* comm;   // bn_tail_kernel only: exchange the sums on `channel` first
  float eps, momentum;
  int clamp_mode, channel;
  int enabled;
};

static inline BnTail bn_tail_from(const s2r_bn_tail* t) {
  BnTail b = {};
  if (!t) return b;
  b.inv_count = 1.0 / t->count;
  b.unbias = t->count > 1 ? t->count / (t->count - 1) : 1.0;
  b.sums = t->sums;
  b.gamma = t->gamma; b.beta = t->beta;
  b.running_mean = t->running_mean; b.running_var = t->running_var;
  b.mean_invstd = t->mean_invstd; b.scale_shift = t->scale_shift;
  b.eps = t->eps; b.momentum = t->momentum;
  b.clamp_mode = t->clamp_mode;
  b.channel = t->channel;
  b.comm = nullptr;
  b.enabled = 1;
  return b;
}

// a consumer kernel may derive scale / shift itself unless the sums still have to travel between ranks
// (the fused prologues use 16-byte vector loads of the sums and the affine parameters)
static inline bool bn_tail_fusable(const s2r_bn_tail* t) {
  return t && t->channel < 0 && ((uintptr_t)t->sums | (uintptr_t)t->gamma | (uintptr_t)t->beta) % 16 == 0;
}

// stand-alone tail: [exchange on t->channel] + finalize, one launch of one CTA (comm.cu)
int s2r_bn_tail_launch(const s2r_bn_tail* t, int C, cudaStream_t stream);

// Low-latency ("LL") wire format: every 8-byte word carries 32 payload bits and the 32-bit sequence number of
// the exchange, so data and flag arrive in ONE atomic store -- no fence, no separate flag write, no second
// NVLink round trip.  A double travels as two such words.  The reader polls each word until its sequence field
// matches; a slot still holding the words of exchange seq - DEPTH can never match.
__device__ __forceinline__ void cm_st_word(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long cm_ld_word(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ double cm_ld_f64(const double* p) {   // L2 read (the values were produced by atomics / other CTAs)
  double v;
  asm volatile("ld.relaxed.gpu.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}

// In-place sum over the ranks of buf[0..n) by ONE CTA (all of its threads call; any block size): contribute to every
// peer's inbox, poll the own inbox, sum in rank order (bitwise identical on all ranks).  Advances the channel's
// sequence counter.  Bounded wait (~20 s), then the sticky error flag.
__device__ __forceinline__ void cm_exchange_cta(const CommDev& c, int ch, double* buf, int n) {
  const int nthr = blockDim.x * blockDim.y * blockDim.z;
  const int tid = threadIdx.x + blockDim.x * (threadIdx.y + blockDim.y * threadIdx.z);
  unsigned long long* seqp = c.seq + ch;
  const unsigned long long seq = *reinterpret_cast<volatile unsigned long long*>(seqp) + 1;
  const unsigned long long tag = (seq & 0xffffffffull) << 32;
  const int d = (int)(seq % CM_DEPTH);
  const size_t slot_words = (size_t)c.slot * 2;
  const size_t ch_words = (size_t)ch * CM_DEPTH * c.world * slot_words;
  for (int i = tid; i < n; i += nthr) {
    const unsigned long long bits = (unsigned long long)__double_as_longlong(cm_ld_f64(buf + i));
    const unsigned long long w0 = tag | (bits & 0xffffffffull), w1 = tag | (bits >> 32);
    for (int r = 0; r < c.world; ++r) {
      if (r == c.rank) continue;
      unsigned long long* dst = reinterpret_cast<unsigned long long*>(c.inbox[r]) + ch_words + ((size_t)d * c.world + c.rank) * slot_words;
      cm_st_word(dst + 2 * i, w0);
      cm_st_word(dst + 2 * i + 1, w1);
    }
  }
  const unsigned long long* in = reinterpret_cast<const unsigned long long*>(c.inbox[c.rank]) + ch_words + (size_t)d * c.world * slot_words;
  bool timed_out = false;
  for (int i = tid; i < n; i += nthr) {
    const double mine = cm_ld_f64(buf + i);
    double acc = 0.0;
    // peers in batches of four: their polls are in flight together (the L2 / NVLink latencies overlap) without
    // holding 2 x world words in registers; the sum stays in rank order
    for (int r0 = 0; r0 < c.world; r0 += 4) {
      unsigned long long w0[4], w1[4];
      const long long t0 = clock64();
      bool all_ok;
      do {
        all_ok = true;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int r = r0 + k;
          w0[k] = w1[k] = tag;
          if (r < c.world && r != c.rank) {
            const unsigned long long* src = in + (size_t)r * slot_words + 2 * i;
            w0[k] = cm_ld_word(src);
            w1[k] = cm_ld_word(src + 1);
          }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if ((((w0[k] ^ tag) | (w1[k] ^ tag)) >> 32) != 0) all_ok = false;
        if (!all_ok && (timed_out || clock64() - t0 > 40000000000ll)) {   // ~20 s: a peer is gone; do not hang the GPU
          timed_out = true;
          break;
        }
      } while (!all_ok);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int r = r0 + k;
        if (r < c.world)
          acc += (r == c.rank) ? mine : __longlong_as_double((long long)((w0[k] & 0xffffffffull) | (w1[k] << 32)));
      }
    }
    buf[i] = acc;
  }
  if (timed_out) {
    *reinterpret_cast<volatile int*>(c.err) = 1;
    __threadfence_system();
  }
  __threadfence();
  __syncthreads();
  if (tid == 0) *reinterpret_cast<volatile unsigned long long*>(seqp) = seq;
}

// scale / shift (and, when publish, mean / inv-std / running statistics to global memory) of channel c from the sums.
// Every consumer CTA runs this for its channels, so it is kept to a dozen instructions per channel: mean and the
// variance numerator in fp64 (the difference of the two sums cancels), everything after that in fp32 with the
// hardware reciprocal square root (2 ulp; the same value is published for the backward pass, so forward and backward
// see one consistent inv-std).
// clamp_mode 0: invstd = (var + eps)^-1/2 (F.batch_norm, batchnorm.py:50-53); 1: max(var, eps)^-1/2 (:125)
__device__ __forceinline__ void bn_fin1(const BnTail& t, int C, int c, double s, double q, float g, float b, bool publish,
                                        float& sc, float& sh) {
  const double mean = s * t.inv_count;
  const double sumvar = q - s * mean;   // batchnorm.py:117-118
  const float var = fmaxf((float)(sumvar * t.inv_count), 0.f);
  const float invstd = rsqrtf(t.clamp_mode ? fmaxf(var, t.eps) : var + t.eps);
  const float meanf = (float)mean;
  sc = g * invstd;
  sh = fmaf(-meanf, sc, b);
  if (publish) {
    if (t.running_mean) {
      t.running_mean[c] = (1.f - t.momentum) * t.running_mean[c] + t.momentum * meanf;
      t.running_var[c] = (1.f - t.momentum) * t.running_var[c] + t.momentum * (var * (float)t.unbias);
    }
    t.mean_invstd[c] = meanf;
    t.mean_invstd[C + c] = invstd;
    t.scale_shift[c] = sc;
    t.scale_shift[C + c] = sh;
  }
}

// four consecutive channels c..c+3 (c % 4 == 0: 16-byte aligned vector loads of the sums and the parameters)
__device__ __forceinline__ void bn_fin4(const BnTail& t, int C, int c, bool publish, float* sc, float* sh) {
  const double2 s01 = *reinterpret_cast<const double2*>(t.sums + c), s23 = *reinterpret_cast<const double2*>(t.sums + c + 2);
  const double2 q01 = *reinterpret_cast<const double2*>(t.sums + C + c), q23 = *reinterpret_cast<const double2*>(t.sums + C + c + 2);
  const float4 g = t.gamma ? __ldg(reinterpret_cast<const float4*>(t.gamma + c)) : make_float4(1.f, 1.f, 1.f, 1.f);
  const float4 b = t.beta ? __ldg(reinterpret_cast<const float4*>(t.beta + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
  bn_fin1(t, C, c, s01.x, q01.x, g.x, b.x, publish, sc[0], sh[0]);
  bn_fin1(t, C, c + 1, s01.y, q01.y, g.y, b.y, publish, sc[1], sh[1]);
  bn_fin1(t, C, c + 2, s23.x, q23.x, g.z, b.z, publish, sc[2], sh[2]);
  bn_fin1(t, C, c + 3, s23.y, q23.y, g.w, b.w, publish, sc[3], sh[3]);
}

#endif  // __CUDACC__
