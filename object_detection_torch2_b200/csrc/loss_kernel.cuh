// Fused MultiBox loss, forward + gradient in ONE launch.  Replaces SSD.loss and everything it calls
// (reference src/model/ssd.py:181-328: _match, _calc_delta, _smooth_l1, _softmax_cross_entropy,
// _split_pos_neg, _k_plus_1_th_value) plus the autograd backward of that graph (src/train.py:121).
//
// Decomposition (B200, sm_100a):
//   * one thread-block CLUSTER per image (4 CTAs x 768 threads, one CTA per SM; 8 x 384 for bigger slabs).
//     The image's contiguous [P, 4+C] slab is cut into 32-row blocks dealt round-robin to the CTAs, so every
//     CTA sees the same mix of prior levels (equal matching work) and keeps its rows in shared memory for
//     the whole kernel: HBM sees each output row exactly once as a read and (with grad) once as a write;
//   * a warp's 32 lanes own one block per row slot: its blocks arrive by TMA bulk copies (cp.async.bulk) on the
//     warp's own mbarriers, issued right after the (tiny, latency-critical) ground-truth rows landed; the IoU
//     matching runs under the bulk load; nothing in the row phases needs a block-wide barrier;
//   * one thread per row: log-sum-exp, positive / negative cross-entropy, smooth-L1 of matched pairs;
//   * hard-negative mining = value threshold at the (k+1)-th largest CE (strict '>', ssd.py:222-223):
//     every CTA PUSHES its 256-bucket histogram (1/16-octave buckets of the CE) into its peers' shared memory
//     (st.async + mbarrier complete_tx: data and signal travel together, no cluster barrier, no remote loads),
//     the selected bucket's candidates are pushed the same way and the exact order statistic is finished locally;
//     only ONE of the two thresholds ever needs a search (see the split logic below);
//   * gradient rows are written in place over the slab and leave by TMA bulk stores;
//   * the last image to finish reduces the per-image losses in a fixed order (deterministic).
#pragma once

#include <cooperative_groups.h>
#include <stdlib.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace ssdh {

constexpr int kSlots = 3;            // rows per thread
constexpr int kBlockRows = 32;       // rows per row slot of a warp
constexpr int kChunkRows = kSlots * kBlockRows;   // rows dealt to one warp: 96 contiguous rows = one TMA copy in, three out
constexpr int kMaxCluster = 8;
constexpr int kMaxLossWarps = 32;
constexpr int kListCap = 256;        // per-CTA candidates of the selected bucket (more: cluster-wide radix fallback)
// Kernel modes (template parameter kMode, a bit mask); the production kernel is mode 0 and carries none of the others' code.
constexpr int kModeTrace = 1;        // phase stamps (tools/trace_loss*.py)
constexpr int kModeForce = 2;        // best-prior-per-ground-truth forcing (north_star extension)
constexpr int kModeExact = 4;        // libdevice expf / logf / IEEE division instead of ex2.approx / lg2.approx / rcp.approx, and an
                                     // optional per-row cross-entropy override: the selection logic under test, fed exact inputs
// Two shapes of the same kernel (template parameters kT = threads per CTA, kCl = CTAs per cluster):
//   <768, 4>: 4 CTAs x ~221 KB per image, one CTA per SM  -> every SM carries the same load (default when it fits)
//   <384, 8>: 8 CTAs x ~110 KB per image                   -> larger images / more ground truth per image

// Workspace record of one image.  The layout is independent of N so that the ticket words of a cached workspace are
// always found zero again (they are reset by their last user), whatever batch size the previous call had.
struct ImageSlot {          // 112 bytes
  double part_loss[kMaxCluster];   // per-CTA partial sums
  double image_loss;               // inv_pos * sum, ssd.py:227 before .mean()
  int part_sel[kMaxCluster];       // selected positives | selected negatives << 16
  unsigned int ticket;             // CTAs of this image that have delivered their partial
  unsigned int pad;
};

struct LossParams {
  const float* outputs;
  const float* targets;
  const float4* priors;
  const float* next_outputs;   // optional: the next micro-batch (same shape), prefetched into L2 once this one is on chip
  const float* next_targets;
  int N, P, C, G;
  float a;
  ThrBand band;
  float inv_n_global;
  float* loss;
  float* grad;
  ssdh_image_stats* stats;
  unsigned int* ticket;   // workspace: zero before first use, left zero
  ImageSlot* slots;       // workspace [N]: per-CTA partial sums of one image + its arrival ticket (left zero)
  int rows_per_cta;       // shared-memory rows reserved per CTA (multiple of kChunkRows)
  int n_chunks;           // ceil(P / kChunkRows)
  int bulk;               // 1: every block is 16-byte aligned/sized -> TMA path
  int early_inputs;       // 1: outputs / targets / priors were complete before the PREVIOUS kernel of the stream started, so they
                          //    may be read under that kernel's tail (programmatic dependent launch); 0: wait for it first
  unsigned long long* trace;   // debug: [grid][kTracePoints] SM clock stamps (NULL in production)
  const float* ce_override;    // kModeExact only: [N, P] cross-entropies to select on (positive CE of matched rows, negative CE of the others)
  ssdh_scalar_exchange xchg;   // world == 0: off.  Otherwise the finalising CTA publishes the loss scalar to every rank's inbox
};

constexpr int kTracePoints = 128;
template <bool kTrace>
__device__ __forceinline__ void trace_point_t(const LossParams& p, int idx) {
  if (kTrace && p.trace != nullptr && threadIdx.x == 0) {
    unsigned long long t;
    if (idx == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    else t = clock64();
    p.trace[static_cast<size_t>(blockIdx.x) * kTracePoints + idx] = t;
  }
}

struct GtRec {             // 48 bytes, one per ground-truth row of the image; three 16-byte groups = three loads
  float x1, x2, y1, y2;    // corners                      (ssd.py:247-248)       -- matching
  float cx, cy, lw, lh;    // lw = log(w) (or w when w <= 0, ssd.py:269)           -- offsets of a matched pair
  float area;              //                                                     -- matching
  float tsum;              // sum of the class vector
  int label;               // class index when the class vector is exactly one-hot, else -1   -- matched pair (with flags)
  int flags;               // bit0: w > 0, bit1: h > 0
};

// What one CTA tells its cluster peers, twice per image.  Both records are whole 16-byte groups: they travel as st.async
// vector stores that complete a transaction count on the RECEIVER's mbarrier.
struct HistRecord {           // exchange #1: 1040 bytes
  uint32_t hist[2][128];      // [set: positives / negatives][256 CE buckets packed as 2 x u16]
  int pos_local;              // rows of this CTA with at least one match
  int pad[3];
};
struct ListRecord {           // exchange #2: 1040 bytes
  uint32_t key[kListCap];     // this CTA's candidates of the selected bucket (any order)
  int cnt;                    // how many it has (may exceed kListCap: then the cluster takes the radix fallback)
  int pad[3];
};
struct ForceRecord {          // forcing exchange (kModeForce only; lives in the bytes of the list records, which are idle then)
  uint32_t any[2];            // ground-truth rows with at least one match in this CTA
  uint32_t pad[2];
  unsigned long long best[kMaxGT];   // per ground-truth row: (IoU bits << 32) | ~row of this CTA's best prior
};
static_assert(sizeof(HistRecord) % 16 == 0 && sizeof(ListRecord) % 16 == 0 && sizeof(ForceRecord) % 16 == 0, "records travel as 16-byte groups");
static_assert(sizeof(ForceRecord) <= sizeof(ListRecord), "the forcing records alias the list records");
static_assert(sizeof(ListRecord) >= 2 * 128 * sizeof(uint32_t), "the cluster-wide totals borrow my own list record's keys");
static_assert(sizeof(HistRecord) >= kMaxLossWarps * (sizeof(double) + sizeof(int)), "the per-warp partial sums borrow a histogram record");

template <int kCluster>
struct LossSharedT {
  HistRecord mine;                    // built locally during the row phase, then pushed to every peer
  HistRecord from[kCluster - 1];      // written by the peers (slot = (sender - receiver - 1) mod kCluster); once the totals
                                      // are known, from[0].hist doubles as the second histogram buffer of the radix passes
  ListRecord lists[kCluster];         // [r]: candidates of rank r; my own slot is built in place, the others arrive by st.async;
                                      // afterwards the keys are compacted to the front of this array ("gathered")
  // (the cluster-wide totals, the counting / radix scratch of the selection and the per-warp partial sums have no storage of
  // their own: each lives in an exchange record that is idle while it is needed -- see tot_a / scratch / wred_* in the kernel)
  unsigned long long mbar[kMaxLossWarps];   // one per warp: its 96-row chunk lands on it
  unsigned long long xbar[3];         // exchange barriers: histograms, candidate lists, forcing
  int soft_labels;
  int pos_raw, k_pos, k_neg, sel_set, need_select, overflow;
  uint32_t sel_prefix, sel_rem;
  uint32_t force_any[2];              // kModeForce: ground-truth rows matched anywhere in the cluster
};

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// Warp-wide float minimum / maximum in one instruction (redux.sync on f32: sm_100a).
__device__ __forceinline__ float warp_min_f32(float v) {
  float r;
  asm volatile("redux.sync.min.f32 %0, %1, 0xffffffff;" : "=f"(r) : "f"(v));
  return r;
}
__device__ __forceinline__ float warp_max_f32(float v) {
  float r;
  asm volatile("redux.sync.max.f32 %0, %1, 0xffffffff;" : "=f"(r) : "f"(v));
  return r;
}

__device__ __forceinline__ void mbar_init(unsigned long long* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gmem_src, uint32_t bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// Same, with an L2 eviction policy: the head output is read exactly once, so its lines should be the first victims
// (keeps the freshly written gradient -- which the backbone's backward reads next -- resident instead).
__device__ __forceinline__ unsigned long long policy_evict_first() {
  unsigned long long pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void bulk_load_hint(void* smem_dst, const void* gmem_src, uint32_t bytes, unsigned long long* bar, unsigned long long pol) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)), "l"(pol)
               : "memory");
}
__device__ __forceinline__ void bulk_store(void* gmem_dst, const void* smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst), "r"(smem_u32(smem_src)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_store_wait() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
// ---- cluster exchange by PUSH ----------------------------------------------------------------------------------------
// A CTA hands data to a peer by writing it INTO THE PEER's shared memory with st.async: the store carries the address of an
// mbarrier in the same (remote) CTA and completes that many bytes of its transaction count when it lands (release at cluster
// scope); the receiver's mbarrier.try_wait (acquire at cluster scope) returns once every expected byte is there.  Data and
// signal travel together, which is exactly the ordering the PTX memory model gives a meaning to -- unlike a relaxed cluster
// barrier behind a CTA-scope fence (round 1), and without the MEMBAR.ALL.GPU that every cluster-scope release fence or
// barrier.cluster.arrive.release costs on this chip (~0.5 k cycles with TMA traffic in flight; SASS checked).  The only
// cluster barrier left is the one at kernel start that publishes the mbarrier initialisation (cutlass' cluster_arrive_relaxed
// + cluster_wait after fence.mbarrier_init); no CTA touches a peer's memory after that peer may have exited, because every
// CTA leaves only after it has RECEIVED everything the others send it.
__device__ __forceinline__ void cluster_arrive_relaxed() { asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait_acquire() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ uint32_t map_to_cta(const void* smem_ptr, uint32_t cta_rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_u32(smem_ptr)), "r"(cta_rank));
  return r;
}
__device__ __forceinline__ void st_async_v4(uint32_t remote_addr, const uint4 v, uint32_t remote_mbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(remote_addr),
               "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "r"(remote_mbar)
               : "memory");
}
// wait for data pushed by OTHER CTAs of the cluster: acquire at cluster scope
__device__ __forceinline__ void mbar_wait_cluster(unsigned long long* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// Push `words16` 16-byte groups starting at `src` (my shared memory) to the same-shaped record `dst_local` (an address in MY
// shared-memory window; the peer's copy of it is addressed through mapa) of every peer, signalling `bar_local` there.
template <int kCluster, int kThreads>
__device__ __forceinline__ void push_to_peers(const void* src, void* dst_local, int words16, unsigned long long* bar_local, int rank, int tid) {
  const uint4* s4 = reinterpret_cast<const uint4*>(src);
  for (int i = tid; i < words16 * (kCluster - 1); i += kThreads) {
    const int q = i / words16, idx = i - q * words16;
    int dst = rank + 1 + q;
    if (dst >= kCluster) dst -= kCluster;
    st_async_v4(map_to_cta(reinterpret_cast<uint4*>(dst_local) + idx, dst), s4[idx], map_to_cta(bar_local, dst));
  }
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// Packed fp32 pairs (FFMA2 / FADD2 on sm_100a): one issue slot for two lanes of the class loop.
__device__ __forceinline__ unsigned long long pack2(float lo, float hi) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(unsigned long long v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ unsigned long long ffma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ unsigned long long fadd2(unsigned long long a, unsigned long long b) {
  unsigned long long d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ unsigned long long fmul2(unsigned long long a, unsigned long long b) {
  unsigned long long d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}

__device__ __forceinline__ float smooth_l1_f(float x) {
  const float ax = fabsf(x);
  return ax < 1.0f ? 0.5f * x * x : ax - 0.5f;
}
__device__ __forceinline__ float clamp1(float x) { return fminf(fmaxf(x, -1.0f), 1.0f); }

// In place: x[c] <- scale * exp(x[c] - lse), the softmax gradient term of one row of logits in shared memory.
template <int kC, bool kExact>
__device__ __forceinline__ void softmax_scaled(float* x, int C, float lse, float scale) {
  constexpr float kLog2e = 1.4426950408889634f;
  const float nls = -(lse * kLog2e);
  if (kExact) {
    for (int c = 0; c < C; ++c) x[c] = scale * expf(x[c] - lse);
  } else if (kC > 0) {
    const unsigned long long l2 = pack2(kLog2e, kLog2e), m2 = pack2(nls, nls), sc2 = pack2(scale, scale);
#pragma unroll
    for (int c = 0; c + 1 < kC; c += 2) {
      float a0, a1;
      unpack2(ffma2(pack2(x[c], x[c + 1]), l2, m2), a0, a1);
      unpack2(fmul2(pack2(ex2_approx(a0), ex2_approx(a1)), sc2), a0, a1);
      x[c] = a0;
      x[c + 1] = a1;
    }
    if (kC & 1) x[kC - 1] = scale * ex2_approx(fmaf(x[kC - 1], kLog2e, nls));
  } else {
    for (int c = 0; c < C; ++c) x[c] = scale * ex2_approx(fmaf(x[c], kLog2e, nls));
  }
}

// Monotone (non-decreasing in the order key) 8-bit bucket: 16 buckets per octave over [2^-12, 2^4), everything
// below / above clamps into the first / last bucket.  Cross-entropies of interest live well inside the window.
__device__ __forceinline__ uint32_t bucket_of(uint32_t key) {
  const int v = static_cast<int>(key >> 19) - (4096 + ((127 - 12) << 4));
  return static_cast<uint32_t>(min(max(v, 0), 255));
}

// One warp: find the bin holding the (rem+1)-th largest element of a 256-bin packed histogram.
// Returns bin in [0,255]; rem is updated to the rank inside that bin.  All lanes get the result.
__device__ __forceinline__ int find_bin_desc(const uint32_t* packed, uint32_t& rem, int lane) {
  // lane l owns bins [248 - 8l, 255 - 8l], i.e. packed words [124 - 4l, 127 - 4l]
  const uint4 w = *reinterpret_cast<const uint4*>(packed + 124 - 4 * lane);
  uint32_t c[8];   // c[0] = highest bin of the lane
  c[0] = w.w >> 16; c[1] = w.w & 0xffffu; c[2] = w.z >> 16; c[3] = w.z & 0xffffu;
  c[4] = w.y >> 16; c[5] = w.y & 0xffffu; c[6] = w.x >> 16; c[7] = w.x & 0xffffu;
  int s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += static_cast<int>(c[i]);
  const int incl = warp_incl_scan(s, lane);
  const uint32_t ballot = __ballot_sync(0xffffffffu, static_cast<uint32_t>(incl) > rem);
  const int owner = ballot ? (__ffs(ballot) - 1) : 31;
  // inside the owner's 8 bins (descending): first i with rem' < c[0] + ... + c[i], branch-free
  const uint32_t r0 = rem - static_cast<uint32_t>(incl - s);
  uint32_t cum = 0, below = 0;
  int idx = 0;
#pragma unroll
  for (int i = 0; i < 7; ++i) {
    cum += c[i];
    const bool past = r0 >= cum;
    idx += past ? 1 : 0;
    below = past ? cum : below;
  }
  int bin = 255 - 8 * lane - idx;
  uint32_t r = r0 - below;
  bin = __shfl_sync(0xffffffffu, bin, owner);
  rem = __shfl_sync(0xffffffffu, r, owner);
  return bin;
}

// Warp-aggregated histogram add: lanes with the same (set, bin) elect one lane to add their count.
__device__ __forceinline__ void hist_add(uint32_t* hist_set0, bool active, int set, uint32_t bin, int lane) {
  const uint32_t tag = active ? (static_cast<uint32_t>(set) << 8 | bin) : 0xffffffffu;
  const uint32_t peers = __match_any_sync(0xffffffffu, tag);
  if (active && lane == __ffs(peers) - 1)
    atomicAdd(hist_set0 + set * 128 + (bin >> 1), static_cast<uint32_t>(__popc(peers)) << (16 * (bin & 1u)));
}

template <int kC, int kLossThreads, int kCluster, int kMode>
__global__ void __launch_bounds__(kLossThreads, 1) multibox_loss_kernel(const LossParams p) {
  constexpr bool kTrace = (kMode & kModeTrace) != 0, kForce = (kMode & kModeForce) != 0, kExact = (kMode & kModeExact) != 0;
  using LossShared = LossSharedT<kCluster>;
  constexpr int kLossWarps = kLossThreads / 32;
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = static_cast<int>(cluster.block_rank());
  const int n = blockIdx.x / kCluster;
  const int C = kC ? kC : p.C;
  const int row = 4 + C;
  const int G = p.G;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t lt_mask = (1u << lane) - 1u;
  // The image is cut into 96-row chunks dealt round-robin to the CTAs of the cluster (every CTA sees the same mix of
  // prior levels); inside a CTA, warp w owns local chunk w: global chunk w * kCluster + rank, rows [32 s, 32 s + 32) of
  // it are the warp's row slot s.  One TMA copy brings the chunk in, one per slot takes the gradient out.
  const int my_chunks = (p.n_chunks - rank + kCluster - 1) / kCluster;
  const int my_chunk = warp * kCluster + rank;                                   // global chunk of this warp
  const int my_rows_w = warp < my_chunks ? min(kChunkRows, p.P - my_chunk * kChunkRows) : 0;   // rows this warp owns
  constexpr float kLog2e = 1.4426950408889634f;

  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* slab = reinterpret_cast<float*>(smem_raw);
  const size_t slab_bytes = (static_cast<size_t>(p.rows_per_cta) * row * sizeof(float) + 15) & ~static_cast<size_t>(15);
  GtRec* gts = reinterpret_cast<GtRec*>(smem_raw + slab_bytes);
  LossShared& sh = *reinterpret_cast<LossShared*>(smem_raw + slab_bytes + ((static_cast<size_t>(G) * sizeof(GtRec) + 15) & ~static_cast<size_t>(15)));

  // Storage borrowed from exchange records while they are idle (their owners' phases never overlap with these uses):
  //   tot_a    cluster-wide bucket totals [2][128], between exchange #1 and the list building, in MY OWN list record -- nobody
  //            else ever writes lists[rank], and its counter word sits behind the 256 key words
  //   scratch  256 words of counting / radix scratch for the selection and the fallback's totals, in from[1] (its peer's
  //            histograms are consumed once the totals exist; from[0] is the second histogram buffer)
  //   wred_*   per-warp partial sums of the masked-sum phase (after the selection), in from[2]
  static_assert(kCluster >= 4, "the borrowed records from[1] and from[2] exist");
  uint32_t* const tot_a = &sh.lists[rank].key[0];
  uint32_t* const scratch = &sh.from[1].hist[0][0];
  double* const wred_loss = reinterpret_cast<double*>(&sh.from[2]);
  int* const wred_a = reinterpret_cast<int*>(wred_loss + kMaxLossWarps);
  const float* img_in = p.outputs + static_cast<size_t>(n) * p.P * row;

  // Programmatic dependent launch: let the NEXT grid in the stream be scheduled as soon as SMs free up.  It may run
  // everything that only reads its own inputs (load, match, CE, selection) under this grid's tail; it blocks at
  // griddepcontrol.wait below, before its first global write (workspace, loss, stats, gradient).
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  // The kernel in front of us in the stream is normally the PRODUCER of outputs / targets (the head's last kernel, a copy,
  // ssdh_expand_targets ...): its writes are only guaranteed visible after griddepcontrol.wait, so unless the caller vouches
  // for the inputs (ssdh_multibox_loss_pipelined) nothing is read from global memory before this point.
  if (!p.early_inputs) asm volatile("griddepcontrol.wait;" ::: "memory");

  // ---- setup ------------------------------------------------------------------------------------
  trace_point_t<kTrace>(p, 0);
  trace_point_t<kTrace>(p, 1);
  float4 pri[kSlots];
  bool valid[kSlots];
#pragma unroll
  for (int s = 0; s < kSlots; ++s) {
    const int grow = my_chunk * kChunkRows + s * kBlockRows + lane;
    valid[s] = s * kBlockRows + lane < my_rows_w;
    pri[s] = __ldg(p.priors + (valid[s] ? grow : 0));
  }
  // The (tiny) ground-truth rows are requested right away: their round trip is the longest chain of this prologue.
  float gt_first = 0.0f;
  if (warp < G) gt_first = __ldg(p.targets + (static_cast<size_t>(n) * G + warp) * row + min(lane, row - 1));
  if (tid == 0) {
    sh.mine.pos_local = 0;
    sh.lists[rank].cnt = 0;
    sh.soft_labels = 0;
    // exchange barriers: one arrival (mine, with the byte count I expect from the peers) + the peers' transaction bytes
    mbar_init(&sh.xbar[0], 1);
    mbar_init(&sh.xbar[1], 1);
    if (kForce) mbar_init(&sh.xbar[2], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    mbar_expect_tx(&sh.xbar[0], static_cast<uint32_t>(sizeof(HistRecord)) * (kCluster - 1));
    mbar_expect_tx(&sh.xbar[1], static_cast<uint32_t>(sizeof(ListRecord)) * (kCluster - 1));
    if (kForce) mbar_expect_tx(&sh.xbar[2], static_cast<uint32_t>(sizeof(ForceRecord)) * (kCluster - 1));
  }
  for (int i = tid; i < 2 * 128; i += kLossThreads) (&sh.mine.hist[0][0])[i] = 0u;
  trace_point_t<kTrace>(p, 48);
  __syncthreads();
  // publish the barrier initialisation to the cluster; the matching wait sits behind the slab / ground-truth requests, where
  // the warps would be waiting for memory anyway
  cluster_arrive_relaxed();

  // ---- slab: every warp fetches its own 96-row chunk with ONE TMA bulk copy onto its own mbarrier; issued right
  // behind the (tiny) ground-truth request, so that it lands while the ground truth is being unpacked -----------------
  float* my_slab = slab + static_cast<size_t>(warp) * kChunkRows * row;
  if (p.bulk) {
    if (lane == 0) {
      mbar_init(&sh.mbar[warp], 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      if (my_rows_w > 0) {
        const uint32_t bytes = static_cast<uint32_t>(my_rows_w) * row * sizeof(float);
        mbar_expect_tx(&sh.mbar[warp], bytes);
        bulk_load_hint(my_slab, img_in + static_cast<size_t>(my_chunk) * kChunkRows * row, bytes, &sh.mbar[warp], policy_evict_first());
      }
    }
    __syncwarp();
  } else {
    const float* src = img_in + static_cast<size_t>(my_chunk) * kChunkRows * row;
#pragma unroll 1
    for (int i = lane; i < my_rows_w * row; i += 32) my_slab[i] = src[i];
    __syncwarp();
  }
  trace_point_t<kTrace>(p, 49);
  cluster_wait_acquire();      // every CTA of the cluster has initialised its exchange barriers (they all started together)
  trace_point_t<kTrace>(p, 50);

  // ---- ground truth of this image -> shared: one warp per row, one coalesced request each -----------------
  for (int g = warp; g < G; g += kLossWarps) {
    const float* tr = p.targets + (static_cast<size_t>(n) * G + g) * row;
    float box = 0.0f, csum = 0.0f;
    int nz = 0, ones = 0, label = -1;
    for (int base = 0; base < row; base += 32) {
      const int c = base + lane;
      const float v = (base == 0 && g == warp) ? (c < row ? gt_first : 0.0f) : (c < row ? __ldg(tr + c) : 0.0f);
      if (base == 0) box = v;
      const bool cls = c >= 4 && c < row;
      const uint32_t b_nz = __ballot_sync(0xffffffffu, cls && v != 0.0f);
      const uint32_t b_one = __ballot_sync(0xffffffffu, cls && v == 1.0f);
      nz += __popc(b_nz);
      ones += __popc(b_one);
      if (b_one) label = base + __ffs(b_one) - 1 - 4;
      csum += cls ? v : 0.0f;
    }
    const bool one_hot = nz == 1 && ones == 1;
    const float tsum = one_hot ? 1.0f : warp_sum(csum);             // a single 1 sums to exactly 1: no reduction needed
    // lanes 2 / 3 hold w / h: both logarithms are taken side by side before lane 0 collects the record
    const float lg = ((lane == 2 || lane == 3) && box > 0.0f) ? logf(box) : box;
    const float gcx = __shfl_sync(0xffffffffu, box, 0), gcy = __shfl_sync(0xffffffffu, box, 1);
    const float gw = __shfl_sync(0xffffffffu, box, 2), gh = __shfl_sync(0xffffffffu, box, 3);
    const float lw = __shfl_sync(0xffffffffu, lg, 2), lh = __shfl_sync(0xffffffffu, lg, 3);
    if (lane == 0) {
      const Corners c = make_corners(gcx, gcy, gw, gh);
      GtRec r;
      r.x1 = c.x1; r.x2 = c.x2; r.y1 = c.y1; r.y2 = c.y2;
      r.area = c.area; r.cx = gcx; r.cy = gcy;
      r.lw = lw;                                                     // log(w), or w itself when w <= 0 (ssd.py:269)
      r.lh = lh;
      r.flags = (gw > 0.0f ? 1 : 0) | (gh > 0.0f ? 2 : 0);
      r.label = one_hot ? label : -1;
      r.tsum = tsum;
      gts[g] = r;
      if (r.label < 0 && c.area > 0.0f) sh.soft_labels = 1;
    }
  }
  trace_point_t<kTrace>(p, 51);
  // ---- corners of my priors and the outer bounds of my whole chunk (box around its priors, smallest / largest prior
  // area), used below to drop ground-truth rows that cannot match ANY of them ------------------------------------
  Corners d[kSlots];
  float cb_x1 = 3e38f, cb_x2 = -3e38f, cb_y1 = 3e38f, cb_y2 = -3e38f, cb_amin = 3e38f, cb_amax = -3e38f;
  bool tame = true;                        // every prior of the chunk has finite positive width and height
#pragma unroll
  for (int s = 0; s < kSlots; ++s) {
    d[s] = make_corners(pri[s].x, pri[s].y, pri[s].z, pri[s].w);
    if (valid[s]) {
      cb_x1 = fminf(cb_x1, d[s].x1); cb_x2 = fmaxf(cb_x2, d[s].x2);
      cb_y1 = fminf(cb_y1, d[s].y1); cb_y2 = fmaxf(cb_y2, d[s].y2);
      cb_amin = fminf(cb_amin, d[s].area); cb_amax = fmaxf(cb_amax, d[s].area);
      tame = tame && pri[s].z > 0.0f && pri[s].w > 0.0f && pri[s].z < 1e18f && pri[s].w < 1e18f && fabsf(pri[s].x) < 1e18f && fabsf(pri[s].y) < 1e18f;
    }
  }
  // warp-wide extremes: one CREDUX each on sm_100a instead of five shuffle rounds
  cb_x1 = warp_min_f32(cb_x1); cb_x2 = warp_max_f32(cb_x2);
  cb_y1 = warp_min_f32(cb_y1); cb_y2 = warp_max_f32(cb_y2);
  cb_amin = warp_min_f32(cb_amin); cb_amax = warp_max_f32(cb_amax);
  tame = __all_sync(0xffffffffu, tame);

  __syncthreads();

  trace_point_t<kTrace>(p, 2);
  // ---- matching: bit g of (mhi:mlo)[s] = IoU(gt g, prior) > thr   (ssd.py:231-250) ----------------------------
  // Dense pass over the fast rows, branch-free and division-free.  e = inter - union * thr is one FMA, so its sign
  // is exact; |e| > union * thr * 2^-20 then decides fl(inter / union) > thr with certainty either way.  The
  // (one in millions) pairs inside that band flag the thread through `amb`, and a flagged thread redoes its pairs
  // with the IEEE division, so the final mask is bit-identical to torch's.  The min / max / compare / logic ops
  // issue at half rate on this SM, hence the care to keep them at 9 per pair.
  uint32_t mlo[kSlots], mhi[kSlots];
#pragma unroll
  for (int s = 0; s < kSlots; ++s) { mlo[s] = 0u; mhi[s] = 0u; }
  {
    const ThrBand band = p.band;
    const float nthr = -band.thr, eps = band.thr * 9.5367431640625e-07f;
    // Which path a ground-truth row takes is decided per warp from the records themselves (lane g looks at rows g and
    // g + 32).  Fast: a normal positive area lets the band test decide IoU > thr without dividing.  Slow: everything else
    // that could still match (degenerate or denormal boxes, exotic thresholds); zero-padding rows take no path at all.
    // Cull (fast rows only): IoU > thr needs inter * (1 + thr) > thr * (area_g + area_p), and inter can exceed neither the
    // overlap of the row with the chunk's outer box nor either area; a row that fails this bound (with a 1e-4 safety
    // margin, far above any rounding) for the chunk's extreme areas matches none of my priors.
    uint32_t fast_lo = 0u, fast_hi = 0u, slow_lo = 0u, slow_hi = 0u, keep_lo = 0u, keep_hi = 0u;
    {
      const float thr = band.thr;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        if (half == 1 && G <= 32) break;                 // uniform
        const int g = 32 * half + lane;
        const bool have = g < G;
        const float4 q = *reinterpret_cast<const float4*>(&gts[have ? g : 0].x1);
        const float garea = gts[have ? g : 0].area;
        const bool fast = have && band.usable && garea >= 1e-30f && garea <= 1e30f;
        const bool slow = have && !fast && (garea > 0.0f || garea > band.thr || !(garea == garea));
        const float w = fmaxf(fminf(q.y, cb_x2) - fmaxf(q.x, cb_x1), 0.0f);
        const float h = fmaxf(fminf(q.w, cb_y2) - fmaxf(q.z, cb_y1), 0.0f);
        const float imax = fminf(fminf(w * h, cb_amax), garea);
        const bool hopeless = imax * (1.0f + thr) <= thr * (garea + cb_amin) * 0.9999f;
        const uint32_t f = __ballot_sync(0xffffffffu, fast), sl = __ballot_sync(0xffffffffu, slow);
        const uint32_t k = __ballot_sync(0xffffffffu, fast && !(tame && hopeless));
        if (half == 0) { fast_lo = f; slow_lo = sl; keep_lo = k; } else { fast_hi = f; slow_hi = sl; keep_hi = k; }
      }
    }
    float amb = 1.0f;                              // min over pairs of |e| - margin; <= 0 means "settle exactly"
    const int n_keep = __popc(keep_lo) + __popc(keep_hi);
#pragma unroll 2
    for (int it = 0; it < n_keep; ++it) {
      int g;
      if (keep_lo) { g = __ffs(keep_lo) - 1; keep_lo &= keep_lo - 1; }
      else { g = 32 + __ffs(keep_hi) - 1; keep_hi &= keep_hi - 1; }
      const float4 q = *reinterpret_cast<const float4*>(&gts[g].x1);
      const float garea = gts[g].area;
      const uint32_t bit = 1u << (g & 31);
      if (g < 32) {
#pragma unroll
        for (int s = 0; s < kSlots; ++s) {
          const float w = fmaxf(fminf(q.y, d[s].x2) - fmaxf(q.x, d[s].x1), 0.0f);
          const float h = fmaxf(fminf(q.w, d[s].y2) - fmaxf(q.z, d[s].y1), 0.0f);
          const float inter = w * h;
          const float uni = (garea + d[s].area) - inter;
          const float e = fmaf(uni, nthr, inter), m = uni * eps;
          amb = fminf(amb, fabsf(e) - m);
          if (e > m) mlo[s] |= bit;
        }
      } else {
#pragma unroll
        for (int s = 0; s < kSlots; ++s) {
          const float w = fmaxf(fminf(q.y, d[s].x2) - fmaxf(q.x, d[s].x1), 0.0f);
          const float h = fmaxf(fminf(q.w, d[s].y2) - fmaxf(q.z, d[s].y1), 0.0f);
          const float inter = w * h;
          const float uni = (garea + d[s].area) - inter;
          const float e = fmaf(uni, nthr, inter), m = uni * eps;
          amb = fminf(amb, fabsf(e) - m);
          if (e > m) mhi[s] |= bit;
        }
      }
    }
    // the band argument needs a positive union: chunks holding a prior of non-positive or non-finite extent settle exactly
    const bool unsure = !(amb > 0.0f) || !tame;
    if (__any_sync(0xffffffffu, unsure) || (slow_lo | slow_hi) != 0u) {
      // exact path: borderline pairs (the thread redoes every fast row), exotic rows / thresholds
      uint32_t ex_lo = slow_lo | (unsure ? fast_lo : 0u), ex_hi = slow_hi | (unsure ? fast_hi : 0u);
      while (ex_lo | ex_hi) {
        int g;
        if (ex_lo) { g = __ffs(ex_lo) - 1; ex_lo &= ex_lo - 1; }
        else { g = 32 + __ffs(ex_hi) - 1; ex_hi &= ex_hi - 1; }
        const float4 q = *reinterpret_cast<const float4*>(&gts[g].x1);
        const float garea = gts[g].area;
        const uint32_t bit = 1u << (g & 31);
#pragma unroll
        for (int s = 0; s < kSlots; ++s) {
          const float w = fmaxf(fminf(q.y, d[s].x2) - fmaxf(q.x, d[s].x1), 0.0f);
          const float h = fmaxf(fminf(q.w, d[s].y2) - fmaxf(q.z, d[s].y1), 0.0f);
          const float inter = w * h;
          const float val = garea > 0.0f ? __fdiv_rn(inter, (garea + d[s].area) - inter) : garea;     // ssd.py:250
          const bool hit = val > band.thr;
          if (g < 32) mlo[s] = hit ? (mlo[s] | bit) : (mlo[s] & ~bit);
          else mhi[s] = hit ? (mhi[s] | bit) : (mhi[s] & ~bit);
        }
      }
    }
  }
  // ---- best-prior-per-ground-truth forcing (north_star extension, off in the reference) -------------------
  // A real ground-truth box that matched some prior already owns its best prior (the arg-max IoU is above the threshold),
  // so forcing only ever adds a bit for boxes WITHOUT any match in the whole image.  Every CTA publishes which rows it
  // matched and, for the rows it did not, its own best prior as a (IoU bits, ~row) key (exact IEEE IoU as in ssd.py:250,
  // maximum = highest IoU, then lowest prior index: torch.argmax's first maximum); the owner of the cluster-wide winner
  // sets the bit.
  if (kForce) {
    ForceRecord* frec = reinterpret_cast<ForceRecord*>(&sh.lists[0]);        // [kCluster] records in the (idle) list bytes
    ForceRecord& fmine = frec[rank];
    if (tid < 2) fmine.any[tid] = 0u;
    for (int g = tid; g < kMaxGT; g += kLossThreads) fmine.best[g] = 0ull;
    __syncthreads();
    uint32_t alo = 0u, ahi = 0u;
#pragma unroll
    for (int s = 0; s < kSlots; ++s) { alo |= valid[s] ? mlo[s] : 0u; ahi |= valid[s] ? mhi[s] : 0u; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { alo |= __shfl_xor_sync(0xffffffffu, alo, o); ahi |= __shfl_xor_sync(0xffffffffu, ahi, o); }
    if (lane == 0) { if (alo) atomicOr(&fmine.any[0], alo); if (ahi) atomicOr(&fmine.any[1], ahi); }
    __syncthreads();
    const uint32_t l_lo = fmine.any[0], l_hi = fmine.any[1];
    for (int g = 0; g < G; ++g) {                                   // uniform for the CTA
      const bool matched = ((g < 32 ? l_lo >> g : l_hi >> (g - 32)) & 1u) != 0u;
      const float garea = gts[g].area;
      if (matched || !(garea > 0.0f)) continue;
      const float4 q = *reinterpret_cast<const float4*>(&gts[g].x1);
      unsigned long long best = 0ull;
#pragma unroll
      for (int s = 0; s < kSlots; ++s) {
        if (!valid[s]) continue;
        const float w = fmaxf(fminf(q.y, d[s].x2) - fmaxf(q.x, d[s].x1), 0.0f);
        const float h = fmaxf(fminf(q.w, d[s].y2) - fmaxf(q.z, d[s].y1), 0.0f);
        const float inter = w * h;
        const float v = __fdiv_rn(inter, (garea + d[s].area) - inter);               // ssd.py:250
        const uint32_t grow = static_cast<uint32_t>(my_chunk * kChunkRows + s * kBlockRows + lane);
        const unsigned long long key = (static_cast<unsigned long long>(__float_as_uint(v)) << 32) | (0xffffffffu - grow);
        if (v > 0.0f && key > best) best = key;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
        best = other > best ? other : best;
      }
      if (lane == 0 && best) atomicMax(&fmine.best[g], best);
    }
    __syncthreads();
    push_to_peers<kCluster, kLossThreads>(&fmine, &fmine, static_cast<int>(sizeof(ForceRecord) / 16), &sh.xbar[2], rank, tid);
    mbar_wait_cluster(&sh.xbar[2], 0);
    uint32_t c_lo = 0u, c_hi = 0u;
#pragma unroll
    for (int r = 0; r < kCluster; ++r) { c_lo |= frec[r].any[0]; c_hi |= frec[r].any[1]; }
    for (int g = 0; g < G; ++g) {
      const bool matched = ((g < 32 ? c_lo >> g : c_hi >> (g - 32)) & 1u) != 0u;
      if (matched || !(gts[g].area > 0.0f)) continue;
      unsigned long long best = 0ull;
#pragma unroll
      for (int r = 0; r < kCluster; ++r) best = frec[r].best[g] > best ? frec[r].best[g] : best;
      if (best == 0ull) continue;                                   // the box overlaps no prior at all
      const int grow = static_cast<int>(0xffffffffu - static_cast<uint32_t>(best & 0xffffffffull));
      const int chunk = grow / kChunkRows, within = grow - chunk * kChunkRows;
      if (chunk == my_chunk && (within & 31) == lane) {
        const int s_hit = within >> 5;
#pragma unroll
        for (int s = 0; s < kSlots; ++s)
          if (s == s_hit) { if (g < 32) mlo[s] |= 1u << g; else mhi[s] |= 1u << (g - 32); }
      }
    }
    __syncthreads();                                                // the list records get their own meaning back
    if (tid == 0) sh.lists[rank].cnt = 0;
  }
  trace_point_t<kTrace>(p, 3);

  // ---- per-row terms ----------------------------------------------------------------------------------
  // ce: positive CE (matched rows) or negative CE (unmatched rows); lloc: smooth-L1 sum; lse: log-sum-exp.
  // Matched rows leave sum_g clamp(l - g_hat, -1, 1) -- the localisation gradient -- in their offset columns.
  float ce[kSlots], lloc[kSlots], lse[kSlots];
#pragma unroll
  for (int s = 0; s < kSlots; ++s) {
    ce[s] = 0.0f; lloc[s] = 0.0f; lse[s] = 0.0f;
    if (s * kBlockRows >= my_rows_w) continue;                      // uniform per warp
    if (p.bulk && s == 0) mbar_wait(&sh.mbar[warp], 0);             // the whole chunk lands on one barrier
    if (s == 0) trace_point_t<kTrace>(p, 4);
    float* rp = my_slab + static_cast<size_t>(s * kBlockRows + lane) * row;
    if (valid[s]) {
      float mx, sum = 0.0f;
      if (kExact) {
        mx = rp[4];
        for (int c = 1; c < C; ++c) mx = fmaxf(mx, rp[4 + c]);
        for (int c = 0; c < C; ++c) sum += expf(rp[4 + c] - mx);
      } else if (kC > 0) {
        float x[kC > 0 ? kC : 1];
#pragma unroll
        for (int c = 0; c < kC; ++c) x[c] = rp[4 + c];
        mx = x[0];
#pragma unroll
        for (int c = 1; c + 1 < kC; c += 2) mx = fmaxf(mx, fmaxf(x[c], x[c + 1]));       // FMNMX3
        if ((kC & 1) == 0) mx = fmaxf(mx, x[kC - 1]);
        const float nmxs = -(mx * kLog2e);
        const unsigned long long l2 = pack2(kLog2e, kLog2e), m2 = pack2(nmxs, nmxs);
        unsigned long long acc2 = pack2(0.0f, 0.0f);
#pragma unroll
        for (int c = 0; c + 1 < kC; c += 2) {                                             // FFMA2 / FADD2: two classes per slot
          float a0, a1;
          unpack2(ffma2(pack2(x[c], x[c + 1]), l2, m2), a0, a1);
          acc2 = fadd2(acc2, pack2(ex2_approx(a0), ex2_approx(a1)));
        }
        float s0, s1;
        unpack2(acc2, s0, s1);
        sum = s0 + s1;
        if (kC & 1) sum += ex2_approx(fmaf(x[kC - 1], kLog2e, nmxs));
      } else {
        mx = rp[4];
        for (int c = 1; c < C; ++c) mx = fmaxf(mx, rp[4 + c]);
        const float mxs = mx * kLog2e;
        for (int c = 0; c < C; ++c) sum += ex2_approx(fmaf(rp[4 + c], kLog2e, -mxs));
      }
      const float ls = kExact ? logf(sum) : __logf(sum);             // lg2.approx: |error| ~1e-7 on sums in [1, C], ample for 1e-5
      lse[s] = mx + ls;
      ce[s] = ls - (rp[4] - mx);                                     // -log_softmax[void]   (ssd.py:212-215)
      if ((mlo[s] | mhi[s]) != 0u) {
        const float4 q = pri[s];
        const float ldw = kExact ? logf(q.z) : __logf(q.z), ldh = kExact ? logf(q.w) : __logf(q.w);   // offsets only feed loss values: fast math is ample
        const float rdw = kExact ? __fdiv_rn(1.0f, q.z) : __fdividef(1.0f, q.z), rdh = kExact ? __fdiv_rn(1.0f, q.w) : __fdividef(1.0f, q.w);
        const float l0 = rp[0], l1 = rp[1], l2 = rp[2], l3 = rp[3];
        // l - g-hat per matched row (ssd.py:202-204, 267-270) = (l + d_c / d_w) - g_c / d_w and (l + log d_w) - log g_w:
        // the prior-only parts are hoisted, leaving one FMA / one add per coordinate and pair
        const float a0 = fmaf(q.x, rdw, l0), a1 = fmaf(q.y, rdh, l1), b2 = l2 + ldw, b3 = l3 + ldh;
        float acc_ce = 0.0f, acc_loc = 0.0f, g0 = 0.0f, g1 = 0.0f, g2 = 0.0f, g3 = 0.0f;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          uint32_t m = half ? mhi[s] : mlo[s];
          while (m) {
            const int g = 32 * half + __ffs(m) - 1;
            m &= m - 1;
            const float4 go = *reinterpret_cast<const float4*>(&gts[g].cx);      // cx, cy, lw, lh
            const int2 lf = *reinterpret_cast<const int2*>(&gts[g].label);         // label, flags
            if (lf.x >= 0) {
              acc_ce += ls - (rp[4 + lf.x] - mx);                    // -log_softmax[label]  (ssd.py:208-209)
            } else {
              const float* tw = p.targets + (static_cast<size_t>(n) * G + g) * row + 4;
              float dot = 0.0f;
#pragma unroll 1
              for (int c = 0; c < C; ++c) dot += tw[c] * ((rp[4 + c] - mx) - ls);      // soft labels: rare, kept small
              acc_ce += -dot;
            }
            const float x0 = fmaf(-go.x, rdw, a0);
            const float x1 = fmaf(-go.y, rdh, a1);
            const float x2 = ((lf.y & 1) ? b2 : l2) - go.z;
            const float x3 = ((lf.y & 2) ? b3 : l3) - go.w;
            // with c = clamp(x, -1, 1): smooth_l1(x) = c * (x - c / 2)  (ssd.py:283) and c is its derivative
            const float c0 = clamp1(x0), c1 = clamp1(x1), c2 = clamp1(x2), c3 = clamp1(x3);
            acc_loc = fmaf(c0, fmaf(-0.5f, c0, x0), acc_loc);
            acc_loc = fmaf(c1, fmaf(-0.5f, c1, x1), acc_loc);
            acc_loc = fmaf(c2, fmaf(-0.5f, c2, x2), acc_loc);
            acc_loc = fmaf(c3, fmaf(-0.5f, c3, x3), acc_loc);
            g0 += c0; g1 += c1; g2 += c2; g3 += c3;
          }
        }
        ce[s] = acc_ce;
        lloc[s] = acc_loc;
        rp[0] = g0; rp[1] = g1; rp[2] = g2; rp[3] = g3;
      }
    }
    if (kExact && p.ce_override != nullptr && valid[s])          // test hook: the selection under test is fed the checker's own cross-entropies
      ce[s] = p.ce_override[static_cast<size_t>(n) * p.P + my_chunk * kChunkRows + s * kBlockRows + lane];
    if (valid[s]) {
      const uint32_t b = bucket_of(float_key(ce[s]));
      atomicAdd(&sh.mine.hist[(mlo[s] | mhi[s]) != 0u ? 0 : 1][b >> 1], 1u << (16 * (b & 1u)));
    }
  }

  {
    int c = 0;
#pragma unroll
    for (int s = 0; s < kSlots; ++s) c += (valid[s] && (mlo[s] | mhi[s]) != 0u) ? 1 : 0;
    c = warp_sum(c);
    if (lane == 0 && c) atomicAdd(&sh.mine.pos_local, c);
  }

  if (kTrace && p.trace != nullptr && lane == 0) p.trace[static_cast<size_t>(blockIdx.x) * kTracePoints + 24 + warp] = clock64();   // per-warp end of the row phase

  // ---- software pipelining across micro-batches: HBM goes quiet from here until the gradient leaves, so every
  // warp now asks the L2 for the same blocks of the NEXT batch (this SM will read them again in the next launch) ------
  if (p.next_outputs != nullptr && p.bulk && lane == 0 && my_rows_w > 0) {
    const uint32_t bytes = static_cast<uint32_t>(my_rows_w) * row * sizeof(float);
    const float* nsrc = p.next_outputs + static_cast<size_t>(n) * p.P * row + static_cast<size_t>(my_chunk) * kChunkRows * row;
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(nsrc), "r"(bytes) : "memory");
  }
  if (p.next_targets != nullptr && rank == 0 && tid == 32 && G > 0) {
    const uint32_t bytes = (static_cast<uint32_t>(G) * row * sizeof(float)) & ~15u;
    const float* nt = p.next_targets + static_cast<size_t>(n) * G * row;
    if (bytes && (reinterpret_cast<uintptr_t>(nt) & 15u) == 0) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(nt), "r"(bytes) : "memory");
  }

  // ---- cluster exchange #1: positives + bucket histograms of both sets, pushed into every peer -----------------------
  trace_point_t<kTrace>(p, 5);
  __syncthreads();                                   // my record is complete
  for (int i = tid; i < static_cast<int>(sizeof(HistRecord) / 16) * (kCluster - 1); i += kLossThreads) {
    constexpr int kW = static_cast<int>(sizeof(HistRecord) / 16);
    const int q = i / kW, idx = i - q * kW;
    int dst = rank + 1 + q;
    if (dst >= kCluster) dst -= kCluster;
    const int slot = kCluster - 2 - q;               // = (rank - dst - 1) mod kCluster: where the receiver files my record
    st_async_v4(map_to_cta(reinterpret_cast<uint4*>(&sh.from[slot]) + idx, dst), reinterpret_cast<const uint4*>(&sh.mine)[idx],
                map_to_cta(&sh.xbar[0], dst));
  }
  mbar_wait_cluster(&sh.xbar[0], 0);                 // the peers' records have landed in my shared memory
  trace_point_t<kTrace>(p, 6);
  if (tid < 256) {
    const int set = tid >> 7, w = tid & 127;
    uint32_t t = sh.mine.hist[set][w];
#pragma unroll
    for (int r = 0; r < kCluster - 1; ++r) t += sh.from[r].hist[set][w];
    tot_a[set * 128 + w] = t;
  } else if (warp == 8) {
    int v = lane == 0 ? sh.mine.pos_local : (lane < kCluster ? sh.from[lane - 1].pos_local : 0);
    v = warp_sum(v);
    if (lane == 0) sh.pos_raw = v;
  }
  __syncthreads();
  if (warp == 0) {
    // 3:1 split (ssd.py:218-220, 310-311) and which threshold needs a search.  Rows outside a set
    // contribute exact zeros to that set's CE array, so with M members and k <= M the (k+1)-th largest
    // is 0 whenever k == M; k < M happens for at most one of the two sets.
    const int pos_raw = sh.pos_raw, neg_raw = p.P - pos_raw;
    const bool crowded = pos_raw * 3 > neg_raw;
    const int k_pos = crowded ? neg_raw / 3 : pos_raw;
    const int k_neg = crowded ? neg_raw : pos_raw * 3;
    int set = -1, k = 0;
    if (k_pos < pos_raw) { set = 0; k = k_pos; }
    else if (k_neg < neg_raw) { set = 1; k = k_neg; }
    uint32_t rem = static_cast<uint32_t>(k);
    int bin = 0;
    if (set >= 0) bin = find_bin_desc(tot_a + set * 128, rem, lane);
    if (lane == 0) {
      sh.k_pos = k_pos; sh.k_neg = k_neg;
      sh.sel_set = set; sh.need_select = set >= 0;
      sh.sel_prefix = static_cast<uint32_t>(bin);       // selected bucket
      sh.sel_rem = rem;                                 // rank of the answer inside the bucket
    }
  }
  __syncthreads();
  // The local radix passes use from[0].hist (the peers' histograms are consumed) and tot[*] as their four zeroed
  // histograms; both are free from here on (the block barriers below order this clear before their reuse).
  uint32_t* const hist1 = &sh.from[0].hist[0][0];      // second histogram buffer: [2][128] words
  if (tid < 256) { scratch[tid] = 0u; hist1[tid] = 0u; }
  trace_point_t<kTrace>(p, 7);

  const int sel_set = sh.sel_set;
  const bool need_select = sh.need_select != 0;
  uint32_t sel_key = 0;          // order key of the searched threshold (identical in every thread of the cluster)
  if (need_select) {
    // my candidates of the selected bucket -> my list record (order is irrelevant)
    const uint32_t bucket = sh.sel_prefix;
    ListRecord& my_list = sh.lists[rank];
    bool member[kSlots];
    uint32_t ballot[kSlots];
    int mine_n = 0;
#pragma unroll
    for (int s = 0; s < kSlots; ++s) {
      member[s] = valid[s] && (((mlo[s] | mhi[s]) != 0u) == (sel_set == 0)) && (bucket_of(float_key(ce[s])) == bucket);
      ballot[s] = __ballot_sync(0xffffffffu, member[s]);
      mine_n += __popc(ballot[s]);
    }
    int base = 0;
    if (lane == 0 && mine_n) base = atomicAdd(&my_list.cnt, mine_n);      // one reservation per warp
    base = __shfl_sync(0xffffffffu, base, 0);
#pragma unroll
    for (int s = 0; s < kSlots; ++s) {
      const int pos = base + __popc(ballot[s] & lt_mask);
      if (member[s] && pos < kListCap) my_list.key[pos] = float_key(ce[s]);
      base += __popc(ballot[s]);
    }
    trace_point_t<kTrace>(p, 16);
    __syncthreads();                                  // my list record is complete
    push_to_peers<kCluster, kLossThreads>(&my_list, &my_list, static_cast<int>(sizeof(ListRecord) / 16), &sh.xbar[1], rank, tid);
    mbar_wait_cluster(&sh.xbar[1], 0);                // everyone's candidates are in my shared memory
    trace_point_t<kTrace>(p, 17);
    // the order statistic is finished locally (and identically) by every CTA
    int cnt[kCluster], total = 0;
    bool over = false;
#pragma unroll
    for (int r = 0; r < kCluster; ++r) {
      cnt[r] = sh.lists[r].cnt;
      over |= cnt[r] > kListCap;
      total += cnt[r];
    }
    uint32_t* const gathered = &sh.lists[0].key[0];   // the keys, compacted to the front of the list records
    uint32_t prefix = 0, rem = sh.sel_rem;
    if (!over) {
      // compaction in place: every element moves towards the front, so all of them are read before any is written
      constexpr int kPerThread = (kCluster * kListCap + kLossThreads - 1) / kLossThreads;
      uint32_t held[kPerThread];
#pragma unroll
      for (int j = 0; j < kPerThread; ++j) {
        const int i = tid + j * kLossThreads;
        int r = 0, off = i;
#pragma unroll
        for (int q = 0; q < kCluster - 1; ++q)
          if (r == q && off >= cnt[q]) { off -= cnt[q]; r = q + 1; }
        held[j] = i < total ? sh.lists[r].key[off] : 0u;
      }
      __syncthreads();
#pragma unroll
      for (int j = 0; j < kPerThread; ++j) {
        const int i = tid + j * kLossThreads;
        if (i < total) gathered[i] = held[j];
      }
      __syncthreads();
      trace_point_t<kTrace>(p, 18);
      if (total <= 256) {
        // few candidates (the usual case): exact order statistic by counting, spread over the whole CTA.  With
        // above(v) = #{candidates > v}, the (rem+1)-th largest is the SMALLEST key whose above() is <= rem.  Thread t
        // compares candidate t % 256 with one part of the list and adds its partial count (tot[] is zero here).
        uint32_t* cnt_above = scratch;
        constexpr int kParts = kLossThreads / 256;                 // 3 with 768 threads, 1 with 384
        const int ci = tid & 255, part = tid >> 8;
        if (tid == 0) sh.sel_prefix = 0xffffffffu;
        if (ci < total && part < kParts) {
          const uint32_t mine = gathered[ci];
          const int per = (((total + kParts - 1) / kParts) + 3) & ~3;
          const int j0 = part * per, j1 = min(total, j0 + per);
          int above = 0;
          int j = j0;
          for (; j + 4 <= j1; j += 4) {
            const uint4 v = *reinterpret_cast<const uint4*>(&gathered[j]);
            above += (v.x > mine) + (v.y > mine) + (v.z > mine) + (v.w > mine);
          }
          for (; j < j1; ++j) above += gathered[j] > mine;
          if (above) atomicAdd(&cnt_above[ci], static_cast<uint32_t>(above));
        }
        __syncthreads();
        if (tid < total && cnt_above[tid] <= rem) atomicMin(&sh.sel_prefix, gathered[tid]);
        __syncthreads();
        prefix = sh.sel_prefix;
      } else {
      // Interior buckets pin the top 13 bits of the key: 19 bits remain -> 3 passes (8 + 8 + 3); the two clamped
      // end buckets span arbitrary keys -> 4 full passes.  One buffer per pass, one barrier per pass; every warp
      // resolves the bin redundantly from the same histogram.
      const bool interior = bucket > 0u && bucket < 255u;
      const int n_pass = interior ? 3 : 4;
      if (interior) prefix = (bucket + 4096u + ((127u - 12u) << 4)) << 19;
      for (int pass = 0; pass < n_pass; ++pass) {
        const int shift = interior ? (pass == 0 ? 11 : (pass == 1 ? 3 : 0)) : 24 - 8 * pass;
        const int bits = (interior && pass == 2) ? 3 : 8;
        const uint32_t himask = (shift + bits >= 32) ? 0u : (0xffffffffu << (shift + bits));
        uint32_t* lh = pass < 2 ? hist1 + 128 * pass : scratch + 128 * (pass - 2);
        for (int i = tid; i < total; i += kLossThreads) {
          const uint32_t key = gathered[i];
          if ((key & himask) == prefix) {
            const uint32_t bin = (key >> shift) & ((1u << bits) - 1u);
            atomicAdd(&lh[bin >> 1], 1u << (16 * (bin & 1u)));
          }
        }
        __syncthreads();
        const int bin = find_bin_desc(lh, rem, lane);
        prefix |= static_cast<uint32_t>(bin) << shift;
      }
      }
    } else {
      // fallback (some CTA holds more than kListCap candidates of the bucket, e.g. thousands of identical CEs):
      // cluster-wide 4-pass radix select on the full key, histograms pulled through distributed shared memory under
      // cooperative groups' cluster.sync() (release / acquire at cluster scope: slow, correct, and rare)
      uint32_t* const hbuf[2] = {&sh.mine.hist[0][0], hist1};
      for (int i = tid; i < 2 * 128; i += kLossThreads) { hbuf[0][i] = 0u; hbuf[1][i] = 0u; }
      __syncthreads();
      rem = static_cast<uint32_t>(sel_set == 0 ? sh.k_pos : sh.k_neg);
      for (int pass = 0; pass < 4; ++pass) {
        const int shift = 24 - 8 * pass;
        const int buf = pass & 1;
        const uint32_t himask = pass == 0 ? 0u : (0xffffffffu << (shift + 8));
#pragma unroll
        for (int s = 0; s < kSlots; ++s) {
          if (s * kBlockRows >= my_rows_w) continue;
          const uint32_t key = float_key(ce[s]);
          const bool member = valid[s] && (((mlo[s] | mhi[s]) != 0u) == (sel_set == 0)) && ((key & himask) == prefix);
          hist_add(hbuf[buf], member, sel_set, (key >> shift) & 255u, lane);
        }
        cluster.sync();
        if (tid < 128) {
          uint32_t t = 0;
#pragma unroll
          for (int r = 0; r < kCluster; ++r) t += *cluster.map_shared_rank(hbuf[buf] + sel_set * 128 + tid, r);
          scratch[tid] = t;
        } else if (tid < 384) {
          hbuf[buf ^ 1][tid - 128] = 0u;
        }
        __syncthreads();
        const int bin = find_bin_desc(scratch, rem, lane);
        prefix |= static_cast<uint32_t>(bin) << shift;
        __syncthreads();
      }
      cluster.sync();               // nobody leaves while a peer may still be reading its histograms
    }
    sel_key = prefix;
    if (tid == 0) sh.overflow = over;
  }
  trace_point_t<kTrace>(p, 8);
  const float thr_sel = need_select ? key_float(sel_key) : 0.0f;
  const float thr_pos = sel_set == 0 ? thr_sel : 0.0f;
  const float thr_neg = sel_set == 1 ? thr_sel : 0.0f;
  const int k_pos = sh.k_pos;
  const float inv_pos = k_pos > 0 ? __fdiv_rn(1.0f, static_cast<float>(k_pos)) : 0.0f;     // ssd.py:226

  // ---- masked sums (ssd.py:227) ---------------------------------------------------------------------------
  // The CTA's partial goes to a global slot; the LAST of the image's CTAs to arrive (ticket) adds the partials in rank
  // order -- deterministic, and nobody waits: the other CTAs go straight on to their gradient rows.
  bool sel[kSlots];
  asm volatile("griddepcontrol.wait;" ::: "memory");     // first global writes below: the previous grid (same workspace) is complete
  {
    float accf = 0.0f;
    int cnt2 = 0;                            // selected positives | selected negatives << 16
#pragma unroll
    for (int s = 0; s < kSlots; ++s) {
      const bool pos = (mlo[s] | mhi[s]) != 0u;
      sel[s] = valid[s] && (ce[s] > (pos ? thr_pos : thr_neg));
      if (sel[s]) {
        accf += pos ? (p.a * lloc[s] + ce[s]) : ce[s];
        cnt2 += pos ? 1 : 65536;
      }
    }
    const double acc = warp_sum(static_cast<double>(accf));
    cnt2 = warp_sum(cnt2);
    if (lane == 0) { wred_loss[warp] = acc; wred_a[warp] = cnt2; }
    __syncthreads();
    if (warp == kLossWarps - 1) {            // the warp with the least row work (none at all for P = 8732): the global
                                             // round trips below stay off the other warps' gradient rows
      double t = lane < kLossWarps ? wred_loss[lane] : 0.0;
      int c2 = lane < kLossWarps ? wred_a[lane] : 0;
      t = warp_sum(t);                      // fixed shuffle tree: deterministic
      c2 = warp_sum(c2);
      ImageSlot* slot = p.slots + n;
      unsigned int arrived = 0u;
      if (lane == 0) {
        slot->part_loss[rank] = t;
        slot->part_sel[rank] = c2;
        __threadfence();
        arrived = atomicAdd(&slot->ticket, 1u);
      }
      arrived = __shfl_sync(0xffffffffu, arrived, 0);
      if (arrived == static_cast<unsigned int>(kCluster) - 1u) {       // last CTA of the image: the whole warp helps
        __threadfence();
        const double pl = lane < kCluster ? __ldcg(&slot->part_loss[lane]) : 0.0;
        const int c = lane < kCluster ? __ldcg(&slot->part_sel[lane]) : 0;
        double total = 0.0;
#pragma unroll
        for (int r = 0; r < kCluster; ++r) total += __shfl_sync(0xffffffffu, pl, r);      // rank order: deterministic
        const int pos_sel = warp_sum(c & 0xffff), neg_sel = warp_sum(c >> 16);
        unsigned int done = 0u;
        if (lane == 0) {
          slot->ticket = 0u;
          const float li = static_cast<float>(total) * inv_pos;
          if (p.stats) {
            ssdh_image_stats st;
            st.loss = li; st.thr_pos = thr_pos; st.thr_neg = thr_neg;
            st.pos_raw = sh.pos_raw; st.k_pos = k_pos; st.k_neg = sh.k_neg; st.pos_sel = pos_sel; st.neg_sel = neg_sel;
            p.stats[n] = st;
          }
          slot->image_loss = static_cast<double>(li);
          __threadfence();
          done = atomicAdd(p.ticket, 1u);
        }
        done = __shfl_sync(0xffffffffu, done, 0);
        if (done == static_cast<unsigned int>(p.N) - 1u) {               // last image of the batch
          __threadfence();
          double acc = 0.0;                 // lane l adds images l, l + 32, ... in order, then a fixed shuffle tree
          for (int i = lane; i < p.N; i += 32) acc += __ldcg(&p.slots[i].image_loss);
          acc = warp_sum(acc);
          const float batch_loss = static_cast<float>(acc * static_cast<double>(p.inv_n_global));
          if (lane == 0) {
            *p.loss = batch_loss;
            *p.ticket = 0u;
          }
          // The collective that follows the step, fused: lane r stores (step << 32 | loss bits) into rank r's inbox -- one 8-byte
          // store per peer over NVLink (the local inbox included), value and step number in ONE word.
          if (p.xchg.world > 0) {
            uint32_t step = 0u;
            if (lane == 0) { step = p.xchg.counters[0] + 1u; p.xchg.counters[0] = step; }
            step = __shfl_sync(0xffffffffu, step, 0);
            if (lane < p.xchg.world) {
              unsigned long long* dst = p.xchg.inbox[lane] + static_cast<size_t>(p.xchg.rank) * p.xchg.ring + ((step - 1u) % p.xchg.ring);
              const unsigned long long word = (static_cast<unsigned long long>(step) << 32) | __float_as_uint(batch_loss);
              asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(dst), "l"(word) : "memory");
            }
          }
        }
      }
    }
  }
  trace_point_t<kTrace>(p, 9);
  // ---- gradient rows, in place over the slab, then out by TMA -------------------------------------------------
  if (p.grad != nullptr) {
    const float sn = inv_pos * p.inv_n_global;          // d loss / d (per-image sum)
    float* img_out = p.grad + static_cast<size_t>(n) * p.P * row;
#pragma unroll
    for (int s = 0; s < kSlots; ++s) {
      if (s * kBlockRows >= my_rows_w) continue;                    // uniform per warp
      float* rp = my_slab + static_cast<size_t>(s * kBlockRows + lane) * row;
      // One code path for every lane: row <- scale * softmax(row) with scale = s_n * (sum of matched class weights)
      // for a selected positive, s_n for a selected negative and 0 for an unselected row (exact zeros); then the
      // one-hot corrections.  Warps without any selected row (the common case late in training) just clear.
      if (!__any_sync(0xffffffffu, sel[s])) {
        if (valid[s])
          for (int c = 0; c < row; ++c) rp[c] = 0.0f;
      } else if (valid[s]) {
        const bool pos = (mlo[s] | mhi[s]) != 0u;
        float tsum = 1.0f;
        if (pos) {
          if (!sh.soft_labels) {
            tsum = static_cast<float>(__popc(mlo[s]) + __popc(mhi[s]));
          } else {
            tsum = 0.0f;
            for (int half = 0; half < 2; ++half) {
              uint32_t m = half ? mhi[s] : mlo[s];
#pragma unroll 1
              while (m) { tsum += gts[32 * half + __ffs(m) - 1].tsum; m &= m - 1; }
            }
          }
        }
        softmax_scaled<kC, kExact>(rp + 4, C, lse[s], sel[s] ? sn * tsum : 0.0f);
        const float as = (sel[s] && pos) ? p.a * sn : 0.0f;       // offset columns of matched rows hold sum_g clamp(l - g_hat)
        rp[0] *= as; rp[1] *= as; rp[2] *= as; rp[3] *= as;
        if (sel[s]) {
          if (!pos) {
            rp[4] -= sn;
          } else {
#pragma unroll
            for (int half = 0; half < 2; ++half) {
              uint32_t m = half ? mhi[s] : mlo[s];
              while (m) {
                const int g = 32 * half + __ffs(m) - 1;
                m &= m - 1;
                const int label = gts[g].label;
                if (label >= 0) {
                  rp[4 + label] -= sn;
                } else {
                  const float* tw = p.targets + (static_cast<size_t>(n) * G + g) * row + 4;
#pragma unroll 1
                  for (int c = 0; c < C; ++c) rp[4 + c] -= sn * tw[c];      // soft labels: rare, kept small
                }
              }
            }
          }
        }
      }
      const int rows_b = min(kBlockRows, my_rows_w - s * kBlockRows);
      float* dst = img_out + (static_cast<size_t>(my_chunk) * kChunkRows + s * kBlockRows) * row;
      const float* srcb = my_slab + static_cast<size_t>(s) * kBlockRows * row;
      if (p.bulk) {
        fence_async_smem();               // my generic-proxy writes -> visible to the TMA engine
        __syncwarp();
        if (lane == 0) {
          bulk_store(dst, srcb, static_cast<uint32_t>(rows_b) * row * sizeof(float));
          bulk_store_commit();
        }
      } else {
        __syncwarp();
#pragma unroll 1
        for (int i = lane; i < rows_b * row; i += 32) dst[i] = srcb[i];
      }
    }
  }

  trace_point_t<kTrace>(p, 10);
  if (p.bulk && p.grad != nullptr && lane == 0) bulk_store_wait();
  // No trailing cluster barrier: peers never READ my shared memory (they were sent copies), and everything they WRITE into it
  // has been received above -- the exchange waits are what lets a CTA retire.
  trace_point_t<kTrace>(p, 11);
  trace_point_t<kTrace>(p, 12);
  if (kTrace && p.trace != nullptr && tid == 0) {
    unsigned int smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    p.trace[static_cast<size_t>(blockIdx.x) * kTracePoints + 13] = smid;
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    p.trace[static_cast<size_t>(blockIdx.x) * kTracePoints + 14] = t;
    p.trace[static_cast<size_t>(blockIdx.x) * kTracePoints + 15] = static_cast<unsigned long long>(sh.overflow);
  }
}

static inline size_t loss_smem_bytes(int rows_per_cta, int row, int G, int cluster) {
  const size_t slab = (static_cast<size_t>(rows_per_cta) * row * sizeof(float) + 15) & ~static_cast<size_t>(15);
  const size_t gt = (static_cast<size_t>(G) * sizeof(GtRec) + 15) & ~static_cast<size_t>(15);
  return slab + gt + (cluster == 4 ? sizeof(LossSharedT<4>) : sizeof(LossSharedT<8>));
}

static inline int chunks_for(int P) { return (P + kChunkRows - 1) / kChunkRows; }

static inline int rows_per_cta_for(int P, int cluster) {          // shared-memory rows of the busiest CTA
  return ((chunks_for(P) + cluster - 1) / cluster) * kChunkRows;
}

constexpr size_t kMaxDynSmem = 227 * 1024;

// Which kernel shape serves (P, C, G): 4 CTAs x 768 threads when the quarter slab fits one SM, else 8 x 384.
struct LossShape { int cluster, threads, rows_per_cta; size_t smem; };

static inline bool pick_shape(int P, int C, int G, LossShape* out) {
  static const int forced = [] { const char* e = getenv("SSDH_LOSS_CLUSTER"); return e ? atoi(e) : 0; }();
  const int row = 4 + C;
  const int shapes[2][2] = {{4, 768}, {8, 384}};
  for (int i = 0; i < 2; ++i) {
    if (forced && shapes[i][0] != forced) continue;
    LossShape s;
    s.cluster = shapes[i][0]; s.threads = shapes[i][1];
    s.rows_per_cta = rows_per_cta_for(P, s.cluster);
    s.smem = loss_smem_bytes(s.rows_per_cta, row, G, s.cluster);
    if (s.rows_per_cta <= kSlots * s.threads && s.smem <= kMaxDynSmem) { *out = s; return true; }
  }
  return false;
}

template <int kC, int kT, int kCl, int kMode>
static int launch_loss(const LossParams& p, size_t smem, cudaStream_t st) {
  if (int e = ensure_dyn_smem(reinterpret_cast<const void*>(multibox_loss_kernel<kC, kT, kCl, kMode>), static_cast<int>(kMaxDynSmem), "ssdh_multibox_loss")) return e;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(static_cast<unsigned>(p.N) * kCl);
  cfg.blockDim = dim3(kT);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kCl;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  static const int pdl = [] { const char* e = getenv("SSDH_LOSS_PDL"); return e ? atoi(e) : 1; }();
  cfg.numAttrs = pdl ? 2 : 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, multibox_loss_kernel<kC, kT, kCl, kMode>, p);
  if (e != cudaSuccess) { set_error("ssdh_multibox_loss: launch: %s", cudaGetErrorString(e)); (void)cudaGetLastError(); return static_cast<int>(e); }
  return 0;
}

// One kernel mode, both class-count instantiations and both shapes.
template <int kMode>
static int launch_loss_mode(const LossParams& p, const LossShape& s, cudaStream_t st) {
  if (p.C == 21) {
    if (s.cluster == 4) return launch_loss<21, 768, 4, kMode>(p, s.smem, st);
    return launch_loss<21, 384, 8, kMode>(p, s.smem, st);
  }
  if (s.cluster == 4) return launch_loss<0, 768, 4, kMode>(p, s.smem, st);
  return launch_loss<0, 384, 8, kMode>(p, s.smem, st);
}

// loss_ext.cu: the opt-in modes (forcing, exact math) live in their own translation unit so that they compile in parallel
// with -- and never perturb the code of -- the production kernel.
int launch_loss_extension(int mode, const LossParams& p, const LossShape& s, cudaStream_t st);

}  // namespace ssdh
