// Fused MultiBox loss, forward + gradient in ONE launch.  Replaces SSD.loss and everything it calls
// (reference src/model/ssd.py:181-328: _match, _calc_delta, _smooth_l1, _softmax_cross_entropy,
// _split_pos_neg, _k_plus_1_th_value) plus the autograd backward of that graph (src/train.py:121).
//
// Decomposition (B200, sm_100a):
//   * one thread-block CLUSTER of 8 CTAs per image; CTA r owns rows [r*R, (r+1)*R) of the image's
//     contiguous [P, 4+C] slab (R = 1092 for P = 8732) and keeps them in shared memory for the whole
//     kernel, so HBM sees each output row exactly once as a read and (with grad) once as a write;
//   * the slab arrives by TMA bulk copies (cp.async.bulk + mbarrier, one chunk per row slot) while the
//     threads compute the IoU match masks, which only need the priors and the ground truth;
//   * one thread per row: log-sum-exp, positive / negative cross-entropy, smooth-L1 of matched pairs;
//   * hard-negative mining = value threshold at the (k+1)-th largest CE (strict '>', ssd.py:222-223),
//     found by an 8-bit radix select whose per-CTA histograms are combined through distributed shared
//     memory; only ONE of the two thresholds ever needs a search (see select logic below);
//   * gradient rows are written in place over the slab and leave by TMA bulk stores;
//   * the last image to finish reduces the per-image losses in a fixed order (deterministic).
#include <cooperative_groups.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace ssdh {

constexpr int kLossThreads = 384;
constexpr int kSlots = 3;            // rows per thread
constexpr int kCluster = 8;
constexpr int kLossWarps = kLossThreads / 32;

struct LossParams {
  const float* outputs;
  const float* targets;
  const float4* priors;
  int N, P, C, G;
  float a;
  ThrBand band;
  float inv_n_global;
  float* loss;
  float* grad;
  ssdh_image_stats* stats;
  unsigned int* ticket;   // workspace: zero before first use, left zero
  double* image_loss;     // workspace [N]
  int rows_per_cta;
  int bulk;               // 1: every chunk is 16-byte aligned/sized -> TMA path
};

struct GtRec {             // 48 bytes, one per ground-truth row of the image
  float x1, x2, y1, y2;    // corners                      (ssd.py:247-248)
  float area, cx, cy, lw;  // lw = log(w) (or w when w <= 0, ssd.py:269)
  float lh;
  int label;               // class index when the class vector is exactly one-hot, else -1
  int flags;               // bit0: w > 0, bit1: h > 0
  float tsum;              // sum of the class vector
};

struct LossShared {
  uint32_t hist[2][2][128];   // [buffer][set][256 bins packed as 2 x u16]
  uint32_t tot[2][128];       // cluster-wide packed totals
  unsigned long long mbar[kSlots];
  unsigned long long deg_mask, deg_hit;   // gt rows with area <= 0 and their constant verdict
  double part_loss[kCluster];             // leader only: written remotely by every CTA
  int part_pos_sel[kCluster], part_neg_sel[kCluster];
  double wred_loss[kLossWarps];
  int wred_a[kLossWarps], wred_b[kLossWarps];
  int pos_local;
  int pos_raw, k_pos, k_neg, sel_set, need_select;
  uint32_t sel_prefix, sel_rem;
};

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
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
__device__ __forceinline__ void bulk_store(void* gmem_dst, const void* smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst), "r"(smem_u32(smem_src)), "r"(bytes)
               : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_store_wait() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ float smooth_l1_f(float x) {
  const float ax = fabsf(x);
  return ax < 1.0f ? 0.5f * x * x : ax - 0.5f;
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
  int bin = 0;
  uint32_t r = 0;
  if (lane == owner) {
    r = rem - static_cast<uint32_t>(incl - s);
    int i = 0;
    for (; i < 7; ++i) {
      if (r < c[i]) break;
      r -= c[i];
    }
    bin = 255 - 8 * lane - i;
  }
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

template <int kC>
__global__ void __launch_bounds__(kLossThreads, 2) multibox_loss_kernel(const LossParams p) {
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = static_cast<int>(cluster.block_rank());
  const int n = blockIdx.x / kCluster;
  const int C = kC ? kC : p.C;
  const int row = 4 + C;
  const int G = p.G;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int row0 = rank * p.rows_per_cta;
  const int my_rows = max(0, min(p.rows_per_cta, p.P - row0));
  constexpr float kLog2e = 1.4426950408889634f;

  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* slab = reinterpret_cast<float*>(smem_raw);
  const size_t slab_bytes = (static_cast<size_t>(p.rows_per_cta) * row * sizeof(float) + 15) & ~static_cast<size_t>(15);
  GtRec* gts = reinterpret_cast<GtRec*>(smem_raw + slab_bytes);
  LossShared& sh = *reinterpret_cast<LossShared*>(smem_raw + slab_bytes + ((static_cast<size_t>(G) * sizeof(GtRec) + 15) & ~static_cast<size_t>(15)));

  const float* src = p.outputs + (static_cast<size_t>(n) * p.P + row0) * row;

  // ---- setup ------------------------------------------------------------------------------------
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < kSlots; ++s) mbar_init(&sh.mbar[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    sh.pos_local = 0;
    sh.deg_mask = 0ull;
    sh.deg_hit = 0ull;
  }
  for (int i = tid; i < 2 * 2 * 128; i += kLossThreads) (&sh.hist[0][0][0])[i] = 0u;
  __syncthreads();

  if (p.bulk) {
    if (tid == 0) {
#pragma unroll
      for (int s = 0; s < kSlots; ++s) {
        const int rows_s = min(kLossThreads, my_rows - s * kLossThreads);
        if (rows_s > 0) {
          const uint32_t bytes = static_cast<uint32_t>(rows_s) * row * sizeof(float);
          mbar_expect_tx(&sh.mbar[s], bytes);
          bulk_load(slab + static_cast<size_t>(s) * kLossThreads * row, src + static_cast<size_t>(s) * kLossThreads * row, bytes, &sh.mbar[s]);
        }
      }
    }
  } else {
    const int total = my_rows * row;
    for (int i = tid; i < total; i += kLossThreads) slab[i] = src[i];
  }

  // ---- ground truth of this image -> shared -------------------------------------------------------
  for (int g = tid; g < G; g += kLossThreads) {
    const float* tr = p.targets + (static_cast<size_t>(n) * G + g) * row;
    const float gcx = tr[0], gcy = tr[1], gw = tr[2], gh = tr[3];
    const Corners c = make_corners(gcx, gcy, gw, gh);
    GtRec r;
    r.x1 = c.x1; r.x2 = c.x2; r.y1 = c.y1; r.y2 = c.y2;
    r.area = c.area; r.cx = gcx; r.cy = gcy;
    r.lw = gw > 0.0f ? logf(gw) : gw;
    r.lh = gh > 0.0f ? logf(gh) : gh;
    r.flags = (gw > 0.0f ? 1 : 0) | (gh > 0.0f ? 2 : 0);
    int nz = 0, label = -1;
    float tsum = 0.0f;
    for (int c2 = 0; c2 < C; ++c2) {
      const float v = tr[4 + c2];
      tsum += v;
      if (v != 0.0f) { ++nz; label = (v == 1.0f) ? c2 : -1 - C; }
    }
    r.label = (nz == 1 && label >= 0) ? label : -1;
    r.tsum = tsum;
    gts[g] = r;
    if (!(c.area > 0.0f)) {
      atomicOr(&sh.deg_mask, 1ull << g);
      if (c.area > p.band.thr) atomicOr(&sh.deg_hit, 1ull << g);
    }
  }

  // ---- priors of my rows ------------------------------------------------------------------------------
  Corners d[kSlots];
  bool valid[kSlots];
#pragma unroll
  for (int s = 0; s < kSlots; ++s) {
    const int lr = s * kLossThreads + tid;
    valid[s] = lr < my_rows;
    const float4 q = p.priors[valid[s] ? row0 + lr : 0];
    d[s] = make_corners(q.x, q.y, q.z, q.w);
  }
  __syncthreads();

  // ---- matching: bit g of mask[s] = IoU(gt g, prior) > thr   (ssd.py:231-250) ---------------------------
  unsigned long long mask[kSlots];
#pragma unroll
  for (int s = 0; s < kSlots; ++s) mask[s] = 0ull;
  {
    const unsigned long long deg_mask = sh.deg_mask, deg_hit = sh.deg_hit;
    const ThrBand band = p.band;
    for (int g = 0; g < G; ++g) {
      const unsigned long long bit = 1ull << g;
      if (deg_mask & bit) {
        if (deg_hit & bit) {
#pragma unroll
          for (int s = 0; s < kSlots; ++s) mask[s] |= bit;
        }
        continue;
      }
      const float4 q = *reinterpret_cast<const float4*>(&gts[g].x1);
      const float garea = gts[g].area;
#pragma unroll
      for (int s = 0; s < kSlots; ++s) {
        const float w = fmaxf(fminf(q.y, d[s].x2) - fmaxf(q.x, d[s].x1), 0.0f);
        const float h = fmaxf(fminf(q.w, d[s].y2) - fmaxf(q.z, d[s].y1), 0.0f);
        const float inter = w * h;
        const float uni = (garea + d[s].area) - inter;
        if (quotient_gt(inter, uni, band)) mask[s] |= bit;
      }
    }
  }
  {
    int c = 0;
#pragma unroll
    for (int s = 0; s < kSlots; ++s) c += (valid[s] && mask[s] != 0ull) ? 1 : 0;
    c = warp_sum(c);
    if (lane == 0 && c) atomicAdd(&sh.pos_local, c);
  }
  if (!p.bulk) __syncthreads();

  // ---- per-row terms ----------------------------------------------------------------------------------
  float ce[kSlots], lloc[kSlots], lse[kSlots];
#pragma unroll
  for (int s = 0; s < kSlots; ++s) {
    ce[s] = 0.0f; lloc[s] = 0.0f; lse[s] = 0.0f;
    const int rows_s = min(kLossThreads, my_rows - s * kLossThreads);
    if (rows_s <= 0) continue;                                      // uniform per CTA
    if (p.bulk) mbar_wait(&sh.mbar[s], 0);
    const int lr = s * kLossThreads + tid;
    const float* rp = slab + static_cast<size_t>(lr) * row;
    if (valid[s]) {
      float mx = rp[4];
      for (int c = 1; c < C; ++c) mx = fmaxf(mx, rp[4 + c]);
      const float mxs = mx * kLog2e;
      float sum = 0.0f;
      for (int c = 0; c < C; ++c) sum += ex2_approx(fmaf(rp[4 + c], kLog2e, -mxs));
      const float ls = logf(sum);
      lse[s] = mx + ls;
      unsigned long long m = mask[s];
      if (m == 0ull) {
        ce[s] = ls - (rp[4] - mx);                                   // -log_softmax[void]   (ssd.py:212-215)
      } else {
        const float4 q = p.priors[row0 + lr];
        const float ldw = logf(q.z), ldh = logf(q.w);
        const float l0 = rp[0], l1 = rp[1], l2 = rp[2], l3 = rp[3];
        float acc_ce = 0.0f, acc_loc = 0.0f;
        while (m) {
          const int g = __ffsll(static_cast<long long>(m)) - 1;
          m &= m - 1;
          const GtRec& r = gts[g];
          if (r.label >= 0) {
            acc_ce += ls - (rp[4 + r.label] - mx);                   // -log_softmax[label]  (ssd.py:208-209)
          } else {
            const float* tw = p.targets + (static_cast<size_t>(n) * G + g) * row + 4;
            float dot = 0.0f;
            for (int c = 0; c < C; ++c) dot += tw[c] * ((rp[4 + c] - mx) - ls);
            acc_ce += -dot;
          }
          const float e0 = __fdiv_rn(r.cx - q.x, q.z);               // g-hat (ssd.py:267-270)
          const float e1 = __fdiv_rn(r.cy - q.y, q.w);
          const float e2 = (r.flags & 1) ? r.lw - ldw : r.lw;
          const float e3 = (r.flags & 2) ? r.lh - ldh : r.lh;
          acc_loc += ((smooth_l1_f(l0 - e0) + smooth_l1_f(l1 - e1)) + smooth_l1_f(l2 - e2)) + smooth_l1_f(l3 - e3);
        }
        ce[s] = acc_ce;
        lloc[s] = acc_loc;
      }
    }
    hist_add(&sh.hist[0][0][0], valid[s], mask[s] != 0ull ? 0 : 1, float_key(ce[s]) >> 24, lane);
  }

  // ---- cluster exchange #1: positives + pass-0 histograms of both sets --------------------------------------
  cluster.sync();
  if (tid < 256) {
    const int set = tid >> 7, w = tid & 127;
    uint32_t t = 0;
#pragma unroll
    for (int r = 0; r < kCluster; ++r) t += *cluster.map_shared_rank(&sh.hist[0][set][w], r);
    sh.tot[set][w] = t;
  } else if (warp == 8) {
    int v = (lane < kCluster) ? *cluster.map_shared_rank(&sh.pos_local, lane) : 0;
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
    if (set >= 0) bin = find_bin_desc(sh.tot[set], rem, lane);
    if (lane == 0) {
      sh.k_pos = k_pos; sh.k_neg = k_neg;
      sh.sel_set = set; sh.need_select = set >= 0;
      sh.sel_prefix = static_cast<uint32_t>(bin) << 24;
      sh.sel_rem = rem;
    }
  }
  __syncthreads();

  const int sel_set = sh.sel_set;
  if (sh.need_select) {
    for (int pass = 1; pass < 4; ++pass) {
      const int shift = 24 - 8 * pass;
      const int buf = pass & 1;
      const uint32_t prefix = sh.sel_prefix;
      const uint32_t himask = 0xffffffffu << (shift + 8);
#pragma unroll
      for (int s = 0; s < kSlots; ++s) {
        if (min(kLossThreads, my_rows - s * kLossThreads) <= 0) continue;
        const uint32_t key = float_key(ce[s]);
        const bool member = valid[s] && ((mask[s] != 0ull) == (sel_set == 0)) && ((key & himask) == prefix);
        hist_add(&sh.hist[buf][0][0], member, sel_set, (key >> shift) & 255u, lane);
      }
      cluster.sync();
      if (tid < 128) {
        uint32_t t = 0;
#pragma unroll
        for (int r = 0; r < kCluster; ++r) t += *cluster.map_shared_rank(&sh.hist[buf][sel_set][tid], r);
        sh.tot[0][tid] = t;
      } else {
        (&sh.hist[buf ^ 1][0][0])[tid - 128] = 0u;      // 256 threads clear the other buffer for the next pass
      }
      __syncthreads();
      if (warp == 0) {
        uint32_t rem = sh.sel_rem;
        const int bin = find_bin_desc(sh.tot[0], rem, lane);
        if (lane == 0) {
          sh.sel_prefix = prefix | (static_cast<uint32_t>(bin) << shift);
          sh.sel_rem = rem;
        }
      }
      __syncthreads();
    }
  }
  const float thr_sel = sh.need_select ? key_float(sh.sel_prefix) : 0.0f;
  const float thr_pos = sel_set == 0 ? thr_sel : 0.0f;
  const float thr_neg = sel_set == 1 ? thr_sel : 0.0f;
  const int k_pos = sh.k_pos;
  const float inv_pos = k_pos > 0 ? __fdiv_rn(1.0f, static_cast<float>(k_pos)) : 0.0f;     // ssd.py:226

  // ---- masked sums (ssd.py:227) ---------------------------------------------------------------------------
  {
    double acc = 0.0;
    int npos = 0, nneg = 0;
#pragma unroll
    for (int s = 0; s < kSlots; ++s) {
      if (!valid[s]) continue;
      if (mask[s] != 0ull) {
        if (ce[s] > thr_pos) { acc += static_cast<double>(p.a * lloc[s] + ce[s]); ++npos; }
      } else if (ce[s] > thr_neg) {
        acc += static_cast<double>(ce[s]);
        ++nneg;
      }
    }
    acc = warp_sum(acc);
    npos = warp_sum(npos);
    nneg = warp_sum(nneg);
    if (lane == 0) { sh.wred_loss[warp] = acc; sh.wred_a[warp] = npos; sh.wred_b[warp] = nneg; }
    __syncthreads();
    if (tid == 0) {
      double t = 0.0;
      int a2 = 0, b2 = 0;
      for (int w = 0; w < kLossWarps; ++w) { t += sh.wred_loss[w]; a2 += sh.wred_a[w]; b2 += sh.wred_b[w]; }
      *cluster.map_shared_rank(&sh.part_loss[rank], 0) = t;
      *cluster.map_shared_rank(&sh.part_pos_sel[rank], 0) = a2;
      *cluster.map_shared_rank(&sh.part_neg_sel[rank], 0) = b2;
    }
  }

  // ---- gradient rows, in place over the slab, then out by TMA -------------------------------------------------
  if (p.grad != nullptr) {
    const float sn = inv_pos * p.inv_n_global;          // d loss / d (per-image sum)
    float* dst = p.grad + (static_cast<size_t>(n) * p.P + row0) * row;
#pragma unroll
    for (int s = 0; s < kSlots; ++s) {
      const int rows_s = min(kLossThreads, my_rows - s * kLossThreads);
      if (rows_s <= 0) continue;
      const int lr = s * kLossThreads + tid;
      float* rp = slab + static_cast<size_t>(lr) * row;
      if (valid[s]) {
        unsigned long long m = mask[s];
        const bool sel_p = (m != 0ull) && (ce[s] > thr_pos);
        const bool sel_n = (m == 0ull) && (ce[s] > thr_neg);
        if (sel_n) {
          const float ls2 = lse[s] * kLog2e;
          for (int c = 0; c < C; ++c) rp[4 + c] = sn * ex2_approx(fmaf(rp[4 + c], kLog2e, -ls2));
          rp[4] -= sn;
          rp[0] = 0.0f; rp[1] = 0.0f; rp[2] = 0.0f; rp[3] = 0.0f;
        } else if (sel_p) {
          const float4 q = p.priors[row0 + lr];
          const float ldw = logf(q.z), ldh = logf(q.w);
          const float l0 = rp[0], l1 = rp[1], l2 = rp[2], l3 = rp[3];
          float tsum = 0.0f, g0 = 0.0f, g1 = 0.0f, g2 = 0.0f, g3 = 0.0f;
          unsigned long long m2 = m;
          while (m2) {
            const int g = __ffsll(static_cast<long long>(m2)) - 1;
            m2 &= m2 - 1;
            const GtRec& r = gts[g];
            tsum += r.tsum;
            const float e0 = __fdiv_rn(r.cx - q.x, q.z);
            const float e1 = __fdiv_rn(r.cy - q.y, q.w);
            const float e2 = (r.flags & 1) ? r.lw - ldw : r.lw;
            const float e3 = (r.flags & 2) ? r.lh - ldh : r.lh;
            g0 += fminf(fmaxf(l0 - e0, -1.0f), 1.0f);
            g1 += fminf(fmaxf(l1 - e1, -1.0f), 1.0f);
            g2 += fminf(fmaxf(l2 - e2, -1.0f), 1.0f);
            g3 += fminf(fmaxf(l3 - e3, -1.0f), 1.0f);
          }
          const float ls2 = lse[s] * kLog2e;
          const float st = sn * tsum;
          for (int c = 0; c < C; ++c) rp[4 + c] = st * ex2_approx(fmaf(rp[4 + c], kLog2e, -ls2));
          while (m) {
            const int g = __ffsll(static_cast<long long>(m)) - 1;
            m &= m - 1;
            const GtRec& r = gts[g];
            if (r.label >= 0) {
              rp[4 + r.label] -= sn;
            } else {
              const float* tw = p.targets + (static_cast<size_t>(n) * G + g) * row + 4;
              for (int c = 0; c < C; ++c) rp[4 + c] -= sn * tw[c];
            }
          }
          const float as = p.a * sn;
          rp[0] = as * g0; rp[1] = as * g1; rp[2] = as * g2; rp[3] = as * g3;
        } else {
          for (int c = 0; c < row; ++c) rp[c] = 0.0f;
        }
      }
      if (p.bulk) {
        fence_async_smem();
        __syncthreads();
        if (tid == 0)
          bulk_store(dst + static_cast<size_t>(s) * kLossThreads * row, slab + static_cast<size_t>(s) * kLossThreads * row,
                     static_cast<uint32_t>(rows_s) * row * sizeof(float));
      }
    }
    if (!p.bulk) {
      __syncthreads();
      const int total = my_rows * row;
      for (int i = tid; i < total; i += kLossThreads) dst[i] = slab[i];
    }
  }

  // ---- cluster exchange #2: per-image loss, stats, batch mean ----------------------------------------------------
  cluster.sync();
  if (rank == 0 && tid == 0) {
    double total = 0.0;
    int pos_sel = 0, neg_sel = 0;
    for (int r = 0; r < kCluster; ++r) { total += sh.part_loss[r]; pos_sel += sh.part_pos_sel[r]; neg_sel += sh.part_neg_sel[r]; }
    const float li = static_cast<float>(total) * inv_pos;
    if (p.stats) {
      ssdh_image_stats st;
      st.loss = li; st.thr_pos = thr_pos; st.thr_neg = thr_neg;
      st.pos_raw = sh.pos_raw; st.k_pos = k_pos; st.k_neg = sh.k_neg; st.pos_sel = pos_sel; st.neg_sel = neg_sel;
      p.stats[n] = st;
    }
    p.image_loss[n] = static_cast<double>(li);
    __threadfence();
    const unsigned int t = atomicAdd(p.ticket, 1u);
    if (t == static_cast<unsigned int>(p.N) - 1u) {
      __threadfence();
      double sum = 0.0;
      for (int i = 0; i < p.N; ++i) sum += __ldcg(p.image_loss + i);      // fixed order -> deterministic
      *p.loss = static_cast<float>(sum * static_cast<double>(p.inv_n_global));
      *p.ticket = 0u;
    }
  }
  if (p.bulk && p.grad != nullptr && tid == 0) bulk_store_wait();
}

static size_t loss_smem_bytes(int rows_per_cta, int row, int G) {
  const size_t slab = (static_cast<size_t>(rows_per_cta) * row * sizeof(float) + 15) & ~static_cast<size_t>(15);
  const size_t gt = (static_cast<size_t>(G) * sizeof(GtRec) + 15) & ~static_cast<size_t>(15);
  return slab + gt + sizeof(LossShared);
}

static int rows_per_cta_for(int P) {
  const int r = (P + kCluster - 1) / kCluster;
  return (r + 3) & ~3;
}

template <int kC>
static int launch_loss(const LossParams& p, size_t smem, cudaStream_t st) {
  static bool configured = false;     // per instantiation; attribute is sticky per device context
  static size_t configured_smem = 0;
  if (!configured || smem > configured_smem) {
    cudaError_t e = cudaFuncSetAttribute(multibox_loss_kernel<kC>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(227 * 1024));
    if (e != cudaSuccess) { set_error("ssdh_multibox_loss: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return static_cast<int>(e); }
    configured = true;
    configured_smem = 227 * 1024;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(static_cast<unsigned>(p.N) * kCluster);
  cfg.blockDim = dim3(kLossThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kCluster;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, multibox_loss_kernel<kC>, p);
  if (e != cudaSuccess) { set_error("ssdh_multibox_loss: launch: %s", cudaGetErrorString(e)); (void)cudaGetLastError(); return static_cast<int>(e); }
  return 0;
}

}  // namespace ssdh

using namespace ssdh;

extern "C" size_t ssdh_multibox_loss_workspace_bytes(int N, int P, int C, int G) {
  (void)P; (void)C; (void)G;
  return 16 + static_cast<size_t>(N > 0 ? N : 0) * sizeof(double);
}

extern "C" int ssdh_multibox_loss(const float* outputs, const float* targets, const float* priors, int N, int P, int C, int G,
                                  float a, float thr, int n_global, float* loss, float* grad, ssdh_image_stats* stats,
                                  void* ws, size_t ws_bytes, ssdh_stream_t stream) {
  if (!outputs || !priors || !loss || N <= 0 || P <= 0 || C <= 0 || G < 0 || n_global <= 0 || (G > 0 && !targets)) {
    set_error("ssdh_multibox_loss: NULL pointer or non-positive dimension");
    return SSDH_E_ARG;
  }
  if (C > kMaxClasses || G > kMaxGT || P > 65535) {
    set_error("ssdh_multibox_loss: limits are C <= %d, G <= %d, P <= 65535 (got C=%d G=%d P=%d)", kMaxClasses, kMaxGT, C, G, P);
    return SSDH_E_LIMIT;
  }
  if (!ws || ws_bytes < ssdh_multibox_loss_workspace_bytes(N, P, C, G)) { set_error("ssdh_multibox_loss: workspace too small"); return SSDH_E_WORKSPACE; }
  if (!aligned16(priors) || !aligned16(ws)) { set_error("ssdh_multibox_loss: priors and ws must be 16-byte aligned"); return SSDH_E_ALIGN; }
  const int row = 4 + C;
  const int rpc = rows_per_cta_for(P);
  if (rpc > kSlots * kLossThreads) {
    set_error("ssdh_multibox_loss: P=%d needs %d rows per CTA, limit %d", P, rpc, kSlots * kLossThreads);
    return SSDH_E_LIMIT;
  }
  const size_t smem = loss_smem_bytes(rpc, row, G);
  if (smem > 227 * 1024) { set_error("ssdh_multibox_loss: image slab of %zu bytes per CTA exceeds shared memory", smem); return SSDH_E_LIMIT; }

  LossParams p;
  p.outputs = outputs; p.targets = targets; p.priors = reinterpret_cast<const float4*>(priors);
  p.N = N; p.P = P; p.C = C; p.G = G;
  p.a = a; p.band = make_band(thr);
  p.inv_n_global = 1.0f / static_cast<float>(n_global);
  p.loss = loss; p.grad = grad; p.stats = stats;
  p.ticket = reinterpret_cast<unsigned int*>(ws);
  p.image_loss = reinterpret_cast<double*>(static_cast<unsigned char*>(ws) + 16);
  p.rows_per_cta = rpc;
  // TMA bulk copies need 16-byte aligned addresses and sizes for every (image, CTA, slot) chunk.
  const bool sizes_ok = (static_cast<long long>(P) * row) % 4 == 0;     // rpc and the slot size are multiples of 4 rows
  p.bulk = sizes_ok && aligned16(outputs) && (grad == nullptr || aligned16(grad));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (C == 21) return launch_loss<21>(p, smem, st);
  return launch_loss<0>(p, smem, st);
}

extern "C" int ssdh_device_info(int* sm_count, int* max_smem_optin, int* loss_cluster_size, int* loss_max_active_clusters) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) { set_error("ssdh_device_info: %s", cudaGetErrorString(e)); return static_cast<int>(e); }
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties(&prop, dev);
  if (e != cudaSuccess) { set_error("ssdh_device_info: %s", cudaGetErrorString(e)); return static_cast<int>(e); }
  if (sm_count) *sm_count = prop.multiProcessorCount;
  if (max_smem_optin) *max_smem_optin = static_cast<int>(prop.sharedMemPerBlockOptin);
  if (loss_cluster_size) *loss_cluster_size = kCluster;
  if (loss_max_active_clusters) {
    cudaFuncSetAttribute(multibox_loss_kernel<21>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(kCluster * 64);
    cfg.blockDim = dim3(kLossThreads);
    cfg.dynamicSmemBytes = loss_smem_bytes(rows_per_cta_for(SSDH_NUM_PRIORS), 25, 20);
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = kCluster; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    int nc = 0;
    e = cudaOccupancyMaxActiveClusters(&nc, multibox_loss_kernel<21>, &cfg);
    if (e != cudaSuccess) { set_error("ssdh_device_info: occupancy: %s", cudaGetErrorString(e)); (void)cudaGetLastError(); return static_cast<int>(e); }
    *loss_max_active_clusters = nc;
  }
  return 0;
}
