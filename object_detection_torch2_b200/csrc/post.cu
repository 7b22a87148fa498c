// Inference-time post-processing: decode (I1), score (I2), pairwise IoU (I3), greedy NMS (I4) and the fused
// decode+score+NMS pass.  Replaces calc_coordicate / calc_score / calc_iou / non_maximum_suppression
// (reference src/utils.py:19-116) as they are chained at src/evaluate.py:129-131 and src/inference.py:67-69.
//
// NMS decomposition: one CTA of 1024 threads per image.
//   A. candidate keys (best non-void score, src/utils.py:99) are compacted in row order,
//   B. block-wide LSD radix sort (8-bit digits, warp match_any ranking) -> descending score, ties by lower row
//      (the stable order; the reference's torch.sort is unstable on ties),
//   C. tiled greedy suppression: every candidate of a 1024-wide tile is tested against the kept boxes held
//      in shared memory (broadcast reads), then the tile is resolved warp by warp with ballots; IoU uses the
//      same fp32 operation sequence as src/utils.py:74-77 so keep lists are bit-identical,
//   D. score columns of rows that were not kept are zeroed (src/utils.py:114).
#include <algorithm>
#include <stdlib.h>

#include <cooperative_groups.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace ssdh {

// ------------------------------------------------------------------------------------------------
// I1 / I2 row math shared by the stand-alone kernels and the fused pass (bitwise identical results)
// ------------------------------------------------------------------------------------------------
// exp(x) as 2^(x * log2 e): one rounded product + ex2.approx (2 ulp).  Relative error <= |x| * 6e-8 + 2.4e-7, i.e. < 2e-6 for
// the |x| <= 20 a head produces -- inside the 1e-5 the decoded extents and scores are held to (north_star), and 5x fewer
// instructions than libdevice's expf, which made this memory-bound pass issue-bound (ncu round 1: 56 % issue-active at 79 %
// of the HBM roofline).  Index decisions (arg-max class, candidate order) never depend on it beyond the score values themselves.
__device__ __forceinline__ float exp_fast(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(__fmul_rn(x, 1.4426950408889634f)));
  return y;
}

__device__ __forceinline__ float4 decode_row(float p0, float p1, float p2, float p3, const float4 d) {
  float4 o;
  o.x = __fadd_rn(__fmul_rn(d.z, p0), d.x);        // src/utils.py:35
  o.y = __fadd_rn(__fmul_rn(d.w, p1), d.y);        // src/utils.py:36
  o.z = __fmul_rn(d.z, exp_fast(p2));              // src/utils.py:37
  o.w = __fmul_rn(d.w, exp_fast(p3));              // src/utils.py:38
  return o;
}

// softmax value at the arg-max class (first max wins) -- the only non-zero of the row, src/utils.py:54-55.
// One routine for the stand-alone and the fused kernels (same operations in the same order: bitwise identical results).
template <typename Load>
__device__ __forceinline__ float best_score(Load logit, int C, int& best) {
  float mx = logit(0);
  best = 0;
#pragma unroll
  for (int c = 1; c < C; ++c) {
    const float v = logit(c);
    if (v > mx) { mx = v; best = c; }
  }
  constexpr float kLog2e = 1.4426950408889634f;
  const float nm = -__fmul_rn(mx, kLog2e);
  float sum = 0.0f;
#pragma unroll
  for (int c = 0; c < C; ++c) {
    float e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(fmaf(logit(c), kLog2e, nm)));
    sum += e;
  }
  return __fdiv_rn(1.0f, sum);
}

__global__ void __launch_bounds__(256)
decode_kernel(const float* __restrict__ pr, int stride, const float4* __restrict__ priors, int P, float4* __restrict__ out, size_t rows) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= rows) return;
  const float* r = pr + i * stride;
  out[i] = decode_row(r[0], r[1], r[2], r[3], priors[i % P]);
}

__global__ void __launch_bounds__(256)
score_kernel(const float* __restrict__ pr, int stride, int C, float* __restrict__ out, size_t rows) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= rows) return;
  const float* r = pr + i * stride + 4;
  int best;
  const float s = best_score([&](int c) { return r[c]; }, C, best);
  float* o = out + i * C;
  for (int c = 0; c < C; ++c) o[c] = (c == best) ? s : 0.0f;
}

// I3  src/utils.py:58-77
__global__ void __launch_bounds__(256)
iou_kernel(const float* __restrict__ t, int t_stride, int T, const float* __restrict__ s, int s_stride, int S,
           float* __restrict__ out, size_t total) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;   // over N*T*S
  if (i >= total) return;
  const int si = static_cast<int>(i % S);
  const size_t nt = i / S;
  const size_t n = nt / T;
  const float* a = t + nt * t_stride;
  const float* b = s + (n * S + si) * s_stride;
  out[i] = iou_value(make_corners(a[0], a[1], a[2], a[3]), make_corners(b[0], b[1], b[2], b[3]));
}

// ------------------------------------------------------------------------------------------------
// Fused pass, kernel 1: decode + score for every row, one read and one write of the slab.
// Writes decoded boxes, ZERO score columns (the NMS kernel scatters the kept scores back) and the
// per-row candidate key / class used by the sort.
// ------------------------------------------------------------------------------------------------
constexpr int kTileRows = 128;

template <int kC>
__global__ void __launch_bounds__(kTileRows)
decode_score_kernel(float* __restrict__ outputs, const float4* __restrict__ priors, int P, int C_rt, size_t total_rows,
                    float* __restrict__ cand_key, uint8_t* __restrict__ cand_cls, int vec_ok) {
  const int C = kC ? kC : C_rt;
  extern __shared__ __align__(16) float tile[];
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");     // let the NMS grid get scheduled early (it waits below)
  const int row = 4 + C;
  const size_t r0 = static_cast<size_t>(blockIdx.x) * kTileRows;
  const int rows = static_cast<int>(min(static_cast<size_t>(kTileRows), total_rows - r0));
  float* base = outputs + r0 * row;
  const int nfl = rows * row;
  if (vec_ok && (nfl & 3) == 0) {
    const float4* b4 = reinterpret_cast<const float4*>(base);
    float4* t4 = reinterpret_cast<float4*>(tile);
    for (int i = threadIdx.x; i < nfl / 4; i += kTileRows) t4[i] = __ldcs(b4 + i);
  } else {
    for (int i = threadIdx.x; i < nfl; i += kTileRows) tile[i] = base[i];
  }
  __syncthreads();
  if (threadIdx.x < rows) {
    float* r = tile + threadIdx.x * row;
    const size_t grow = r0 + threadIdx.x;
    int pi = static_cast<int>(r0 % static_cast<size_t>(P)) + static_cast<int>(threadIdx.x);       // grow % P without a 64-bit division per row
    while (pi >= P) pi -= P;
    const float4 box = decode_row(r[0], r[1], r[2], r[3], priors[pi]);
    int best;
    float s;
    if (kC > 0) {                                    // the VOC head: logits in registers, loops unrolled
      float x[kC > 0 ? kC : 1];
#pragma unroll
      for (int c = 0; c < kC; ++c) x[c] = r[4 + c];
      s = best_score([&](int c) { return x[c]; }, kC, best);
    } else {
      s = best_score([&](int c) { return r[4 + c]; }, C, best);
    }
    r[0] = box.x; r[1] = box.y; r[2] = box.z; r[3] = box.w;
#pragma unroll
    for (int c = 0; c < C; ++c) r[4 + c] = 0.0f;
    cand_key[grow] = best != 0 ? s : 0.0f;          // max over non-void columns, src/utils.py:99
    cand_cls[grow] = static_cast<uint8_t>(best);
  }
  __syncthreads();
  if (vec_ok && (nfl & 3) == 0) {
    float4* b4 = reinterpret_cast<float4*>(base);
    const float4* t4 = reinterpret_cast<const float4*>(tile);
    for (int i = threadIdx.x; i < nfl / 4; i += kTileRows) b4[i] = t4[i];
  } else {
    for (int i = threadIdx.x; i < nfl; i += kTileRows) base[i] = tile[i];
  }
}

// Stand-alone NMS, kernel 1: candidate key / class from already scored rows (general rows: any number of
// positive classes; key = max over columns 5.., class = its arg-max, first max wins).
__global__ void __launch_bounds__(256)
candidate_kernel(const float* __restrict__ outputs, int C, size_t total_rows, float* __restrict__ cand_key,
                 uint8_t* __restrict__ cand_cls) {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const int row = 4 + C;
  const int lane = threadIdx.x & 31;
  const size_t warp_global = (static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const size_t n_warps = (static_cast<size_t>(gridDim.x) * blockDim.x) >> 5;
  // one warp per row: lanes read the row's score columns (coalesced 4*(C-1) bytes), shuffle arg-max
  for (size_t r = warp_global; r < total_rows; r += n_warps) {
    const float* p = outputs + r * row + 5;
    float best = -INFINITY;
    int bc = 0x7fffffff;
    for (int c = lane; c < C - 1; c += 32) {
      const float v = p[c];
      if (v > best) { best = v; bc = c; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, best, o);
      const int oc = __shfl_xor_sync(0xffffffffu, bc, o);
      if (ov > best || (ov == best && oc < bc)) { best = ov; bc = oc; }
    }
    if (lane == 0) {
      cand_key[r] = (C > 1) ? best : 0.0f;
      cand_cls[r] = static_cast<uint8_t>(C > 1 ? bc + 1 : 0);
    }
  }
}

// Branch-free pair test for the NMS inner loops.  Mirrors src/utils.py:74-77 followed by `> thr`:
//   inter = clamp(w) * clamp(h);  value = inter > 0 ? inter / union : inter;  hit = value > thr.
// e = inter - union * thr is a single FMA, so its sign is exact; when |e| clears the 2^-20 band the rounded
// quotient is on the same side of thr with certainty.  Pairs inside the band (or with a degenerate union) report
// `amb` and are settled by the caller with the IEEE division (iou_gt), which keeps keep-lists bit-identical.
struct PairThr {
  float thr, nthr, eps;
  bool usable;
};
__host__ __device__ inline PairThr make_pair_thr(float thr) {
  PairThr t;
  t.thr = thr; t.nthr = -thr; t.eps = thr * 9.5367431640625e-07f;
  t.usable = thr >= 1e-6f && thr <= 1e6f;
  return t;
}
__device__ __forceinline__ bool pair_hit_fast(const float4 a, float a_area, const float4 b, float b_area, const PairThr& t, bool& amb) {
  // boxes are (x1, x2, y1, y2)
  const float w = fmaxf(fminf(a.y, b.y) - fmaxf(a.x, b.x), 0.0f);
  const float h = fmaxf(fminf(a.w, b.w) - fmaxf(a.z, b.z), 0.0f);
  const float inter = w * h;
  const float uni = (a_area + b_area) - inter;
  const float e = fmaf(uni, t.nthr, inter), m = uni * t.eps;
  // inter == 0 gives e = -union * thr < -m for any sane union; a non-positive / non-finite union lands in the band
  amb = !(fabsf(e) > m) || !(uni >= 1e-30f);
  return e > m;
}

// ------------------------------------------------------------------------------------------------
// NMS kernel: one CTA per image
// ------------------------------------------------------------------------------------------------
constexpr int kNmsThreads = 1024;
constexpr int kNmsWarps = 32;
constexpr int kMaxRounds = 10;                 // candidates per thread in the sort -> P <= 10240

struct NmsParams {
  float* outputs;
  int N, P, C;
  const float* cand_key;      // [N, P]
  const uint8_t* cand_cls;    // [N, P]
  ThrBand band;
  float score_thr;
  int top_k, per_class, scatter;   // scatter = 1: fused pass (write kept scores), 0: zero rows not kept
  int32_t* order;             // [N, P] (workspace or user)
  int32_t* order_cnt;         // [N] or NULL
  int32_t* keep;              // [N, P] (workspace or user)
  int32_t* keep_cnt;          // [N] or NULL
  unsigned long long* trace;  // debug: [N][16] SM clock stamps, NULL in production
};

constexpr int kTileC = 256;                        // candidates settled per round of the greedy suppression
constexpr int kSplit = kNmsThreads / kTileC;       // threads sharing one candidate in the sweep over the kept list
constexpr int kTileWords = kTileC / 32;
constexpr int kNmsCluster = 4;                    // CTAs (SMs) sharing a dense image: the kept list is dealt round-robin to them

struct NmsTile {                                   // the round's surviving candidates, in score order
  float4 box[kTileC];                              // x1, x2, y1, y2
  float area[kTileC];
  int row[kTileC];
  uint32_t ov_in[kTileC][kTileWords];              // ov_in[c] bit r (r < c): the earlier (higher-score) survivor r overlaps c above the threshold
  uint8_t cls[kTileC];
  uint8_t alive[kTileC];
  uint32_t alive_w[kTileWords], keep_w[kTileWords], supp_w[kTileWords];
  int undecided[2];
  uint32_t alive_x[2][kNmsCluster][kTileWords];    // per round parity: every CTA's verdict on the round's candidates (written remotely)
};

struct NmsShared {
  int warp_total[kNmsWarps];
  int warp_alive[kNmsWarps];
  int part[4][256];
  int digit_base[256];
  int n_cand, kept, uniform_digit, stop, odd_kept;
  uint32_t todo[64];                 // which of this cluster's images (cluster_id + k * n_clusters) are left to this kernel
};

// Debug stamps (ssdh_debug_set_nms_trace): [image][16] SM clocks.
#define NMS_TRACE(idx) do { if (p.trace != nullptr && threadIdx.x == 0 && trace_on) p.trace[trace_slot * 16 + (idx)] = clock64(); } while (0)
#define NMS_ACC(var, t0) do { if (p.trace != nullptr && threadIdx.x == 0 && trace_on) { const long long now_ = clock64(); var += now_ - (t0); t0 = now_; } } while (0)

// kCl = 1: one CTA per image.  kCl = kNmsCluster: a cluster of CTAs per image; every CTA runs the same rounds on the same
// candidates but sweeps only ITS share of the kept list (entry g lives in CTA g % kCl), the verdicts are AND-ed through
// distributed shared memory once per round, and the (cheap) matrix / walk steps run redundantly so that nothing else has to
// be exchanged.  CTA 0 owns every global write.
template <bool kPerClass, int kCl>
__global__ void __launch_bounds__(kNmsThreads, 1) nms_kernel(const NmsParams p, const int32_t* __restrict__ large) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int P = p.P, row = 4 + p.C;
  const int rank = kCl > 1 ? static_cast<int>(cg::this_cluster().block_rank()) : 0;
  const int cluster_id = blockIdx.x / kCl, n_clusters = gridDim.x / kCl;
  const bool trace_on = rank == 0;
  size_t trace_slot = 0;
  asm volatile("griddepcontrol.wait;" ::: "memory");
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t lt_mask = (1u << lane) - 1u;

  // layout: [NmsShared][keep bitmap: P bits][union{ sort buffers | kept list }]
  NmsShared& sh = *reinterpret_cast<NmsShared*>(smem_raw);
  size_t off = (sizeof(NmsShared) + 15) & ~static_cast<size_t>(15);
  uint32_t* keep_bits = reinterpret_cast<uint32_t*>(smem_raw + off);
  const int bit_words = (P + 31) / 32;
  off += (static_cast<size_t>(bit_words) * 4 + 15) & ~static_cast<size_t>(15);
  unsigned char* region = smem_raw + off;
  const size_t Pp = (static_cast<size_t>(P) + 15) & ~static_cast<size_t>(15);
  // sort view
  uint32_t* keys_a = reinterpret_cast<uint32_t*>(region);
  uint32_t* keys_b = keys_a + Pp;
  uint16_t* idx_a = reinterpret_cast<uint16_t*>(keys_b + Pp);
  uint16_t* idx_b = idx_a + Pp;
  uint16_t* whist = idx_b + Pp;                 // [32 warps][256 digits]
  // kept-list view (aliases the sort buffers, used after the order has been written to global memory)
  float4* k_box = reinterpret_cast<float4*>(region);          // x1, x2, y1, y2
  float* k_area = reinterpret_cast<float*>(k_box + Pp);
  uint8_t* k_cls = reinterpret_cast<uint8_t*>(k_area + Pp);
  const size_t sort_bytes = Pp * (4 + 4 + 2 + 2) + kNmsWarps * 256 * 2, kept_bytes = Pp * (16 + 4 + 1);
  NmsTile& tile = *reinterpret_cast<NmsTile*>(region + (((sort_bytes > kept_bytes ? sort_bytes : kept_bytes) + 15) & ~static_cast<size_t>(15)));

  // The grid is a few dozen clusters, each walking over the images cluster_id, cluster_id + n_clusters, ...; the flags of
  // those images (1 = too many candidates for nms_small_kernel, left to this kernel) are fetched in one parallel pass.
  const int mine = p.N > cluster_id ? (p.N - cluster_id + n_clusters - 1) / n_clusters : 0;      // <= 2048 (host check)
  for (int k0 = 0; k0 < mine; k0 += kNmsThreads) {
    const int k = k0 + tid;
    const bool todo = k < mine && (large == nullptr || large[cluster_id + k * n_clusters] != 0);
    const uint32_t word = __ballot_sync(0xffffffffu, todo);
    if (lane == 0) sh.todo[(k0 >> 5) + warp] = word;
  }
  __syncthreads();
  for (int k = 0; k < mine; ++k) {
  if (!((sh.todo[k >> 5] >> (k & 31)) & 1u)) continue;      // uniform for the whole cluster
  const int n = cluster_id + k * n_clusters;
  trace_slot = static_cast<size_t>(n);
  const float* key_in = p.cand_key + static_cast<size_t>(n) * P;
  const uint8_t* cls_in = p.cand_cls + static_cast<size_t>(n) * P;
  int32_t* order = p.order + static_cast<size_t>(n) * P;
  int32_t* keep = p.keep + static_cast<size_t>(n) * P;
  float* img = p.outputs + static_cast<size_t>(n) * P * row;

  __syncthreads();                                          // the previous image of this CTA is finished with shared memory
  if (tid == 0) { sh.n_cand = 0; sh.kept = 0; sh.stop = 0; sh.odd_kept = 0; }
  NMS_TRACE(0);
  for (int i = tid; i < bit_words; i += kNmsThreads) keep_bits[i] = 0u;
  __syncthreads();

  // ---- A. compaction in row order ---------------------------------------------------------------------
  for (int base = 0; base < P; base += kNmsThreads) {
    const int r = base + tid;
    float key = 0.0f;
    if (r < P) key = key_in[r];
    const bool cand = (r < P) && (key > p.score_thr);
    const uint32_t ballot = __ballot_sync(0xffffffffu, cand);
    if (lane == 0) sh.warp_total[warp] = __popc(ballot);
    __syncthreads();
    int before = sh.n_cand;
    for (int w = 0; w < warp; ++w) before += sh.warp_total[w];
    if (cand) {
      const int pos = before + __popc(ballot & lt_mask);
      keys_a[pos] = ~float_key(key);              // ascending sort of ~key = descending score
      idx_a[pos] = static_cast<uint16_t>(r);
    }
    __syncthreads();
    if (tid == 0) {
      int t = 0;
      for (int w = 0; w < kNmsWarps; ++w) t += sh.warp_total[w];
      sh.n_cand += t;
    }
    __syncthreads();
  }
  const int K = sh.n_cand;
  NMS_TRACE(1);

  // ---- B. stable LSD radix sort of (key, idx) -------------------------------------------------------------
  {
    const int seg = ((K + kNmsWarps * 32 - 1) / (kNmsWarps * 32)) * 32;     // per-warp segment, multiple of 32
    const int rounds = seg / 32;                                              // <= kMaxRounds
    uint32_t* kin = keys_a; uint32_t* kout = keys_b;
    uint16_t* iin = idx_a;  uint16_t* iout = idx_b;
    for (int pass = 0; pass < 4 && K > 1; ++pass) {
      const int shift = 8 * pass;
      for (int i = tid; i < kNmsWarps * 256 / 2; i += kNmsThreads) reinterpret_cast<uint32_t*>(whist)[i] = 0u;
      __syncthreads();
      int loc[kMaxRounds];
      uint16_t* my_hist = whist + warp * 256;
#pragma unroll
      for (int r = 0; r < kMaxRounds; ++r) {
        loc[r] = 0;
        if (r < rounds) {
          const int i = warp * seg + r * 32 + lane;
          const bool act = i < K;
          const uint32_t d = act ? ((kin[i] >> shift) & 255u) : 0xffffffffu;
          const uint32_t peers = __match_any_sync(0xffffffffu, d);
          int prev = 0;
          if (act) prev = my_hist[d];
          __syncwarp();
          if (act && lane == __ffs(peers) - 1) my_hist[d] = static_cast<uint16_t>(prev + __popc(peers));
          __syncwarp();
          loc[r] = prev + __popc(peers & lt_mask);
        }
      }
      __syncthreads();
      // exclusive scan over (digit major, warp minor)
      {
        const int d = tid & 255, q = tid >> 8;      // q handles warps [8q, 8q+8)
        int s = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) s += whist[(8 * q + w) * 256 + d];
        sh.part[q][d] = s;
      }
      __syncthreads();
      if (warp == 0) {
        int t[8], s = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int d = lane * 8 + j;
          t[j] = sh.part[0][d] + sh.part[1][d] + sh.part[2][d] + sh.part[3][d];
          s += t[j];
        }
        const int incl = warp_incl_scan(s, lane);
        int run = incl - s;
        bool uni = false;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          sh.digit_base[lane * 8 + j] = run;
          run += t[j];
          uni |= (t[j] == K);
        }
        const uint32_t any_uni = __ballot_sync(0xffffffffu, uni);
        if (lane == 0) sh.uniform_digit = any_uni != 0u;
      }
      __syncthreads();
      if (sh.uniform_digit) { __syncthreads(); continue; }     // every key has the same digit: nothing moves
      {
        const int d = tid & 255, q = tid >> 8;
        int run = sh.digit_base[d];
        for (int qq = 0; qq < q; ++qq) run += sh.part[qq][d];
#pragma unroll
        for (int w = 0; w < 8; ++w) {
          const int c = whist[(8 * q + w) * 256 + d];
          whist[(8 * q + w) * 256 + d] = static_cast<uint16_t>(run);
          run += c;
        }
      }
      __syncthreads();
#pragma unroll
      for (int r = 0; r < kMaxRounds; ++r) {
        if (r < rounds) {
          const int i = warp * seg + r * 32 + lane;
          if (i < K) {
            const uint32_t key = kin[i];
            const int pos = my_hist[(key >> shift) & 255u] + loc[r];
            kout[pos] = key;
            iout[pos] = iin[i];
          }
        }
      }
      __syncthreads();
      uint32_t* tk = kin; kin = kout; kout = tk;
      uint16_t* ti = iin; iin = iout; iout = ti;
    }
    // the sorted order leaves shared memory here; the kept list takes over the region
    for (int i = tid; i < K; i += kNmsThreads) order[i] = iin[i];
    __syncthreads();
  }

  NMS_TRACE(2);
  long long t_sweep = 0, t_matrix = 0, t_walk = 0, t_mark = p.trace != nullptr ? clock64() : 0;
  // ---- C. greedy suppression -----------------------------------------------------------------------------
  const ThrBand band = p.band;
  const PairThr pt = make_pair_thr(band.thr);
  const int limit = p.top_k > 0 ? p.top_k : 0x7fffffff;
  // pair test: fast band decision, the (rare) borderline pair is settled with the IEEE division
  auto suppresses = [&](const float4 kb, float ka, const float4 mb, float ma) -> bool {
    bool amb;
    bool hit = pair_hit_fast(kb, ka, mb, ma, pt, amb);
    if (amb || !pt.usable) {
      Corners o, m;
      o.x1 = kb.x; o.x2 = kb.y; o.y1 = kb.z; o.y2 = kb.w; o.area = ka;
      m.x1 = mb.x; m.x2 = mb.y; m.y1 = mb.z; m.y2 = mb.w; m.area = ma;
      hit = iou_gt(o, m, band);
    }
    return hit;
  };
  // Rounds of kTileC candidates in score order.  (1) every candidate is checked against everything kept so far
  // (kSplit threads share the sweep over the kept list, broadcast reads); (2) the survivors are compacted and their
  // mutual overlaps go into a bit matrix, warp per row, one ballot per 32 columns; (3) one warp walks the rows in score
  // order, OR-ing the masks of the rows it keeps into a register-resident `removed` set -- the sequential greedy rule
  // (src/utils.py:102-108) at a few instructions per candidate; (4) the kept rows join the kept list.
  for (int base = 0; base < K; base += kTileC) {
    if (sh.stop) break;
    const int c = tid / kSplit, part = tid % kSplit;
    const int i = base + c;
    bool alive = i < K;
    int my_row = 0, my_cls = 0;
    Corners me = {0.f, 0.f, 0.f, 0.f, 0.f};
    if (alive) {
      my_row = order[i];
      const float* b = img + static_cast<size_t>(my_row) * row;
      me = make_corners(b[0], b[1], b[2], b[3]);
      if (kPerClass) my_cls = cls_in[my_row];
    }
    const float4 me4 = make_float4(me.x1, me.x2, me.y1, me.y2);
    const int kept_before = sh.kept;
    const int kept_local = kept_before > rank ? (kept_before - rank + kCl - 1) / kCl : 0;      // my share of the kept list
    {
      // fast sweep: no per-pair branch; the band test needs positive finite areas on both sides (odd boxes are flagged
      // when they enter the kept list) and flags a borderline pair through `slack`; both cases redo the sweep exactly
      const bool tame = pt.usable && sh.odd_kept == 0 && me.area >= 1e-30f && me.area <= 1e30f;
      // Per pair two FMAs with exact signs: hi = inter - union * thr(1 + 2^-20), lo = inter - union * thr(1 - 2^-20).
      // hi > 0: the rounded quotient is certainly above thr (suppressed); lo <= 0: certainly not; in between the pair is
      // settled exactly below.  Both verdicts are accumulated as running maxima (no per-pair predicates), and only ONE of
      // the two extents is clamped at zero: with w clamped, a negative h makes inter <= 0 and both values negative, the same
      // verdict as the reference's doubly clamped product -- 16 instructions per pair instead of 23.
      float acc_hi = -1.0f, acc_lo = -1.0f;
      const float n_hi = -band.hi, n_lo = -band.lo;
      for (int j0 = 0; j0 < kept_local; j0 += 32 * kSplit) {
        if (!__any_sync(0xffffffffu, alive && !(acc_hi > 0.0f))) break;
        const int j1 = min(j0 + 32 * kSplit, kept_local);
        int j = j0 + part;
#pragma unroll 2
        for (; j < j1; j += kSplit) {
          const float4 kb = k_box[j];
          const float wd = fmaxf(fminf(kb.y, me4.y) - fmaxf(kb.x, me4.x), 0.0f);
          const float ht = fminf(kb.w, me4.w) - fmaxf(kb.z, me4.z);
          const float inter = wd * ht;
          const float uni = (k_area[j] + me.area) - inter;
          float hi = fmaf(uni, n_hi, inter), lo = fmaf(uni, n_lo, inter);
          if (kPerClass) { const bool same = k_cls[j] == my_cls; hi = same ? hi : -1.0f; lo = same ? lo : -1.0f; }
          acc_hi = fmaxf(acc_hi, hi);
          acc_lo = fmaxf(acc_lo, lo);
        }
      }
      bool dead = acc_hi > 0.0f;
      const bool maybe = acc_lo > 0.0f;
      const bool redo = alive && (!tame || (maybe && !dead));
      if (__any_sync(0xffffffffu, redo)) {     // rare: settle this candidate with the IEEE division
        if (redo) {
          dead = false;
          for (int j = part; j < kept_local && !dead; j += kSplit) {
            bool hit = suppresses(k_box[j], k_area[j], me4, me.area);
            if (kPerClass) hit = hit && (k_cls[j] == my_cls);
            dead = hit;
          }
        }
      }
      alive = alive && !dead;
    }
#pragma unroll
    for (int o = 1; o < kSplit; o <<= 1) alive = (__shfl_xor_sync(0xffffffffu, alive ? 1 : 0, o) != 0) && alive;
    if (part == 0) tile.alive[c] = alive ? 1 : 0;
    __syncthreads();
    if (warp < kTileWords) {
      const uint32_t w = __ballot_sync(0xffffffffu, tile.alive[32 * warp + lane] != 0);
      if (lane == 0) tile.alive_w[warp] = w;
    }
    __syncthreads();
    if (kCl > 1) {
      // AND of the verdicts of all CTAs: everyone stores its words into everyone's buffer, one cluster barrier
      cg::cluster_group cluster = cg::this_cluster();
      const int par = (base / kTileC) & 1;
      if (tid < kCl * kTileWords) {
        const int dst = tid / kTileWords, w = tid % kTileWords;
        cluster.map_shared_rank(&tile.alive_x[par][rank][w], dst)[0] = tile.alive_w[w];
      }
      cluster.sync();
      if (tid < kTileWords) {
        uint32_t a = 0xffffffffu;
#pragma unroll
        for (int q = 0; q < kCl; ++q) a &= tile.alive_x[par][q][tid];
        tile.alive_w[tid] = a;
      }
      __syncthreads();
      alive = alive && ((tile.alive_w[c >> 5] >> (c & 31)) & 1u) != 0u;
    }
    int M = 0, my_idx = 0;
#pragma unroll
    for (int w = 0; w < kTileWords; ++w) {
      const uint32_t aw = tile.alive_w[w];
      if (w < (c >> 5)) my_idx += __popc(aw);
      else if (w == (c >> 5)) my_idx += __popc(aw & ((1u << (c & 31)) - 1u));
      M += __popc(aw);
    }
    if (alive && part == 0) {
      tile.box[my_idx] = me4;
      tile.area[my_idx] = me.area;
      tile.cls[my_idx] = static_cast<uint8_t>(my_cls);
      tile.row[my_idx] = my_row;
    }
    __syncthreads();
    NMS_ACC(t_sweep, t_mark);
    const int Mw = (M + 31) >> 5;
    // (2) mutual overlaps of the survivors, recorded at the SUPPRESSED side (sparse: a few hits per row): warp per earlier
    // row r, lane per later column j
    for (int i = tid; i < M * kTileWords; i += kNmsThreads) (&tile.ov_in[0][0])[i] = 0u;
    if (tid < kTileWords) { tile.keep_w[tid] = 0u; tile.supp_w[tid] = 0u; }
    if (tid < 2) tile.undecided[tid] = 0;
    __syncthreads();
    for (int r = warp; r < M; r += kNmsWarps) {
      const float4 bi = tile.box[r];
      const float ai = tile.area[r];
      const int ci = tile.cls[r];
      const uint32_t rbit = 1u << (r & 31);
      for (int w = r >> 5; w < Mw; ++w) {
        const int j = 32 * w + lane;
        if (j > r && j < M) {
          const float4 bj = tile.box[j];
          const float aj = tile.area[j];
          const float wd = fmaxf(fminf(bi.y, bj.y) - fmaxf(bi.x, bj.x), 0.0f);
          const float ht = fminf(bi.w, bj.w) - fmaxf(bi.z, bj.z);
          const float inter = wd * ht;
          const float uni = (ai + aj) - inter;
          bool hit = fmaf(uni, -band.hi, inter) > 0.0f;                      // certain (see the sweep above)
          const bool sure = hit || !(fmaf(uni, -band.lo, inter) > 0.0f);
          const bool tame_pair = pt.usable && ai >= 1e-30f && ai <= 1e30f && aj >= 1e-30f && aj <= 1e30f;
          if (!sure || !tame_pair) hit = suppresses(bi, ai, bj, aj);        // rare: band or odd boxes -> IEEE division
          if (kPerClass) hit = hit && (tile.cls[j] == ci);
          if (hit) atomicOr(&tile.ov_in[j][r >> 5], rbit);
        }
      }
    }
    __syncthreads();
    NMS_ACC(t_matrix, t_mark);
    // (3) the sequential greedy rule (src/utils.py:102-108) as a parallel fixed point over the tile: survivor j is
    //   suppressed as soon as an earlier overlapping survivor is known kept,
    //   kept       as soon as every earlier overlapping survivor is known suppressed.
    // The first undecided survivor in score order always resolves, so every round makes progress and the fixed point is
    // exactly the sequential result; tiles settle in a handful of rounds instead of a 256-step walk on one warp.
    {
      int state = tid < M ? 0 : 2;             // 0 undecided, 1 kept, 2 suppressed / not a survivor
      uint32_t in_w[kTileWords];
#pragma unroll
      for (int w = 0; w < kTileWords; ++w) in_w[w] = (tid < M && w < Mw) ? tile.ov_in[tid][w] : 0u;
      for (int round = 0; round <= M; ++round) {
        if (state == 0) {
          bool any_kept = false, all_supp = true;
#pragma unroll
          for (int w = 0; w < kTileWords; ++w) {
            any_kept |= (in_w[w] & tile.keep_w[w]) != 0u;
            all_supp &= (in_w[w] & ~tile.supp_w[w]) == 0u;
          }
          if (any_kept) state = 2;
          else if (all_supp) state = 1;
        }
        __syncthreads();                        // everyone has read the word arrays (and the previous round's flag)
        if (tid == 0) tile.undecided[(round + 1) & 1] = 0;
        if (tid < M) {
          const uint32_t bit = 1u << (tid & 31);
          if (state == 1 && !(tile.keep_w[tid >> 5] & bit)) atomicOr(&tile.keep_w[tid >> 5], bit);
          if (state == 2 && !(tile.supp_w[tid >> 5] & bit)) atomicOr(&tile.supp_w[tid >> 5], bit);
          if (state == 0) tile.undecided[round & 1] = 1;
        }
        __syncthreads();
        if (!tile.undecided[round & 1]) break;  // block-uniform
      }
    }
    if (warp == 0) {
      // top_k: only the first (limit - kept so far) kept survivors of the tile stay; totals for the next round
      int cnt = lane < Mw ? __popc(tile.keep_w[lane]) : 0;
      const int incl = warp_incl_scan(cnt, lane);
      const int room = limit - kept_before;                       // > 0 here (the loop stops once the limit is reached)
      const int before = incl - cnt;
      if (lane < Mw && incl > room) {
        uint32_t bits = tile.keep_w[lane], out = 0u;
        int left = room - before;
        while (bits && left > 0) { const uint32_t b = bits & (0u - bits); out |= b; bits ^= b; --left; }
        tile.keep_w[lane] = out;
      }
      const int total_tile = __shfl_sync(0xffffffffu, incl, 31);
      if (lane == 0) {
        const int total = kept_before + min(total_tile, room);
        sh.kept = total;
        if (total >= limit) sh.stop = 1;
      }
    }
    __syncthreads();
    if (tid < M) {
      const uint32_t kw = tile.keep_w[tid >> 5];
      if ((kw >> (tid & 31)) & 1u) {
        int pos = kept_before + __popc(kw & ((1u << (tid & 31)) - 1u));
        for (int w = 0; w < (tid >> 5); ++w) pos += __popc(tile.keep_w[w]);
        const int r = tile.row[tid];
        if (pos % kCl == rank) {                  // my share of the kept list
          const int lp = pos / kCl;
          k_box[lp] = tile.box[tid];
          k_area[lp] = tile.area[tid];
          if (!(tile.area[tid] >= 1e-30f && tile.area[tid] <= 1e30f)) sh.odd_kept = 1;
          k_cls[lp] = tile.cls[tid];
        }
        if (rank == 0) {
          keep[pos] = r;
          atomicOr(&keep_bits[r >> 5], 1u << (r & 31));
        }
      }
    }
    __syncthreads();
    NMS_ACC(t_walk, t_mark);
  }
  __syncthreads();
  if (kCl > 1) cg::this_cluster().sync();         // no CTA moves on while a peer may still store into its buffers
  if (rank != 0) continue;                        // CTA 0 owns the global writes of the image
  const int kept = sh.kept;
  NMS_TRACE(3);
  if (p.trace != nullptr && tid == 0) {
    p.trace[trace_slot * 16 + 9] = t_sweep;
    p.trace[trace_slot * 16 + 10] = t_matrix;
    p.trace[trace_slot * 16 + 11] = t_walk;
  }
  if (tid == 0) {
    if (p.keep_cnt) p.keep_cnt[n] = kept;
    if (p.order_cnt) p.order_cnt[n] = K;
  }

  // ---- D. apply ------------------------------------------------------------------------------------------
  if (p.scatter) {
    for (int j = tid; j < kept; j += kNmsThreads) {
      const int r = keep[j];
      img[static_cast<size_t>(r) * row + 4 + cls_in[r]] = key_in[r];
    }
  } else {
    // outputs[:, :, 4:] *= mask (src/utils.py:114): rows not kept lose every score, void column included
    for (int r = warp; r < P; r += kNmsWarps) {
      if ((keep_bits[r >> 5] >> (r & 31)) & 1u) continue;
      float* dst = img + static_cast<size_t>(r) * row + 4;
      for (int c = lane; c < p.C; c += 32) dst[c] = 0.0f;
    }
  }
  }   // images of this cluster
}


// ------------------------------------------------------------------------------------------------
// NMS, small-K kernel: images with at most kSmallCap candidates (every realistic, trained-like image).
// 512 threads, ~64 KB of shared memory -> three images per SM.  Same results as nms_kernel:
//   compaction in row order -> stable radix sort -> K x K suppression bitmask built in parallel (warp per row,
//   lane per column, ballot per 32 columns) -> one warp walks the rows in score order, OR-ing the masks of
//   the rows it keeps into a register-resident `removed` bitset.
// Images with more candidates are flagged in `large[]` and left to nms_kernel.
// ------------------------------------------------------------------------------------------------
constexpr int kSmallThreads = 512;
constexpr int kSmallWarps = kSmallThreads / 32;
constexpr int kSmallCap = 512;
constexpr int kSmallWords = kSmallCap / 32;
constexpr int kSmallRounds = 18;         // rows per lane in the compaction -> P <= 16 warps * 32 * 18 = 9216

struct SmallShared {
  uint32_t keys[2][kSmallCap];
  uint16_t idx[2][kSmallCap];
  uint16_t whist[kSmallWarps * 256];
  int part[2][256];
  int digit_base[256];
  float4 box[kSmallCap];          // x1, x2, y1, y2 of the sorted candidates
  float area[kSmallCap];
  uint8_t cls[kSmallCap];
  uint32_t ov_in[kSmallCap][kSmallWords];  // ov_in[j] bit i (i < j): an earlier (higher-score) candidate i overlaps j above the threshold
  uint32_t kept_words[kSmallWords], supp_words[kSmallWords];
  int undecided[2];
  int warp_total[kSmallWarps];
  int n_cand, uniform_digit, kept;
};

static unsigned long long* g_nms_trace = nullptr;

template <bool kPerClass>
__global__ void __launch_bounds__(kSmallThreads, 3) nms_small_kernel(const NmsParams p, int32_t* __restrict__ large) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SmallShared& sh = *reinterpret_cast<SmallShared*>(smem_raw);
  const bool trace_on = true;
  const size_t trace_slot = blockIdx.x;
  uint32_t* keep_bits = reinterpret_cast<uint32_t*>(smem_raw + ((sizeof(SmallShared) + 15) & ~static_cast<size_t>(15)));   // P bits
  const int n = blockIdx.x, P = p.P, row = 4 + p.C;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t lt_mask = (1u << lane) - 1u;
  const float* key_in = p.cand_key + static_cast<size_t>(n) * P;
  const uint8_t* cls_in = p.cand_cls + static_cast<size_t>(n) * P;
  int32_t* order = p.order + static_cast<size_t>(n) * P;
  int32_t* keep = p.keep + static_cast<size_t>(n) * P;
  float* img = p.outputs + static_cast<size_t>(n) * P * row;

  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");                  // candidate keys / decoded boxes of the previous grid are complete
  NMS_TRACE(0);
  // ---- A. candidates in row order: warp w owns a contiguous chunk of rows; all of a lane's loads are issued before
  // the first ballot so the chunk costs one memory latency, not one per 32 rows -----------------------------------------
  const int chunk = ((P + kSmallWarps * 32 - 1) / (kSmallWarps * 32)) * 32;        // rows per warp, multiple of 32
  const int c0 = warp * chunk, c1 = min(P, c0 + chunk);
  float kv[kSmallRounds];
#pragma unroll
  for (int q = 0; q < kSmallRounds; ++q) {
    const int r = c0 + 32 * q + lane;
    kv[q] = (32 * q < chunk && r < c1) ? key_in[r] : 0.0f;
  }
  uint32_t cand_bits[kSmallRounds];
  int mine = 0;                                    // warp-uniform count of candidates in my chunk
#pragma unroll
  for (int q = 0; q < kSmallRounds; ++q) {
    cand_bits[q] = __ballot_sync(0xffffffffu, kv[q] > p.score_thr);
    mine += __popc(cand_bits[q]);
  }
  if (lane == 0) sh.warp_total[warp] = mine;
  NMS_TRACE(12);
  for (int i = tid; i < kSmallCap * kSmallWords; i += kSmallThreads) (&sh.ov_in[0][0])[i] = 0u;
  if (tid < kSmallWords) { sh.kept_words[tid] = 0u; sh.supp_words[tid] = 0u; }
  if (tid < 2) sh.undecided[tid] = 0;
  NMS_TRACE(13);
  __syncthreads();
  NMS_TRACE(14);
  int before = 0, total = 0;
  for (int w = 0; w < kSmallWarps; ++w) {
    const int t = sh.warp_total[w];
    before += w < warp ? t : 0;
    total += t;
  }
  const int K = total;
  if (K > kSmallCap) {                       // uniform: leave this image to the general kernel
    if (tid == 0) large[n] = 1;
    return;
  }
  if (tid == 0) large[n] = 0;
#pragma unroll
  for (int q = 0; q < kSmallRounds; ++q) {
    if ((cand_bits[q] >> lane) & 1u) {
      const int pos = before + __popc(cand_bits[q] & lt_mask);
      sh.keys[0][pos] = ~float_key(kv[q]);
      sh.idx[0][pos] = static_cast<uint16_t>(c0 + 32 * q + lane);
    }
    before += __popc(cand_bits[q]);
  }
  if (!p.scatter)
    for (int i = tid; i < (P + 31) / 32; i += kSmallThreads) keep_bits[i] = 0u;
  __syncthreads();

  NMS_TRACE(1);
  // ---- B. stable LSD radix sort, one key per thread -------------------------------------------------------------
  int cur = 0;
  for (int pass = 0; pass < 4 && K > 1; ++pass) {
    const int shift = 8 * pass;
    for (int i = tid; i < kSmallWarps * 256 / 2; i += kSmallThreads) reinterpret_cast<uint32_t*>(sh.whist)[i] = 0u;
    __syncthreads();
    uint16_t* my_hist = sh.whist + warp * 256;
    const bool act = tid < K;
    const uint32_t key = act ? sh.keys[cur][tid] : 0u;
    const uint32_t d = act ? ((key >> shift) & 255u) : 0xffffffffu;
    const uint32_t peers = __match_any_sync(0xffffffffu, d);
    if (act && lane == __ffs(peers) - 1) my_hist[d] = static_cast<uint16_t>(__popc(peers));
    const int loc = __popc(peers & lt_mask);
    __syncthreads();
    {
      const int dd = tid & 255, q = tid >> 8;          // q handles warps [8q, 8q + 8)
      int sum = 0;
#pragma unroll
      for (int w = 0; w < 8; ++w) sum += sh.whist[(8 * q + w) * 256 + dd];
      sh.part[q][dd] = sum;
    }
    __syncthreads();
    if (warp == 0) {
      int t[8], sum = 0;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        t[j] = sh.part[0][lane * 8 + j] + sh.part[1][lane * 8 + j];
        sum += t[j];
      }
      const int inc = warp_incl_scan(sum, lane);
      int run = inc - sum;
      bool uni = false;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        sh.digit_base[lane * 8 + j] = run;
        run += t[j];
        uni |= (t[j] == K);
      }
      const uint32_t any_uni = __ballot_sync(0xffffffffu, uni);
      if (lane == 0) sh.uniform_digit = any_uni != 0u;
    }
    __syncthreads();
    if (sh.uniform_digit) continue;                     // block-uniform; rewritten only two barriers later
    {
      const int dd = tid & 255, q = tid >> 8;
      int run = sh.digit_base[dd] + (q ? sh.part[0][dd] : 0);
#pragma unroll
      for (int w = 0; w < 8; ++w) {
        const int c = sh.whist[(8 * q + w) * 256 + dd];
        sh.whist[(8 * q + w) * 256 + dd] = static_cast<uint16_t>(run);
        run += c;
      }
    }
    __syncthreads();
    if (act) {
      const int pos = my_hist[d] + loc;
      sh.keys[cur ^ 1][pos] = key;
      sh.idx[cur ^ 1][pos] = sh.idx[cur][tid];
    }
    __syncthreads();
    cur ^= 1;
  }

  NMS_TRACE(2);
  // ---- sorted candidates: order list, boxes ---------------------------------------------------------------------
  bool odd_box = false;            // an area that is not a normal positive number: the whole image takes the exact path
  if (tid < K) {
    const int r = sh.idx[cur][tid];
    order[tid] = r;
    const float* b = img + static_cast<size_t>(r) * row;
    const Corners c = make_corners(b[0], b[1], b[2], b[3]);
    sh.box[tid] = make_float4(c.x1, c.x2, c.y1, c.y2);
    sh.area[tid] = c.area;
    sh.cls[tid] = kPerClass ? cls_in[r] : 0;
    odd_box = !(c.area >= 1e-30f && c.area <= 1e30f);
  } else if (tid < ((K + 31) & ~31)) {
    // padding up to a full word of columns: an empty box far away never overlaps anything (and is never ambiguous)
    sh.box[tid] = make_float4(3e38f, 3e38f, 3e38f, 3e38f);
    sh.area[tid] = 1.0f;
    sh.cls[tid] = -1;
  }
  const bool all_tame = __syncthreads_and(!odd_box) != 0;

  NMS_TRACE(3);
  // ---- C1. overlaps: warp per row i, lane per later column j; the (sparse) hits are recorded at the SUPPRESSED side,
  // ov_in[j] |= bit i, which is the orientation the resolution below reads --------------------------------------------
  const ThrBand band = p.band;
  const PairThr pt = make_pair_thr(band.thr);
  const int Kw = (K + 31) / 32;
  const bool fast_ok = pt.usable && all_tame;      // then union >= max(area) > 0 and only |e| <= m is ambiguous
  for (int i = warp; i < K; i += kSmallWarps) {
    const float4 bi = sh.box[i];
    const float ai = sh.area[i];
    const int ci = sh.cls[i];
    const uint32_t ibit = 1u << (i & 31);
    const int w0 = i >> 5;
    bool row_amb = !fast_ok;
    if (fast_ok) {
      // columns are padded to whole words with inert boxes, so the only bound to respect is j > i in the first word
      float slack = 1.0f;                              // min over pairs of |e| - m; <= 0: some pair sits in the band
      auto pair = [&](int w, bool allowed) {
        const int j = 32 * w + lane;
        const float4 bj = sh.box[j];
        const float wd = fmaxf(fminf(bi.y, bj.y) - fmaxf(bi.x, bj.x), 0.0f);
        const float ht = fmaxf(fminf(bi.w, bj.w) - fmaxf(bi.z, bj.z), 0.0f);
        const float inter = wd * ht;
        const float uni = (ai + sh.area[j]) - inter;
        const float e = fmaf(uni, pt.nthr, inter), m = uni * pt.eps;
        slack = fminf(slack, fabsf(e) - m);
        bool hit = allowed && e > m;
        if (kPerClass) hit = hit && sh.cls[j] == ci;
        if (hit) atomicOr(&sh.ov_in[j][w0], ibit);
      };
      // two words of columns per trip, both computed before either records its hits (independent chains: this loop is
      // latency-bound, a CTA has 16 rows in flight)
      auto pair2 = [&](int w) {
        const int j0 = 32 * w + lane, j1 = j0 + 32;
        const float4 b0 = sh.box[j0], b1 = sh.box[j1];
        const float a0 = sh.area[j0], a1 = sh.area[j1];
        const float wd0 = fmaxf(fminf(bi.y, b0.y) - fmaxf(bi.x, b0.x), 0.0f), wd1 = fmaxf(fminf(bi.y, b1.y) - fmaxf(bi.x, b1.x), 0.0f);
        const float ht0 = fmaxf(fminf(bi.w, b0.w) - fmaxf(bi.z, b0.z), 0.0f), ht1 = fmaxf(fminf(bi.w, b1.w) - fmaxf(bi.z, b1.z), 0.0f);
        const float in0 = wd0 * ht0, in1 = wd1 * ht1;
        const float un0 = (ai + a0) - in0, un1 = (ai + a1) - in1;
        const float e0 = fmaf(un0, pt.nthr, in0), m0 = un0 * pt.eps, e1 = fmaf(un1, pt.nthr, in1), m1 = un1 * pt.eps;
        slack = fminf(slack, fminf(fabsf(e0) - m0, fabsf(e1) - m1));
        bool hit0 = e0 > m0, hit1 = e1 > m1;
        if (kPerClass) { hit0 = hit0 && sh.cls[j0] == ci; hit1 = hit1 && sh.cls[j1] == ci; }
        if (hit0) atomicOr(&sh.ov_in[j0][w0], ibit);
        if (hit1) atomicOr(&sh.ov_in[j1][w0], ibit);
      };
      pair(w0, lane > (i & 31));
      int w = w0 + 1;
      for (; w + 1 < Kw; w += 2) pair2(w);
      if (w < Kw) pair(w, true);
      row_amb = !(slack > 0.0f);
    }
    if (__any_sync(0xffffffffu, row_amb)) {          // rare: redo the row with the exact division (set or clear each bit)
      Corners a;
      a.x1 = bi.x; a.x2 = bi.y; a.y1 = bi.z; a.y2 = bi.w; a.area = ai;
      for (int w = w0; w < Kw; ++w) {
        const int jj = 32 * w + lane;
        if (jj > i && jj < K) {
          const float4 bj = sh.box[jj];
          Corners o;
          o.x1 = bj.x; o.x2 = bj.y; o.y1 = bj.z; o.y2 = bj.w; o.area = sh.area[jj];
          const bool hit = iou_gt(a, o, band) && (sh.cls[jj] == ci);
          if (hit) atomicOr(&sh.ov_in[jj][w0], ibit);
          else atomicAnd(&sh.ov_in[jj][w0], ~ibit);
        }
      }
    }
  }
  __syncthreads();

  NMS_TRACE(4);
  // ---- C2. greedy NMS as a parallel fixed point.  Candidate j (thread j) is
  //   suppressed as soon as an earlier overlapping candidate is known kept,
  //   kept       as soon as every earlier overlapping candidate is known suppressed.
  // The first undecided candidate in score order always resolves, so each round makes progress; the result is exactly
  // the sequential greedy keep set (src/utils.py:102-108).  Typical images settle in a handful of rounds. -----------------
  {
    int state = tid < K ? 0 : 2;             // 0 undecided, 1 kept, 2 suppressed / not a candidate
    uint32_t in_w[kSmallWords];
#pragma unroll
    for (int w = 0; w < kSmallWords; ++w) in_w[w] = (tid < K && w < Kw) ? sh.ov_in[tid][w] : 0u;
    for (int round = 0; round <= K; ++round) {
      if (state == 0) {
        bool any_kept = false, all_supp = true;
#pragma unroll
        for (int w = 0; w < kSmallWords; ++w) {
          any_kept |= (in_w[w] & sh.kept_words[w]) != 0u;
          all_supp &= (in_w[w] & ~sh.supp_words[w]) == 0u;
        }
        if (any_kept) state = 2;
        else if (all_supp) state = 1;
      }
      __syncthreads();                        // everyone has read the word arrays (and the previous round's flag)
      if (tid == 0) sh.undecided[(round + 1) & 1] = 0;
      if (tid < K) {
        const uint32_t bit = 1u << (tid & 31);
        if (state == 1 && !(sh.kept_words[tid >> 5] & bit)) atomicOr(&sh.kept_words[tid >> 5], bit);
        if (state == 2 && !(sh.supp_words[tid >> 5] & bit)) atomicOr(&sh.supp_words[tid >> 5], bit);
        if (state == 0) sh.undecided[round & 1] = 1;
      }
      __syncthreads();
      if (!sh.undecided[round & 1]) { if (p.trace != nullptr && tid == 0) p.trace[static_cast<size_t>(blockIdx.x) * 16 + 8] = round + 1; break; }    // block-uniform
    }
    if (tid == 0) {
      int kept = 0;
      for (int w = 0; w < Kw; ++w) kept += __popc(sh.kept_words[w]);
      sh.kept = kept;
    }
    __syncthreads();
    if (p.top_k > 0 && sh.kept > p.top_k) {   // keep only the first top_k kept candidates (block-uniform branch)
      if (tid == 0) {
        int left = p.top_k;
        for (int w = 0; w < Kw; ++w) {
          uint32_t bits = sh.kept_words[w], out = 0u;
          while (bits && left > 0) { const uint32_t b = bits & (0u - bits); out |= b; bits ^= b; --left; }
          sh.kept_words[w] = out;
        }
        sh.kept = p.top_k;
      }
      __syncthreads();
    }
  }

  NMS_TRACE(5);
  // ---- D. keep list (score order), counts, apply ----------------------------------------------------------------------
  const int kept = sh.kept;
  if (tid < K) {
    const uint32_t wbits = sh.kept_words[tid >> 5];
    if ((wbits >> (tid & 31)) & 1u) {
      int pos = __popc(wbits & ((1u << (tid & 31)) - 1u));
      for (int w = 0; w < (tid >> 5); ++w) pos += __popc(sh.kept_words[w]);
      const int r = sh.idx[cur][tid];
      keep[pos] = r;
      if (p.scatter) img[static_cast<size_t>(r) * row + 4 + cls_in[r]] = key_in[r];
      else atomicOr(&keep_bits[r >> 5], 1u << (r & 31));
    }
  }
  if (tid == 0) {
    if (p.keep_cnt) p.keep_cnt[n] = kept;
    if (p.order_cnt) p.order_cnt[n] = K;
  }
  NMS_TRACE(6);
  if (!p.scatter) {
    __syncthreads();
    for (int r = warp; r < P; r += kSmallWarps) {
      if ((keep_bits[r >> 5] >> (r & 31)) & 1u) continue;
      float* dst = img + static_cast<size_t>(r) * row + 4;
      for (int c = lane; c < p.C; c += 32) dst[c] = 0.0f;
    }
  }
}

// Compact detection lists from the keep lists (one warp per detection slot).
__global__ void __launch_bounds__(256)
gather_detections_kernel(const float* __restrict__ outputs, const int32_t* __restrict__ keep, const int32_t* __restrict__ keep_cnt,
                         int P, int C, int max_det, float* __restrict__ dets, int32_t* __restrict__ det_cnt) {
  const int n = blockIdx.y, row = 4 + C;
  const int lane = threadIdx.x & 31;
  const int slot = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (slot >= max_det) return;
  const int cnt = min(keep_cnt[n], max_det);
  if (slot == 0 && lane == 0) det_cnt[n] = cnt;
  float* out = dets + (static_cast<size_t>(n) * max_det + slot) * 6;
  if (slot >= cnt) {
    if (lane < 6) out[lane] = 0.0f;
    return;
  }
  const int r = keep[static_cast<size_t>(n) * P + slot];
  const float* src = outputs + (static_cast<size_t>(n) * P + r) * row;
  // the kept row carries exactly one positive score among the non-void classes: find it with a ballot
  float best = 0.0f;
  int label = 0;
  for (int c = 1 + lane; c < C; c += 32) {
    const float v = src[4 + c];
    if (v > best) { best = v; label = c; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, best, o);
    const int ol = __shfl_xor_sync(0xffffffffu, label, o);
    if (ov > best || (ov == best && ol != 0 && (label == 0 || ol < label))) { best = ov; label = ol; }
  }
  if (lane < 4) out[lane] = src[lane];
  if (lane == 4) out[4] = best;
  if (lane == 5) out[5] = static_cast<float>(label);
}

static size_t nms_smem_bytes(int P) {
  const size_t Pp = (static_cast<size_t>(P) + 15) & ~static_cast<size_t>(15);
  const size_t head = ((sizeof(NmsShared) + 15) & ~static_cast<size_t>(15)) + ((static_cast<size_t>((P + 31) / 32) * 4 + 15) & ~static_cast<size_t>(15));
  const size_t sort_bytes = Pp * (4 + 4 + 2 + 2) + kNmsWarps * 256 * 2;
  const size_t kept_bytes = Pp * (16 + 4 + 1);
  return head + (((sort_bytes > kept_bytes ? sort_bytes : kept_bytes) + 15) & ~static_cast<size_t>(15)) + sizeof(NmsTile) + 16;
}

struct NmsWorkspace {
  float* cand_key;
  uint8_t* cand_cls;
  int32_t* order;
  int32_t* keep;
  int32_t* large;
};

static size_t nms_ws_layout(int N, int P, void* ws, NmsWorkspace* out) {
  const size_t np = static_cast<size_t>(N) * P;
  size_t off = 0;
  auto take = [&](size_t bytes) { const size_t o = off; off += (bytes + 255) & ~static_cast<size_t>(255); return o; };
  const size_t o_key = take(np * 4), o_ord = take(np * 4), o_keep = take(np * 4), o_cls = take(np), o_large = take(static_cast<size_t>(N) * 4);
  if (out && ws) {
    unsigned char* b = static_cast<unsigned char*>(ws);
    out->cand_key = reinterpret_cast<float*>(b + o_key);
    out->order = reinterpret_cast<int32_t*>(b + o_ord);
    out->keep = reinterpret_cast<int32_t*>(b + o_keep);
    out->cand_cls = b + o_cls;
    out->large = reinterpret_cast<int32_t*>(b + o_large);
  }
  return off;
}

static int run_nms(float* outputs, const float* priors, int N, int P, int C, float iou_thr, float score_thr, int top_k,
                   int per_class, int32_t* order, int32_t* order_cnt, int32_t* keep, int32_t* keep_cnt, void* ws,
                   size_t ws_bytes, cudaStream_t st, bool fused, const char* fn) {
  if (!outputs || N <= 0 || P <= 0 || C <= 0 || (fused && !priors)) { set_error("%s: NULL pointer or non-positive dimension", fn); return SSDH_E_ARG; }
  if (N > 65535) { set_error("%s: N <= 65535 per call", fn); return SSDH_E_LIMIT; }
  if (C > kMaxClasses || P > kMaxRounds * kNmsThreads) { set_error("%s: limits are C <= %d, P <= %d", fn, kMaxClasses, kMaxRounds * kNmsThreads); return SSDH_E_LIMIT; }
  const size_t smem = nms_smem_bytes(P);
  if (smem > 227 * 1024) { set_error("%s: P=%d needs %zu bytes of shared memory", fn, P, smem); return SSDH_E_LIMIT; }
  if (!ws || ws_bytes < nms_ws_layout(N, P, nullptr, nullptr)) { set_error("%s: workspace too small", fn); return SSDH_E_WORKSPACE; }
  if (!aligned16(ws) || (fused && !aligned16(priors))) { set_error("%s: ws/priors must be 16-byte aligned", fn); return SSDH_E_ALIGN; }
  NmsWorkspace w;
  nms_ws_layout(N, P, ws, &w);
  const int row = 4 + C;
  if (int e = ensure_dyn_smem(reinterpret_cast<const void*>(decode_score_kernel<21>), 96 * 1024, fn)) return e;
  if (int e = ensure_dyn_smem(reinterpret_cast<const void*>(decode_score_kernel<0>), 96 * 1024, fn)) return e;
  if (int e = ensure_dyn_smem(reinterpret_cast<const void*>(nms_kernel<false, kNmsCluster>), 227 * 1024, fn)) return e;
  if (int e = ensure_dyn_smem(reinterpret_cast<const void*>(nms_kernel<true, kNmsCluster>), 227 * 1024, fn)) return e;
  if (int e = ensure_dyn_smem(reinterpret_cast<const void*>(nms_small_kernel<false>), 100 * 1024, fn)) return e;
  if (int e = ensure_dyn_smem(reinterpret_cast<const void*>(nms_small_kernel<true>), 100 * 1024, fn)) return e;
  NmsParams p;
  p.outputs = outputs; p.N = N; p.P = P; p.C = C;
  p.cand_key = w.cand_key; p.cand_cls = w.cand_cls;
  p.band = make_band(iou_thr);
  p.score_thr = score_thr; p.top_k = top_k; p.per_class = per_class; p.scatter = fused ? 1 : 0;
  p.order = order ? order : w.order; p.order_cnt = order_cnt;
  p.keep = keep ? keep : w.keep; p.keep_cnt = keep_cnt;
  p.trace = g_nms_trace;
  const size_t small_smem = ((sizeof(SmallShared) + 15) & ~static_cast<size_t>(15)) + (static_cast<size_t>((P + 31) / 32) * 4 + 16);
  const bool use_small = small_smem <= 100 * 1024 && P <= kSmallWarps * 32 * kSmallRounds;

  // Launch helper.  With `pdl` the grid may be scheduled while its predecessor in the stream is still running
  // (programmatic dependent launch); kernels that consume the predecessor's results block in griddepcontrol.wait.
  int cluster_dim = 1;                 // set before launching a clustered kernel
  auto launch = [&](auto kernel, unsigned grid, int threads, size_t dyn_smem, bool pdl, auto... args) -> int {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(threads);
    cfg.dynamicSmemBytes = dyn_smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    int na = 0;
    if (pdl) {
      attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      attr[na].val.programmaticStreamSerializationAllowed = 1;
      ++na;
    }
    if (cluster_dim > 1) {
      attr[na].id = cudaLaunchAttributeClusterDimension;
      attr[na].val.clusterDim.x = static_cast<unsigned>(cluster_dim); attr[na].val.clusterDim.y = 1; attr[na].val.clusterDim.z = 1;
      ++na;
    }
    cfg.attrs = attr;
    cfg.numAttrs = na;
    const cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, args...);
    if (e != cudaSuccess) { set_error("%s: launch: %s", fn, cudaGetErrorString(e)); (void)cudaGetLastError(); return static_cast<int>(e); }
    return 0;
  };

  // The batch goes through in image-aligned parts: (decode+score | candidates)(k) -> nms_small(k).  The first stage of part
  // k+1 does not depend on nms_small(k) (different images), so with programmatic launch the memory-bound pass over part
  // k+1 runs on the same SMs under the compute-bound suppression of part k.  The very first launch is an ordinary one
  // (it depends on whatever produced `outputs`), and so is the final tiled kernel, which reads the flags of every part.
  // Measured on B200 (batch 256): splitting is a loss -- 131 us as one part, 148 / 189 / 222 us in 2 / 3 / 4 parts (round 2, with
  // the lighter decode pass; 140 us in 2 parts when every kernel asks for the same shared-memory carve-out, which in turn
  // costs the single-part chain 2 us): kernels with different carve-outs do not share an SM, so the parts mostly serialise
  // and every extra launch adds its ramp -- the batch goes through as one part.
  const int parts = 1;
  for (int k = 0; k < parts; ++k) {
    const int n0 = static_cast<int>(static_cast<long long>(N) * k / parts), n1 = static_cast<int>(static_cast<long long>(N) * (k + 1) / parts);
    const int nk = n1 - n0;
    if (nk <= 0) continue;
    const size_t rows_k = static_cast<size_t>(nk) * P, off = static_cast<size_t>(n0) * P;
    float* out_k = outputs + off * row;
    if (fused) {
      const unsigned blocks = static_cast<unsigned>((rows_k + kTileRows - 1) / kTileRows);
      const size_t tile_bytes = static_cast<size_t>(kTileRows) * row * sizeof(float);
      const int vec_ok = aligned16(out_k) && ((static_cast<size_t>(kTileRows) * row) % 4 == 0);
      if (int e = C == 21 ? launch(decode_score_kernel<21>, blocks, kTileRows, tile_bytes, k > 0, out_k, reinterpret_cast<const float4*>(priors), P, C, rows_k,
                                   w.cand_key + off, w.cand_cls + off, vec_ok)
                          : launch(decode_score_kernel<0>, blocks, kTileRows, tile_bytes, k > 0, out_k, reinterpret_cast<const float4*>(priors), P, C, rows_k,
                                   w.cand_key + off, w.cand_cls + off, vec_ok)) return e;
    } else {
      const unsigned blocks = static_cast<unsigned>(std::min<size_t>((rows_k + 7) / 8, 148 * 16));
      if (int e = launch(candidate_kernel, blocks, 256, 0, k > 0, static_cast<const float*>(out_k), C, rows_k, w.cand_key + off, w.cand_cls + off)) return e;
    }
    if (use_small) {
      NmsParams pk = p;
      pk.outputs = out_k; pk.cand_key = w.cand_key + off; pk.cand_cls = w.cand_cls + off;
      pk.order = p.order + off; pk.keep = p.keep + off;
      pk.order_cnt = p.order_cnt ? p.order_cnt + n0 : nullptr;
      pk.keep_cnt = p.keep_cnt ? p.keep_cnt + n0 : nullptr;
      if (int e = per_class ? launch(nms_small_kernel<true>, static_cast<unsigned>(nk), kSmallThreads, small_smem, true, pk, w.large + n0)
                            : launch(nms_small_kernel<false>, static_cast<unsigned>(nk), kSmallThreads, small_smem, true, pk, w.large + n0)) return e;
    }
  }
  cluster_dim = kNmsCluster;           // dense images: a cluster of CTAs each; 36 clusters (144 of the 148 SMs) walk over the batch
  const int dense_clusters = N < 36 ? N : 36;
  // programmatic launch here too: the kernel's first statement is griddepcontrol.wait (everything before it in the stream is
  // complete and visible), so only its launch latency and block scheduling overlap with the tail of its predecessor
  if (per_class) return launch(nms_kernel<true, kNmsCluster>, static_cast<unsigned>(dense_clusters) * kNmsCluster, kNmsThreads, smem, parts == 1, p, static_cast<const int32_t*>(use_small ? w.large : nullptr));
  return launch(nms_kernel<false, kNmsCluster>, static_cast<unsigned>(dense_clusters) * kNmsCluster, kNmsThreads, smem, parts == 1, p, static_cast<const int32_t*>(use_small ? w.large : nullptr));
}

}  // namespace ssdh

using namespace ssdh;

extern "C" int ssdh_decode(const float* pr, int pr_row_stride, const float* priors, int N, int P, float* out, ssdh_stream_t stream) {
  if (!pr || !priors || !out || N <= 0 || P <= 0 || pr_row_stride < 4) { set_error("ssdh_decode: bad argument"); return SSDH_E_ARG; }
  if (!aligned16(priors) || !aligned16(out)) { set_error("ssdh_decode: priors/out must be 16-byte aligned"); return SSDH_E_ALIGN; }
  const size_t rows = static_cast<size_t>(N) * P;
  decode_kernel<<<static_cast<unsigned>((rows + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      pr, pr_row_stride, reinterpret_cast<const float4*>(priors), P, reinterpret_cast<float4*>(out), rows);
  return cuda_status("ssdh_decode");
}

extern "C" int ssdh_score(const float* pr, int pr_row_stride, int N, int P, int C, float* out, ssdh_stream_t stream) {
  if (!pr || !out || N <= 0 || P <= 0 || C <= 0 || pr_row_stride < 4 + C) { set_error("ssdh_score: bad argument"); return SSDH_E_ARG; }
  const size_t rows = static_cast<size_t>(N) * P;
  score_kernel<<<static_cast<unsigned>((rows + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(pr, pr_row_stride, C, out, rows);
  return cuda_status("ssdh_score");
}

extern "C" int ssdh_iou(const float* t, int t_row_stride, int T, const float* s, int s_row_stride, int S, int N, float* out,
                        ssdh_stream_t stream) {
  if (!t || !s || !out || N <= 0 || T <= 0 || S <= 0) { set_error("ssdh_iou: bad argument"); return SSDH_E_ARG; }
  const size_t total = static_cast<size_t>(N) * T * S;
  iou_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(t, t_row_stride, T, s, s_row_stride, S, out, total);
  return cuda_status("ssdh_iou");
}

extern "C" __attribute__((visibility("default"))) void ssdh_debug_set_nms_trace(unsigned long long* buf) { g_nms_trace = buf; }

extern "C" int ssdh_gather_detections(const float* outputs, const int32_t* keep, const int32_t* keep_cnt, int N, int P, int C,
                                      int max_det, float* dets, int32_t* det_cnt, ssdh_stream_t stream) {
  if (!outputs || !keep || !keep_cnt || !dets || !det_cnt || N <= 0 || P <= 0 || C <= 1 || max_det <= 0) { set_error("ssdh_gather_detections: bad argument"); return SSDH_E_ARG; }
  const dim3 grid(static_cast<unsigned>((static_cast<size_t>(max_det) * 32 + 255) / 256), static_cast<unsigned>(N));
  gather_detections_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(outputs, keep, keep_cnt, P, C, max_det, dets, det_cnt);
  return cuda_status("ssdh_gather_detections");
}

extern "C" size_t ssdh_nms_workspace_bytes(int N, int P, int C) {
  (void)C;
  if (N <= 0 || P <= 0) return 0;
  return nms_ws_layout(N, P, nullptr, nullptr);
}

extern "C" int ssdh_nms(float* outputs, int N, int P, int C, float iou_thr, float score_thr, int top_k, int per_class,
                        int32_t* order, int32_t* order_cnt, int32_t* keep, int32_t* keep_cnt, void* ws, size_t ws_bytes,
                        ssdh_stream_t stream) {
  return run_nms(outputs, nullptr, N, P, C, iou_thr, score_thr, top_k, per_class, order, order_cnt, keep, keep_cnt, ws, ws_bytes,
                 static_cast<cudaStream_t>(stream), false, "ssdh_nms");
}

extern "C" int ssdh_postprocess(float* outputs, const float* priors, int N, int P, int C, float iou_thr, float score_thr,
                                int top_k, int per_class, int32_t* order, int32_t* order_cnt, int32_t* keep, int32_t* keep_cnt,
                                void* ws, size_t ws_bytes, ssdh_stream_t stream) {
  return run_nms(outputs, priors, N, P, C, iou_thr, score_thr, top_k, per_class, order, order_cnt, keep, keep_cnt, ws, ws_bytes,
                 static_cast<cudaStream_t>(stream), true, "ssdh_postprocess");
}
