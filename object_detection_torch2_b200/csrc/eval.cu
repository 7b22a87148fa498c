// E1 + E2: per-class TP/FP assignment accumulated as integer tallies.  Replaces get_order and the loop body of
// the evaluation script (reference src/evaluate.py:31-42 and :132-151).
//
// The reference sorts the detections of a class by score and lets each one claim its arg-max-IoU ground truth
// (valid when IoU > 0.5); a detection is a true positive iff it is the FIRST claimant of that box
// (evaluate.py:146-148).  "First in descending score order, ties by lower row" is a minimum over the key
// (~score_bits, row), so no sort is needed: one atomicMin per valid detection, then one compare.
// The reference's AP (evaluate.py:45-67) reduces to TP / #gt, so {TP, detections, gt} per class
// are sufficient statistics and sum exactly across images and GPUs.
#include "common.cuh"

namespace ssdh {

constexpr int kEvalThreads = 512;

struct EvalParams {
  const float* outputs;
  const float* gts;
  int P, C, G;
  ThrBand band;
  unsigned long long* tallies;   // [C-1, 3]
  uint8_t* tp_flags;             // [N, P] or NULL
  int* status;                   // workspace: set to 1 if an image had more than P detections
  const int32_t* keep;           // [N, P] kept rows in score order (ssdh_nms / ssdh_postprocess), or NULL: scan the slab
  const int32_t* keep_cnt;       // [N]
};

struct Det {          // 4 bytes: the score is re-read from the slab when a claim is made (keeps five images per SM)
  uint16_t row;
  uint8_t cls;      // 0-based non-void class
  uint8_t best_g;   // 255 = no valid claim
};

// kKept = false: the detections are found by scanning the whole [P, 4+C] slab (any rows, several positive classes per row
// allowed): S bytes per image.  kKept = true: the detections ARE the kept list the NMS pass has just written -- only those
// rows are touched (a few hundred x 100 bytes per image instead of 873 KB); rows must carry one positive class at most.
template <bool kKept>
__global__ void __launch_bounds__(kEvalThreads) eval_kernel(const EvalParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int n = blockIdx.x, P = p.P, C = p.C, G = p.G, row = 4 + C, NC = C - 1;
  const int tid = threadIdx.x;
  // layout
  unsigned long long* claim = reinterpret_cast<unsigned long long*>(smem_raw);          // [NC][G]
  Corners* gt_box = reinterpret_cast<Corners*>(claim + static_cast<size_t>(NC) * G);     // [G]
  float* gt_val = reinterpret_cast<float*>(gt_box + G);                                  // [NC][G] class column values
  uint8_t* gt_order = reinterpret_cast<uint8_t*>(gt_val + static_cast<size_t>(NC) * G);  // [NC][G]
  int* gt_cnt = reinterpret_cast<int*>(gt_order + ((static_cast<size_t>(NC) * G + 3) & ~static_cast<size_t>(3)));  // [NC]
  int* tp_cnt = gt_cnt + NC;
  int* det_cnt = tp_cnt + NC;
  int* n_det = det_cnt + NC;
  Det* dets = reinterpret_cast<Det*>(n_det + 4);                                         // [P]

  const float* img = p.outputs + static_cast<size_t>(n) * P * row;
  const float* gimg = p.gts + static_cast<size_t>(n) * G * row;

  for (int i = tid; i < NC * G; i += kEvalThreads) {
    claim[i] = ~0ull;
    gt_val[i] = gimg[static_cast<size_t>(i % G) * row + 5 + i / G];
  }
  for (int g = tid; g < G; g += kEvalThreads) {
    const float* r = gimg + static_cast<size_t>(g) * row;
    gt_box[g] = make_corners(r[0], r[1], r[2], r[3]);
  }
  for (int c = tid; c < NC; c += kEvalThreads) { tp_cnt[c] = 0; det_cnt[c] = 0; }
  if (tid == 0) *n_det = 0;
  if (p.tp_flags)
    for (int i = tid; i < P; i += kEvalThreads) p.tp_flags[static_cast<size_t>(n) * P + i] = 255;
  __syncthreads();

  // get_order on the ground truth (evaluate.py:41-42): rows with a positive class column, by that value
  // descending, ties by lower row.  G <= 64: insertion by one thread per class.
  for (int c = tid; c < NC; c += kEvalThreads) {
    int cnt = 0;
    for (int g = 0; g < G; ++g) {
      const float v = gt_val[c * G + g];
      if (!(v > 0.0f)) continue;
      int j = cnt++;
      while (j > 0 && gt_val[c * G + gt_order[c * G + j - 1]] < v) { gt_order[c * G + j] = gt_order[c * G + j - 1]; --j; }
      gt_order[c * G + j] = static_cast<uint8_t>(g);
    }
    gt_cnt[c] = cnt;
  }

  // detections: every positive entry of the score columns 5.. (coalesced flat scan of the image slab, 16 bytes per
  // load when the slab allows).  The decoded box columns are positive in every row, so the column of each element is
  // tracked incrementally (no division) and only score columns are examined; after NMS a few hundred entries per image
  // are positive and only those pay for the row / column split.
  if (!kKept) {
  const int total = P * row;
  auto take = [&](int i, float v) {
    if (!(v > 0.0f)) return;
    const int r = i / row, c = i - r * row;
    if (c < 5) return;
    const int slot = atomicAdd(n_det, 1);
    if (slot < P) {
      Det d;
      d.row = static_cast<uint16_t>(r); d.cls = static_cast<uint8_t>(c - 5); d.best_g = 255;
      dets[slot] = d;
    }
  };
  if ((reinterpret_cast<uintptr_t>(img) & 15u) == 0 && (total & 3) == 0 && row >= 8) {
    const float4* img4 = reinterpret_cast<const float4*>(img);
    const int total4 = total >> 2;
    constexpr int kBatch = 8;                                  // loads in flight per thread before any of them is examined
    const int step = (4 * kEvalThreads) % row;                 // column advance between consecutive loads of a thread
    int col = (4 * tid) % row;                                 // column of the first element of the next load
    for (int q0 = tid; q0 < total4; q0 += kBatch * kEvalThreads) {
      float4 v[kBatch];
#pragma unroll
      for (int k = 0; k < kBatch; ++k) {
        const int q = q0 + k * kEvalThreads;
        v[k] = q < total4 ? __ldg(img4 + q) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int k = 0; k < kBatch; ++k) {
        const int c0 = col;
        col += step;
        if (col >= row) col -= row;
        // score columns among the four elements: column c0 + j (mod row) >= 5
        const int c1 = c0 + 1 >= row ? c0 + 1 - row : c0 + 1, c2 = c0 + 2 >= row ? c0 + 2 - row : c0 + 2, c3 = c0 + 3 >= row ? c0 + 3 - row : c0 + 3;
        const bool hit = (v[k].x > 0.0f && c0 >= 5) || (v[k].y > 0.0f && c1 >= 5) || (v[k].z > 0.0f && c2 >= 5) || (v[k].w > 0.0f && c3 >= 5);
        if (hit) {
          const int q = q0 + k * kEvalThreads;
          take(4 * q + 0, v[k].x); take(4 * q + 1, v[k].y); take(4 * q + 2, v[k].z); take(4 * q + 3, v[k].w);
        }
      }
    }
  } else {
    for (int i = tid; i < total; i += kEvalThreads) {
      const float v = img[i];
      if (v > 0.0f && i % row >= 5) take(i, v);
    }
  }
  } else {
    // kept list -> detections: thread per kept row; its class is the (single) positive score column
    const int K = min(p.keep_cnt[n], P);
    const int32_t* kp = p.keep + static_cast<size_t>(n) * P;
    for (int i = tid; i < K; i += kEvalThreads) {
      const int r = kp[i];
      const float* sc = img + static_cast<size_t>(r) * row + 5;
      int cls = -1;
      for (int c = 0; c < NC; ++c)
        if (cls < 0 && sc[c] > 0.0f) cls = c;
      if (cls < 0) continue;                       // a kept row always has a positive score; tolerate foreign lists
      const int slot = atomicAdd(n_det, 1);
      Det d;
      d.row = static_cast<uint16_t>(r); d.cls = static_cast<uint8_t>(cls); d.best_g = 255;
      dets[slot] = d;
    }
  }
  __syncthreads();
  int D = *n_det;
  if (D > P) {
    if (tid == 0) atomicExch(p.status, 1);
    D = P;
  }

  for (int i = tid; i < D; i += kEvalThreads) {
    Det d = dets[i];
    const int c = d.cls;
    const float* b = img + static_cast<size_t>(d.row) * row;
    const Corners me = make_corners(b[0], b[1], b[2], b[3]);
    const int cnt = gt_cnt[c];
    float best = -INFINITY;
    int bg = -1;
    for (int j = 0; j < cnt; ++j) {                       // arg-max over the class's gt in get_order order, first max wins
      const int g = gt_order[c * G + j];
      const float v = iou_value(me, gt_box[g]);
      if (v > best) { best = v; bg = g; }
    }
    if (bg >= 0 && best > p.band.thr) {                   // evaluate.py:147
      const unsigned long long key = (static_cast<unsigned long long>(~float_key(b[5 + c])) << 32) | d.row;
      atomicMin(&claim[c * G + bg], key);
      dets[i].best_g = static_cast<uint8_t>(bg);
    }
    atomicAdd(&det_cnt[c], 1);
  }
  __syncthreads();
  for (int i = tid; i < D; i += kEvalThreads) {
    const Det d = dets[i];
    bool tp = false;
    if (d.best_g != 255) {
      const float score = img[static_cast<size_t>(d.row) * row + 5 + d.cls];
      const unsigned long long key = (static_cast<unsigned long long>(~float_key(score)) << 32) | d.row;
      tp = claim[d.cls * G + d.best_g] == key;            // first claimant in score order, evaluate.py:148
    }
    if (tp) atomicAdd(&tp_cnt[d.cls], 1);
    if (p.tp_flags) p.tp_flags[static_cast<size_t>(n) * P + d.row] = tp ? 1 : 0;
  }
  __syncthreads();
  for (int c = tid; c < NC; c += kEvalThreads) {
    if (tp_cnt[c]) atomicAdd(&p.tallies[c * 3 + 0], static_cast<unsigned long long>(tp_cnt[c]));
    if (det_cnt[c]) atomicAdd(&p.tallies[c * 3 + 1], static_cast<unsigned long long>(det_cnt[c]));
    if (gt_cnt[c]) atomicAdd(&p.tallies[c * 3 + 2], static_cast<unsigned long long>(gt_cnt[c]));
  }
}

static size_t eval_smem_bytes(int P, int C, int G) {
  const size_t NC = C - 1;
  size_t b = NC * G * 8 + static_cast<size_t>(G) * sizeof(Corners) + NC * G * 4 + ((NC * G + 3) & ~static_cast<size_t>(3)) + (3 * NC + 4) * 4;
  b = (b + 15) & ~static_cast<size_t>(15);
  return b + static_cast<size_t>(P) * sizeof(Det) + 16;
}

}  // namespace ssdh

using namespace ssdh;

extern "C" size_t ssdh_eval_workspace_bytes(int N, int P, int C, int G) {
  (void)N; (void)P; (void)C; (void)G;
  return 256;
}

static int eval_impl(const float* outputs, const int32_t* keep, const int32_t* keep_cnt, const float* gts, int N, int P, int C, int G,
                     float iou_thr, int64_t* tallies, uint8_t* tp_flags, void* ws, size_t ws_bytes, ssdh_stream_t stream) {
  if (!outputs || !tallies || N <= 0 || P <= 0 || C <= 1 || G < 0 || (G > 0 && !gts)) { set_error("ssdh_eval_accumulate: bad argument"); return SSDH_E_ARG; }
  if (C > kMaxClasses || G > kMaxGT || P > 65535) { set_error("ssdh_eval_accumulate: limits are C <= %d, G <= %d, P <= 65535", kMaxClasses, kMaxGT); return SSDH_E_LIMIT; }
  if (!ws || ws_bytes < ssdh_eval_workspace_bytes(N, P, C, G)) { set_error("ssdh_eval_accumulate: workspace too small"); return SSDH_E_WORKSPACE; }
  const size_t smem = eval_smem_bytes(P, C, G);
  if (smem > 227 * 1024) { set_error("ssdh_eval_accumulate: needs %zu bytes of shared memory", smem); return SSDH_E_LIMIT; }
  if (int e = ensure_dyn_smem(reinterpret_cast<const void*>(eval_kernel<false>), 227 * 1024, "ssdh_eval_accumulate")) return e;
  if (int e = ensure_dyn_smem(reinterpret_cast<const void*>(eval_kernel<true>), 227 * 1024, "ssdh_eval_accumulate")) return e;
  EvalParams p;
  p.outputs = outputs; p.gts = gts; p.P = P; p.C = C; p.G = G;
  p.band = make_band(iou_thr);
  p.tallies = reinterpret_cast<unsigned long long*>(tallies);
  p.tp_flags = tp_flags;
  p.status = reinterpret_cast<int*>(ws);
  p.keep = keep; p.keep_cnt = keep_cnt;
  if (keep) eval_kernel<true><<<N, kEvalThreads, smem, static_cast<cudaStream_t>(stream)>>>(p);
  else eval_kernel<false><<<N, kEvalThreads, smem, static_cast<cudaStream_t>(stream)>>>(p);
  return cuda_status("ssdh_eval_accumulate");
}

extern "C" int ssdh_eval_accumulate(const float* outputs, const float* gts, int N, int P, int C, int G, float iou_thr,
                                    int64_t* tallies, uint8_t* tp_flags, void* ws, size_t ws_bytes, ssdh_stream_t stream) {
  return eval_impl(outputs, nullptr, nullptr, gts, N, P, C, G, iou_thr, tallies, tp_flags, ws, ws_bytes, stream);
}

extern "C" int ssdh_eval_accumulate_kept(const float* outputs, const int32_t* keep, const int32_t* keep_cnt, const float* gts,
                                         int N, int P, int C, int G, float iou_thr, int64_t* tallies, uint8_t* tp_flags,
                                         void* ws, size_t ws_bytes, ssdh_stream_t stream) {
  if (!keep || !keep_cnt) { set_error("ssdh_eval_accumulate_kept: keep / keep_cnt are required"); return SSDH_E_ARG; }
  return eval_impl(outputs, keep, keep_cnt, gts, N, P, C, G, iou_thr, tallies, tp_flags, ws, ws_bytes, stream);
}

extern "C" int ssdh_eval_status(const void* ws, int* status_host, ssdh_stream_t stream) {
  if (!ws || !status_host) { set_error("ssdh_eval_status: NULL pointer"); return SSDH_E_ARG; }
  // the one synchronising entry point of the library, deliberately separate from the launches: 4 bytes D2H
  cudaError_t e = cudaMemcpyAsync(status_host, ws, sizeof(int), cudaMemcpyDeviceToHost, static_cast<cudaStream_t>(stream));
  if (e == cudaSuccess) e = cudaStreamSynchronize(static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) { set_error("ssdh_eval_status: %s", cudaGetErrorString(e)); (void)cudaGetLastError(); return static_cast<int>(e); }
  return 0;
}
