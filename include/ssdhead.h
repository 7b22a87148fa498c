/*
 * ssdhead.h -- C ABI of libssdhead.so: the SSD300 detection-head hot path on B200 (sm_100a).
 *
 * The reference (rs1004/object-detection-torch2) has no FFI / plugin layer: its boundary for this path
 * is the Python surface of src/model/ssd.py, src/utils.py and src/evaluate.py.  Each entry point below
 * names the reference function (file:line under /root/reference) whose arithmetic it replaces; the
 * Python mirror in object_detection_torch2_b200/ binds them with ctypes.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer on the caller's current device unless marked "host";
 *   - all tensors are dense row-major fp32 unless a stride argument says otherwise;
 *   - the library never allocates, frees or synchronises: outputs and scratch are caller-owned, every
 *     call only enqueues work on `stream` (a cudaStream_t passed as void*), so calls are CUDA-graph
 *     capturable and thread-safe per stream;
 *   - return value: 0 = OK, <0 = SSDH_E_* argument error, >0 = cudaError_t of a failed launch;
 *     ssdh_last_error() returns a thread-local message for the last non-zero return;
 *   - limits: 1 <= C <= 64 classes (incl. void at index 0), G <= 64 ground-truth rows per image,
 *     one image's [P, 4+C] slab must fit the cluster's shared memory (P*(4+C)*4 <= ~1.7 MB), P <= 65535;
 *     NMS / post-processing: N <= 65535 images per call, P <= 10240 priors per image.
 */
#ifndef SSDHEAD_H_
#define SSDHEAD_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define SSDH_API __attribute__((visibility("default")))
#else
#define SSDH_API
#endif

#define SSDH_VERSION 100

#define SSDH_E_ARG       (-1)  /* null pointer / non-positive dimension                    */
#define SSDH_E_LIMIT     (-2)  /* dimension above a documented limit (C, G, slab size)     */
#define SSDH_E_ALIGN     (-3)  /* pointer not 16-byte aligned                              */
#define SSDH_E_WORKSPACE (-4)  /* workspace missing or too small                           */

typedef void* ssdh_stream_t;   /* cudaStream_t */

/* Per-image by-products of the MultiBox loss (all written by ssdh_multibox_loss). */
typedef struct ssdh_image_stats {
  float loss;     /* inv_pos * sum over selected rows, src/model/ssd.py:227 before .mean()      */
  float thr_pos;  /* (k_pos+1)-th largest positive CE, src/model/ssd.py:222                     */
  float thr_neg;  /* (k_neg+1)-th largest negative CE, src/model/ssd.py:223                     */
  int32_t pos_raw;  /* priors with >= 1 match, src/model/ssd.py:218                               */
  int32_t k_pos;    /* after the 3:1 split, src/model/ssd.py:310-311                              */
  int32_t k_neg;
  int32_t pos_sel;  /* rows with ce_pos > thr_pos (can be < k_pos on ties)                        */
  int32_t neg_sel;  /* rows with ce_neg > thr_neg                                                 */
} ssdh_image_stats;

SSDH_API int ssdh_version(void);
SSDH_API const char* ssdh_last_error(void);

/* Static facts about the loaded build and the current device (host out-params, may be NULL). */
SSDH_API int ssdh_device_info(int* sm_count, int* max_smem_optin, int* loss_cluster_size, int* loss_max_active_clusters);

/* P1  SSD._get_default_bboxes, src/model/ssd.py:108-133.  out: [8732, 4] = [cx, cy, w, h]. */
SSDH_API int ssdh_default_boxes(float* out, ssdh_stream_t stream);
#define SSDH_NUM_PRIORS 8732

/* L1  SSD._match, src/model/ssd.py:231-250 (IoU > thr, no forcing).
 * gt: [N, G, gt_row_stride] rows, columns 0..3 read.  priors: [P, 4].
 * match_bits [N, P] u64 (bit g = prior matches gt g)            -- may be NULL
 * match_mask [N, P, G] u8 0/1, the reference's bool layout      -- may be NULL
 * best_gt [N, P] i32 / best_iou [N, P]: arg-max IoU over gt per prior (first max), may be NULL
 * best_prior [N, G] i32 / best_prior_iou [N, G]: arg-max IoU over priors per gt (lowest index on
 * ties), the hook for best-prior forcing (north_star extension), may be NULL. */
SSDH_API int ssdh_match(const float* gt, int gt_row_stride, int N, int G, const float* priors, int P, float thr,
               uint64_t* match_bits, uint8_t* match_mask, int32_t* best_gt, float* best_iou,
               int32_t* best_prior, float* best_prior_iou, ssdh_stream_t stream);

/* L2  SSD._calc_delta, src/model/ssd.py:252-272.  out: [N, P, G, 4]. */
SSDH_API int ssdh_encode(const float* gt, int gt_row_stride, int N, int G, const float* priors, int P, float* out,
                ssdh_stream_t stream);

/* L3  SSD._smooth_l1, src/model/ssd.py:274-283, elementwise over n floats. */
SSDH_API int ssdh_smooth_l1(const float* x, float* out, size_t n, ssdh_stream_t stream);

/* L4  SSD._softmax_cross_entropy, src/model/ssd.py:285-298.
 * pr: [N, P, pr_row_stride] with C logits from column 0; gt: [N, G, gt_row_stride] with C weights from
 * column 0; out: [N, P, G]. */
SSDH_API int ssdh_softmax_cross_entropy(const float* pr, int pr_row_stride, const float* gt, int gt_row_stride,
                               int N, int P, int G, int C, float* out, ssdh_stream_t stream);

/* L5  SSD._split_pos_neg, src/model/ssd.py:300-311 (int64 in / out, n entries). */
SSDH_API int ssdh_split_pos_neg(const int64_t* pos, const int64_t* neg, int64_t* pos_out, int64_t* neg_out, int n,
                       ssdh_stream_t stream);

/* L6  SSD._k_plus_1_th_value, src/model/ssd.py:313-328, batched: values [rows, len], k [rows] (device i64),
 * out [rows] = (k+1)-th largest of each row (k = 0 -> max).  Radix select, no sort. */
SSDH_API int ssdh_kplus1_value(const float* values, int rows, int len, const int64_t* k, float* out, ssdh_stream_t stream);

/* L1-L7 fused  SSD.loss, src/model/ssd.py:181-229, forward and gradient in one launch.
 * outputs [N, P, 4+C], targets [N, G, 4+C], priors [P, 4]; `a` = localisation weight; thr = match IoU.
 * n_global: the divisor of the batch mean (= N on one GPU, the global batch when images are sharded).
 * loss [1]: sum_i stats[i].loss / n_global.   grad [N, P, 4+C] = d loss / d outputs, or NULL (forward only).
 * stats [N] or NULL.  ws: ssdh_multibox_loss_workspace_bytes() bytes, ZEROED ONCE by the caller before the
 * first call (the kernel leaves it zeroed). */
SSDH_API size_t ssdh_multibox_loss_workspace_bytes(int N, int P, int C, int G);
SSDH_API int ssdh_multibox_loss(const float* outputs, const float* targets, const float* priors, int N, int P, int C, int G,
                       float a, float thr, int n_global, float* loss, float* grad, ssdh_image_stats* stats,
                       void* ws, size_t ws_bytes, ssdh_stream_t stream);

/* Same as ssdh_multibox_loss for callers that run it back to back on buffers that are ALREADY COMPLETE (a training loop
 * over micro-batches staged earlier, gradient accumulation, a benchmark): by calling this entry point the caller vouches
 * that outputs / targets / priors were not written by the kernel that precedes this call in the stream.  That allows
 *   - programmatic dependent launch with early reads: the grid loads, matches and selects under the previous grid's tail
 *     and only waits for it before its first global write (ssdh_multibox_loss waits before its first global READ, because
 *     its predecessor is normally the producer of `outputs`);
 *   - software pipelining across micro-batches: once this batch's slabs are on chip (HBM is then idle until the gradient
 *     is written) every CTA asks the L2 for the blocks it will read from next_outputs / next_targets (same shapes; either
 *     may be NULL) in the NEXT call.  Pure hint: no data is changed. */
SSDH_API int ssdh_multibox_loss_pipelined(const float* outputs, const float* targets, const float* priors, int N, int P, int C, int G,
                       float a, float thr, int n_global, float* loss, float* grad, ssdh_image_stats* stats,
                       void* ws, size_t ws_bytes, ssdh_stream_t stream, const float* next_outputs, const float* next_targets);

/* Extended form: everything above plus the opt-in behaviours, selected through a versioned options record (NULL = the
 * reference behaviour of ssdh_multibox_loss).
 *   force_best_prior  north_star's "best-prior-per-GT forcing" (the reference has none, src/model/ssd.py:231-250):
 *                     every real ground-truth box additionally claims its arg-max-IoU prior (lowest index on ties) when that
 *                     IoU is positive.  0 = the reference.
 *   inputs_stable     the contract of ssdh_multibox_loss_pipelined.
 *   exact_math        libdevice exp / log and IEEE division instead of the approximate units (slower; for parity studies).
 *   exchange          the step's scalar all-reduce fused into the kernel's epilogue (NVLink stores into the peers' inboxes).
 *   ce_override       test hook (needs exact_math): [N, P] cross-entropies the hard-negative selection is run on -- positive CE
 *                     for matched priors, negative CE for the others -- so the selection logic can be checked on the checker's
 *                     own numbers. */
/* Step-scalar exchange over NVLink (the one collective of sharded training), see csrc/exchange.cu.
 * Every rank owns an INBOX of world x SSDH_XCHG_RING 64-bit words (+ two 32-bit counters behind them) in device memory that
 * its peers map through CUDA IPC.  When ssdh_multibox_loss_ex is given an exchange, the CTA that finalises step s stores
 * (s << 32 | bits(loss)) into slot [rank][(s - 1) % ring] of EVERY rank's inbox -- one 8-byte store per peer over NVLink, no
 * kernel launch, no NCCL.  ssdh_scalar_exchange_reduce(count) waits for the next `count` steps of all ranks and writes their
 * sums (added in rank order: bit-identical on every rank) to out[count]; call it every <= ring / 2 steps, anywhere in the
 * stream (it is CUDA-graph capturable; a peer that never delivers sets *status = 1 and yields NaN instead of hanging).
 *   create:  allocates + zeroes this rank's inbox, returns it with its IPC handle (exchange the handles with any transport)
 *   open:    maps a PEER's inbox (enables peer access lazily); close / destroy undo open / create. */
#define SSDH_MAX_RANKS 16
#define SSDH_XCHG_RING 256
typedef struct ssdh_ipc_handle { unsigned char bytes[64]; } ssdh_ipc_handle;
typedef struct ssdh_scalar_exchange {
  int32_t world, rank;
  uint32_t ring;                                  /* = SSDH_XCHG_RING */
  uint32_t reserved;
  unsigned long long* inbox[SSDH_MAX_RANKS];      /* inbox[r]: rank r's inbox as seen from THIS device (inbox[rank] = the local one) */
  uint32_t* counters;                             /* local inbox + world * ring words: [0] steps pushed, [1] steps consumed */
} ssdh_scalar_exchange;
SSDH_API size_t ssdh_scalar_exchange_bytes(int world);
SSDH_API int ssdh_scalar_exchange_create(int world, void** inbox, ssdh_ipc_handle* handle);
SSDH_API int ssdh_scalar_exchange_open(const ssdh_ipc_handle* handle, void** peer_inbox);
SSDH_API int ssdh_scalar_exchange_close(void* peer_inbox);
SSDH_API int ssdh_scalar_exchange_destroy(void* inbox);
SSDH_API int ssdh_scalar_exchange_reduce(const ssdh_scalar_exchange* x, int count, float* out, int* status, ssdh_stream_t stream);

typedef struct ssdh_loss_options {
  uint32_t struct_bytes;       /* = sizeof(ssdh_loss_options) */
  int32_t force_best_prior;
  int32_t inputs_stable;
  int32_t exact_math;
  const float* next_outputs;   /* as in ssdh_multibox_loss_pipelined, may be NULL */
  const float* next_targets;
  const float* ce_override;    /* may be NULL */
  const ssdh_scalar_exchange* exchange;   /* may be NULL: publish this step's loss scalar to every rank's inbox (see above) */
} ssdh_loss_options;
SSDH_API int ssdh_multibox_loss_ex(const float* outputs, const float* targets, const float* priors, int N, int P, int C, int G,
                       float a, float thr, int n_global, float* loss, float* grad, ssdh_image_stats* stats,
                       void* ws, size_t ws_bytes, ssdh_stream_t stream, const ssdh_loss_options* options);

/* Software pipelining hook: pull [ptr, ptr + bytes) from HBM into the L2 (cp.async.bulk.prefetch.L2), e.g. the NEXT
 * micro-batch's head output on a side stream while ssdh_multibox_loss works on the current one.  Reads nothing into
 * the SMs, changes no data; ptr must be 16-byte aligned. */
SSDH_API int ssdh_prefetch_l2(const void* ptr, size_t bytes, ssdh_stream_t stream);

/* SURVEY 8f-1 -- head-output producer, the tail of SSD.forward (src/model/ssd.py:96-104): the n_levels detector outputs
 * levels[k] (N, ch[k], hw[k]) fp32 NCHW-contiguous, ch[k] = anchors_k * width, are written as
 * permute(0, 2, 3, 1).reshape(N, -1, width) blocks, concatenated along dim 1, into outputs (N, P, width) in ONE pass
 * (the reference makes seven copies).  levels / ch / hw are HOST arrays; P must equal sum_k hw[k] * ch[k] / width. */
SSDH_API int ssdh_pack_head(const float* const* levels, const int* ch, const int* hw, int n_levels, int N, int width,
                            float* outputs, int P, ssdh_stream_t stream);

/* Backward of ssdh_pack_head: scatters grad_outputs (N, P, width) into the detectors' NCHW gradients level_grads[k]. */
SSDH_API int ssdh_unpack_head(const float* grad_outputs, float* const* level_grads, const int* ch, const int* hw, int n_levels,
                              int N, int width, int P, ssdh_stream_t stream);

/* The same two passes for channels-last producers: levels[k] / level_grads[k] are (N, hw[k], ch[k]) contiguous -- the memory of
 * an (N, ch, H, W) tensor in torch.channels_last format, which is what cuDNN's tensor-core convolutions prefer to write.  Every
 * (level, image) block is then already in slab order and the pass is a plain copy (permute(0, 2, 3, 1) is free, reshape + cat
 * of src/model/ssd.py:103-104 is the copy). */
SSDH_API int ssdh_pack_head_nhwc(const float* const* levels, const int* ch, const int* hw, int n_levels, int N, int width,
                                 float* outputs, int P, ssdh_stream_t stream);
SSDH_API int ssdh_unpack_head_nhwc(const float* grad_outputs, float* const* level_grads, const int* ch, const int* hw, int n_levels,
                                   int N, int width, int P, ssdh_stream_t stream);

/* SURVEY 8f-3 -- ground-truth ingest, the device side of collate_fn (src/utils.py:8-16): compact rows
 * [cx, cy, w, h, label] (N, G, 5) fp32 and the per-image row counts lengths[N] (NULL: every row is real) are expanded
 * into the dense zero-padded one-hot tensor targets (N, G, 4 + C) that pad_sequence would have produced.  label is a
 * class index 0..C-1 stored as a float (0 = void). */
SSDH_API int ssdh_expand_targets(const float* compact, const int* lengths, int N, int G, int C, float* targets, ssdh_stream_t stream);

/* grad *= *scale (device scalar), skipped entirely when *scale == 1: the autograd chain-rule hook. */
SSDH_API int ssdh_scale_inplace(float* x, size_t n, const float* scale, ssdh_stream_t stream);

/* I1  calc_coordicate, src/utils.py:19-40.  pr: [N, P, pr_row_stride] (cols 0..3), out: [N, P, 4]. */
SSDH_API int ssdh_decode(const float* pr, int pr_row_stride, const float* priors, int N, int P, float* out, ssdh_stream_t stream);

/* I2  calc_score, src/utils.py:43-55.  pr: [N, P, pr_row_stride] with C logits from column 4; out: [N, P, C]. */
SSDH_API int ssdh_score(const float* pr, int pr_row_stride, int N, int P, int C, float* out, ssdh_stream_t stream);

/* I3  calc_iou, src/utils.py:58-77.  t: [N, T, t_row_stride], s: [N, S, s_row_stride] (cols 0..3); out [N, T, S]. */
SSDH_API int ssdh_iou(const float* t, int t_row_stride, int T, const float* s, int s_row_stride, int S, int N, float* out,
             ssdh_stream_t stream);

/* I4  non_maximum_suppression, src/utils.py:80-116, in place on outputs [N, P, 4+C] (decoded + scored).
 * Reference behaviour: score_thr = 0, top_k = 0 (off), per_class = 0, iou_thr = 0.5.
 * order [N, P] i32 / order_cnt [N]: candidate rows by descending best non-void score (stable) -- may be NULL
 * keep  [N, P] i32 / keep_cnt  [N]: kept rows in that order                                  -- may be NULL
 * Score columns of every row that is not kept are set to 0; box columns are untouched. */
SSDH_API size_t ssdh_nms_workspace_bytes(int N, int P, int C);
SSDH_API int ssdh_nms(float* outputs, int N, int P, int C, float iou_thr, float score_thr, int top_k, int per_class,
             int32_t* order, int32_t* order_cnt, int32_t* keep, int32_t* keep_cnt,
             void* ws, size_t ws_bytes, ssdh_stream_t stream);

/* I1+I2+I4 fused: the three calls at src/evaluate.py:129-131 / src/inference.py:67-69 in one pass over
 * outputs [N, P, 4+C] (raw head output in, decoded boxes + NMS-masked scores out, in place). */
SSDH_API int ssdh_postprocess(float* outputs, const float* priors, int N, int P, int C, float iou_thr, float score_thr,
                     int top_k, int per_class, int32_t* order, int32_t* order_cnt, int32_t* keep, int32_t* keep_cnt,
                     void* ws, size_t ws_bytes, ssdh_stream_t stream);

/* "Next" row: compact detection lists instead of the dense in-place tensor.  For each image the kept rows
 * (keep [N, P] i32 / keep_cnt [N] as written by ssdh_nms / ssdh_postprocess, score order) become rows
 * [cx, cy, w, h, score, label] of dets [N, max_det, 6]; det_cnt [N] = min(keep_cnt, max_det); unused rows are zeroed.
 * label is the 1-based class (column - 4); replaces the 8732-row Python loop of src/inference.py:77-81. */
SSDH_API int ssdh_gather_detections(const float* outputs, const int32_t* keep, const int32_t* keep_cnt, int N, int P, int C,
                           int max_det, float* dets, int32_t* det_cnt, ssdh_stream_t stream);

/* E1+E2  TP/FP assignment, src/evaluate.py:31-42 and :132-151, accumulated as sufficient statistics.
 * outputs [N, P, 4+C] after NMS, gts [N, G, 4+C].  tallies [C-1, 3] i64 += {TP, detections, ground truths}
 * per class (atomic adds: zero it before the first batch).  tp_flags [N, P] u8 or NULL: 1 = true positive,
 * 0 = false positive, 255 = row is not a detection.  Rows must carry at most one positive class score
 * (always true after calc_score). */
SSDH_API size_t ssdh_eval_workspace_bytes(int N, int P, int C, int G);
SSDH_API int ssdh_eval_accumulate(const float* outputs, const float* gts, int N, int P, int C, int G, float iou_thr,
                         int64_t* tallies, uint8_t* tp_flags, void* ws, size_t ws_bytes, ssdh_stream_t stream);

/* Same tallies, fed from the kept lists the NMS pass has just produced (keep [N, P] i32 / keep_cnt [N] of ssdh_nms /
 * ssdh_postprocess) instead of re-scanning the dense tensor: per image only the kept rows are read (a few hundred x
 * (4+C)*4 bytes instead of the whole P*(4+C)*4 slab), which is what makes decode + NMS + tallies cost 2*S + O(kept) per
 * image rather than 3*S.  Rows must be calc_score rows (one positive class column at most). */
SSDH_API int ssdh_eval_accumulate_kept(const float* outputs, const int32_t* keep, const int32_t* keep_cnt, const float* gts,
                              int N, int P, int C, int G, float iou_thr, int64_t* tallies, uint8_t* tp_flags,
                              void* ws, size_t ws_bytes, ssdh_stream_t stream);

/* Reads back the status word of an eval workspace (host out-param): 0 = fine, 1 = some image handed ssdh_eval_accumulate
 * more than P positive score entries (rows with several positive classes) and detections were dropped.  The ONLY
 * entry point that synchronises (4 bytes device -> host on `stream`); call it once after the last batch. */
SSDH_API int ssdh_eval_status(const void* ws, int* status_host, ssdh_stream_t stream);

/* "Next" row: the true PASCAL VOC average precision, opt-in beside the reference's recall-style AP
 * (src/evaluate.py:45-67 collapses to TP / #gt, see the tallies above).  One (score, tp flag, class) triple per detection of
 * the whole dataset (tp: 1 = true positive, anything else = false positive, as tp_flags of ssdh_eval_accumulate; cls: 0-based
 * non-void class), D of them, and the tallies [NC, 3] for the per-class ground-truth counts.  Detections are ranked per class
 * by descending score (ties: input order); ap_out [NC] = area under the monotone precision envelope (VOC2010+) or the
 * 11-point mean (use_07_metric), NaN for classes without ground truth.  fp64 inside, as the host restatement. */
SSDH_API size_t ssdh_voc_ap_workspace_bytes(int D, int NC);
SSDH_API int ssdh_voc_ap(const float* scores, const uint8_t* tp, const int32_t* cls, int D, const int64_t* tallies, int NC,
                int use_07_metric, float* ap_out, void* ws, size_t ws_bytes, ssdh_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* SSDHEAD_H_ */
