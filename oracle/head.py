"""CPU restatement of the reference SSD300 head math.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

All tensors are fp32 on the CPU.  Where the reference materialises (N, P, G, 4) and
(N, P, G, C) temporaries, this restatement walks the ground-truth axis in a Python loop
and keeps only (N, P) accumulators; the elementary fp32 operations applied to each
(prior, ground-truth) pair are the same and in the same order, so boolean / index
results are identical and floating-point sums agree to summation-order rounding.

Reference line numbers are relative to /root/reference/.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import torch

# ----------------------------------------------------------------------------------------
# P1  default boxes                                             src/model/ssd.py:108-133
# ----------------------------------------------------------------------------------------
LEVELS: Tuple[Tuple[int, int], ...] = ((38, 4), (19, 6), (10, 6), (5, 6), (3, 4), (1, 4))
S_MIN, S_MAX, N_LEVELS = 0.2, 0.9, 6


def _scale(k: int) -> float:
    # src/model/ssd.py:114-115 (k is 1-based; k = 7 gives 1.04 and is used by the extra box)
    return S_MIN + (S_MAX - S_MIN) * (k - 1) / (N_LEVELS - 1)


def level_shapes(k: int, n_anchor: int) -> List[Tuple[float, float]]:
    """(w, h) of the anchors of level k in row order.  src/model/ssd.py:121-129."""
    ratios = [1.0, 2.0, 0.5] if n_anchor == 4 else [1.0, 2.0, 0.5, 3.0, 1.0 / 3.0]
    s = _scale(k)
    shapes = [(s * (a ** 0.5), s * ((1.0 / a) ** 0.5)) for a in ratios]
    extra = (s * _scale(k + 1)) ** 0.5
    shapes.append((extra, extra))
    return shapes


def default_boxes() -> torch.Tensor:
    """(8732, 4) fp32 priors ``[cx, cy, w, h]``.

    Row index = level offset + (i * m + j) * A + a with cx = (i + .5) / m and
    cy = (j + .5) / m, i.e. the OUTER loop index drives cx (src/model/ssd.py:122-130).
    Values are computed in float64 and rounded once to fp32, as ``torch.Tensor([[...]])``
    does with Python floats (src/model/ssd.py:130).
    """
    chunks = []
    for k, (m, n_anchor) in enumerate(LEVELS, start=1):
        shapes = torch.tensor(level_shapes(k, n_anchor), dtype=torch.float64)        # (A, 2)
        centre = (torch.arange(m, dtype=torch.float64) + 0.5) / m                     # (m,)
        cx = centre.view(m, 1, 1).expand(m, m, n_anchor)
        cy = centre.view(1, m, 1).expand(m, m, n_anchor)
        w = shapes[:, 0].view(1, 1, n_anchor).expand(m, m, n_anchor)
        h = shapes[:, 1].view(1, 1, n_anchor).expand(m, m, n_anchor)
        chunks.append(torch.stack([cx, cy, w, h], dim=-1).reshape(-1, 4))
    return torch.cat(chunks).to(torch.float32)


# ----------------------------------------------------------------------------------------
# L1  matching                                                  src/model/ssd.py:231-250
# ----------------------------------------------------------------------------------------
def _overlap_1d(c_a, s_a, c_b, s_b):
    # clamp(min(a_hi, b_hi) - max(a_lo, b_lo), 0)          src/model/ssd.py:247-248
    return (torch.minimum(c_a + s_a / 2, c_b + s_b / 2) - torch.maximum(c_a - s_a / 2, c_b - s_b / 2)).clamp(min=0)


def pair_iou_match(gt_box: torch.Tensor, priors: torch.Tensor) -> torch.Tensor:
    """IoU-like value the reference thresholds in ``_match``: (N, P) for ONE gt column.

    ``gt_box`` is (N, 4); ``priors`` (P, 4).  Rows with zero gt area (padding) yield the
    gt area itself (= 0), src/model/ssd.py:250.
    """
    g = gt_box.unsqueeze(1)                     # (N, 1, 4)
    d = priors.unsqueeze(0)                     # (1, P, 4)
    w = _overlap_1d(g[..., 0], g[..., 2], d[..., 0], d[..., 2])
    h = _overlap_1d(g[..., 1], g[..., 3], d[..., 1], d[..., 3])
    g_area = g[..., 2] * g[..., 3]
    inter = w * h
    iou = inter / (g_area + d[..., 2] * d[..., 3] - inter)
    return torch.where(g_area > 0, iou, g_area.expand_as(iou))


def match_mask(targets: torch.Tensor, priors: torch.Tensor, threshold: float = 0.25,
               force_best_prior: bool = False) -> torch.Tensor:
    """(N, P, G) bool, ``iou > threshold``.  src/model/ssd.py:231-250 (the reference does NO best-prior forcing).

    ``force_best_prior`` is the north_star extension, the identity when off: every real ground-truth
    box (area > 0) additionally claims the prior it overlaps best (first maximum = lowest prior index on ties,
    ``torch.argmax``), provided that best IoU is positive -- the SSD paper's "match each ground truth box to the default
    box with the best jaccard overlap", expressed on the reference's multi-match mask."""
    N, G = targets.shape[0], targets.shape[1]
    out = torch.zeros(N, priors.shape[0], G, dtype=torch.bool)
    for g in range(G):
        iou = pair_iou_match(targets[:, g, :4], priors)
        out[:, :, g] = iou > threshold
        if force_best_prior:
            best_iou, best_prior = iou.max(dim=1)
            real = (targets[:, g, 2] * targets[:, g, 3] > 0) & (best_iou > 0)
            rows = torch.nonzero(real).flatten()
            out[rows, best_prior[rows], g] = True
    return out


# ----------------------------------------------------------------------------------------
# L2 / L3  offset encoding and smooth-L1                        src/model/ssd.py:252-283
# ----------------------------------------------------------------------------------------
def encode_offsets(gt_box: torch.Tensor, priors: torch.Tensor) -> torch.Tensor:
    """g-hat for ONE gt column: (N, P, 4).  No variances; log guarded by ``> 0``.  ssd.py:267-270."""
    g = gt_box.unsqueeze(1)
    d = priors.unsqueeze(0)
    e_cx = (g[..., 0] - d[..., 0]) / d[..., 2]
    e_cy = (g[..., 1] - d[..., 1]) / d[..., 3]
    gw = g[..., 2].expand_as(e_cx)
    gh = g[..., 3].expand_as(e_cx)
    e_w = torch.where(gw > 0, torch.log(gw / d[..., 2]), gw)
    e_h = torch.where(gh > 0, torch.log(gh / d[..., 3]), gh)
    return torch.stack([e_cx, e_cy, e_w, e_h], dim=-1)


def smooth_l1(x: torch.Tensor) -> torch.Tensor:
    """|x| < 1 ? x^2/2 : |x| - 1/2.  src/model/ssd.py:274-283."""
    ax = x.abs()
    return torch.where(ax < 1, 0.5 * x * x, ax - 0.5)


# ----------------------------------------------------------------------------------------
# L5 / L6  3:1 split and (k+1)-th value threshold               src/model/ssd.py:300-328
# ----------------------------------------------------------------------------------------
def split_pos_neg(pos_raw: torch.Tensor, n_priors: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """(k_pos, k_neg) int64 per image.  src/model/ssd.py:218-220, 310-311."""
    neg_raw = n_priors - pos_raw
    crowded = pos_raw * 3 > neg_raw
    k_pos = torch.where(crowded, torch.div(neg_raw, 3, rounding_mode="floor"), pos_raw)
    k_neg = torch.where(crowded, neg_raw, pos_raw * 3)
    return k_pos, k_neg


def kplus1_threshold(values: torch.Tensor, k: int) -> torch.Tensor:
    """(k+1)-th largest of a 1-D tensor (k = 0 -> the max).  src/model/ssd.py:313-328."""
    ordered = torch.sort(values, descending=True).values
    return ordered[int(k)]


# ----------------------------------------------------------------------------------------
# L4 + L7  MultiBox loss                                        src/model/ssd.py:181-229
# ----------------------------------------------------------------------------------------
def multibox_loss(outputs: torch.Tensor, targets: torch.Tensor, priors: torch.Tensor,
                  a: float = 1.0, threshold: float = 0.25, want_grad: bool = False,
                  force_best_prior: bool = False) -> Dict[str, torch.Tensor]:
    """Restated ``SSD.loss``; returns the scalar and every per-image intermediate.

    Keys: loss (0-d), loss_per_image (N,), match (N,P,G) bool, pos_raw, k_pos, k_neg (N,) int64,
    thr_pos, thr_neg (N,), pos_sel, neg_sel (N,) int64 (#rows actually selected), ce_pos, ce_neg,
    l_loc (N,P), pos_valid, neg_valid (N,P) bool, and grad (N,P,4+C) if ``want_grad``.
    """
    outputs = outputs.detach().clone().requires_grad_(want_grad)
    N, P, _ = outputs.shape
    G = targets.shape[1]
    loc = outputs[:, :, :4]
    logp = torch.log_softmax(outputs[:, :, 4:], dim=2)                  # ssd.py:298

    match = match_mask(targets, priors, threshold, force_best_prior)    # ssd.py:199 (+ opt-in forcing, off = reference)
    l_loc = torch.zeros(N, P)
    ce_pos = torch.zeros(N, P)
    for g in range(G):
        m = match[:, :, g].to(torch.float32)
        delta = loc - encode_offsets(targets[:, g, :4], priors)         # ssd.py:202-204
        l_loc = l_loc + smooth_l1(delta).sum(dim=2) * m
        ce_g = -(targets[:, g, 4:].unsqueeze(1) * logp).sum(dim=2)      # ssd.py:208, 298
        ce_pos = ce_pos + ce_g * m                                      # ssd.py:209
    n_match = match.sum(dim=2)
    unmatched = n_match == 0                                            # ssd.py:214
    ce_neg = -logp[:, :, 0] * unmatched                                 # ssd.py:212-215

    pos_raw = (n_match != 0).sum(dim=1)                                 # ssd.py:218
    k_pos, k_neg = split_pos_neg(pos_raw, P)                            # ssd.py:219-220
    thr_pos = torch.stack([kplus1_threshold(ce_pos[i].detach(), k_pos[i]) for i in range(N)])   # ssd.py:222
    thr_neg = torch.stack([kplus1_threshold(ce_neg[i].detach(), k_neg[i]) for i in range(N)])   # ssd.py:223
    pos_valid = ce_pos > thr_pos.unsqueeze(1)
    neg_valid = ce_neg > thr_neg.unsqueeze(1)

    inv = torch.where(k_pos > 0, 1.0 / k_pos.float(), k_pos.float())    # ssd.py:226
    per_image = ((a * l_loc + ce_pos) * pos_valid + ce_neg * neg_valid).sum(dim=1) * inv
    loss = per_image.mean()                                             # ssd.py:227

    res = dict(loss=loss.detach(), loss_per_image=per_image.detach(), match=match, pos_raw=pos_raw,
               k_pos=k_pos, k_neg=k_neg, thr_pos=thr_pos, thr_neg=thr_neg,
               pos_sel=pos_valid.sum(dim=1), neg_sel=neg_valid.sum(dim=1),
               ce_pos=ce_pos.detach(), ce_neg=ce_neg.detach(), l_loc=l_loc.detach(),
               pos_valid=pos_valid, neg_valid=neg_valid)
    if want_grad:
        loss.backward()
        res["grad"] = outputs.grad.detach()
    return res


# ----------------------------------------------------------------------------------------
# I1 / I2  decode and score                                     src/utils.py:19-55
# ----------------------------------------------------------------------------------------
def decode_boxes(pr: torch.Tensor, priors: torch.Tensor) -> torch.Tensor:
    """(N, P, 4) ``[cx, cy, w, h]``: cx = dw*p0 + dcx, w = dw*exp(p2).  src/utils.py:35-38."""
    d = priors.unsqueeze(0)
    cx = d[..., 2] * pr[..., 0] + d[..., 0]
    cy = d[..., 3] * pr[..., 1] + d[..., 1]
    w = d[..., 2] * torch.exp(pr[..., 2])
    h = d[..., 3] * torch.exp(pr[..., 3])
    return torch.stack([cx, cy, w, h], dim=2)


def class_scores(pr: torch.Tensor) -> torch.Tensor:
    """(N, P, C): softmax kept only at each row's arg-max class, 0 elsewhere.  src/utils.py:52-55."""
    logits = pr[:, :, 4:]
    best = logits.max(dim=2).indices
    keep = torch.zeros_like(logits)
    keep.scatter_(2, best.unsqueeze(2), 1.0)
    return torch.softmax(logits, dim=2) * keep


# ----------------------------------------------------------------------------------------
# I3  pairwise IoU                                              src/utils.py:58-77
# ----------------------------------------------------------------------------------------
def pair_iou(t: torch.Tensor, s: torch.Tensor) -> torch.Tensor:
    """(N, T, S) IoU of centre-form boxes; 0-overlap pairs return the overlap (0).  src/utils.py:74-77."""
    a = t[:, :, None, :4]
    b = s[:, None, :, :4]
    w = _overlap_1d(a[..., 0], a[..., 2], b[..., 0], b[..., 2])
    h = _overlap_1d(a[..., 1], a[..., 3], b[..., 1], b[..., 3])
    inter = w * h
    return torch.where(inter > 0, inter / (a[..., 2] * a[..., 3] + b[..., 2] * b[..., 3] - inter), inter)


# ----------------------------------------------------------------------------------------
# I4  greedy NMS                                                src/utils.py:80-116
# ----------------------------------------------------------------------------------------
def greedy_nms(rows: torch.Tensor, iou_thresh: float = 0.5, score_thresh: float = 0.0,
               top_k: Optional[int] = None, per_class: bool = False) -> Tuple[torch.Tensor, torch.Tensor]:
    """One image.  Returns (order, keep): candidate row indices by descending best non-void
    score (ties: lower row first), and the kept subset in that order.

    Defaults are the reference: candidates have ``max(row[5:]) > 0`` (src/utils.py:99-100),
    class-agnostic suppression at ``IoU > iou_thresh`` (src/utils.py:102-108).
    Extensions (north_star): ``score_thresh`` replaces the ``> 0`` cut, ``per_class`` only lets
    boxes of the same arg-max class suppress each other, ``top_k`` truncates the kept list.
    """
    key, cls = rows[:, 5:].max(dim=1)
    vals, idx = torch.sort(key, descending=True, stable=True)
    order = idx[vals > score_thresh]
    K = order.numel()
    alive = torch.ones(K, dtype=torch.bool)
    boxes = rows[order, :4].unsqueeze(0)
    kcls = cls[order]
    for i in range(K - 1):
        if not alive[i]:
            continue
        iou = pair_iou(boxes[:, i:i + 1], boxes[:, i + 1:])[0, 0]
        hit = iou > iou_thresh
        if per_class:
            hit &= kcls[i + 1:] == kcls[i]
        alive[i + 1:] &= ~hit
    keep = order[alive]
    if top_k is not None:
        keep = keep[:top_k]
    return order, keep


def nms_inplace(outputs: torch.Tensor, iou_thresh: float = 0.5, **kw) -> Tuple[torch.Tensor, List[torch.Tensor]]:
    """Batch NMS with the reference's in-place contract: score columns 4: of every row that
    is not kept become 0, box columns are untouched, the SAME tensor is returned.  utils.py:111-116."""
    keeps = []
    for n in range(outputs.shape[0]):
        _, keep = greedy_nms(outputs[n], iou_thresh, **kw)
        mask = torch.zeros(outputs.shape[1], 1)
        mask[keep] = 1.0
        outputs[n, :, 4:] = outputs[n, :, 4:] * mask
        keeps.append(keep)
    return outputs, keeps


# ----------------------------------------------------------------------------------------
# E1 / E2 / E3  evaluation                                      src/evaluate.py:31-67, 132-159
# ----------------------------------------------------------------------------------------
def class_order(rows: torch.Tensor, class_id: int) -> torch.Tensor:
    """Indices of rows whose column 5+class_id is > 0, by that column descending.  evaluate.py:41-42."""
    vals, idx = torch.sort(rows[:, 5 + class_id], descending=True, stable=True)
    return idx[vals > 0]


def eval_image_class(output: torch.Tensor, gt: torch.Tensor, class_id: int, iou_thresh: float = 0.5):
    """TP flags for one image and class.  Returns (det_order, correct (D,), n_gt) or None when
    neither detections nor ground truth exist (src/evaluate.py:137-150).

    Each detection goes to its arg-max-IoU ground truth of the class; it is a true positive
    when that IoU is > iou_thresh and no higher-scored detection already claimed that box.
    """
    det = class_order(output, class_id)
    gto = class_order(gt, class_id)
    if det.numel() == 0 and gto.numel() == 0:
        return None
    if det.numel() == 0:
        return det, torch.zeros(0), int(gto.numel())
    correct = torch.zeros(det.numel())
    if gto.numel() > 0:
        iou = pair_iou(output[det].unsqueeze(0), gt[gto].unsqueeze(0))[0]      # (D, Gc)
        best_iou, best_gt = iou.max(dim=1)
        claimed = set()
        for d in range(det.numel()):
            if best_iou[d] > iou_thresh:
                gidx = int(best_gt[d])
                if gidx not in claimed:
                    claimed.add(gidx)
                    correct[d] = 1.0
    return det, correct, int(gto.numel())


def eval_batch(outputs: torch.Tensor, gts: torch.Tensor, n_classes: int = 20, iou_thresh: float = 0.5):
    """Loop body of src/evaluate.py:134-151 over a batch.

    Returns (tallies int64 (n_classes, 3) = [TP, detections, ground truths], results) where
    results[c] is a list of (correct, score) float tensors, one per image that had detections.
    """
    tallies = torch.zeros(n_classes, 3, dtype=torch.int64)
    results: Dict[int, List[torch.Tensor]] = {c: [] for c in range(n_classes)}
    for n in range(outputs.shape[0]):
        for c in range(n_classes):
            r = eval_image_class(outputs[n], gts[n], c, iou_thresh)
            if r is None:
                continue
            det, correct, n_gt = r
            tallies[c, 2] += n_gt
            if det.numel() == 0:
                continue
            tallies[c, 0] += int(correct.sum())
            tallies[c, 1] += det.numel()
            results[c].append(torch.stack([correct, outputs[n, det, 5 + c]], dim=1))
    return tallies, results


def average_precision(result: torch.Tensor, count) -> torch.Tensor:
    """The reference's AP.  src/evaluate.py:55-67 sorts the two columns of ``result`` independently,
    so the TP flags are ordered on their own (ones first); the interpolated precision is then 1
    over the first TP entries and the area under the curve collapses to TP / count."""
    flags = torch.sort(result[:, 0], descending=True).values
    tp = torch.cumsum(flags == 1.0, dim=0)
    fp = torch.cumsum(flags == 0.0, dim=0)
    precision = tp / (tp + fp)
    recall = tp / count
    zero = torch.zeros(1)
    padded = torch.cat([zero, precision, zero])
    envelope = torch.flip(torch.cummax(torch.flip(padded, dims=[0]), dim=0).values, dims=[0])
    rec = torch.cat([zero, recall, torch.ones(1)])
    return torch.sum(envelope[1:] * (rec[1:] - rec[:-1]))


# ----------------------------------------------------------------------------------------
# Opt-in extension: true PASCAL VOC AP (not in the reference).  Plain numpy.
# ----------------------------------------------------------------------------------------
def voc_ap_numpy(scores, tp, n_gt: int, use_07_metric: bool = False) -> float:
    import numpy as np
    scores, tp = np.asarray(scores, dtype=np.float64), np.asarray(tp, dtype=np.float64)
    if n_gt <= 0:
        return float("nan")
    if scores.size == 0:
        return 0.0
    order = np.argsort(-scores, kind="stable")
    hit = tp[order]
    tps, fps = np.cumsum(hit), np.cumsum(1.0 - hit)
    rec, prec = tps / n_gt, tps / np.maximum(tps + fps, 1e-12)
    if use_07_metric:
        return float(sum((prec[rec >= t].max() if (rec >= t).any() else 0.0) for t in np.arange(0.0, 1.1, 0.1)) / 11.0)
    mrec = np.concatenate([[0.0], rec, [1.0]])
    mpre = np.concatenate([[0.0], prec, [0.0]])
    for i in range(mpre.size - 2, -1, -1):
        mpre[i] = max(mpre[i], mpre[i + 1])
    idx = np.where(mrec[1:] != mrec[:-1])[0]
    return float(((mrec[idx + 1] - mrec[idx]) * mpre[idx + 1]).sum())
