"""Evaluation core with the reference's function names (``src/evaluate.py``): per-class ordering, TP/FP tallies
and the reference's AP formula.  The per-image / per-class Python loop of src/evaluate.py:134-151 is one kernel
launch per batch (``accumulate``); AP comes from the integer tallies."""
from __future__ import annotations

from typing import Optional

import torch

from . import ops


def get_order(t: torch.Tensor, class_id: int) -> torch.Tensor:
    """Rows whose column ``5 + class_id`` is positive, by that value descending (ties: lower row).
    Reference src/evaluate.py:31-42.  Host-side convenience; the kernel path does not need it."""
    vals, indices = torch.sort(t[:, 5 + class_id], descending=True, stable=True)
    return indices[vals > 0.]


def accumulate(outputs: torch.Tensor, gts: torch.Tensor, tallies: Optional[torch.Tensor] = None, iou_thresh: float = 0.5,
               want_flags: bool = False, keep: Optional[torch.Tensor] = None, keep_cnt: Optional[torch.Tensor] = None,
               check_status: bool = True):
    """One batch of src/evaluate.py:132-151: adds {TP, detections, ground truths} per class to ``tallies``.
    ``keep`` / ``keep_cnt`` (from ``ops.postprocess_(..., want_lists=True)``) let the kernel read only the kept rows."""
    return ops.eval_accumulate(outputs, gts, tallies, iou_thresh, want_flags, keep, keep_cnt, check_status)


def average_precision_from_tallies(tallies: torch.Tensor) -> torch.Tensor:
    """(C-1,) AP per class.  The reference's AP sorts the TP column independently of the scores
    (src/evaluate.py:55), which makes it exactly TP / #gt; classes without ground truth give NaN as there."""
    t = tallies.to(torch.float32)
    return t[:, 0] / t[:, 2]


def calc_average_precision(result: torch.Tensor, count) -> torch.Tensor:
    """The reference AP on an explicit (X, 2) list of [correct, score] rows, src/evaluate.py:45-67."""
    n_true = (result[:, 0] == 1.).sum()
    n_rows = result.shape[0]
    rank = torch.arange(1, n_rows + 1, device=result.device)
    tp = torch.minimum(rank, n_true)             # TP flags sorted on their own: all ones first
    precision = tp.float() / rank.float()
    recall = tp.float() / count
    zero = torch.zeros(1, device=result.device)
    envelope = torch.flip(torch.cummax(torch.flip(torch.cat([zero, precision, zero]), dims=[0]), dim=0).values, dims=[0])
    rec = torch.cat([zero, recall, torch.ones(1, device=result.device)])
    return torch.sum(envelope[1:] * (rec[1:] - rec[:-1]))


# ---------------------------------------------------------------------------------------------------------------------
# "Next" row: the true VOC average precision, opt-in beside the reference's recall-style AP.
# ---------------------------------------------------------------------------------------------------------------------
def voc_average_precision(scores: torch.Tensor, tp: torch.Tensor, n_gt: int, use_07_metric: bool = False) -> torch.Tensor:
    """PASCAL VOC AP of one class from its detections over the WHOLE dataset: ``scores`` (D,), ``tp`` (D,) 0/1 flags
    (first-claimant assignment, as the kernel produces them), ``n_gt`` ground-truth boxes.  Detections are ranked
    jointly by score (stable: earlier entries first on ties); AP is the 11-point mean (VOC2007) or the area under the
    monotone precision envelope (VOC2010+).  Pure tensor math, runs wherever the inputs live: the host-side restatement of
    ``ops.voc_ap`` (the kernel ``DetectionEvaluator.compute`` uses), kept as its checker and for CPU-side callers."""
    if n_gt <= 0:
        return torch.tensor(float("nan"), device=scores.device)
    if scores.numel() == 0:
        return torch.zeros((), device=scores.device)
    order = torch.sort(scores, descending=True, stable=True).indices
    hit = tp[order].to(torch.float64)
    tps = torch.cumsum(hit, 0)
    fps = torch.cumsum(1.0 - hit, 0)
    recall = tps / float(n_gt)
    precision = tps / torch.clamp(tps + fps, min=1e-12)
    if use_07_metric:
        ap = torch.zeros((), dtype=torch.float64, device=scores.device)
        for t in range(11):
            sel = recall >= t * 0.1              # same float thresholds as the canonical np.arange(0., 1.1, 0.1)
            ap = ap + (precision[sel].max() if bool(sel.any()) else 0.0) / 11.0
        return ap.to(torch.float32)
    zero = torch.zeros(1, dtype=torch.float64, device=scores.device)
    mrec = torch.cat([zero, recall, torch.ones(1, dtype=torch.float64, device=scores.device)])
    mpre = torch.cat([zero, precision, zero])
    mpre = torch.flip(torch.cummax(torch.flip(mpre, dims=[0]), dim=0).values, dims=[0])
    return torch.sum((mrec[1:] - mrec[:-1]) * mpre[1:]).to(torch.float32)


class DetectionEvaluator:
    """Streams batches through ``ssdh_eval_accumulate`` and keeps what both metrics need:
    the int64 (C-1, 3) tallies (the reference's AP = TP / #gt) and, per detection, (class, score, TP flag) for the true
    VOC AP.  ``compute(group)`` all-reduces the tallies and all-gathers the detection lists across ranks."""

    def __init__(self, num_classes: int = 20, iou_thresh: float = 0.5):
        self.num_classes, self.iou_thresh = num_classes, iou_thresh
        self.tallies: Optional[torch.Tensor] = None
        self.cls, self.score, self.tp = [], [], []

    def update(self, outputs: torch.Tensor, gts: torch.Tensor, keep: Optional[torch.Tensor] = None,
               keep_cnt: Optional[torch.Tensor] = None) -> None:
        """``outputs`` (N, P, 4+C) after NMS, ``gts`` (N, G, 4+C); optionally the kept lists of the NMS pass."""
        self.tallies, flags = ops.eval_accumulate(outputs, gts, self.tallies, self.iou_thresh, want_flags=True, keep=keep,
                                                  keep_cnt=keep_cnt)
        rows = flags != 255
        scores, cls = outputs[:, :, 5:][rows].max(dim=1)
        self.cls.append(cls.to(torch.int32))
        self.score.append(scores)
        self.tp.append(flags[rows].to(torch.float32))

    def compute(self, group=None, use_07_metric: bool = False):
        import torch.distributed as dist
        tallies = self.tallies.clone()
        cls, score, tp = torch.cat(self.cls), torch.cat(self.score), torch.cat(self.tp)
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(tallies, op=dist.ReduceOp.SUM, group=group)
            n = torch.tensor([cls.numel()], device=cls.device)
            sizes = [torch.zeros_like(n) for _ in range(dist.get_world_size(group))]
            dist.all_gather(sizes, n, group=group)
            cap = int(max(int(x) for x in sizes))
            packed = torch.zeros(cap, 3, device=cls.device)
            packed[: cls.numel()] = torch.stack([cls.float(), score, tp], dim=1)
            parts = [torch.zeros_like(packed) for _ in sizes]
            dist.all_gather(parts, packed, group=group)          # the one bandwidth-class collective of the eval path
            packed = torch.cat([p[: int(k)] for p, k in zip(parts, sizes)])
            cls, score, tp = packed[:, 0].to(torch.int32), packed[:, 1], packed[:, 2]
        ap_ref = average_precision_from_tallies(tallies)
        # one stable radix sort + one CTA per class (ssdh_voc_ap); ``voc_average_precision`` above is its host restatement
        ap_voc = ops.voc_ap(score, tp, cls, tallies, use_07_metric)
        return dict(tallies=tallies, ap_reference=ap_ref, ap_voc=ap_voc)
