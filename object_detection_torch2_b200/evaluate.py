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
               want_flags: bool = False):
    """One batch of src/evaluate.py:132-151: adds {TP, detections, ground truths} per class to ``tallies``."""
    return ops.eval_accumulate(outputs, gts, tallies, iou_thresh, want_flags)


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
