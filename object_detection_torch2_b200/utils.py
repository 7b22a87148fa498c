"""Drop-in post-processing functions with the names, argument order, defaults and aliasing of the reference
``src/utils.py`` (``calc_coordicate`` keeps the reference's spelling), running on libssdhead.so kernels."""
from __future__ import annotations

from typing import Optional

import torch

from . import ops


def calc_coordicate(pr: torch.Tensor, df: torch.Tensor) -> torch.Tensor:
    """Offsets -> centre-form boxes, new (N, P, 4) tensor.  Reference src/utils.py:19-40."""
    return ops.decode(pr, df)


def calc_score(pr: torch.Tensor) -> torch.Tensor:
    """Softmax kept at each row's arg-max class only, new (N, P, C-4) tensor.  Reference src/utils.py:43-55."""
    return ops.score(pr)


def calc_iou(t: torch.Tensor, s: torch.Tensor) -> torch.Tensor:
    """Pairwise IoU (N, T, S) of centre-form boxes.  Reference src/utils.py:58-77."""
    return ops.iou(t, s)


def non_maximum_suppression(outputs: torch.Tensor, iou_thresh: float = 0.5, score_thresh: float = 0.0,
                            top_k: Optional[int] = None, per_class: bool = False) -> torch.Tensor:
    """Greedy class-agnostic NMS; zeroes ``outputs[:, :, 4:]`` of suppressed rows IN PLACE and returns the same
    tensor (reference src/utils.py:80-116).  ``score_thresh`` / ``top_k`` / ``per_class`` are opt-in extensions whose
    defaults are the reference behaviour."""
    if outputs.dtype == torch.float32 and outputs.is_contiguous():
        ops.nms_(outputs, iou_thresh, score_thresh, top_k, per_class)
        return outputs
    work = outputs.float().contiguous()          # keep the aliasing contract for odd layouts: copy back in place
    ops.nms_(work, iou_thresh, score_thresh, top_k, per_class)
    outputs.copy_(work)
    return outputs


def postprocess(outputs: torch.Tensor, df: torch.Tensor, iou_thresh: float = 0.5, score_thresh: float = 0.0,
                top_k: Optional[int] = None, per_class: bool = False) -> torch.Tensor:
    """The three calls of reference src/evaluate.py:129-131 / src/inference.py:67-69 as one in-place pass:
    raw head output in, decoded boxes + NMS-masked scores out (one read and one write of the tensor)."""
    if outputs.dtype == torch.float32 and outputs.is_contiguous():
        ops.postprocess_(outputs, df, iou_thresh, score_thresh, top_k, per_class)
        return outputs
    work = outputs.float().contiguous()
    ops.postprocess_(work, df, iou_thresh, score_thresh, top_k, per_class)
    outputs.copy_(work)
    return outputs


def detect(outputs: torch.Tensor, df: torch.Tensor, iou_thresh: float = 0.5, score_thresh: float = 0.0, top_k: Optional[int] = 200,
           per_class: bool = False):
    """Raw head output -> compact detections: ``postprocess`` in place, then the kept rows of every image
    as ``dets (N, top_k, 6) = [cx, cy, w, h, score, label]`` in descending score order plus ``det_cnt (N,)``.
    Replaces the dense tensor walk of reference src/inference.py:71-81."""
    res = ops.postprocess_(outputs, df, iou_thresh, score_thresh, top_k, per_class, want_lists=True)
    return ops.gather_detections(outputs, res.keep, res.keep_cnt, max_det=int(top_k or outputs.shape[1]))


def collate_fn_compact(batch, pin_memory: bool = True):
    """Drop-in for the reference's ``collate_fn`` (src/utils.py:8-16) that ships the ground truth compactly.

    ``batch`` is the dataset's list of ``(image, gt)`` with ``gt`` of shape (G_i, 4 + C) one-hot rows.  Returns
    ``(images, compact, lengths)``: ``compact`` (N, G, 5) rows ``[cx, cy, w, h, label]`` zero-padded to the batch
    maximum and ``lengths`` (N,) int32, both in pinned host memory so the H2D copies can be asynchronous.  On the device
    ``targets_from_compact`` rebuilds exactly the tensor ``pad_sequence`` would have produced (100 -> 20 bytes per row)."""
    images = torch.stack([img for img, _ in batch], dim=0)
    gmax = max((int(gt.shape[0]) for _, gt in batch), default=0)
    compact = torch.zeros((len(batch), gmax, 5), dtype=torch.float32)
    lengths = torch.zeros((len(batch),), dtype=torch.int32)
    for n, (_, gt) in enumerate(batch):
        g = int(gt.shape[0])
        lengths[n] = g
        if g:
            compact[n, :g, :4] = gt[:, :4]
            compact[n, :g, 4] = gt[:, 4:].argmax(dim=1).to(torch.float32)
    if pin_memory and torch.cuda.is_available():
        compact, lengths = compact.pin_memory(), lengths.pin_memory()
    return images, compact, lengths


def targets_from_compact(compact: torch.Tensor, lengths: torch.Tensor, num_classes: int, device=None) -> torch.Tensor:
    """Asynchronous H2D of the compact ground truth + on-device expansion to the dense (N, G, 4 + C) targets of SSD.loss."""
    device = device if device is not None else (compact.device if compact.is_cuda else "cuda")
    return ops.expand_targets(compact.to(device, non_blocking=True), lengths.to(device, non_blocking=True), num_classes)
