"""Seeded synthetic SSD300 head inputs (VOC-shaped), used by tests and bench.py.

Shapes follow the reference's data contract:
  * ground truth rows are ``[cx, cy, w, h, one-hot(21)]`` with class ``id + 1`` set
    (reference ``src/dataset.py:92-116``), zero-padded per batch to the longest image
    exactly as ``pad_sequence`` does in ``collate_fn`` (reference ``src/utils.py:8-16``);
  * head output rows are ``[dcx, dcy, dw, dh, 21 logits]`` (reference ``src/model/ssd.py:86-106``).

Everything is generated on the CPU with an explicit ``torch.Generator`` so the same seed
gives the same bytes in this container and on the GPU box (same torch build).
"""
from __future__ import annotations

import torch

NUM_PRIORS = 8732
NUM_CLASSES = 21          # 20 VOC classes + void at index 0 (reference src/train.py:77)
ROW = 4 + NUM_CLASSES


def make_targets(n_images: int, seed: int, max_boxes: int = 20, min_boxes: int = 1,
                 num_classes: int = NUM_CLASSES) -> torch.Tensor:
    """Random ground truth, ``(N, G, 4 + C)`` fp32, G = longest image in the batch."""
    g = torch.Generator().manual_seed(1000003 * seed + 17)
    counts = torch.randint(min_boxes, max_boxes + 1, (n_images,), generator=g)
    G = int(counts.max().item()) if n_images > 0 else 0
    out = torch.zeros(n_images, G, 4 + num_classes, dtype=torch.float32)
    for i in range(n_images):
        k = int(counts[i])
        if k == 0:
            continue
        wh = 0.05 + 0.60 * torch.rand(k, 2, generator=g)
        ctr = wh / 2 + torch.rand(k, 2, generator=g) * (1 - wh)
        lab = torch.randint(1, num_classes, (k,), generator=g)
        out[i, :k, 0:2] = ctr
        out[i, :k, 2:4] = wh
        out[i, torch.arange(k), 4 + lab] = 1.0
    return out


def make_outputs(n_images: int, seed: int, dist: str = "D1", num_priors: int = NUM_PRIORS,
                 num_classes: int = NUM_CLASSES) -> torch.Tensor:
    """Random head outputs ``(N, P, 4 + C)`` fp32.

    ``D1`` ("random-init"): N(0,1) in every column -> ~95 % of priors are NMS candidates.
    ``D2`` ("trained-like"): offsets 0.1*N(0,1), void logit +4 -> a few hundred candidates.
    """
    g = torch.Generator().manual_seed(7919 * seed + 3)
    x = torch.randn(n_images, num_priors, 4 + num_classes, generator=g, dtype=torch.float32)
    if dist == "D1":
        return x
    if dist == "D2":
        x[:, :, :4] *= 0.1
        x[:, :, 4] += 4.0
        return x
    raise ValueError(f"unknown distribution {dist!r}")


def make_batch(n_images: int, seed: int, dist: str = "D1", max_boxes: int = 20):
    return make_outputs(n_images, seed, dist), make_targets(n_images, seed, max_boxes)


def plant_detections(outputs: torch.Tensor, targets: torch.Tensor, priors: torch.Tensor, seed: int,
                     per_gt: int = 3, jitter: float = 0.15) -> torch.Tensor:
    """Overwrite a few rows per ground-truth box so that they decode close to that box with its label.

    Random logits almost never produce true positives, which would leave the TP/FP bookkeeping of
    the evaluation path (reference src/evaluate.py:146-148) unexercised.  For every real gt row this
    picks ``per_gt`` priors near the box centre, writes offsets that decode to the box (plus jitter,
    so the duplicates overlap each other and NMS has work to do) and a confident logit for the label.
    """
    g = torch.Generator().manual_seed(104729 * seed + 11)
    out = outputs.clone()
    N, G = targets.shape[0], targets.shape[1]
    for n in range(N):
        for j in range(G):
            box = targets[n, j, :4]
            if float(box[2] * box[3]) <= 0:
                continue
            label = int(targets[n, j, 4:].argmax())
            dist = (priors[:, 0] - box[0]).abs() + (priors[:, 1] - box[1]).abs() + (priors[:, 2] - box[2]).abs()
            rows = torch.topk(dist, per_gt, largest=False).indices
            for r in rows.tolist():
                d = priors[r]
                noise = jitter * (torch.rand(4, generator=g) - 0.5)
                out[n, r, 0] = (box[0] - d[0]) / d[2] + noise[0]
                out[n, r, 1] = (box[1] - d[1]) / d[3] + noise[1]
                out[n, r, 2] = torch.log(box[2] / d[2]) + noise[2]
                out[n, r, 3] = torch.log(box[3] / d[3]) + noise[3]
                out[n, r, 4:] = -2.0
                out[n, r, 4 + label] = 4.0 + 4.0 * float(torch.rand(1, generator=g))
    return out


# ---------------------------------------------------------------------------------------------------------------------
# Device-side generators for the large benchmark legs (1024-image and 4952-image configs): the same distributions, drawn
# with a CUDA generator so that gigabytes of inputs do not have to come through the host.  Deterministic for a given
# (seed, shape, GPU model, torch build) -- every rank of one box regenerates the same block from the same seed.
# ---------------------------------------------------------------------------------------------------------------------
def make_outputs_device(n_images: int, seed: int, dist: str, device, num_priors: int = NUM_PRIORS,
                        num_classes: int = NUM_CLASSES) -> torch.Tensor:
    g = torch.Generator(device=device).manual_seed(7919 * seed + 3)
    x = torch.randn(n_images, num_priors, 4 + num_classes, generator=g, dtype=torch.float32, device=device)
    if dist == "D2":
        x[:, :, :4] *= 0.1
        x[:, :, 4] += 4.0
    elif dist != "D1":
        raise ValueError(f"unknown distribution {dist!r}")
    return x


def pad_targets(targets: torch.Tensor, g_rows: int) -> torch.Tensor:
    """Zero-pad the ground-truth axis to ``g_rows`` rows (padding rows are inert, reference ssd.py:250,269)."""
    n, g, row = targets.shape
    if g >= g_rows:
        return targets
    return torch.cat([targets, targets.new_zeros(n, g_rows - g, row)], dim=1)


def plant_detections_device(outputs: torch.Tensor, targets: torch.Tensor, priors: torch.Tensor, per_gt: int = 2,
                            chunk: int = 64) -> torch.Tensor:
    """Vectorised, in-place variant of ``plant_detections`` for tensors that live on the GPU: for every real ground-truth
    box the ``per_gt`` nearest priors are rewritten so that they decode onto the box (the second one shifted by 4 % of
    the box size, so the duplicates overlap and NMS / first-claimant bookkeeping have work to do) with a confident logit
    for its label.  Deterministic (no random numbers)."""
    N, G = targets.shape[0], targets.shape[1]
    pr = priors.to(outputs.device)
    for n0 in range(0, N, chunk):
        t = targets[n0:n0 + chunk].to(outputs.device)
        box = t[:, :, :4]                                                   # (n, G, 4)
        real = (box[:, :, 2] * box[:, :, 3]) > 0
        d = ((pr[None, None, :, 0] - box[:, :, None, 0]).abs() + (pr[None, None, :, 1] - box[:, :, None, 1]).abs()
             + (pr[None, None, :, 2] - box[:, :, None, 2]).abs())           # (n, G, P)
        rows = torch.topk(d, per_gt, dim=2, largest=False).indices          # (n, G, per_gt)
        label = t[:, :, 4:].argmax(dim=2)                                   # (n, G)
        for j in range(per_gt):
            nn, gg = torch.nonzero(real, as_tuple=True)
            r = rows[nn, gg, j]
            dp = pr[r]
            b = box[nn, gg]
            shift = 0.04 * j
            vals = torch.full((nn.numel(), outputs.shape[2]), -2.0, device=outputs.device)
            vals[:, 0] = (b[:, 0] + shift * b[:, 2] - dp[:, 0]) / dp[:, 2]
            vals[:, 1] = (b[:, 1] - shift * b[:, 3] - dp[:, 1]) / dp[:, 3]
            vals[:, 2] = torch.log(b[:, 2] / dp[:, 2])
            vals[:, 3] = torch.log(b[:, 3] / dp[:, 3])
            vals[torch.arange(nn.numel(), device=outputs.device), 4 + label[nn, gg]] = 6.0 - 1.5 * j
            # two boxes may pick the same prior: an indexed store with duplicate targets is not deterministic on the GPU,
            # so only the first claimant (lowest image / box index) of every (image, prior) writes
            key = nn * pr.shape[0] + r
            order = torch.arange(key.numel(), device=key.device)
            uniq, inv = torch.unique(key, return_inverse=True)
            first = torch.full((uniq.numel(),), key.numel(), device=key.device, dtype=order.dtype).scatter_reduce(0, inv, order, reduce="amin")
            outputs[n0 + nn[first], r[first]] = vals[first]
    return outputs
