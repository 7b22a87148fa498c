"""Generate tests/golden/ssd_head_golden.npz by running the UNMODIFIED reference on CPU.

Run in the build container only (needs /root/reference):  python tests/golden/make_golden.py
Inputs are NOT stored: they are regenerated from seeds by object_detection_torch2_b200.synth and
their SHA-256 is stored so that a drifting generator is detected instead of silently mis-compared.
Everything stored below is an output of the reference's own functions
(src/model/ssd.py, src/utils.py, src/evaluate.py) or of the harness in oracle/ref_loader.py that
drives those functions through the script-only loop body of src/evaluate.py:132-151.
"""
from __future__ import annotations

import hashlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from object_detection_torch2_b200 import synth  # noqa: E402
from oracle import ref_loader  # noqa: E402

GRAD_STRIDE = 101

# (name, N, seed, dist, max_boxes, a, special)
LOSS_CASES = [
    ("d1_s0", 4, 0, "D1", 20, 1.0, None),
    ("d1_s1", 3, 1, "D1", 20, 1.0, None),
    ("d2_s2", 4, 2, "D2", 20, 1.0, None),
    ("d1_s3_a2", 2, 3, "D1", 6, 2.0, None),
    ("d2_s4_emptyimg", 3, 4, "D2", 5, 1.0, "empty_image_1"),
    ("d1_s5_one", 1, 5, "D1", 1, 1.0, None),
]
# (name, N, seed, dist, iou_thresh)
POST_CASES = [
    ("d2_s0", 3, 0, "D2", 0.5),
    ("d2_s1_t45", 2, 1, "D2", 0.45),
    ("d1_s2", 1, 2, "D1", 0.5),
    ("d2_s6_planted", 4, 6, "D2+planted", 0.5),
]


def sha(t: torch.Tensor) -> str:
    return hashlib.sha256(t.contiguous().numpy().tobytes()).hexdigest()


_PRIORS = None


def synth_priors():
    """Priors for planting detections: the reference's own (this script only runs beside the reference)."""
    global _PRIORS
    if _PRIORS is None:
        ref = ref_loader.load()
        _PRIORS = ref.SSD._get_default_bboxes(ref.net)
    return _PRIORS


def case_inputs(N, seed, dist, max_boxes, special=None):
    planted = dist.endswith("+planted")
    o, t = synth.make_batch(N, seed, dist.split("+")[0], max_boxes)
    if planted:
        o = synth.plant_detections(o, t, synth_priors(), seed)
    if special == "empty_image_1":
        t[1] = 0.0
    return o, t


def main():
    torch.manual_seed(0)
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    ref = ref_loader.load()
    net = ref.net
    out = {}

    df = ref.SSD._get_default_bboxes(net)
    out["priors"] = df.numpy()

    for name, N, seed, dist, mb, a, special in LOSS_CASES:
        o, t = case_inputs(N, seed, dist, mb, special)
        k = f"loss/{name}/"
        out[k + "sha_outputs"], out[k + "sha_targets"] = sha(o), sha(t)
        m = net._match(gt=t, df=df)
        out[k + "match_bits"] = np.packbits(m.numpy().astype(np.uint8), axis=None)
        out[k + "match_shape"] = np.array(m.shape)
        pos_raw = (m.sum(dim=2) != 0).sum(dim=1)
        kp, kn = net._split_pos_neg(pos_raw, o.shape[1] - pos_raw)
        out[k + "pos_raw"], out[k + "k_pos"], out[k + "k_neg"] = pos_raw.numpy(), kp.numpy(), kn.numpy()
        delta = net._calc_delta(gt=t, df=df)
        out[k + "delta_sample"] = delta.reshape(-1)[::GRAD_STRIDE].numpy()
        x = o.clone().requires_grad_(True)
        loss = net.loss(outputs=x, targets=t, default_bboxes=df, a=a)
        loss.backward()
        out[k + "loss"] = loss.detach().numpy()
        g = x.grad
        out[k + "grad_sample"] = g.reshape(-1)[::GRAD_STRIDE].numpy()
        out[k + "grad_abs_sum"] = g.abs().sum(dim=(1, 2)).double().numpy()
        out[k + "grad_row_nonzero"] = np.packbits((g.abs().sum(dim=2) > 0).numpy().astype(np.uint8), axis=None)
        with torch.no_grad():       # validation-pass call convention, src/train.py:128-139
            out[k + "loss_nograd"] = net.loss(outputs=o, targets=t, default_bboxes=df, a=a).numpy()
        print(name, "loss", float(loss), "pos_raw", pos_raw.tolist())

    for name, N, seed, dist, thr in POST_CASES:
        o, t = case_inputs(N, seed, dist, 20)
        k = f"post/{name}/"
        out[k + "sha_outputs"], out[k + "sha_targets"] = sha(o), sha(t)
        x = o.clone()
        box = ref.utils.calc_coordicate(pr=x, df=df)
        x[:, :, :4] = box
        sc = ref.utils.calc_score(pr=x)
        x[:, :, 4:] = sc
        out[k + "box_sample"] = box.reshape(-1)[::7].numpy()
        out[k + "score_argmax"] = sc.argmax(dim=2).numpy().astype(np.uint8)
        out[k + "score_max"] = sc.max(dim=2).values.numpy()
        scored = x.clone()
        y = ref.utils.non_maximum_suppression(outputs=x, iou_thresh=thr)
        assert y is x
        kept_rows = (y[:, :, 4:].sum(dim=2) > 0)
        out[k + "kept_rows"] = np.packbits(kept_rows.numpy().astype(np.uint8), axis=None)
        out[k + "kept_count"] = kept_rows.sum(dim=1).numpy()
        out[k + "sha_after_nms"] = sha(y)
        # IoU of the first 64 scored rows against the ground truth (calc_iou, src/utils.py:58-77)
        out[k + "iou_gt_64"] = ref.utils.calc_iou(scored[:, :64], t).numpy()
        rc, cnt = ref_loader.reference_eval_loop(ref, y, t)
        tallies = np.zeros((20, 3), dtype=np.int64)
        flags = []
        aps = np.full(20, np.nan, dtype=np.float32)
        for c in range(20):
            rows = [rc[i][c] for i in sorted(rc) if c in rc[i]]
            tallies[c, 2] = cnt[c]
            if rows:
                r = torch.cat(rows)
                tallies[c, 0] = int(r[:, 0].sum())
                tallies[c, 1] = r.shape[0]
                flags.append(r[:, 0].numpy().astype(np.uint8))
                aps[c] = float(ref.evaluate.calc_average_precision(result=r, count=cnt[c]))
        out[k + "tallies"] = tallies
        out[k + "tp_flags"] = np.concatenate(flags) if flags else np.zeros(0, np.uint8)
        out[k + "ap"] = aps
        print(name, "kept", out[k + "kept_count"].tolist(), "TP", tallies[:, 0].sum())

    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ssd_head_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
