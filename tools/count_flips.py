"""Hard-negative selection flips of ssdh_multibox_loss against the oracle, batch 32, seeds 0-9, dist D1 and D2:
per-image flip counts and the distance of every flipped row's (oracle) cross-entropy from the (oracle) threshold in ulp.
python tools/count_flips.py > profiles/r02_selection_flips.txt"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ.setdefault("PYTEST_DISABLE_PLUGIN_AUTOLOAD", "1")
from object_detection_torch2_b200 import ops, synth  # noqa: E402
from oracle import head  # noqa: E402
import test_gpu_loss as T  # noqa: E402

dev = "cuda"
priors_gpu = ops.default_boxes(dev)
priors_cpu = head.default_boxes()
print("# selection flips vs the oracle (torch CPU fp32), batch 32; gap = |CE_oracle - threshold_oracle| / ulp(threshold)")
print("# kernel            dist seed  images_with_flips  total_flips  max_per_image  max_gap_ulp  gaps_ulp")
for name, kw in (("production", {}), ("exact_math", {"exact_math": True})):
    tot, worst = 0, 0.0
    for dist in ("D1", "D2"):
        for seed in range(10):
            o, t = synth.make_batch(32, seed, dist)
            per_image, gaps = T._flip_report(o, t, priors_cpu, priors_gpu, **kw)
            tot += int(per_image.sum())
            worst = max(worst, float(gaps.max()) if gaps.numel() else 0.0)
            print(f"{name:18s} {dist}   {seed:2d}   {int((per_image > 0).sum()):8d}  {int(per_image.sum()):11d}  {int(per_image.max()):13d}  "
                  f"{(float(gaps.max()) if gaps.numel() else 0.0):11.2f}  {[round(float(x), 2) for x in gaps]}")
    print(f"# {name}: {tot} flips over 20 batches x 32 images x 8732 rows; largest gap {worst:.2f} ulp")
