"""Build container only: the oracle port (oracle/head.py) against the UNMODIFIED reference (/root/reference/src) on the
headline step -- SSD.loss forward + backward, batch 32, dist D1, all host threads -- so that bench.py's CPU arm
(`kind: "port"`, the reference itself does not travel to the GPU box) can state how the port relates to the real thing.

    python tools/time_port_vs_reference.py [passes]      -> profiles/r02_port_vs_reference.json
"""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from object_detection_torch2_b200 import synth  # noqa: E402
from oracle import head, ref_loader  # noqa: E402

passes = int(sys.argv[1]) if len(sys.argv) > 1 else 5
threads = os.cpu_count() or 1
torch.set_num_threads(threads)
ref = ref_loader.load()
priors = head.default_boxes()
o, t = synth.make_batch(32, 0, "D1")


def port():
    head.multibox_loss(o, t, priors, want_grad=True)


def reference():
    x = o.clone().requires_grad_(True)
    ref.net.loss(outputs=x, targets=t, default_bboxes=priors).backward()


def rate(fn):
    fn()
    best = 1e9
    for _ in range(passes):
        t0 = time.perf_counter()
        fn()
        best = min(best, time.perf_counter() - t0)
    return 32 / best


res = {"workload": "SSD.loss forward + backward, batch 32, G<=20, dist D1 (seed 0)", "cores": threads, "passes": passes,
       "port_images_per_s": rate(port), "reference_images_per_s": rate(reference)}
res["port_over_reference"] = res["port_images_per_s"] / res["reference_images_per_s"]
res["where"] = "build container (no GPU), torch %s CPU" % torch.__version__
print(json.dumps(res, indent=1))
with open(os.path.join(ROOT, "profiles", "r02_port_vs_reference.json"), "w") as f:
    json.dump(res, f, indent=1)
