"""Small driver for profiling the post-processing kernels: python tools/post_probe.py [N] [D1|D2] [reps]"""
import sys, os
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from object_detection_torch2_b200 import ops, synth
N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
dist = sys.argv[2] if len(sys.argv) > 2 else "D2"
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
dev = torch.device("cuda")
priors = ops.default_boxes(dev)
src = synth.make_outputs(N, 5, dist).to(dev)
buf = src.clone()
for r in range(reps):
    buf.copy_(src)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = ops.postprocess_(buf, priors, iou_thresh=0.45, want_lists=True)
    e1.record()
    torch.cuda.synchronize()
    print(f"postprocess N={N} {dist}: {e0.elapsed_time(e1)*1e3:.1f} us, cand/img {float(out.order_cnt.float().mean()):.0f}, kept/img {float(out.keep_cnt.float().mean()):.0f}")
