"""Steady-state timing of ssdh_postprocess from a CUDA graph (restore copy outside the timed region): python tools/time_post.py [N] [D1|D2] [reps]"""
import os, statistics, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from object_detection_torch2_b200 import ops, synth
N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
dist = sys.argv[2] if len(sys.argv) > 2 else "D2"
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 20
dev = torch.device("cuda")
priors = ops.default_boxes(dev)
src = synth.make_outputs_device(N, 5, dist, dev)
bufs = [src.clone() for _ in range(2)]
graphs = []
s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    for b in bufs:
        ops.postprocess_(b, priors, iou_thresh=0.45, want_lists=True); b.copy_(src)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            out = ops.postprocess_(b, priors, iou_thresh=0.45, want_lists=True)
        graphs.append((g, out))
torch.cuda.current_stream().wait_stream(s); torch.cuda.synchronize()
times = []
for r in range(reps + 3):
    b = bufs[r % 2]; g, out = graphs[r % 2]
    b.copy_(src); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    if r >= 3: times.append(e0.elapsed_time(e1))
ms = statistics.median(times)
print(f"postprocess N={N} {dist} parts={os.environ.get('SSDH_POST_PARTS', '1')}: median {ms*1e3:.1f} us (min {min(times)*1e3:.1f}), hbm frac {N*2*873200/(ms*1e-3)/1e9/6538.3:.4f}, kept/img {float(out.keep_cnt.float().mean()):.1f}")
