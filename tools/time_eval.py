"""configs[4] per-GPU share: 619 images (4952 / 8) through decode + score + NMS + TP/FP tallies.  python tools/time_eval.py [N]"""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from object_detection_torch2_b200 import evaluate, ops, synth, utils
N = int(sys.argv[1]) if len(sys.argv) > 1 else 619
dev = torch.device("cuda")
priors_cpu = ops.default_boxes(dev).cpu()
priors = priors_cpu.to(dev)
o = synth.make_outputs(N, 7, "D2")
t = synth.make_targets(N, 7, 20)
o[:32] = synth.plant_detections(o[:32], t[:32], priors_cpu, seed=7)        # some true positives for the bookkeeping
src, gts = o.to(dev), t.to(dev)
buf = src.clone()
tallies = torch.zeros(20, 3, dtype=torch.int64, device=dev)
def run():
    out = utils.postprocess(buf, priors, iou_thresh=0.5)                   # evaluate.py:129-131, in place
    return evaluate.accumulate(out, gts, tallies)                          # evaluate.py:132-151
for _ in range(2):
    buf.copy_(src); run()
times = []
for _ in range(5):
    buf.copy_(src); tallies.zero_(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); run(); e1.record(); torch.cuda.synchronize()
    times.append(e0.elapsed_time(e1))
ms = sorted(times)[len(times) // 2]
parts = []
for _ in range(5):
    buf.copy_(src); tallies.zero_(); torch.cuda.synchronize()
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e0.record(); out = utils.postprocess(buf, priors, iou_thresh=0.5); e1.record(); evaluate.accumulate(out, gts, tallies); e2.record(); torch.cuda.synchronize()
    parts.append((e0.elapsed_time(e1), e1.elapsed_time(e2)))
parts.sort()
print(f"  post-processing {parts[2][0]*1e3:.0f} us, tallies {parts[2][1]*1e3:.0f} us")
byts = N * (2 * 873200 + 873200 + t.shape[1] * 100)
ap = evaluate.average_precision_from_tallies(tallies)
print(f"eval N={N}: {ms*1e3:.0f} us = {N/ms/1e3:.3f} M images/s, {byts/ms/1e6:.0f} GB/s = {byts/ms/1e6/6538.3*100:.1f}% of the HBM roofline "
      f"(3 S per image: post-processing in place + tally pass); TP {int(tallies[:, 0].sum())}, detections {int(tallies[:, 1].sum())}, gt {int(tallies[:, 2].sum())}, mAP {float(ap.nanmean()):.4f}")
