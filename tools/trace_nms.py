import ctypes, sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from object_detection_torch2_b200 import ops, synth, _lib
N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
dist = sys.argv[2] if len(sys.argv) > 2 else "D2"
dev = torch.device("cuda"); lib = _lib.load(); lib.ssdh_debug_set_nms_trace.argtypes = [ctypes.c_void_p]
priors = ops.default_boxes(dev)
src = synth.make_outputs(N, 5, dist).to(dev); buf = src.clone()
for _ in range(2):
    buf.copy_(src); ops.postprocess_(buf, priors, iou_thresh=0.45)
trace = torch.zeros(N, 16, dtype=torch.int64, device=dev)
buf.copy_(src); torch.cuda.synchronize()
lib.ssdh_debug_set_nms_trace(trace.data_ptr()); ops.postprocess_(buf, priors, iou_thresh=0.45); torch.cuda.synchronize(); lib.ssdh_debug_set_nms_trace(None)
t = trace.cpu().numpy().astype(np.float64)
names = ["A compaction", "B sort", "boxes", "C1 overlaps", "C2 fixed point", "D apply"]
for i, nm in enumerate(names):
    d = t[:, i + 1] - t[:, i]
    print(f"{nm:16s} median {np.median(d):8.0f}  p90 {np.percentile(d, 90):8.0f} cycles")
print("A detail (median since start): keys loaded + ballots %.0f, ov_in cleared %.0f, barrier %.0f" % tuple(np.median(t[:, 12:15] - t[:, :1], axis=0)))
print("total median", np.median(t[:, 6] - t[:, 0]), "rounds median", np.median(t[:, 8]), "max", t[:, 8].max())
