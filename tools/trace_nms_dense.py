"""Phase timeline of nms_kernel (dense images): python tools/trace_nms_dense.py [N]"""
import ctypes, sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from object_detection_torch2_b200 import ops, synth, _lib
N = int(sys.argv[1]) if len(sys.argv) > 1 else 32
dev = torch.device("cuda"); lib = _lib.load(); lib.ssdh_debug_set_nms_trace.argtypes = [ctypes.c_void_p]
priors = ops.default_boxes(dev)
src = synth.make_outputs(N, 5, "D1").to(dev); buf = src.clone()
for _ in range(2):
    buf.copy_(src); ops.postprocess_(buf, priors, iou_thresh=0.45)
trace = torch.zeros(N, 16, dtype=torch.int64, device=dev)
buf.copy_(src); torch.cuda.synchronize()
lib.ssdh_debug_set_nms_trace(trace.data_ptr()); ops.postprocess_(buf, priors, iou_thresh=0.45); torch.cuda.synchronize(); lib.ssdh_debug_set_nms_trace(None)
t = trace.cpu().numpy().astype(np.float64)
for i, nm in enumerate(["A compaction", "B sort", "C greedy rounds"]):
    print(f"{nm:18s} median {np.median(t[:, i + 1] - t[:, i]):9.0f} cycles")
for i, nm in zip((9, 10, 11), ("  sweep over kept list", "  bit matrix", "  walk + append")):
    print(f"{nm:22s} median {np.median(t[:, i]):9.0f} cycles")
