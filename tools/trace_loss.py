"""Phase timeline of multibox_loss_kernel (debug hook ssdh_debug_set_loss_trace): per-CTA SM-clock stamps."""
import ctypes, sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from object_detection_torch2_b200 import ops, synth, _lib

N = int(sys.argv[1]) if len(sys.argv) > 1 else 32
dist = sys.argv[2] if len(sys.argv) > 2 else "D1"
dev = torch.device("cuda")
lib = _lib.load()
lib.ssdh_debug_set_loss_trace.argtypes = [ctypes.c_void_p]
priors = ops.default_boxes(dev)
o, t = synth.make_batch(N, 0, dist)
o, t = o.to(dev), t.to(dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for _ in range(3):
    ops.multibox_loss_raw(o, t, priors)
CL = ops.device_info()["loss_cluster_size"]
trace = torch.zeros(N * CL, 128, dtype=torch.int64, device=dev)
flush.fill_(1)
torch.cuda.synchronize()
lib.ssdh_debug_set_loss_trace(trace.data_ptr())
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
ops.multibox_loss_raw(o, t, priors)
e1.record()
torch.cuda.synchronize()
lib.ssdh_debug_set_loss_trace(None)
tr = trace.cpu().numpy()
print("kernel ms", e0.elapsed_time(e1))
names = ["start", "setup", "gt+tma issue", "match", "slab0 arrived", "rows done", "csync1", "gather0", "select done", "sums+csync2", "grad+stores issued", "stores read out", "end"]
clk = tr[:, 1:13].astype(np.float64)
rel = (clk - clk[:, :1])
g0, g1 = tr[:, 0], tr[:, 14]
print("globaltimer: first start -> last end = %.2f us; start spread %.2f us" % ((g1.max() - g0.min()) / 1e3, (g0.max() - g0.min()) / 1e3))
print("per-CTA duration (globaltimer) median %.2f us max %.2f us" % (np.median(g1 - g0) / 1e3, (g1 - g0).max() / 1e3))
print("%-22s %10s %10s %10s   (cycles since CTA start; delta median)" % ("phase", "median", "p10", "max"))
prev = np.zeros(len(rel))
for i, nm in enumerate(names[1:]):
    col = rel[:, i]
    print("%-22s %10.0f %10.0f %10.0f   +%.0f" % (nm, np.median(col), np.percentile(col, 10), col.max(), np.median(col - prev)))
    prev = col
sel = tr[:, 16:19].astype(np.float64) - clk[:, :1]
print("select detail (median cycles since CTA start): list built %.0f, lists received %.0f, gathered %.0f" % tuple(np.median(sel, axis=0)))
print("overflow fallbacks:", int(tr[:, 15].sum()))
sm = tr[:, 13]
cnt = np.bincount(sm.astype(int), minlength=148)
print("CTAs per SM histogram:", np.bincount(cnt))
# clusters: are the 8 CTAs of an image spread over distinct SMs?
for n in range(min(N, 3)):
    print("image", n, "SMs", sm[n * CL:(n + 1) * CL].tolist(), "end cycles", rel[n * CL:(n + 1) * CL, -1].astype(int).tolist())
