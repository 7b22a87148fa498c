"""Phase timeline of multibox_loss_kernel in the steady-state loop (CUDA graph over rotating buffers)."""
import ctypes, sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from object_detection_torch2_b200 import ops, synth, _lib
N = 32; ROT = 12
pipelined = len(sys.argv) > 1 and sys.argv[1] == "pipelined"
dev = torch.device("cuda"); lib = _lib.load(); lib.ssdh_debug_set_loss_trace.argtypes = [ctypes.c_void_p]
priors = ops.default_boxes(dev)
outs, tgts = [], []; G = 0
for r in range(ROT):
    o, t = synth.make_batch(N, r, "D1"); G = max(G, t.shape[1]); outs.append(o); tgts.append(t)
tgts = [torch.cat([t, torch.zeros(N, G - t.shape[1], 25)], 1).to(dev).contiguous() for t in tgts]
outs = [o.to(dev) for o in outs]; grads = [torch.empty_like(o) for o in outs]; loss = torch.zeros(ROT, device=dev)
CL = ops.device_info()["loss_cluster_size"]
traces = [torch.zeros(N * CL, 128, dtype=torch.int64, device=dev) for _ in range(ROT)]
def step(i):
    lib.ssdh_debug_set_loss_trace(traces[i].data_ptr())
    nxt = (i + 1) % ROT
    ops.multibox_loss_raw(outs[i], tgts[i], priors, loss_out=loss[i], grad_out=grads[i],
                          next_outputs=outs[nxt] if pipelined else None, next_targets=tgts[nxt] if pipelined else None)
s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    for i in range(ROT): step(i)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=s):
        for i in range(ROT): step(i)
lib.ssdh_debug_set_loss_trace(None)
torch.cuda.current_stream().wait_stream(s); torch.cuda.synchronize()
for _ in range(5): g.replay()
torch.cuda.synchronize()
names = ["setup", "gt+tma issue", "match", "slab0 arrived", "rows done", "csync1", "gather0", "select done", "sums+csync2", "grad+stores issued", "stores read out", "end"]
tr = traces[6].cpu().numpy(); tr5 = traces[5].cpu().numpy()
clk = tr[:, 1:13].astype(np.float64); rel = clk - clk[:, :1]
prev = np.zeros(len(rel))
for i, nm in enumerate(names):
    col = rel[:, i]
    print("%-22s median %8.0f  p10 %8.0f  max %8.0f   +%.0f" % (nm, np.median(col), np.percentile(col, 10), col.max(), np.median(col - prev)))
    prev = col
sel = tr[:, 16:19].astype(np.float64) - clk[:, :1]
print("select detail (median cycles since CTA start): list built %.0f, lists received %.0f, gathered %.0f" % tuple(np.median(sel, axis=0)))
st = tr[:, 48:52].astype(np.float64) - clk[:, :1]
print("startup detail (median): smem init %.0f, TMA issued %.0f, cluster wait done %.0f, ground truth parsed %.0f" % tuple(np.median(st, axis=0)))
w = tr[:, 24:48].astype(np.float64) - clk[:, :1]
print("row phase end per warp (median over CTAs, cycles since CTA start):", np.median(w, axis=0).astype(int).tolist())
print("  per CTA: max over warps median %.0f, median over warps median %.0f" % (np.median(w.max(axis=1)), np.median(np.median(w, axis=1))))
g0, g1 = tr[:, 0], tr[:, 14]
print("step 6: first start -> last end %.2f us; gap from step 5 last end to step 6 first start %.2f us" % ((g1.max() - g0.min()) / 1e3, (g0.min() - tr5[:, 14].max()) / 1e3))
