"""One ssdh_multibox_loss launch at a time (the regime of a training step): cold L2, and right behind ssdh_pack_head.
Same method as bench.py's `isolated` leg: (graph of K x [flush, (producer,) loss] - graph of K x [flush(, producer)]) / K.

    python tools/time_loss_isolated.py [N] [dist]        (A/B: python tools/with_lib.py <lib.so> tools/time_loss_isolated.py)
"""
import os, statistics, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from object_detection_torch2_b200 import ops, synth
N = int(sys.argv[1]) if len(sys.argv) > 1 else 32
dist = sys.argv[2] if len(sys.argv) > 2 else "D1"
K, ROT, ROW = 12, 4, 25
dev = torch.device("cuda")
priors = ops.default_boxes(dev)
outs, tgts, G = [], [], 0
for r in range(ROT):
    o, t = synth.make_batch(N, r, dist); G = max(G, t.shape[1]); outs.append(o); tgts.append(t)
tgts = [torch.cat([t, torch.zeros(N, G - t.shape[1], 25)], 1).to(dev).contiguous() for t in tgts]
outs = [o.to(dev) for o in outs]
grads = [torch.empty_like(o) for o in outs]
loss = torch.zeros(K, device=dev)
flush = torch.empty(512 * 1024 * 1024 // 4, dtype=torch.float32, device=dev).normal_()
sink = torch.zeros((), device=dev)
levels = [torch.randn(N, a * ROW, m, m, device=dev) for m, a in ((38, 4), (19, 6), (10, 6), (5, 6), (3, 4), (1, 4))]


def capture(body):
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s), torch.no_grad():
        body()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            body()
    torch.cuda.current_stream().wait_stream(s); torch.cuda.synchronize()
    return g


def timed(g):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)


def bodies(producer, with_loss):
    def body():
        for k in range(K):
            torch.sum(flush, dim=0, out=sink)
            o = ops.pack_head(levels, ROW) if producer else outs[k % ROT]
            if with_loss:
                ops.multibox_loss_raw(o, tgts[k % ROT], priors, n_global=N, want_grad=True, loss_out=loss[k], grad_out=grads[k % ROT])
    return body


for producer in (False, True):
    ga, gb = capture(bodies(producer, True)), capture(bodies(producer, False))
    ta, tb = [], []
    for _ in range(9):
        ta.append(timed(ga)); tb.append(timed(gb))
    us = (statistics.median(ta) - statistics.median(tb)) / K * 1e3
    byts = N * (873200 * 2 + G * 100) + 139712
    print(f"N={N} {dist} {'behind ssdh_pack_head (input in L2)' if producer else 'cold L2'}: {us:.2f} us per launch = {byts / us / 1e3 / 6538.3 * 100:.1f}% of peak")
