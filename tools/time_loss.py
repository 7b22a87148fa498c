"""Steady-state timing of ssdh_multibox_loss variants: python tools/time_loss.py [N] [dist] [pad ground truth to G rows] [stable]
("stable": the pipelined entry point -- inputs vouched for, next batch prefetched -- as bench.py uses it)"""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from object_detection_torch2_b200 import ops, synth
N = int(sys.argv[1]) if len(sys.argv) > 1 else 32
dist = sys.argv[2] if len(sys.argv) > 2 else "D1"
pad_g = int(sys.argv[3]) if len(sys.argv) > 3 else 0
stable = len(sys.argv) > 4 and sys.argv[4] == "stable"
ROT = max(2, 384 // N)
dev = torch.device("cuda")
priors = ops.default_boxes(dev)
outs, tgts = [], []
G = pad_g
for r in range(ROT):
    o, t = synth.make_batch(N, r, dist); G = max(G, t.shape[1]); outs.append(o); tgts.append(t)
tgts = [torch.cat([t, torch.zeros(N, G - t.shape[1], 25)], 1).to(dev).contiguous() for t in tgts]
outs = [o.to(dev) for o in outs]
grads = [torch.empty_like(o) for o in outs]
loss = torch.zeros(ROT, device=dev)
for want_grad in (True, False):
    def step(i):
        nxt = (i + 1) % ROT
        ops.multibox_loss_raw(outs[i], tgts[i], priors, want_grad=want_grad, loss_out=loss[i], grad_out=grads[i] if want_grad else None,
                              next_outputs=outs[nxt] if stable else None, next_targets=tgts[nxt] if stable else None)
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for i in range(ROT): step(i)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for i in range(ROT): step(i)
    torch.cuda.current_stream().wait_stream(s); torch.cuda.synchronize()
    for _ in range(3): g.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = max(4, 4000 // (ROT * max(1, N // 32)))
    torch.cuda.synchronize(); e0.record()
    for _ in range(reps): g.replay()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / (reps * ROT)
    byts = N * (873200 * (2 if want_grad else 1) + G * 100) + 139712
    print(f"N={N} {dist} G={G} {'stable' if stable else 'plain'} want_grad={want_grad}: {us:.2f} us/step, {N/us:.3f} M img/s, {byts/us/1e3:.0f} GB/s = {byts/us/1e3/6538.3*100:.1f}% of peak")
