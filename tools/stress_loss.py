"""Determinism stress for ssdh_multibox_loss: replay the steady-state graph many times, every replay must reproduce the
first one bit for bit (loss, per-image statistics, a checksum of every gradient).  python tools/stress_loss.py [replays] [N]"""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from object_detection_torch2_b200 import ops, synth
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 200
N = int(sys.argv[2]) if len(sys.argv) > 2 else 32
ROT = 12
dev = torch.device("cuda")
priors = ops.default_boxes(dev)
outs, tgts, G = [], [], 0
for r in range(ROT):
    o, t = synth.make_batch(N, 100 + r, "D1" if r % 2 == 0 else "D2"); G = max(G, t.shape[1]); outs.append(o); tgts.append(t)
tgts = [torch.cat([t, torch.zeros(N, G - t.shape[1], 25)], 1).to(dev).contiguous() for t in tgts]
outs = [o.to(dev) for o in outs]
grads = [torch.empty_like(o) for o in outs]
loss = torch.zeros(ROT, device=dev)
stats = [torch.zeros(N, 32, dtype=torch.uint8, device=dev) for _ in range(ROT)]
def step(i):
    nxt = (i + 1) % ROT
    ops.multibox_loss_raw(outs[i], tgts[i], priors, loss_out=loss[i], grad_out=grads[i], want_stats=True, stats_out=stats[i],
                          next_outputs=outs[nxt], next_targets=tgts[nxt])
def digest():
    return (loss.clone(), torch.stack([g.view(torch.int32).sum(dtype=torch.int64) for g in grads]), torch.stack(stats).clone())
s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    for i in range(ROT): step(i)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=s):
        for i in range(ROT): step(i)
torch.cuda.current_stream().wait_stream(s); torch.cuda.synchronize()
g.replay(); torch.cuda.synchronize()
ref = digest()
bad = 0
for k in range(reps):
    for gr in grads: gr.fill_(float("nan"))
    g.replay(); torch.cuda.synchronize()
    cur = digest()
    bits = lambda t: t.view(torch.int32) if t.dtype == torch.float32 else t
    if not all(torch.equal(bits(a), bits(b)) for a, b in zip(cur, ref)):
        bad += 1
print(f"stress: {reps} replays x {ROT} batches of {N}: {bad} replays differed from the first")
sys.exit(1 if bad else 0)
