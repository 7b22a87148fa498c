"""Steady-state timing with the next batch's slab prefetched into L2 on a side stream (2-stream CUDA graph)."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from object_detection_torch2_b200 import ops, synth
N = int(sys.argv[1]) if len(sys.argv) > 1 else 32
ROT = 12
dev = torch.device("cuda")
priors = ops.default_boxes(dev)
outs, tgts = [], []
G = 0
for r in range(ROT):
    o, t = synth.make_batch(N, r, "D1"); G = max(G, t.shape[1]); outs.append(o); tgts.append(t)
tgts = [torch.cat([t, torch.zeros(N, G - t.shape[1], 25)], 1).to(dev).contiguous() for t in tgts]
outs = [o.to(dev) for o in outs]
grads = [torch.empty_like(o) for o in outs]
loss = torch.zeros(ROT, device=dev)
def step(i, pipelined=False):
    nxt = (i + 1) % ROT
    ops.multibox_loss_raw(outs[i], tgts[i], priors, loss_out=loss[i], grad_out=grads[i],
                          next_outputs=outs[nxt] if pipelined else None, next_targets=tgts[nxt] if pipelined else None)
for mode in ("plain", "prefetch", "in-kernel"):
    main, side = torch.cuda.Stream(), torch.cuda.Stream()
    main.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(main):
        for i in range(ROT): step(i)
        ops.prefetch_l2(outs[0])
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=main):
            for i in range(ROT):
                if mode == "prefetch":
                    side.wait_stream(main)                    # fork: the prefetch of batch i+1 runs beside loss(i)
                    with torch.cuda.stream(side):
                        ops.prefetch_l2(outs[(i + 1) % ROT])
                        ops.prefetch_l2(tgts[(i + 1) % ROT])
                step(i, mode == "in-kernel")
                if mode == "prefetch":
                    main.wait_stream(side)                    # join
    torch.cuda.current_stream().wait_stream(main); torch.cuda.synchronize()
    for _ in range(3): g.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 40
    torch.cuda.synchronize(); e0.record()
    for _ in range(reps): g.replay()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / (reps * ROT)
    byts = N * (873200 * 2 + G * 100) + 139712
    print(f"N={N} {mode}: {us:.2f} us/step, {N/us:.3f} M img/s, {byts/us/1e3:.0f} GB/s = {byts/us/1e3/6538.3*100:.1f}% of peak, loss {float(loss.mean()):.4f}")
