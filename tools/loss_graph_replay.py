"""Steady-state replay of the headline step graph (48 steps over 12 rotating buffer pairs, batch 32, dist D1) for profilers:

    ncu --graph-profiling graph --cache-control none --clock-control none --profile-from-start off \
        --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --csv --log-file out.csv \
        python tools/loss_graph_replay.py

With --graph-profiling graph one whole graph launch (48 kernels) is ONE profiled workload, so the DRAM counters are the
steady-state traffic of 48 consecutive steps: every step's slab is read once (from HBM, or through the previous step's
L2 prefetch), every gradient is written back as the rotation pushes it out of the L2.  bytes / 48 = traffic per launch.
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from object_detection_torch2_b200 import ops, synth  # noqa: E402

N, ROT, STEPS = 32, 12, 48
dev = torch.device("cuda")
priors = ops.default_boxes(dev)
outs, tgts = [], []
for r in range(ROT):
    o, t = synth.make_batch(N, r, "D1")
    outs.append(o.to(dev))
    tgts.append(synth.pad_targets(t, 20).to(dev).contiguous())
grads = [torch.empty_like(o) for o in outs]
loss = torch.zeros(STEPS, device=dev)


def step(k):
    i, nxt = k % ROT, (k + 1) % ROT
    ops.multibox_loss_raw(outs[i], tgts[i], priors, n_global=N, loss_out=loss[k], grad_out=grads[i], next_outputs=outs[nxt], next_targets=tgts[nxt])


s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    for i in range(ROT):
        step(i)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=s):
        for i in range(STEPS):
            step(i)
torch.cuda.current_stream().wait_stream(s)
torch.cuda.synchronize()
for _ in range(4):
    g.replay()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    g.replay()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("loss", float(loss.mean()))
