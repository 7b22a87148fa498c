"""Head producer: ssdh_pack_head against the reference's permute / reshape / cat tail.  python tools/time_pack.py [N]"""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from object_detection_torch2_b200 import ops
N = int(sys.argv[1]) if len(sys.argv) > 1 else 32
dev = torch.device("cuda")
LEVELS = [(38, 4), (19, 6), (10, 6), (5, 6), (3, 4), (1, 4)]
ROT = max(2, 512 // N)
sets = [[torch.randn(N, a * 25, m, m, device=dev) for m, a in LEVELS] for _ in range(ROT)]
def ref(xs): return torch.cat([t.permute(0, 2, 3, 1).reshape(N, -1, 25) for t in xs], dim=1)
sets_cl = [[t.contiguous(memory_format=torch.channels_last) for t in xs] for xs in sets]
for name, fn, sets in (("ssdh_pack_head", lambda xs: ops.pack_head(xs, 25), sets), ("permute+reshape+cat", ref, sets),
                       ("ssdh_pack_head (channels-last inputs)", lambda xs: ops.pack_head(xs, 25), sets_cl), ("permute+reshape+cat (channels-last inputs)", ref, sets_cl)):
    for xs in sets: fn(xs)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 20
    e0.record()
    for _ in range(reps):
        for xs in sets: fn(xs)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / (reps * ROT)
    byts = 2 * N * 873200
    print(f"N={N} {name}: {us:.1f} us, {byts/us/1e3:.0f} GB/s = {byts/us/1e3/6538.3*100:.1f}% of the HBM roofline (2 S per image)")
