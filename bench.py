#!/usr/bin/env python
"""Benchmark of the SSD300 detection-head hot path on B200 (BASELINE.json metric: images/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Headline workload (BASELINE.json configs[1]): one training step of the head on a batch of 32 images per GPU --
IoU matching + hard-negative-mined MultiBox loss, forward AND gradient w.r.t. the (32, 8732, 25) head output --
on synthetic VOC-shaped inputs (random GT, 1-20 boxes per image; N(0,1) "random-init" head outputs, dist D1).
A step is one launch of ssdh_multibox_loss through the C ABI.  Images are sharded by GPU (weak scaling: 32 per
GPU); the only collective is an all-reduce of the packed per-step loss scalars, batched per graph replay on a
side stream.

L2 policy: the timed loop rotates over ROT distinct (outputs, grad) buffer pairs whose footprint (ROT x 56 MB) is
several times the 126 MB L2, so every step streams its slab from HBM ("inputs larger than L2").
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

P, C, ROW = 8732, 21, 25
BATCH = 32                  # per GPU (configs[1])
POST_BATCH = 256            # configs[2]
ROT = 12                    # rotating buffer pairs: 12 x (27.9 + 27.9) MB = 671 MB >> 126 MB L2
GRAPH_STEPS = 48            # steps per CUDA-graph replay (4 rotations): the hand-over between replays is not pipelined
SLAB = P * ROW * 4          # 873 200 B per image


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:  # noqa: BLE001
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                          "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:  # noqa: BLE001
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for ln in self.lines:
            parts = [x.strip() for x in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for name, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------------------------------------------------
def cpu_reference_rate(steps: int, warmup: int, budget_s: float = 20.0):
    """The reference's CPU path for the headline step (match + MultiBox loss forward + backward), timed on this
    box's host cores with all threads.  /root/reference is pure Python and does not travel to the GPU box, so this is
    the committed oracle port (oracle/head.py, proven equal to the reference in tests/)."""
    from object_detection_torch2_b200 import synth
    from oracle import head
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    priors = head.default_boxes()
    o, t = synth.make_batch(BATCH, 0, "D1")
    t0 = time.perf_counter()
    head.multibox_loss(o[:2], t[:2], priors, want_grad=True)
    per_img = (time.perf_counter() - t0) / 2
    per_step_budget = budget_s / max(1, steps + warmup)
    n = int(max(1, min(BATCH, per_step_budget / max(per_img, 1e-4))))
    for _ in range(warmup):
        head.multibox_loss(o[:n], t[:n], priors, want_grad=True)
    t0 = time.perf_counter()
    for _ in range(steps):
        head.multibox_loss(o[:n], t[:n], priors, want_grad=True)
    dt = time.perf_counter() - t0
    return n * steps / dt, dt / steps * 1e3, threads, f"{n} of the {BATCH} images of the step per pass, {steps} passes, torch CPU fp32"


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    rate, ms, threads, sample = cpu_reference_rate(args.steps, args.warmup, budget_s=90.0)
    line = {"impl": "reference", "metric": "images/s for SSD300 match+MultiBox loss (fwd+grad) training step", "value": rate,
            "unit": "images/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "SSD300 head training step: IoU match + MultiBox loss fwd+grad, batch 32, G<=20, dist D1",
                       "device": "host CPU"},
            "cpu_baseline": {"value": rate, "unit": "images/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": rate, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=4800)
    ap.add_argument("--warmup", type=int, default=48)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--dist", default="D1", choices=["D1", "D2"])
    ap.add_argument("--no-extras", action="store_true", help="skip the post-processing and CPU-baseline legs")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        run_reference(args)
        return

    import torch.distributed as dist
    from object_detection_torch2_b200 import ops, parallel, synth
    from object_detection_torch2_b200.model import SSD

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n_global = BATCH * world

    priors = ops.default_boxes(dev)
    # ROT distinct synthetic batches (seed differs per rank and slot), resident in HBM before timing starts
    outs, tgts, grads = [], [], []
    G = 0
    for r in range(ROT):
        o, t = synth.make_batch(BATCH, 1000 * rank + r, args.dist)
        G = max(G, t.shape[1])
        outs.append(o)
        tgts.append(t)
    tgts = [torch.cat([t, torch.zeros(BATCH, G - t.shape[1], ROW)], dim=1) if t.shape[1] < G else t for t in tgts]
    outs = [o.to(dev) for o in outs]
    tgts = [t.to(dev).contiguous() for t in tgts]
    grads = [torch.empty_like(o) for o in outs]
    losses = torch.zeros(GRAPH_STEPS, dtype=torch.float32, device=dev)

    def step(k):
        # software pipelining: while batch i is on chip the kernel asks the L2 for batch i+1 (HBM is idle then)
        i, nxt = k % ROT, (k + 1) % ROT
        ops.multibox_loss_raw(outs[i], tgts[i], priors, a=1.0, threshold=0.25, n_global=n_global, want_grad=True,
                              loss_out=losses[k], grad_out=grads[i], next_outputs=outs[nxt], next_targets=tgts[nxt])

    # one CUDA graph = GRAPH_STEPS consecutive steps cycling through the ROT buffer pairs
    cap = torch.cuda.Stream(device=dev)
    cap.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(cap):
        for i in range(ROT):
            step(i)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=cap):
            for i in range(GRAPH_STEPS):
                step(i)
    torch.cuda.current_stream().wait_stream(cap)
    torch.cuda.synchronize()

    replays = max(1, -(-args.steps // GRAPH_STEPS))
    steps = replays * GRAPH_STEPS              # exactly `steps` timed steps (rounded up to whole replays)
    warm_replays = max(1, -(-args.warmup // GRAPH_STEPS))
    reducer = parallel.ScalarAllReducer(width=GRAPH_STEPS, window=1, device=dev, dtype=torch.float32)

    def run(n_replays):
        for _ in range(n_replays):
            graph.replay()
            if world > 1:
                reducer.push(losses)           # one packed all-reduce per replay, async on the side stream
        if world > 1:
            return reducer.flush()
        return None

    run(warm_replays)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    reduced = run(replays)
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    clocks = sampler.stop() if rank == 0 else None
    elapsed_ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(elapsed_ms, op=dist.ReduceOp.MAX)
    elapsed_ms = float(elapsed_ms)
    ms_per_step = elapsed_ms / steps
    value = n_global * steps / (elapsed_ms * 1e-3)
    loss_value = float(reduced[-1].sum()) / GRAPH_STEPS if reduced is not None else float(losses.mean())

    # ---- roofline of the dominant (only) kernel of the step ----------------------------------------------------
    peak, peak_src = measured_peak()
    alg_bytes = BATCH * (2 * SLAB + G * ROW * 4) + P * 16        # read slab + write grad + GT rows, priors once
    achieved = alg_bytes / (ms_per_step * 1e-3) / 1e9
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "loss_kernel_dram.json")) as f:
            traffic = json.load(f).get("dram_bytes_per_launch")
    except Exception:  # noqa: BLE001
        pass
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "kernel": "multibox_loss_kernel<21>", "algorithmic_bytes_per_launch": alg_bytes, "peak_source": peak_src,
                "avg_launch_us": ms_per_step * 1e3}

    # ---- e2e: public API (SSD.loss + backward) from pinned HOST buffers, copies inside the timed region ------------
    e2e_steps = max(8, min(steps, 64))
    net = SSD.__new__(SSD)
    torch.nn.Module.__init__(net)
    h_out = [o.cpu().pin_memory() for o in outs[:4]]
    h_tgt = [t.cpu().pin_memory() for t in tgts[:4]]
    h_loss = torch.zeros(e2e_steps, dtype=torch.float32).pin_memory()
    d_out = [torch.empty_like(outs[0]) for _ in range(2)]
    d_tgt = [torch.empty_like(tgts[0]) for _ in range(2)]

    # double buffering: the H2D copy of step i + 1 runs on its own stream under the compute of step i
    copy_s = torch.cuda.Stream(device=dev)
    ev_ready = [torch.cuda.Event() for _ in range(2)]
    ev_free = [torch.cuda.Event() for _ in range(2)]
    for ev in ev_free:
        ev.record()

    def e2e_step(i, record=True):
        b = i & 1
        cur = torch.cuda.current_stream()
        with torch.cuda.stream(copy_s):
            copy_s.wait_event(ev_free[b])                 # the step that last read this buffer pair has finished
            d_out[b].copy_(h_out[i % 4], non_blocking=True)
            d_tgt[b].copy_(h_tgt[i % 4], non_blocking=True)
            ev_ready[b].record(copy_s)
        cur.wait_event(ev_ready[b])
        x = d_out[b].detach().requires_grad_(True)
        loss = net.loss(outputs=x, targets=d_tgt[b], default_bboxes=priors)
        loss.backward()
        ev_free[b].record(cur)
        if record:
            h_loss[i].copy_(loss.detach(), non_blocking=True)
        return x.grad

    for i in range(3):
        e2e_step(i, record=False)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0.record()
    for i in range(e2e_steps):
        e2e_step(i)
    e1.record()
    torch.cuda.synchronize()
    e2e_ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
    e2e = {"value": n_global * e2e_steps / (float(e2e_ms) * 1e-3), "unit": "images/s",
           "h2d_bytes_per_step": outs[0].numel() * 4 + tgts[0].numel() * 4, "d2h_bytes_per_step": 4, "steps": e2e_steps,
           "api": "SSD.loss(outputs, targets, default_bboxes) + loss.backward(), pinned host -> device copy per step (double-buffered on a copy stream)",
           "loss_check": float(h_loss[-1])}

    line = {"metric": "images/s for SSD300 match+MultiBox loss (fwd+grad) training step", "value": value, "unit": "images/s",
            "n_gpus": world, "steps": steps, "warmup": warm_replays * GRAPH_STEPS, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "SSD300 head training step: IoU match + MultiBox loss fwd+grad, batch 32 per GPU, G<=20, dist " + args.dist,
                       "global_batch": n_global, "priors": P, "classes": C, "gt_rows": G,
                       "l2": f"inputs larger than L2: rotation over {ROT} (outputs, grad) buffer pairs = {ROT * 2 * BATCH * SLAB / 1e6:.0f} MB",
                       "launch": f"CUDA graph of {GRAPH_STEPS} steps, C ABI ssdh_multibox_loss_pipelined (in-kernel L2 prefetch of the next batch, programmatic dependent launch between steps)", "parallelism": f"dp{world} (images sharded, scalar all-reduce per replay)"},
            "roofline": roofline, "e2e": e2e, "gpu_launches": steps, "clocks": clocks, "loss": loss_value}

    if rank == 0 and world == 1 and not args.no_extras:
        line["post"] = bench_post(ops, synth, priors, dev, peak)
        line["pack_head"] = bench_pack(ops, dev, peak)
        rate, ms, threads, sample = cpu_reference_rate(3, 1, budget_s=15.0)
        line["cpu_baseline"] = {"value": rate, "unit": "images/s", "cores": threads, "kind": "port", "sample": sample}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def bench_post(ops, synth, priors, dev, peak):
    """configs[2]: batch-256 inference post-processing (decode + score + NMS fused, in place).  D2 = trained-like
    logits (a few hundred candidates per image, HBM-bound), D1 = random-init logits (~8 300 candidates per image, the
    greedy suppression is fp32-compute-bound by construction -- SURVEY 7.3-4).  Each call is replayed from a CUDA
    graph (3 kernels, no host gaps) on a freshly restored input; the restore copy is outside the timed region."""
    res = {}
    # D2 / D1: the reference's own call (class-agnostic, no score cut, no top-k; iou_thresh passed explicitly);
    # D2_north_star: the wording of configs[2] -- score threshold 0.01, per-class NMS 0.45, top-200 (opt-in kwargs)
    variants = (("D2", "D2", 20, {}), ("D2_north_star", "D2", 20, {"score_thresh": 0.01, "top_k": 200, "per_class": True}), ("D1", "D1", 3, {}))
    for key, dist_name, reps, kw in variants:
        n = POST_BATCH if dist_name == "D2" else 32
        src = synth.make_outputs(n, 5, dist_name).to(dev)
        bufs = [src.clone() for _ in range(2 if dist_name == "D2" else 1)]        # 2 x 224 MB > L2
        order = torch.empty(n, P, dtype=torch.int32, device=dev)
        graphs = []
        s = torch.cuda.Stream(device=dev)
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for b in bufs:
                out = ops.postprocess_(b, priors, iou_thresh=0.45, want_lists=True, **kw)      # warm-up + workspace allocation
                b.copy_(src)
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=s):
                    out = ops.postprocess_(b, priors, iou_thresh=0.45, want_lists=True, **kw)
                graphs.append((g, out))
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        times = []
        for r in range(reps + 3):
            b = bufs[r % len(bufs)]
            g, out = graphs[r % len(bufs)]
            b.copy_(src)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            g.replay()
            e1.record()
            torch.cuda.synchronize()
            if r >= 3:
                times.append(e0.elapsed_time(e1))
        ms = statistics.median(times)
        alg = n * 2 * SLAB
        res[key] = {"batch": n, "ms": ms, "images_per_s": n / (ms * 1e-3), "hbm_frac": alg / (ms * 1e-3) / 1e9 / peak,
                          "candidates_per_image": float(out.order_cnt.float().mean()), "kept_per_image": float(out.keep_cnt.float().mean()),
                          "launches_per_call": 3}
    return res


def bench_pack(ops, dev, peak):
    """the six detector outputs -> (N, 8732, 25) in one pass (ssdh_pack_head) against the reference's
    permute / reshape / cat tail (ssd.py:103-104), batch 256, rotating inputs larger than L2."""
    n = POST_BATCH
    levels = [(38, 4), (19, 6), (10, 6), (5, 6), (3, 4), (1, 4)]
    sets = [[torch.randn(n, a * ROW, m, m, device=dev) for m, a in levels] for _ in range(2)]
    res = {"batch": n}
    for name, fn in (("ms", lambda xs: ops.pack_head(xs, ROW)),
                     ("torch_ms", lambda xs: torch.cat([t.permute(0, 2, 3, 1).reshape(n, -1, ROW) for t in xs], dim=1))):
        for xs in sets:
            fn(xs)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 10
        e0.record()
        for _ in range(reps):
            for xs in sets:
                fn(xs)
        e1.record()
        torch.cuda.synchronize()
        res[name] = e0.elapsed_time(e1) / (reps * len(sets))
    res["hbm_frac"] = n * 2 * SLAB / (res["ms"] * 1e-3) / 1e9 / peak
    return res


if __name__ == "__main__":
    main()
