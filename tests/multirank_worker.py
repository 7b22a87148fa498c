"""Worker of tests/test_gpu_multirank.py: runs under torch.distributed.run with one process per GPU (NCCL) and checks that
sharding the batch by image over the ranks reproduces the single-GPU results (SURVEY 8e, section 4 "multi-GPU" level):

  * per-image loss / thresholds / selection counts of every shard == the same images of a single-GPU launch, bit for bit;
  * the all-reduced loss scalar == the single-GPU scalar (fp32 summation order differs: 1e-6 relative);
  * every rank's gradient == the corresponding slice of the single-GPU gradient, bit for bit;
  * all-reduced int64 TP / detection / ground-truth tallies == the single-GPU tallies, exactly;
  * DetectionEvaluator.compute (tally all-reduce + detection all-gather over NCCL) == the single-GPU evaluator.

Exit code 0 = every check passed on every rank.
"""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from object_detection_torch2_b200 import evaluate, ops, parallel, synth  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    priors = ops.default_boxes(dev)
    N = 8 * world + 3                                    # uneven shards on purpose
    fails = []

    def check(ok, what):
        if not bool(ok):
            fails.append(what)

    # ---- training step ------------------------------------------------------------------------------------------
    o, t = synth.make_batch(N, 77, "D1")
    o, t = o.to(dev), t.to(dev).contiguous()
    lo, hi = parallel.shard_bounds(N, world, rank)
    so, st = parallel.shard_batch(o, t, world, rank)
    loss_l, grad_l, stats_l = ops.multibox_loss_raw(so.contiguous(), st.contiguous(), priors, n_global=N, want_grad=True, want_stats=True)
    loss_g = parallel.global_loss(loss_l)
    loss_1, grad_1, stats_1 = ops.multibox_loss_raw(o, t, priors, n_global=N, want_grad=True, want_stats=True)
    check(torch.equal(stats_l, stats_1[lo:hi]), "per-image stats of the shard differ from the single-GPU launch")
    check(torch.equal(grad_l, grad_1[lo:hi]), "shard gradient differs from the single-GPU slice")
    rel = abs(float(loss_g) - float(loss_1)) / max(abs(float(loss_1)), 1e-12)
    check(rel <= 1e-6, f"all-reduced loss {float(loss_g)} vs single-GPU {float(loss_1)} (rel {rel:.2e})")
    # fp64 sum of the per-image losses in image order reproduces the single-GPU scalar's own definition
    per_img = torch.from_numpy(ops.stats_to_numpy(stats_l)["loss"].copy()).to(dev).double()
    gathered = [torch.zeros(parallel.shard_bounds(N, world, r)[1] - parallel.shard_bounds(N, world, r)[0], dtype=torch.float64, device=dev)
                for r in range(world)]
    dist.all_gather(gathered, per_img)
    total64 = float(torch.cat(gathered).sum() / N)
    check(abs(total64 - float(loss_1)) <= 1e-6 * abs(float(loss_1)), "fp64 sum of the gathered per-image losses differs")

    # the packed asynchronous reducer used by training loops
    red = parallel.ScalarAllReducer(width=2, window=2, device=dev)
    for k in range(3):
        red.push(torch.stack([loss_l.double() * (k + 1), torch.tensor(float(hi - lo), device=dev, dtype=torch.float64)]))
    rows = red.flush()
    check(rows.shape == (3, 2) and all(abs(float(rows[k, 0]) - (k + 1) * float(loss_1)) <= 2e-6 * abs(float(loss_1)) for k in range(3))
          and rows[:, 1].tolist() == [float(N)] * 3, "ScalarAllReducer rows are wrong")

    # the same scalars through the NVLink exchange fused into the loss kernel (csrc/exchange.cu): rank-ordered sums, identical
    # on every rank bit for bit, equal to the all-reduced loss
    try:
        xchg = parallel.ScalarExchange(dev)
    except RuntimeError as exc:
        xchg = None
        if rank == 0:
            print(f"scalar exchange skipped: {exc}", flush=True)
    if xchg is not None:
        want = []
        for k in range(5):
            sub = slice(k, k + 4)
            l_k, _, _ = ops.multibox_loss_raw(so[sub].contiguous(), st[sub].contiguous(), priors, n_global=N, want_grad=False, exchange=xchg)
            want.append(parallel.global_loss(l_k))
        got = xchg.reduce(5)
        torch.cuda.synchronize()
        check(xchg.ok(), "scalar exchange gave up on a peer")
        check(torch.allclose(got, torch.stack(want), rtol=1e-6, atol=0), f"scalar exchange {got.tolist()} vs all-reduce {[float(w) for w in want]}")
        everyone = [torch.zeros_like(got) for _ in range(world)]
        dist.all_gather(everyone, got)
        check(all(torch.equal(e, everyone[0]) for e in everyone), "exchange sums are not bit-identical across ranks")
        dist.barrier()
        xchg.close()

    # ---- evaluation ------------------------------------------------------------------------------------------------
    t2 = synth.make_targets(N, 78)
    o2 = synth.plant_detections(synth.make_outputs(N, 78, "D2"), t2, priors.cpu(), 78).to(dev)
    t2 = t2.to(dev).contiguous()
    full = o2.clone()
    r1 = ops.postprocess_(full, priors, want_lists=True)
    tall_1, _ = evaluate.accumulate(full, t2, keep=r1.keep, keep_cnt=r1.keep_cnt)
    tall_dense, _ = evaluate.accumulate(full, t2)
    check(torch.equal(tall_1, tall_dense), "kept-list tallies differ from the dense scan")
    mine = o2[lo:hi].clone()
    r2 = ops.postprocess_(mine, priors, want_lists=True)
    check(torch.equal(mine, full[lo:hi]), "post-processed shard differs from the single-GPU slice")
    tall_l, _ = evaluate.accumulate(mine, t2[lo:hi].contiguous(), keep=r2.keep, keep_cnt=r2.keep_cnt)
    parallel.all_reduce_tallies(tall_l)
    check(torch.equal(tall_l, tall_1), "all-reduced tallies differ from the single-GPU tallies")
    check(int(tall_1[:, 0].sum()) > 0, "no true positives in the test data (planting failed)")
    ev1 = evaluate.DetectionEvaluator()
    ev1.update(full, t2)
    single = {"tallies": ev1.tallies.clone(), "cls": torch.cat(ev1.cls), "score": torch.cat(ev1.score), "tp": torch.cat(ev1.tp)}
    ap1 = torch.stack([evaluate.voc_average_precision(single["score"][single["cls"] == c], single["tp"][single["cls"] == c],
                                                      int(single["tallies"][c, 2])) for c in range(20)])
    evw = evaluate.DetectionEvaluator()
    evw.update(mine, t2[lo:hi].contiguous(), keep=r2.keep, keep_cnt=r2.keep_cnt)
    out = evw.compute()                                   # NCCL: all-reduce of the tallies + all-gather of the detection lists
    check(torch.equal(out["tallies"], single["tallies"]), "DetectionEvaluator tallies differ")
    check(torch.allclose(out["ap_voc"], ap1, rtol=1e-6, atol=1e-7, equal_nan=True), "VOC AP over NCCL differs from the single-GPU evaluator")
    check(torch.allclose(out["ap_reference"], evaluate.average_precision_from_tallies(tall_1), equal_nan=True), "reference-formula AP differs")

    bad = torch.tensor([len(fails)], device=dev)
    dist.all_reduce(bad)
    for f in fails:
        print(f"[rank {rank}] FAIL: {f}", flush=True)
    if rank == 0:
        print(f"multirank world={world} N={N}: loss {float(loss_g):.6f} (single {float(loss_1):.6f}), tallies TP/det/gt "
              f"{[int(x) for x in tall_1.sum(dim=0)]}, failures {int(bad)}", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(1 if int(bad) else 0)


if __name__ == "__main__":
    main()
