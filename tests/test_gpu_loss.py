"""GPU parity: default boxes, matching, the loss probes and the fused MultiBox loss against the oracle and the
reference-generated fixtures.  Everything goes through the C ABI (ctypes) of libssdhead.so.

Tolerances (north_star): match masks / counts bit-exact; loss, thresholds, gradients and encoded offsets within
1e-5 relative in fp32.  The hard-negative SELECTION is a value threshold on cross-entropies that differ in the last
ulp between torch's CPU exp/log and the device's, so rows whose CE sits within 1e-5 of the threshold may flip; those
rows are excluded from the element-wise gradient comparison and bounded in number.
"""
import numpy as np
import pytest
import torch

import cases
from conftest import sha, unpack_bits
from object_detection_torch2_b200 import ops, synth
from object_detection_torch2_b200.model import SSD
from oracle import head

pytestmark = pytest.mark.gpu
DEV = "cuda"
RTOL = 1e-5


@pytest.fixture(scope="module")
def priors_gpu():
    return ops.default_boxes(DEV)


def test_default_boxes_bit_exact(golden, priors_cpu, priors_gpu):
    got = priors_gpu.cpu()
    assert np.array_equal(got.numpy().view(np.uint32), golden["priors"].view(np.uint32))
    assert torch.equal(got, priors_cpu)


def test_device_info():
    info = ops.device_info()
    assert info["sm_count"] >= 100 and info["loss_cluster_size"] in (4, 8) and info["loss_max_active_clusters"] >= 1


@pytest.mark.parametrize("case", cases.LOSS_CASES, ids=[c[0] for c in cases.LOSS_CASES])
def test_match_golden(case, golden, priors_cpu, priors_gpu):
    k = f"loss/{case[0]}/"
    o, t = cases.loss_inputs(case, priors_cpu)
    assert sha(t) == str(golden[k + "sha_targets"])
    r = ops.match(t.to(DEV), priors_gpu, want_bits=True, want_mask=True, want_best_gt=True, want_best_prior=True)
    want = unpack_bits(golden[k + "match_bits"], golden[k + "match_shape"])
    assert torch.equal(r.mask.cpu(), want)
    G = t.shape[1]
    bits = r.bits.cpu()
    rebuilt = torch.stack([(bits >> g) & 1 for g in range(G)], dim=2).bool()
    assert torch.equal(rebuilt, want)
    # arg-max extensions against the oracle's IoU table
    iou = torch.stack([head.pair_iou_match(t[:, g, :4], priors_cpu) for g in range(G)], dim=2)      # (N, P, G)
    bv, bg = iou.max(dim=2)
    assert torch.equal(r.best_iou.cpu(), bv)
    assert torch.equal(iou.gather(2, r.best_gt.cpu().long().unsqueeze(2)).squeeze(2), bv)
    pv, _ = iou.max(dim=1)
    assert torch.equal(r.best_prior_iou.cpu(), pv)
    assert torch.equal(iou.gather(1, r.best_prior.cpu().long().unsqueeze(1)).squeeze(1), pv)


@pytest.mark.parametrize("seed", [21, 22])
def test_match_batch32_vs_oracle(seed, priors_cpu, priors_gpu):
    t = synth.make_targets(32, seed)
    got = SSD._match(None, t.to(DEV), priors_gpu)
    assert got.dtype == torch.bool
    assert torch.equal(got.cpu(), head.match_mask(t, priors_cpu))
    got = SSD._match(None, t.to(DEV), priors_gpu, threshold=0.5)
    assert torch.equal(got.cpu(), head.match_mask(t, priors_cpu, 0.5))


def test_probe_helpers(priors_cpu, priors_gpu, golden):
    o, t = synth.make_batch(2, 41, "D1", 7)
    od, td = o.to(DEV), t.to(DEV)
    # _calc_delta: divisions are IEEE (bit-exact columns 0-1), log differs by an ulp
    got = SSD._calc_delta(None, td, priors_gpu).cpu()
    want = torch.stack([head.encode_offsets(t[:, g, :4], priors_cpu) for g in range(t.shape[1])], dim=2)
    assert torch.equal(got[..., :2], want[..., :2])
    torch.testing.assert_close(got[..., 2:], want[..., 2:], rtol=RTOL, atol=1e-6)
    # _smooth_l1
    x = torch.randn(10000, generator=torch.Generator().manual_seed(1)) * 2
    assert torch.equal(SSD._smooth_l1(None, x.to(DEV)).cpu(), head.smooth_l1(x))
    # _softmax_cross_entropy, both call shapes of ssd.py:208 and :213
    want = -(t[:, None, :, 4:] * torch.log_softmax(o[:, :, None, 4:], dim=3)).sum(dim=3)
    torch.testing.assert_close(SSD._softmax_cross_entropy(None, od[:, :, 4:], td[:, :, 4:]).cpu(), want, rtol=RTOL, atol=1e-6)
    void = torch.eye(21)[0].view(1, 1, 21)
    want = -(void[:, None] * torch.log_softmax(o[:, :, None, 4:], dim=3)).sum(dim=3)
    torch.testing.assert_close(SSD._softmax_cross_entropy(None, od[:, :, 4:], void.to(DEV)).cpu(), want, rtol=RTOL, atol=1e-6)
    # _split_pos_neg
    pos = torch.tensor([0, 10, 2183, 2184, 5000, 8732])
    kp, kn = SSD._split_pos_neg(None, pos.to(DEV), (8732 - pos).to(DEV))
    wp, wn = head.split_pos_neg(pos, 8732)
    assert torch.equal(kp.cpu(), wp) and torch.equal(kn.cpu(), wn)
    # _k_plus_1_th_value: exact order statistic, incl. ties, k = 0 and negative values
    g = torch.Generator().manual_seed(3)
    v = torch.randn(8732, generator=g)
    v[100:200] = v[50]
    for k in (0, 1, 77, 4000, 8731):
        got = SSD._k_plus_1_th_value(None, v.to(DEV), torch.tensor(k, device=DEV))
        assert float(got) == float(head.kplus1_threshold(v, k))
    rows = torch.rand(5, 1000, generator=g)
    ks = torch.tensor([0, 5, 500, 998, 999])
    got = ops.kplus1_value(rows.to(DEV), ks.to(DEV)).cpu()
    assert torch.equal(got, torch.stack([head.kplus1_threshold(rows[i], int(ks[i])) for i in range(5)]))


# ------------------------------------------------------------------------------------------------------------------
def run_and_compare(o, t, priors_cpu, priors_gpu, a=1.0, thr=0.25, max_flips=4, check_ambiguous=True):
    """Fused kernel vs oracle on the same inputs; returns (gpu stats, oracle result)."""
    ref = head.multibox_loss(o, t, priors_cpu, a=a, threshold=thr, want_grad=True)
    loss, grad, stats = ops.multibox_loss_raw(o.to(DEV).contiguous(), t.to(DEV).contiguous(), priors_gpu, a=a, threshold=thr,
                                              want_grad=True, want_stats=True)
    st = ops.stats_to_numpy(stats)
    assert np.array_equal(st["pos_raw"], ref["pos_raw"].numpy())
    assert np.array_equal(st["k_pos"], ref["k_pos"].numpy())
    assert np.array_equal(st["k_neg"], ref["k_neg"].numpy())
    np.testing.assert_allclose(st["thr_pos"], ref["thr_pos"].numpy(), rtol=RTOL, atol=1e-7)
    np.testing.assert_allclose(st["thr_neg"], ref["thr_neg"].numpy(), rtol=RTOL, atol=1e-7)
    assert np.abs(st["pos_sel"] - ref["pos_sel"].numpy()).max() <= max_flips
    assert np.abs(st["neg_sel"] - ref["neg_sel"].numpy()).max() <= max_flips
    np.testing.assert_allclose(st["loss"], ref["loss_per_image"].numpy(), rtol=RTOL, atol=1e-7)
    np.testing.assert_allclose(float(loss), float(ref["loss"]), rtol=RTOL, atol=1e-7)
    # gradient: skip rows whose CE is within tolerance of its threshold (selection may legitimately flip)
    near_p = (ref["ce_pos"] - ref["thr_pos"][:, None]).abs() <= 1e-5 * ref["thr_pos"][:, None].clamp(min=1.0)
    near_n = (ref["ce_neg"] - ref["thr_neg"][:, None]).abs() <= 1e-5 * ref["thr_neg"][:, None].clamp(min=1.0)
    member_p = ref["match"].any(dim=2)
    ambiguous = (near_p & member_p) | (near_n & ~member_p)
    if check_ambiguous:
        assert int(ambiguous.sum()) <= 8 * o.shape[0] + 16
    g = grad.cpu()
    want = ref["grad"]
    ok = ~ambiguous
    scale = want.abs().max().clamp(min=1e-12)
    err = ((g - want).abs() * ok[:, :, None]).max()
    assert float(err) <= 1e-5 * float(scale) + 1e-10, f"grad error {float(err)} vs scale {float(scale)}"
    # element-wise relative check on the well-conditioned entries ((softmax - 1) * s cancels for confident rows)
    rel_rows = ok[:, :, None] & (want.abs() > 1e-2 * scale)
    if bool(rel_rows.any()):
        assert float(((g - want).abs() / want.abs().clamp(min=1e-30))[rel_rows].max()) <= 1e-4
    # rows that are selected by neither side carry exactly zero gradient
    unselected = ~(ref["pos_valid"] | ref["neg_valid"]) & ok
    assert float(g[unselected].abs().max() if unselected.any() else 0.0) == 0.0
    return st, ref


@pytest.mark.parametrize("case", cases.LOSS_CASES, ids=[c[0] for c in cases.LOSS_CASES])
def test_loss_golden(case, golden, priors_cpu, priors_gpu):
    k = f"loss/{case[0]}/"
    o, t = cases.loss_inputs(case, priors_cpu)
    assert sha(o) == str(golden[k + "sha_outputs"]) and sha(t) == str(golden[k + "sha_targets"]), "input generator drifted"
    a = case[5]
    st, ref = run_and_compare(o, t, priors_cpu, priors_gpu, a=a)
    # against the reference's own numbers
    x = o.to(DEV).requires_grad_(True)
    net = SSD.__new__(SSD)
    loss = SSD.loss(net, outputs=x, targets=t.to(DEV), default_bboxes=priors_gpu, a=a)
    assert loss.dim() == 0
    np.testing.assert_allclose(float(loss.detach()), float(golden[k + "loss"]), rtol=RTOL)
    with torch.autograd.set_detect_anomaly(True):        # the reference trains under anomaly detection (train.py:102)
        loss.backward()
    g = x.grad.cpu()
    np.testing.assert_allclose(g.abs().sum(dim=(1, 2)).double().numpy(), golden[k + "grad_abs_sum"], rtol=1e-4)
    sample = g.reshape(-1)[::cases.GRAD_STRIDE].numpy()
    want = golden[k + "grad_sample"]
    bad = np.abs(sample - want) > 1e-5 * np.abs(want).max() + 1e-10
    assert bad.sum() <= 2, f"{bad.sum()} sampled gradient entries differ"
    rows = unpack_bits(golden[k + "grad_row_nonzero"], g.shape[:2])
    flips = int(((g.abs().sum(dim=2) > 0) != rows).sum())
    assert flips <= 4 * o.shape[0], f"{flips} rows changed selection"
    with torch.no_grad():                                # validation pass (train.py:128-139)
        val = SSD.loss(net, outputs=o.to(DEV), targets=t.to(DEV), default_bboxes=priors_gpu, a=a)
    np.testing.assert_allclose(float(val), float(golden[k + "loss_nograd"]), rtol=RTOL)


@pytest.mark.parametrize("seed,dist", [(51, "D1"), (52, "D2")])
def test_loss_batch32_vs_oracle(seed, dist, priors_cpu, priors_gpu):
    o, t = synth.make_batch(32, seed, dist)
    run_and_compare(o, t, priors_cpu, priors_gpu)


def test_loss_edge_cases(priors_cpu, priors_gpu):
    # every ground-truth row is padding -> loss 0, gradient 0 (ssd.py:226)
    o = synth.make_outputs(2, 61)
    t = torch.zeros(2, 3, 25)
    loss, grad, stats = ops.multibox_loss_raw(o.to(DEV), t.to(DEV), priors_gpu, want_stats=True)
    assert float(loss) == 0.0 and float(grad.abs().max()) == 0.0
    st = ops.stats_to_numpy(stats)
    assert st["pos_raw"].tolist() == [0, 0] and st["neg_sel"].tolist() == [0, 0]
    # G = 0 (no ground-truth rows at all)
    loss, grad, _ = ops.multibox_loss_raw(o.to(DEV), torch.zeros(2, 0, 25, device=DEV), priors_gpu)
    assert float(loss) == 0.0 and float(grad.abs().max()) == 0.0
    # all-equal logits: every negative ties at the threshold, strict '>' selects none (ssd.py:223)
    o = torch.zeros(1, 8732, 25)
    t = synth.make_targets(1, 62, 4)
    st, ref = run_and_compare(o, t, priors_cpu, priors_gpu, check_ambiguous=False)
    assert int(st["neg_sel"][0]) == 0 and int(ref["neg_sel"][0]) == 0
    # crowded image: 3 * pos > neg, only the neg // 3 hardest positives keep their terms (ssd.py:310-311)
    t = synth.make_targets(2, 63, 20, min_boxes=20)
    o = synth.make_outputs(2, 63)
    st, ref = run_and_compare(o, t, priors_cpu, priors_gpu)
    assert (st["pos_raw"] * 3 > 8732 - st["pos_raw"]).any()
    # other match thresholds and loss weights
    run_and_compare(o, t, priors_cpu, priors_gpu, a=0.5, thr=0.5)


@pytest.mark.parametrize("thr", [0.02, 0.25, 0.6, 0.95])
def test_loss_ground_truth_culling_is_conservative(thr, priors_cpu, priors_gpu):
    """The fused kernel drops ground-truth rows per 96-prior chunk with an outer-box / area bound before matching;
    whatever the threshold and the box sizes, the positives must stay exactly the oracle's (ssd.py:231-250)."""
    g = torch.Generator().manual_seed(int(thr * 1000))
    t = synth.make_targets(4, 90, 20, min_boxes=12)
    # add extreme boxes: tiny, full-image, thin slivers, one far outside the image
    t[0, 0, :4] = torch.tensor([0.5, 0.5, 1.0, 1.0])
    t[0, 1, :4] = torch.tensor([0.31, 0.72, 0.004, 0.004])
    t[1, 0, :4] = torch.tensor([0.5, 0.1, 0.98, 0.02])
    t[1, 1, :4] = torch.tensor([3.0, 3.0, 0.3, 0.3])
    t[2, 0, :4] = torch.tensor([0.02, 0.5, 0.02, 0.9])
    o = torch.randn(4, 8732, 25, generator=g)
    run_and_compare(o, t, priors_cpu, priors_gpu, thr=thr, max_flips=8)


def test_loss_odd_priors_skip_the_culling(priors_cpu, priors_gpu):
    """Priors with a non-positive or huge extent disable the bound for their chunk; results still follow the oracle."""
    pri = priors_cpu.clone()
    pri[5, 2] = 0.0                      # zero width
    pri[700, 3] = -0.2                   # negative height
    pri[4000, 2:] = torch.tensor([1e20, 1e20])
    pri[8731, 0] = 50.0                  # far away
    o, t = synth.make_batch(2, 93, "D1")
    # (the reference's dense encode turns such priors into NaN losses -- 0 * log(negative) -- so only the matching is compared)
    ref = head.multibox_loss(o, t, pri, a=1.0, threshold=0.25, want_grad=False)
    _, _, stats = ops.multibox_loss_raw(o.to(DEV), t.to(DEV), pri.to(DEV), want_grad=False, want_stats=True)
    st = ops.stats_to_numpy(stats)
    for k in ("pos_raw", "k_pos", "k_neg"):
        assert np.array_equal(st[k], ref[k].numpy()), k


def test_loss_generic_shapes(priors_cpu, priors_gpu):
    g = torch.Generator().manual_seed(7)
    # fewer priors (bulk path, ragged last CTA) and a row count that defeats 16-byte chunks (fallback copy path)
    for P in (600, 601, 37):
        pri = priors_cpu[torch.randperm(8732, generator=g)[:P]].contiguous()
        o = torch.randn(3, P, 25, generator=g)
        t = synth.make_targets(3, 70 + P % 7, 6)
        run_and_compare(o, t, pri, pri.to(DEV))
    # class counts other than 21 take the runtime-C instantiation; soft labels take the generic CE path
    P = 900
    pri = priors_cpu[::9][:P].contiguous()
    for C in (5, 33):
        o = torch.randn(2, P, 4 + C, generator=g)
        t = torch.zeros(2, 5, 4 + C)
        base = synth.make_targets(2, 80 + C, 5)
        t[:, :base.shape[1], :4] = base[:, :, :4]
        t[:, :, 4:] = torch.rand(2, 5, C, generator=g) * (t[:, :, 2:3] > 0)
        run_and_compare(o, t, pri, pri.to(DEV))


def test_loss_autograd_contract(priors_gpu):
    o, t = synth.make_batch(2, 91, "D2", 5)
    net = SSD.__new__(SSD)
    x = o.to(DEV).requires_grad_(True)
    loss = net.loss(outputs=x, targets=t.to(DEV), default_bboxes=priors_gpu)      # keyword call, train.py:119-120
    (loss * 3.0).backward()
    y = o.to(DEV).requires_grad_(True)
    net.loss(outputs=y, targets=t.to(DEV), default_bboxes=priors_gpu).backward()
    torch.testing.assert_close(x.grad, 3.0 * y.grad, rtol=1e-6, atol=0)
    # non-contiguous input: gradient flows back through the copy
    z = o.to(DEV).transpose(0, 1).contiguous().transpose(0, 1).requires_grad_(True)
    net.loss(outputs=z, targets=t.to(DEV), default_bboxes=priors_gpu).backward()
    torch.testing.assert_close(z.grad, y.grad, rtol=0, atol=0)
    with pytest.raises(RuntimeError):
        net.loss(outputs=o, targets=t, default_bboxes=priors_gpu.cpu())              # CPU tensors: no fallback


def test_loss_deterministic_and_graph_capturable(priors_gpu):
    o, t = synth.make_batch(8, 93, "D1")
    od, td = o.to(DEV), t.to(DEV)
    l1, g1, _ = ops.multibox_loss_raw(od, td, priors_gpu)
    l2, g2, _ = ops.multibox_loss_raw(od, td, priors_gpu)
    assert torch.equal(l1, l2) and torch.equal(g1, g2)
    loss = torch.empty((), device=DEV)
    grad = torch.empty_like(od)
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        ops.multibox_loss_raw(od, td, priors_gpu, loss_out=loss, grad_out=grad)      # warm up on the capture stream
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=s):
            ops.multibox_loss_raw(od, td, priors_gpu, loss_out=loss, grad_out=grad)
    torch.cuda.current_stream().wait_stream(s)
    loss.zero_()
    grad.zero_()
    graph.replay()
    torch.cuda.synchronize()
    assert torch.equal(loss, l1) and torch.equal(grad, g1)


def test_sharded_loss_equals_whole_batch(priors_gpu):
    # images are independent: two shards called with n_global = 8 sum to the whole-batch mean
    from object_detection_torch2_b200 import parallel
    o, t = synth.make_batch(8, 95, "D1")
    od, td = o.to(DEV), t.to(DEV)
    whole, gw, _ = ops.multibox_loss_raw(od, td, priors_gpu)
    parts, grads = [], []
    for r in range(2):
        so, stt = parallel.shard_batch(od, td, 2, r)
        l, g, _ = ops.multibox_loss_raw(so.contiguous(), stt.contiguous(), priors_gpu, n_global=8)
        parts.append(l)
        grads.append(g)
    torch.testing.assert_close(parts[0] + parts[1], whole, rtol=1e-6, atol=0)
    assert torch.equal(torch.cat(grads), gw)


@pytest.mark.parametrize("g_rows", [42, 64])
def test_loss_many_ground_truth_rows(g_rows, priors_cpu, priors_gpu):
    # VOC images carry up to 42 boxes; G > 32 exercises the high mask word, large G the 8-CTA kernel shape
    t = synth.make_targets(2, 140 + g_rows, max_boxes=g_rows, min_boxes=g_rows)
    assert t.shape[1] == g_rows
    t[:, :, 2:4] *= 0.4                      # smaller boxes keep the image out of the crowded regime for one of them
    o = synth.make_outputs(2, 141, "D2")
    run_and_compare(o, t, priors_cpu, priors_gpu)
    assert torch.equal(ops.match(t.to(DEV), priors_gpu).mask.cpu(), head.match_mask(t, priors_cpu))


def test_loss_full_size_properties(priors_gpu):
    # config 4 of BASELINE.json per-GPU width (128 images): size-independent properties instead of the slow oracle
    N = 128
    o, t = synth.make_batch(N, 151, "D2")
    od, td = o.to(DEV), t.to(DEV)
    loss, grad, stats = ops.multibox_loss_raw(od, td, priors_gpu, want_stats=True)
    st = ops.stats_to_numpy(stats)
    # batch mean == mean of the per-image losses; images are independent (a sub-batch gives the same per-image numbers)
    np.testing.assert_allclose(float(loss), st["loss"].astype(np.float64).mean(), rtol=1e-6)
    sub = slice(40, 48)
    l2, g2, s2 = ops.multibox_loss_raw(od[sub].contiguous(), td[sub].contiguous(), priors_gpu, n_global=N, want_stats=True)
    st2 = ops.stats_to_numpy(s2)
    assert np.array_equal(st2["loss"], st["loss"][sub]) and np.array_equal(st2["pos_raw"], st["pos_raw"][sub])
    assert torch.equal(g2, grad[sub])
    # split arithmetic, selection counts and gradient support
    pos, kp, kn = st["pos_raw"].astype(np.int64), st["k_pos"], st["k_neg"]
    crowded = pos * 3 > 8732 - pos
    assert np.array_equal(kp, np.where(crowded, (8732 - pos) // 3, pos)) and np.array_equal(kn, np.where(crowded, 8732 - pos, pos * 3))
    assert (st["pos_sel"] <= kp).all() and (st["neg_sel"] <= kn).all()
    nonzero_rows = (grad.abs().sum(dim=2) > 0).sum(dim=1).cpu().numpy()
    assert (nonzero_rows <= st["pos_sel"] + st["neg_sel"]).all() and (nonzero_rows >= st["pos_sel"] + st["neg_sel"] - 2).all()
    # every selected row's class gradient sums to ~0 for negatives (softmax - e0) -- spot check via the void column sign
    assert float(grad[:, :, 4:].sum(dim=2).abs().max()) < 1e-3
    assert torch.isfinite(grad).all() and torch.isfinite(loss)


def test_pipelined_call_and_prefetch_are_pure_hints(priors_gpu):
    # ssdh_multibox_loss_pipelined / ssdh_prefetch_l2 only warm the L2 for the next micro-batch: results are unchanged
    o, t = synth.make_batch(6, 161, "D1")
    o2, t2 = synth.make_batch(6, 162, "D2")
    G = max(t.shape[1], t2.shape[1])                # the next micro-batch has the shapes of this one
    od, td, o2d, t2d = o.to(DEV), synth.pad_targets(t, G).to(DEV), o2.to(DEV), synth.pad_targets(t2, G).to(DEV)
    base_l, base_g, _ = ops.multibox_loss_raw(od, td, priors_gpu)
    keep = o2d.clone()
    l, g, _ = ops.multibox_loss_raw(od, td, priors_gpu, next_outputs=o2d, next_targets=t2d)
    assert torch.equal(l, base_l) and torch.equal(g, base_g) and torch.equal(o2d, keep)
    ops.prefetch_l2(o2d)
    ops.prefetch_l2(t2d[:, :1])                     # odd sizes / small tensors are fine
    torch.cuda.synchronize()
    assert torch.equal(o2d, keep)
    # back-to-back launches overlap through programmatic dependent launch: same workspace, same result every time
    outs = [ops.multibox_loss_raw(od, td, priors_gpu) for _ in range(8)]
    torch.cuda.synchronize()
    assert all(torch.equal(x[0], base_l) and torch.equal(x[1], base_g) for x in outs)


def test_loss_right_behind_its_producers(priors_cpu, priors_gpu):
    """The kernel in front of the loss in the stream is normally the PRODUCER of its inputs (ssdh_pack_head at the end of
    SSD.forward, ssdh_expand_targets, a copy): the loss is launched with programmatic stream serialization and must not
    read `outputs` / `targets` before that producer's writes are visible.  Many back-to-back rounds on the same buffers
    with changing contents, no synchronisation in between; a stale read shows up as the previous round's loss."""
    levels_shape = [(38, 4), (19, 6), (10, 6), (5, 6), (3, 4), (1, 4)]
    N = 8
    g = torch.Generator().manual_seed(5)
    rounds = []
    for r in range(6):
        lv = [torch.randn(N, a * 25, m, m, generator=g) for m, a in levels_shape]
        t = synth.make_targets(N, 200 + r, 12)
        rounds.append((lv, t))
    # expected values, computed with full synchronisation and freshly allocated tensors
    want = []
    for lv, t in rounds:
        o = torch.cat([x.permute(0, 2, 3, 1).reshape(N, -1, 25) for x in lv], dim=1).to(DEV).contiguous()
        l, gr, _ = ops.multibox_loss_raw(o, t.to(DEV).contiguous(), priors_gpu)
        torch.cuda.synchronize()
        want.append((l.clone(), gr.abs().sum().clone()))
    # now the pipelined stream: stage -> pack -> expand -> loss, all on one stream, buffers reused every round
    dev_lv = [torch.empty(N, a * 25, m, m, device=DEV) for m, a in levels_shape]
    G = max(t.shape[1] for _, t in rounds)
    compact = torch.empty(N, G, 5, device=DEV)
    lengths = torch.empty(N, dtype=torch.int32, device=DEV)
    got = []
    for rep in range(3):
        for (lv, t) in rounds:
            tc = synth.pad_targets(t, G)
            real = (tc[:, :, 2] * tc[:, :, 3]) > 0
            comp = torch.cat([tc[:, :, :4], tc[:, :, 4:].argmax(dim=2, keepdim=True).float()], dim=2) * real[:, :, None]
            for d, s in zip(dev_lv, lv):
                d.copy_(s.to(DEV))
            compact.copy_(comp.to(DEV))
            lengths.copy_(real.sum(dim=1).to(torch.int32).to(DEV))
            o = ops.pack_head(dev_lv, 25)                       # producer 1: last kernel before the loss is NOT this one ...
            tg = ops.expand_targets(compact, lengths, 21)       # ... producer 2 is (targets), pack_head two launches back
            l, gr, _ = ops.multibox_loss_raw(o, tg, priors_gpu)
            got.append((l, gr.abs().sum()))
            tg2 = ops.expand_targets(compact, lengths, 21)
            o2 = ops.pack_head(dev_lv, 25)                      # and the other order: pack_head directly in front of the loss
            l2, gr2, _ = ops.multibox_loss_raw(o2, tg2, priors_gpu)
            got.append((l2, gr2.abs().sum()))
    torch.cuda.synchronize()
    for i, (l, s) in enumerate(got):
        wl, ws = want[(i // 2) % len(rounds)]
        assert torch.equal(l, wl), f"round {i}: loss {float(l)} vs {float(wl)} (stale input read under the producer's tail?)"
        assert torch.equal(s, ws), f"round {i}: gradient checksum differs"


def test_loss_backward_twice_raises_and_args_are_checked(priors_gpu):
    o, t = synth.make_batch(2, 171, "D2", 4)
    net = SSD.__new__(SSD)
    x = o.to(DEV).requires_grad_(True)
    loss = net.loss(outputs=x, targets=t.to(DEV), default_bboxes=priors_gpu)
    loss.backward(retain_graph=True)
    first = x.grad.clone()
    with pytest.raises(RuntimeError, match="second time"):
        loss.backward()
    assert torch.equal(x.grad, first)
    od, td = o.to(DEV), t.to(DEV)
    with pytest.raises(ValueError, match="targets"):
        ops.multibox_loss_raw(od, td[:, :, :24].contiguous(), priors_gpu)            # class count mismatch
    with pytest.raises(ValueError, match="priors"):
        ops.multibox_loss_raw(od, td, priors_gpu[:100].contiguous())
    with pytest.raises(ValueError, match="contiguous"):
        ops.multibox_loss_raw(od.transpose(0, 1).contiguous().transpose(0, 1), td, priors_gpu)
    with pytest.raises(ValueError, match="next_targets"):
        ops.multibox_loss_raw(od, td, priors_gpu, next_outputs=od, next_targets=td[:, :1].contiguous())


# ------------------------------------------------------------------------------------------------------------------
# north_star extension: best-prior-per-ground-truth forcing; off = the reference
# ------------------------------------------------------------------------------------------------------------------
def _small_box_batch(n, seed, scale):
    t = synth.make_targets(n, seed)
    t[:, :, 2:4] *= scale                                # small boxes: many have no prior above the threshold
    return t


@pytest.mark.parametrize("thr,scale", [(0.25, 1.0), (0.5, 0.3), (0.7, 0.5)])
def test_force_best_prior_matching(thr, scale, priors_cpu, priors_gpu):
    t = _small_box_batch(6, 301, scale)
    want0 = head.match_mask(t, priors_cpu, thr)
    want1 = head.match_mask(t, priors_cpu, thr, force_best_prior=True)
    got0 = SSD._match(None, t.to(DEV), priors_gpu, threshold=thr)
    got1 = SSD._match(None, t.to(DEV), priors_gpu, threshold=thr, force_best_prior=True)
    assert torch.equal(got0.cpu(), want0) and torch.equal(got1.cpu(), want1)
    forced = int((want1 & ~want0).sum())
    real = (t[:, :, 2] * t[:, :, 3]) > 0
    assert forced == int((real & ~want0.any(dim=1)).sum())             # exactly the boxes that had no match at all
    if scale < 1.0:
        assert forced > 0
    r = ops.match(t.to(DEV), priors_gpu, thr, want_bits=True, force_best_prior=True)
    rebuilt = torch.stack([(r.bits.cpu() >> g) & 1 for g in range(t.shape[1])], dim=2).bool()
    assert torch.equal(rebuilt, want1)


@pytest.mark.parametrize("thr,scale,dist", [(0.5, 0.3, "D1"), (0.25, 1.0, "D2"), (0.7, 0.5, "D2")])
def test_force_best_prior_loss(thr, scale, dist, priors_cpu, priors_gpu):
    """The fused kernel with forcing against the oracle extension (which is the identity when off): counts bit-exact,
    loss / thresholds / gradient 1e-5 -- i.e. the forced priors are the oracle's, bit for bit (pos_raw, k_pos and the
    gradient support would differ otherwise)."""
    N = 6
    t = _small_box_batch(N, 311, scale)
    o = synth.make_outputs(N, 311, dist)
    ref = head.multibox_loss(o, t, priors_cpu, threshold=thr, want_grad=True, force_best_prior=True)
    base = head.multibox_loss(o, t, priors_cpu, threshold=thr)
    loss, grad, stats = ops.multibox_loss_raw(o.to(DEV), t.to(DEV).contiguous(), priors_gpu, threshold=thr, want_stats=True,
                                              force_best_prior=True)
    st = ops.stats_to_numpy(stats)
    assert np.array_equal(st["pos_raw"], ref["pos_raw"].numpy()) and np.array_equal(st["k_pos"], ref["k_pos"].numpy())
    assert np.array_equal(st["k_neg"], ref["k_neg"].numpy())
    np.testing.assert_allclose(st["loss"], ref["loss_per_image"].numpy(), rtol=RTOL, atol=1e-7)
    np.testing.assert_allclose(float(loss), float(ref["loss"]), rtol=RTOL)
    if scale < 1.0:
        assert int((ref["pos_raw"] - base["pos_raw"]).sum()) > 0       # forcing did add positives
    # the rows that forcing turned positive carry a positive-type gradient (localisation columns non-zero) when selected
    newly = ref["match"].any(dim=2) & ~base["match"].any(dim=2)
    g = grad.cpu()
    sel = newly & ref["pos_valid"]
    near = (ref["ce_pos"] - ref["thr_pos"][:, None]).abs() <= 1e-5 * ref["thr_pos"][:, None].clamp(min=1.0)
    sure = sel & ~near
    if bool(sure.any()):
        assert bool((g[:, :, :4].abs().sum(dim=2)[sure] > 0).all())
        torch.testing.assert_close(g[sure], ref["grad"][sure], rtol=1e-4, atol=1e-5 * float(ref["grad"].abs().max()))
    # off = the reference, through the same entry point
    l0, g0, _ = ops.multibox_loss_raw(o.to(DEV), t.to(DEV).contiguous(), priors_gpu, threshold=thr)
    l1, g1, _ = ops.multibox_loss_raw(o.to(DEV), t.to(DEV).contiguous(), priors_gpu, threshold=thr, exact_math=False, force_best_prior=False)
    assert torch.equal(l0, l1) and torch.equal(g0, g1)
    # public surface (SSD.loss matches at the reference's fixed 0.25, ssd.py:231)
    if thr == 0.25:
        net = SSD.__new__(SSD)
        x = o.to(DEV).requires_grad_(True)
        lf = net.loss(outputs=x, targets=t.to(DEV), default_bboxes=priors_gpu, force_best_prior=True)
        np.testing.assert_allclose(float(lf.detach()), float(ref["loss"]), rtol=RTOL)


# ------------------------------------------------------------------------------------------------------------------
# Hard-negative selection: exact on exact inputs, and every production flip is a tie at the threshold
# ------------------------------------------------------------------------------------------------------------------
def _selection_of(grad, match_any):
    """(pos_sel, neg_sel) row masks recovered from the gradient's support: unselected rows carry exact zeros, a selected
    row never does (its softmax terms cannot all vanish)."""
    nz = grad.abs().sum(dim=2) > 0
    return nz & match_any, nz & ~match_any


@pytest.mark.parametrize("seed,dist,n", [(51, "D1", 8), (52, "D2", 8), (53, "D1", 32)])
def test_selection_is_exact_on_the_oracles_cross_entropies(seed, dist, n, priors_cpu, priors_gpu):
    """The selection logic in isolation (bucket histograms, cluster exchange, order statistic, strict '>'): the exact-math
    instantiation is fed the oracle's own cross-entropies and must then select EXACTLY the oracle's rows -- zero flips,
    thresholds bit-identical.  (With its own CE the kernel differs from torch in the last ulps, see the next test.)"""
    o, t = synth.make_batch(n, seed, dist)
    ref = head.multibox_loss(o, t, priors_cpu, want_grad=True)
    member = ref["match"].any(dim=2)
    ce = torch.where(member, ref["ce_pos"], ref["ce_neg"]).contiguous()
    loss, grad, stats = ops.multibox_loss_raw(o.to(DEV), t.to(DEV).contiguous(), priors_gpu, want_stats=True, exact_math=True,
                                              ce_override=ce.to(DEV))
    st = ops.stats_to_numpy(stats)
    assert np.array_equal(st["thr_pos"].view(np.uint32), ref["thr_pos"].numpy().view(np.uint32))
    assert np.array_equal(st["thr_neg"].view(np.uint32), ref["thr_neg"].numpy().view(np.uint32))
    assert np.array_equal(st["pos_sel"], ref["pos_sel"].numpy()) and np.array_equal(st["neg_sel"], ref["neg_sel"].numpy())
    ps, ns = _selection_of(grad.cpu(), member)
    assert torch.equal(ps, ref["pos_valid"]) and torch.equal(ns, ref["neg_valid"])
    np.testing.assert_allclose(float(loss), float(ref["loss"]), rtol=RTOL)


def _flip_report(o, t, priors_cpu, priors_gpu, **kw):
    """Rows whose selection differs between the kernel and the oracle, with the distance of the oracle's CE from the
    oracle's threshold in units of the threshold's ulp."""
    ref = head.multibox_loss(o, t, priors_cpu, want_grad=False)
    _, grad, _ = ops.multibox_loss_raw(o.to(DEV), t.to(DEV).contiguous(), priors_gpu, **kw)
    member = ref["match"].any(dim=2)
    ps, ns = _selection_of(grad.cpu(), member)
    flips_p = ps != ref["pos_valid"]
    flips_n = ns != ref["neg_valid"]
    def ulps(ce, thr, flips):
        thr_b = thr[:, None].expand_as(ce)[flips]
        gap = (ce[flips].double() - thr_b.double()).abs()
        ulp = torch.from_numpy(np.spacing(np.abs(thr_b.numpy()).astype(np.float32))).double()
        return gap / ulp
    return flips_p.sum(dim=1) + flips_n.sum(dim=1), torch.cat([ulps(ref["ce_pos"], ref["thr_pos"], flips_p), ulps(ref["ce_neg"], ref["thr_neg"], flips_n)])


@pytest.mark.parametrize("seed,dist", [(0, "D1"), (1, "D2"), (2, "D1"), (3, "D2")])
def test_selection_flips_are_ties_at_the_threshold(seed, dist, priors_cpu, priors_gpu):
    """Batch 32: a row may only change sides when the oracle's CE sits within a few ulp of the oracle's threshold (the
    kernel's exp / log differ from torch's in the last bits).  Bounds measured over seeds 0-9 (tools/count_flips.py,
    DESIGN.md): the production kernel (ex2.approx / lg2.approx) stays within 16 ulp, the exact-math instantiation within 4."""
    o, t = synth.make_batch(32, seed, dist)
    per_image, gaps = _flip_report(o, t, priors_cpu, priors_gpu)
    assert int(per_image.max()) <= 4 and (gaps.numel() == 0 or float(gaps.max()) <= 16.0), (per_image.tolist(), gaps.tolist())
    per_image, gaps = _flip_report(o, t, priors_cpu, priors_gpu, exact_math=True)
    assert int(per_image.max()) <= 2 and (gaps.numel() == 0 or float(gaps.max()) <= 4.0), (per_image.tolist(), gaps.tolist())
