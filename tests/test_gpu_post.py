"""GPU parity: decode, score, IoU, NMS (stand-alone and fused) and the evaluation tallies, through the C ABI.

Index results (sort order, keep lists, TP flags, tallies) are compared bit-exactly ON IDENTICAL INPUTS: the NMS and
evaluation kernels are fed the oracle's decoded + scored tensor, as north_star words it.  Decoded boxes and scores
themselves are held to 1e-5 relative (exp differs in the last ulp between torch's CPU kernels and the device).
"""
import numpy as np
import pytest
import torch

import cases
from conftest import sha, unpack_bits
from object_detection_torch2_b200 import evaluate, ops, synth, utils
from oracle import head

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(scope="module")
def priors_gpu():
    return ops.default_boxes(DEV)


def scored_cpu(o, priors_cpu):
    x = o.clone()
    x[:, :, :4] = head.decode_boxes(x, priors_cpu)
    x[:, :, 4:] = head.class_scores(x)
    return x


def lists(res, n):
    order = res.order[n, : int(res.order_cnt[n])].cpu().long()
    keep = res.keep[n, : int(res.keep_cnt[n])].cpu().long()
    return order, keep


@pytest.mark.parametrize("case", cases.POST_CASES, ids=[c[0] for c in cases.POST_CASES])
def test_post_golden(case, golden, priors_cpu, priors_gpu):
    k = f"post/{case[0]}/"
    o, t = cases.post_inputs(case, priors_cpu)
    assert sha(o) == str(golden[k + "sha_outputs"]) and sha(t) == str(golden[k + "sha_targets"]), "input generator drifted"
    od = o.to(DEV)
    # I1 / I2 with the reference's call pattern (evaluate.py:129-130)
    box = utils.calc_coordicate(pr=od, df=priors_gpu)
    np.testing.assert_allclose(box.cpu().reshape(-1)[::7].numpy(), golden[k + "box_sample"], rtol=1e-5, atol=1e-7)
    assert torch.equal(box[..., :2].cpu(), head.decode_boxes(o, priors_cpu)[..., :2])         # mul + add: bit-exact
    od[:, :, :4] = box
    sc = utils.calc_score(pr=od)
    assert sc.shape == (o.shape[0], 8732, 21)
    assert np.array_equal(sc.argmax(dim=2).cpu().numpy().astype(np.uint8), golden[k + "score_argmax"])
    np.testing.assert_allclose(sc.max(dim=2).values.cpu().numpy(), golden[k + "score_max"], rtol=1e-5)
    assert int((sc > 0).sum(dim=2).max()) == 1
    # I3 + I4 + E on the oracle's scored tensor (identical inputs -> bit-exact indices)
    x = scored_cpu(o, priors_cpu)
    xd = x.to(DEV)
    assert torch.equal(utils.calc_iou(xd[:, :64], t.to(DEV)).cpu(), torch.from_numpy(golden[k + "iou_gt_64"]))
    y = utils.non_maximum_suppression(outputs=xd, iou_thresh=case[4])
    assert y is xd
    kept_rows = (y[:, :, 4:].sum(dim=2) > 0).cpu()
    assert torch.equal(kept_rows, unpack_bits(golden[k + "kept_rows"], kept_rows.shape))
    assert sha(y.cpu()) == str(golden[k + "sha_after_nms"])
    tallies, flags = evaluate.accumulate(y, t.to(DEV), want_flags=True)
    assert np.array_equal(tallies.cpu().numpy(), golden[k + "tallies"])
    # TP flags in the reference's order: class-major, then image, then descending score
    yc, fl = y.cpu(), flags.cpu()
    got = []
    for c in range(20):
        for n in range(o.shape[0]):
            det = head.class_order(yc[n], c)
            got.append(fl[n, det].numpy())
    got = np.concatenate(got) if got else np.zeros(0, np.uint8)
    assert np.array_equal(got, golden[k + "tp_flags"])
    ap = evaluate.average_precision_from_tallies(tallies).cpu().numpy()
    want = golden[k + "ap"]
    has = ~np.isnan(want)
    np.testing.assert_allclose(ap[has], want[has], rtol=1e-6)


@pytest.mark.parametrize("seed,dist,n", [(101, "D2", 4), (102, "D1", 2)])
def test_nms_lists_vs_oracle(seed, dist, n, priors_cpu):
    o, t = synth.make_batch(n, seed, dist)
    o = synth.plant_detections(o, t, priors_cpu, seed)
    x = scored_cpu(o, priors_cpu)
    for kw in (dict(iou_thresh=0.5), dict(iou_thresh=0.45, score_thresh=0.3), dict(iou_thresh=0.45, top_k=200),
               dict(iou_thresh=0.3, per_class=True), dict(iou_thresh=0.45, score_thresh=0.01, top_k=200, per_class=True)):
        xd = x.to(DEV)
        res = ops.nms_(xd, want_lists=True, **kw)
        want = x.clone()
        want, _ = head.nms_inplace(want, **kw)
        for i in range(n):
            order, keep = lists(res, i)
            o_want, k_want = head.greedy_nms(x[i], **kw)
            assert torch.equal(order, o_want), kw
            assert torch.equal(keep, k_want), kw
        assert torch.equal(xd.cpu(), want), kw


def test_nms_edge_cases():
    rows = torch.zeros(1, 64, 25)
    xd = rows.to(DEV)
    res = ops.nms_(xd, want_lists=True)                                  # no candidates at all
    assert int(res.order_cnt[0]) == 0 and int(res.keep_cnt[0]) == 0 and float(xd.abs().max()) == 0.0
    rows[0, :, :4] = torch.tensor([.5, .5, .2, .2])
    rows[0, :, 6] = 0.9                                                  # 64 identical boxes, identical scores
    rows[0, 10, 4] = 1.0
    rows[0, 10, 6] = 0.0                                                 # void row among them
    xd = rows.to(DEV)
    res = ops.nms_(xd, want_lists=True)
    order, keep = lists(res, 0)
    assert order.tolist() == [i for i in range(64) if i != 10]            # stable: ties keep row order
    assert keep.tolist() == [0]
    assert float(xd[0, 10, 4]) == 0.0                                    # void score of a non-candidate is masked too
    single = torch.zeros(1, 5, 25)
    single[0, 3, :4] = torch.tensor([.3, .3, .1, .1])
    single[0, 3, 9] = 0.7
    sd = single.to(DEV)
    res = ops.nms_(sd, want_lists=True)
    assert lists(res, 0)[1].tolist() == [3] and torch.equal(sd.cpu(), single)
    with pytest.raises(RuntimeError):
        ops.nms_(torch.zeros(1, 4, 25))                                  # CPU tensor: no fallback


def test_fused_postprocess_equals_staged(priors_cpu, priors_gpu):
    o, t = synth.make_batch(6, 111, "D2")
    o = synth.plant_detections(o, t, priors_cpu, 111)
    staged = o.to(DEV)
    staged[:, :, :4] = utils.calc_coordicate(pr=staged, df=priors_gpu)
    staged[:, :, 4:] = utils.calc_score(pr=staged)
    scored = staged.clone()
    staged = utils.non_maximum_suppression(outputs=staged, iou_thresh=0.45)
    fused = o.to(DEV)
    out = utils.postprocess(fused, priors_gpu, iou_thresh=0.45)
    assert out is fused
    assert torch.equal(fused, staged)
    # and both equal the oracle's NMS applied to the device's own decoded + scored tensor
    want, _ = head.nms_inplace(scored.cpu(), iou_thresh=0.45)
    assert torch.equal(fused.cpu(), want)
    # end to end against the pure-CPU oracle pipeline: boxes within 1e-5, same rows kept
    pure, keeps = head.nms_inplace(scored_cpu(o, priors_cpu), iou_thresh=0.45)
    torch.testing.assert_close(fused[:, :, :4].cpu(), pure[:, :, :4], rtol=1e-5, atol=1e-7)
    same = ((fused[:, :, 4:].sum(dim=2) > 0).cpu() == (pure[:, :, 4:].sum(dim=2) > 0)).float().mean()
    assert float(same) > 0.9995


def test_nms_properties_full_size(priors_gpu):
    # config 3 of BASELINE.json at full width: batch 256; size-independent invariants instead of the (slow) oracle
    o = synth.make_outputs(256, 121, "D2").to(DEV)
    res = ops.postprocess_(o, priors_gpu, iou_thresh=0.45, want_lists=True)
    kc, oc = res.keep_cnt.cpu(), res.order_cnt.cpu()
    assert int(oc.min()) > 0 and bool((kc <= oc).all()) and bool((kc > 0).all())
    scores = o[:, :, 4:]
    assert int((scores > 0).sum(dim=2).max()) == 1                                   # at most one class per row
    assert torch.equal((scores.sum(dim=2) > 0).sum(dim=1).cpu().int(), kc)           # exactly the kept rows keep a score
    assert float(scores[:, :, 0].abs().max()) == 0.0                                 # void column always masked
    # kept boxes are mutually non-overlapping above the threshold (spot-check 8 images)
    for n in range(0, 256, 32):
        keep = res.keep[n, : int(kc[n])].long()
        boxes = o[n, keep][None, :, :4]
        iou = utils.calc_iou(boxes, boxes)[0]
        iou.fill_diagonal_(0)
        assert float(iou.max()) <= 0.45
        key = o[n, keep, 5:].max(dim=1).values
        assert bool((key[:-1] >= key[1:]).all())                                     # kept list is in score order
    # idempotence: NMS of an NMS result changes nothing
    again = o.clone()
    ops.nms_(again, iou_thresh=0.45)
    assert torch.equal(again, o)


def test_eval_planted_vs_oracle(priors_cpu):
    o, t = synth.make_batch(8, 131, "D2", 12)
    o = synth.plant_detections(o, t, priors_cpu, 131, per_gt=4, jitter=0.3)
    x, _ = head.nms_inplace(scored_cpu(o, priors_cpu))
    want, _ = head.eval_batch(x, t)
    tallies = torch.zeros(20, 3, dtype=torch.int64, device=DEV)
    for lo in (0, 4):                                                   # two batches accumulate into one tally
        evaluate.accumulate(x[lo:lo + 4].to(DEV), t[lo:lo + 4].to(DEV), tallies)
    assert torch.equal(tallies.cpu(), want)
    assert int(want[:, 0].sum()) > 20
    # an image without ground truth and an image without detections
    x2 = x[:2].clone()
    t2 = t[:2].clone()
    t2[0] = 0
    x2[1, :, 4:] = 0
    want2, _ = head.eval_batch(x2, t2)
    got2, _ = evaluate.accumulate(x2.to(DEV), t2.to(DEV))
    assert torch.equal(got2.cpu(), want2)
    res = torch.tensor([[1., .9], [0., .8], [1., .7], [0., .6]])
    assert float(evaluate.calc_average_precision(res.to(DEV), 4)) == pytest.approx(float(head.average_precision(res, 4)))
    assert evaluate.get_order(x[0].to(DEV), 3).cpu().tolist() == head.class_order(x[0], 3).tolist()


def test_compact_detections_and_voc_ap(priors_cpu, priors_gpu):
    # "next" rows: compact detection lists and the true VOC AP, against the oracle pipeline
    o, t = synth.make_batch(6, 171, "D2", 10)
    o = synth.plant_detections(o, t, priors_cpu, 171, per_gt=4, jitter=0.3)
    x, keeps = head.nms_inplace(scored_cpu(o, priors_cpu))
    xd = x.to(DEV)
    res = ops.nms_(scored_cpu(o, priors_cpu).to(DEV), want_lists=True)
    dets, cnt = ops.gather_detections(xd, res.keep, res.keep_cnt, max_det=100)
    for n in range(6):
        k = min(len(keeps[n]), 100)
        assert int(cnt[n]) == k
        rows = x[n, keeps[n][:k]]
        assert torch.equal(dets[n, :k, :4].cpu(), rows[:, :4])
        assert torch.equal(dets[n, :k, 4].cpu(), rows[:, 5:].max(dim=1).values)
        assert torch.equal(dets[n, :k, 5].cpu().long(), rows[:, 4:].argmax(dim=1))
        assert float(dets[n, k:].abs().sum()) == 0.0
    d2, c2 = utils.detect(o.to(DEV), priors_gpu, top_k=50)
    assert d2.shape == (6, 50, 6) and int(c2.max()) <= 50 and bool((d2[:, :, 4][:, :-1] >= d2[:, :, 4][:, 1:]).all())
    ev = evaluate.DetectionEvaluator()
    ev.update(xd[:3], t[:3].to(DEV))
    ev.update(xd[3:], t[3:].to(DEV))
    out = ev.compute()
    tallies, results = head.eval_batch(x, t)
    assert torch.equal(out["tallies"].cpu(), tallies)
    for c in range(20):
        if results[c]:
            r = torch.cat(results[c])
            want = head.voc_ap_numpy(r[:, 1].numpy(), r[:, 0].numpy(), int(tallies[c, 2]))
        else:
            want = float("nan") if int(tallies[c, 2]) == 0 else 0.0
        got = float(out["ap_voc"][c])
        assert (np.isnan(want) and np.isnan(got)) or got == pytest.approx(want, rel=1e-5, abs=1e-6), (c, got, want)


# ------------------------------------------------------------------------------------------------ 8f-1 head producer
SSD_LEVELS = [(38, 4), (19, 6), (10, 6), (5, 6), (3, 4), (1, 4)]


def _reference_tail(levels, width):
    n = levels[0].shape[0]
    return torch.cat([t.permute(0, 2, 3, 1).reshape(n, -1, width) for t in levels], dim=1)      # ssd.py:103-104


@pytest.mark.gpu
@pytest.mark.parametrize("n,width,levels", [(3, 25, SSD_LEVELS), (2, 9, [(7, 3), (2, 1), (33, 2)]), (1, 68, [(5, 6)])])
def test_pack_head_matches_permute_reshape_cat(n, width, levels):
    """ssdh_pack_head is a bit-exact copy of the reference's forward tail, and its backward routes gradients the same way."""
    g = torch.Generator().manual_seed(n * 100 + width)
    xs = [torch.randn(n, a * width, m, m, generator=g).to(DEV).requires_grad_(True) for m, a in levels]
    out = ops.pack_head(xs, width)
    want = _reference_tail([x.detach() for x in xs], width)
    assert out.shape == want.shape and torch.equal(out, want)
    w = torch.randn(out.shape, generator=g).to(DEV)
    (out * w).sum().backward()
    ys = [x.detach().clone().requires_grad_(True) for x in xs]
    (_reference_tail(ys, width) * w).sum().backward()
    for x, y in zip(xs, ys):
        assert torch.equal(x.grad, y.grad)


@pytest.mark.gpu
@pytest.mark.parametrize("n,width,levels", [(3, 25, SSD_LEVELS), (2, 9, [(7, 3), (2, 1), (33, 2)]), (5, 7, [(3, 1), (4, 5)])])
def test_pack_head_channels_last_producers(n, width, levels):
    """Detector outputs in torch.channels_last format (what cuDNN's tensor-core convolutions write) take the copy path
    (ssdh_pack_head_nhwc / ssdh_unpack_head_nhwc): same values as the reference tail, gradients in channels_last too."""
    g = torch.Generator().manual_seed(n * 1000 + width)
    xs = [torch.randn(n, a * width, m, m, generator=g).to(DEV).contiguous(memory_format=torch.channels_last).requires_grad_(True) for m, a in levels]
    assert all(x.is_contiguous(memory_format=torch.channels_last) for x in xs)
    out = ops.pack_head(xs, width)
    want = _reference_tail([x.detach() for x in xs], width)
    assert out.shape == want.shape and torch.equal(out, want)
    w = torch.randn(out.shape, generator=g).to(DEV)
    (out * w).sum().backward()
    ys = [x.detach().clone().requires_grad_(True) for x in xs]
    (_reference_tail(ys, width) * w).sum().backward()
    for x, y in zip(xs, ys):
        assert torch.equal(x.grad, y.grad)
        assert x.grad.is_contiguous(memory_format=torch.channels_last)
    # a mixed list (one level NCHW) falls back to the transposing kernel and still agrees
    mixed = [x.detach() if i else x.detach().contiguous() for i, x in enumerate(xs)]
    assert torch.equal(ops.pack_head(mixed, width), want)


@pytest.mark.gpu
def test_pack_head_rejects_bad_shapes():
    x = torch.zeros(2, 26, 3, 3, device=DEV)
    with pytest.raises(ValueError):
        ops.pack_head([x], 25)


# ------------------------------------------------------------------------------------------------ 8f-3 ground-truth ingest
@pytest.mark.gpu
def test_compact_ground_truth_round_trip_equals_pad_sequence():
    """collate_fn_compact + ssdh_expand_targets rebuild exactly what the reference's collate_fn (pad_sequence) builds."""
    from torch.nn.utils.rnn import pad_sequence
    g = torch.Generator().manual_seed(3)
    batch = []
    for n, rows in enumerate([3, 0, 7, 1]):
        gt = torch.zeros(rows, 25)
        gt[:, :4] = torch.rand(rows, 4, generator=g)
        labels = torch.randint(1, 21, (rows,), generator=g)
        gt[torch.arange(rows), 4 + labels] = 1.0
        batch.append((torch.zeros(3, 4, 4), gt))
    want = pad_sequence([gt for _, gt in batch], batch_first=True)                 # src/utils.py:15
    images, compact, lengths = utils.collate_fn_compact(batch)
    assert images.shape == (4, 3, 4, 4) and compact.shape == (4, 7, 5) and lengths.tolist() == [3, 0, 7, 1]
    got = utils.targets_from_compact(compact, lengths, 21)
    assert torch.equal(got.cpu(), want)
    # without lengths every row is taken as real: zero rows then carry the void label like any other row would
    full = ops.expand_targets(compact.to(DEV), None, 21).cpu()
    assert torch.equal(full[2], want[2]) and float(full[1, 0, 4]) == 1.0


@pytest.mark.gpu
def test_dense_batch_matches_per_image_calls(priors_gpu):
    """More dense images than clusters: every cluster of the dense NMS kernel walks over several images (and skips the
    trained-like ones nms_small has already settled).  Each image must come out exactly as when it is processed alone."""
    n = 44
    o = synth.make_outputs(n, 131, "D1")
    sparse = synth.make_outputs(n, 132, "D2")
    mix = [3, 17, 40]                                        # a few images nms_small handles, in between the dense ones
    o[mix] = sparse[mix]
    batch = o.to(DEV)
    res = ops.postprocess_(batch, priors_gpu, iou_thresh=0.45, want_lists=True)
    assert int((res.order_cnt > 512).sum()) == n - len(mix)
    for i in (0, 3, 16, 17, 36, 37, 43):
        single = o[i:i + 1].to(DEV)
        r1 = ops.postprocess_(single, priors_gpu, iou_thresh=0.45, want_lists=True)
        o_b, k_b = lists(res, i)
        o_s, k_s = lists(r1, 0)
        assert torch.equal(o_b, o_s) and torch.equal(k_b, k_s), i
        assert torch.equal(batch[i], single[0]), i


@pytest.mark.gpu
def test_ssd_forward_uses_the_head_producer(priors_cpu):
    """SSD.forward (reference ssd.py:96-104): (N, 3, 300, 300) -> (N, 8732, 25) through ssdh_pack_head, identical in
    value and gradient to the reference's permute / reshape / cat tail over the same detector outputs."""
    from object_detection_torch2_b200.model import SSD
    torch.manual_seed(0)
    net = SSD(num_classes=21).to(DEV).eval()
    assert torch.equal(net.default_bboxes, priors_cpu)
    x = torch.rand(2, 3, 300, 300, device=DEV)
    y = net(x)
    assert y.shape == (2, 8732, 25)
    # the same network with the reference tail
    levels, h = [], net.normalize(x)
    for name, layer in net.features.items():
        h = layer(h)
        if name.startswith("act") and "det" + name[3:] in net.detectors:
            levels.append(net.detectors["det" + name[3:]](h))
    want = _reference_tail(levels, 25)
    assert torch.equal(y, want)
    w = torch.randn_like(y)
    g1 = torch.autograd.grad((y * w).sum(), net.detectors["det_4_3"].weight, retain_graph=True)[0]
    g2 = torch.autograd.grad((want * w).sum(), net.detectors["det_4_3"].weight)[0]
    torch.testing.assert_close(g1, g2, rtol=1e-5, atol=1e-6)


@pytest.mark.gpu
def test_training_step_through_the_whole_module():
    """The reference's training step (src/train.py:118-121): forward, SSD.loss with kwargs, backward -- gradients reach
    the trainable layers through ssdh_multibox_loss (fused backward) and ssdh_unpack_head."""
    from object_detection_torch2_b200.model import SSD
    torch.manual_seed(1)
    net = SSD(num_classes=21).to(DEV).train()
    images = torch.rand(2, 3, 300, 300, device=DEV)
    targets = synth.make_targets(2, 141, 5).to(DEV)
    outputs = net(images)
    loss = net.loss(outputs=outputs, targets=targets, default_bboxes=net.default_bboxes.to(DEV))
    assert loss.dim() == 0 and torch.isfinite(loss)
    loss.backward()
    g = net.detectors["det_7_1"].weight.grad
    assert g is not None and torch.isfinite(g).all() and float(g.abs().sum()) > 0.0
    assert net.features["conv_1_1"].weight.grad is None            # the VGG trunk is frozen (reference ssd.py:31-32)
    assert list(net.train_params())                                 # optimiser parameter groups as in the reference


@pytest.mark.gpu
def test_eval_from_kept_lists_equals_dense_scan(priors_cpu, priors_gpu):
    """ssdh_eval_accumulate_kept (tallies fed from the NMS pass's kept lists, O(kept) rows per image) against the dense scan
    and the oracle: tallies and TP flags identical, for the reference call and for the opt-in NMS variants."""
    o, t = synth.make_batch(8, 231, "D2", 12)
    o = synth.plant_detections(o, t, priors_cpu, 231, per_gt=4, jitter=0.3)
    td = t.to(DEV)
    for kw in ({}, {"iou_thresh": 0.45, "per_class": True, "score_thresh": 0.01, "top_k": 200}, {"top_k": 5}):
        x = o.to(DEV)
        res = ops.postprocess_(x, priors_gpu, want_lists=True, **kw)
        dense, dflags = evaluate.accumulate(x, td, want_flags=True)
        kept, kflags = evaluate.accumulate(x, td, want_flags=True, keep=res.keep, keep_cnt=res.keep_cnt)
        assert torch.equal(dense, kept) and torch.equal(dflags, kflags)
        want, _ = head.eval_batch(x.cpu(), t)
        assert torch.equal(kept.cpu(), want)
        assert int(kept[:, 0].sum()) > 0
    # dense random-init image (thousands of kept rows) and an image with nothing kept
    o1 = synth.make_outputs(2, 232, "D1")
    o1[1, :, 4] += 30.0                                              # every row is void: no candidates
    x = o1.to(DEV)
    res = ops.postprocess_(x, priors_gpu, want_lists=True)
    t1 = synth.make_targets(2, 232).to(DEV)
    dense, _ = evaluate.accumulate(x, t1)
    kept, _ = evaluate.accumulate(x, t1, keep=res.keep, keep_cnt=res.keep_cnt)
    assert torch.equal(dense, kept) and int(res.keep_cnt[1]) == 0 and int(kept[:, 1].sum()) == int(res.keep_cnt[0])
    with pytest.raises(ValueError):
        evaluate.accumulate(x, t1, keep=res.keep)                    # keep without keep_cnt


@pytest.mark.gpu
def test_eval_overflow_is_reported():
    """More than P positive score entries in one image (rows with several positive classes: input that skipped calc_score)
    cannot be tallied by the dense scan; the status word is read back and raised instead of returning wrong tallies."""
    from object_detection_torch2_b200 import _lib
    x = torch.rand(1, 64, 25, device=DEV) + 0.1                      # 64 rows x 20 positive classes > P = 64
    t = synth.make_targets(1, 7, 3).to(DEV)
    with pytest.raises(_lib.SsdHeadError, match="more than P"):
        evaluate.accumulate(x, t)
    ok = torch.zeros(1, 64, 25, device=DEV)
    ok[0, :, 6] = 0.5
    evaluate.accumulate(ok, t)                                        # the status word was reset: the next call is clean


@pytest.mark.gpu
def test_voc_ap_kernel_matches_numpy_oracle():
    """ssdh_voc_ap (one radix sort + one CTA per class) against the numpy restatement of the VOC devkit formula and the torch
    restatement, both metrics: ties in score, classes without detections / without ground truth, segments longer than one
    1024-wide chunk, all-TP and all-FP classes."""
    g = torch.Generator().manual_seed(17)
    NC = 20
    sizes = [0, 1, 5, 300, 1024, 1025, 2500, 0, 40, 7, 64, 1, 2, 3, 900, 33, 0, 10, 4096, 12]
    cls = torch.cat([torch.full((n,), c, dtype=torch.int32) for c, n in enumerate(sizes)])
    D = cls.numel()
    perm = torch.randperm(D, generator=g)
    cls = cls[perm]
    scores = torch.rand(D, generator=g)
    scores[::7] = scores[3]                                       # plenty of exact ties: stable order decides
    tp = (torch.rand(D, generator=g) > 0.55).to(torch.uint8)
    tp[cls == 4] = 1                                              # all true positives
    tp[cls == 5] = 0                                              # all false positives
    tallies = torch.zeros(NC, 3, dtype=torch.int64)
    tallies[:, 2] = torch.randint(1, 3000, (NC,), generator=g)
    tallies[9, 2] = 0                                             # detections but no ground truth -> NaN
    tallies[16, 2] = 0                                            # neither
    for m07 in (False, True):
        got = ops.voc_ap(scores.to(DEV), tp.to(DEV), cls.to(DEV), tallies.to(DEV), use_07_metric=m07).cpu()
        for c in range(NC):
            sel = cls == c
            want = head.voc_ap_numpy(scores[sel].numpy(), tp[sel].float().numpy(), int(tallies[c, 2]), m07)
            host = float(evaluate.voc_average_precision(scores[sel], tp[sel].float(), int(tallies[c, 2]), m07))
            if np.isnan(want):
                assert np.isnan(float(got[c])) and np.isnan(host), c
            else:
                assert float(got[c]) == pytest.approx(want, rel=1e-6, abs=1e-7), (c, m07, float(got[c]), want)
                assert host == pytest.approx(want, rel=1e-6, abs=1e-7)
    # no detections at all
    empty = ops.voc_ap(torch.zeros(0, device=DEV), torch.zeros(0, dtype=torch.uint8, device=DEV), torch.zeros(0, dtype=torch.int32, device=DEV),
                       tallies.to(DEV)).cpu()
    assert torch.isnan(empty[9]) and float(empty[0]) == 0.0
