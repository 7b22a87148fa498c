"""CPU, build container only: the oracle restatement against the LIVE reference (/root/reference/src).

Skipped wherever the reference tree is absent (e.g. the GPU box); tests/test_oracle_golden.py covers
the same ground there through the committed fixtures.
"""
import pytest
import torch

from object_detection_torch2_b200 import synth
from oracle import head, ref_loader

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="reference tree not present")


@pytest.fixture(scope="module")
def ref():
    return ref_loader.load()


def test_default_boxes(ref, priors_cpu):
    assert torch.equal(ref.SSD._get_default_bboxes(ref.net), priors_cpu)


@pytest.mark.parametrize("seed,n,dist,mb", [(11, 2, "D1", 20), (12, 3, "D2", 9), (13, 1, "D1", 3)])
def test_loss_and_grad(ref, priors_cpu, seed, n, dist, mb):
    o, t = synth.make_batch(n, seed, dist, mb)
    assert torch.equal(ref.net._match(gt=t, df=priors_cpu), head.match_mask(t, priors_cpu))
    x = o.clone().requires_grad_(True)
    loss = ref.net.loss(outputs=x, targets=t, default_bboxes=priors_cpu)
    loss.backward()
    r = head.multibox_loss(o, t, priors_cpu, want_grad=True)
    torch.testing.assert_close(r["loss"], loss.detach(), rtol=1e-6, atol=0)
    torch.testing.assert_close(r["grad"], x.grad, rtol=1e-5, atol=1e-9)
    kp, kn = ref.net._split_pos_neg(r["pos_raw"], 8732 - r["pos_raw"])
    assert torch.equal(kp, r["k_pos"]) and torch.equal(kn, r["k_neg"])
    for i in range(n):
        assert torch.equal(ref.net._k_plus_1_th_value(r["ce_neg"][i], r["k_neg"][i]), r["thr_neg"][i])


def test_all_padding_batch(ref, priors_cpu):
    o = synth.make_outputs(2, 21)
    t = torch.zeros(2, 3, 25)
    want = ref.net.loss(outputs=o, targets=t, default_bboxes=priors_cpu)
    got = head.multibox_loss(o, t, priors_cpu)
    assert float(want) == 0.0 and float(got["loss"]) == 0.0


def test_tied_negatives_select_nothing(ref, priors_cpu):
    # all-equal logits -> every negative has the same CE -> strict '>' selects none of them
    o = torch.zeros(1, 8732, 25)
    t = synth.make_targets(1, 31, 4)
    want = ref.net.loss(outputs=o.clone(), targets=t, default_bboxes=priors_cpu)
    got = head.multibox_loss(o, t, priors_cpu)
    torch.testing.assert_close(got["loss"], want, rtol=1e-6, atol=0)
    assert int(got["neg_sel"][0]) == 0


@pytest.mark.parametrize("seed,dist,thr", [(14, "D2", 0.5), (15, "D2", 0.3)])
def test_postprocess(ref, priors_cpu, seed, dist, thr):
    o, t = synth.make_batch(2, seed, dist)
    o = synth.plant_detections(o, t, priors_cpu, seed)
    a, b = o.clone(), o.clone()
    a[:, :, :4] = ref.utils.calc_coordicate(pr=a, df=priors_cpu)
    a[:, :, 4:] = ref.utils.calc_score(pr=a)
    b[:, :, :4] = head.decode_boxes(b, priors_cpu)
    b[:, :, 4:] = head.class_scores(b)
    assert torch.equal(a, b)
    assert torch.equal(ref.utils.calc_iou(a[:, :200], t), head.pair_iou(b[:, :200], t))
    a = ref.utils.non_maximum_suppression(outputs=a, iou_thresh=thr)
    b, _ = head.nms_inplace(b, iou_thresh=thr)
    assert torch.equal(a, b)
    rc, cnt = ref_loader.reference_eval_loop(ref, a, t)
    tallies, results = head.eval_batch(b, t)
    for c in range(20):
        rows = [rc[i][c] for i in sorted(rc) if c in rc[i]]
        assert len(rows) == len(results[c]) and all(torch.equal(p, q) for p, q in zip(rows, results[c]))
        assert cnt[c] == int(tallies[c, 2])
        assert torch.equal(ref.evaluate.get_order(a[0], c), head.class_order(b[0], c))
        if rows and cnt[c] > 0:
            want = ref.evaluate.calc_average_precision(result=torch.cat(rows), count=cnt[c])
            assert torch.equal(want, head.average_precision(torch.cat(results[c]), cnt[c]))
