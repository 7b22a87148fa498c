"""CPU: the C-ABI library builds for sm_100a, loads, and exports every symbol include/ssdhead.h declares
(no compute calls -- those need a GPU); host-side helpers; world_size-2 gloo run of the sharding logic."""
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_header_symbols():
    from object_detection_torch2_b200 import _lib, build
    build.build()
    lib = _lib.load()
    header = open(os.path.join(ROOT, "include", "ssdhead.h")).read()
    declared = set(re.findall(r"SSDH_API[^;(]*?\b(ssdh_\w+)\s*\(", header))
    assert len(declared) >= 20
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.ssdh_version() == 100
    # pure host functions are callable without a GPU
    assert lib.ssdh_multibox_loss_workspace_bytes(32, 8732, 21, 20) >= 32 * 8
    assert lib.ssdh_nms_workspace_bytes(4, 8732, 21) >= 4 * 8732 * 13
    # argument validation happens before any CUDA call
    assert lib.ssdh_default_boxes(None, None) == -1
    assert b"NULL" in lib.ssdh_last_error()
    assert lib.ssdh_multibox_loss(None, None, None, 1, 1, 1, 1, 1.0, 0.25, 1, None, None, None, None, 0, None) == -1
    # the "next row" entry points validate their arguments before touching the device as well
    import ctypes
    assert lib.ssdh_pack_head(None, None, None, 1, 1, 25, None, 1, None) == -1
    ptrs, ch, hw = (ctypes.c_void_p * 1)(1), (ctypes.c_int * 1)(26), (ctypes.c_int * 1)(4)
    assert lib.ssdh_pack_head(ptrs, ch, hw, 1, 1, 25, 1, 4, None) == -1 and b"multiple of the row width" in lib.ssdh_last_error()
    ch[0] = 50
    assert lib.ssdh_pack_head(ptrs, ch, hw, 1, 1, 25, 1, 9, None) == -1 and b"rows" in lib.ssdh_last_error()
    assert lib.ssdh_pack_head(ptrs, ch, hw, 9, 1, 25, 1, 8, None) == -2          # SSDH_E_LIMIT: more than 8 levels
    assert lib.ssdh_unpack_head(None, None, None, None, 1, 1, 25, 1, None) == -1
    assert lib.ssdh_pack_head_nhwc(None, None, None, 1, 1, 25, None, 1, None) == -1
    assert lib.ssdh_pack_head_nhwc(ptrs, ch, hw, 1, 1, 25, 1, 9, None) == -1 and b"rows" in lib.ssdh_last_error()
    assert lib.ssdh_unpack_head_nhwc(None, None, None, None, 1, 1, 25, 1, None) == -1
    assert lib.ssdh_expand_targets(None, None, 2, 3, 21, None, None) == -1
    assert lib.ssdh_expand_targets(None, None, 0, 3, 21, None, None) == 0         # nothing to do
    # round-2 entry points: kept-list evaluation, VOC AP, extended loss, scalar exchange
    assert lib.ssdh_eval_accumulate_kept(1, None, None, 1, 1, 8, 21, 1, 0.5, 1, None, 1, 256, None) == -1 and b"keep" in lib.ssdh_last_error()
    assert lib.ssdh_eval_status(None, None, None) == -1
    assert lib.ssdh_voc_ap(None, None, None, 5, None, 20, 0, None, None, 0, None) == -1
    assert lib.ssdh_voc_ap(None, None, None, 0, 1, 300, 0, 1, None, 0, None) == -1            # more than 255 classes
    opt = _lib.LossOptions()
    opt.struct_bytes = 8                                                           # a caller built against another layout
    assert lib.ssdh_multibox_loss_ex(1, 1, 1, 1, 8, 21, 1, 1.0, 0.25, 1, 1, None, None, 1, 1 << 20, None, ctypes.byref(opt)) == -1
    assert b"struct_bytes" in lib.ssdh_last_error()
    assert ctypes.sizeof(_lib.LossOptions) == 48 and ctypes.sizeof(_lib.ScalarExchange) == 16 + 8 * _lib.MAX_RANKS + 8
    assert lib.ssdh_scalar_exchange_bytes(8) == 8 * _lib.XCHG_RING * 8 + 64 and lib.ssdh_scalar_exchange_bytes(0) == 0
    assert lib.ssdh_scalar_exchange_bytes(_lib.MAX_RANKS + 1) == 0
    assert lib.ssdh_scalar_exchange_reduce(None, 4, None, None, None) == -1
    x = _lib.ScalarExchange()
    x.world, x.rank, x.ring = 2, 0, _lib.XCHG_RING
    assert lib.ssdh_scalar_exchange_reduce(ctypes.byref(x), 4, 1, None, None) == -1          # no counters / inboxes
    assert lib.ssdh_scalar_exchange_create(0, None, None) == -1 and lib.ssdh_scalar_exchange_open(None, None) == -1
    assert lib.ssdh_scalar_exchange_close(None) == 0 and lib.ssdh_scalar_exchange_destroy(None) == 0


def test_stats_struct_layout():
    import ctypes
    from object_detection_torch2_b200 import _lib, ops
    assert ctypes.sizeof(_lib.ImageStats) == 32 == ops.STATS_DTYPE.itemsize
    assert [f[0] for f in _lib.ImageStats._fields_] == list(ops.STATS_DTYPE.names)


def test_no_cpu_fallback():
    from object_detection_torch2_b200 import ops, utils
    from object_detection_torch2_b200.model import SSD
    with pytest.raises(RuntimeError, match="CUDA tensors only"):
        utils.calc_score(torch.zeros(1, 4, 25))
    with pytest.raises(RuntimeError, match="CUDA tensors only"):
        SSD.loss(None, outputs=torch.zeros(1, 8, 25), targets=torch.zeros(1, 1, 25), default_bboxes=torch.rand(8, 4))
    with pytest.raises(RuntimeError):
        ops.nms_(torch.zeros(1, 4, 25))
    # the product package never imports the oracle
    import object_detection_torch2_b200 as pkg
    pkg_dir = os.path.dirname(pkg.__file__)
    for dirpath, _, files in os.walk(pkg_dir):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, f


def test_host_side_eval_helpers(priors_cpu):
    from object_detection_torch2_b200 import evaluate
    from oracle import head
    g = torch.Generator().manual_seed(5)
    for _ in range(5):
        n = int(torch.randint(1, 40, (1,), generator=g))
        res = torch.stack([(torch.rand(n, generator=g) > 0.6).float(), torch.rand(n, generator=g)], dim=1)
        cnt = int(torch.randint(1, 30, (1,), generator=g))
        assert float(evaluate.calc_average_precision(res, cnt)) == pytest.approx(float(head.average_precision(res, cnt)), rel=1e-6)
    t = torch.tensor([[10, 40, 100], [0, 5, 7]])
    ap = evaluate.average_precision_from_tallies(t)
    assert ap.tolist() == pytest.approx([0.1, 0.0])
    rows = torch.zeros(6, 25)
    rows[[1, 3, 4], 8] = torch.tensor([0.2, 0.9, 0.2])
    assert evaluate.get_order(rows, 3).tolist() == [3, 1, 4] == head.class_order(rows, 3).tolist()


def test_voc_average_precision_matches_numpy_oracle():
    from object_detection_torch2_b200 import evaluate
    from oracle import head
    g = torch.Generator().manual_seed(9)
    for trial in range(12):
        n = int(torch.randint(0, 60, (1,), generator=g))
        scores = torch.rand(n, generator=g)
        if n > 4:
            scores[1] = scores[3]                      # a tie: stable order
        tp = (torch.rand(n, generator=g) > 0.5).float()
        n_gt = int(torch.randint(1, 40, (1,), generator=g))
        for m07 in (False, True):
            got = float(evaluate.voc_average_precision(scores, tp, n_gt, m07))
            assert got == pytest.approx(head.voc_ap_numpy(scores.numpy(), tp.numpy(), n_gt, m07), rel=1e-6, abs=1e-7)
    assert torch.isnan(evaluate.voc_average_precision(torch.rand(3), torch.ones(3), 0))
    # perfect ranking -> AP 1; all misses -> 0
    assert float(evaluate.voc_average_precision(torch.tensor([.9, .8]), torch.ones(2), 2)) == pytest.approx(1.0)
    assert float(evaluate.voc_average_precision(torch.tensor([.9, .8]), torch.zeros(2), 2)) == 0.0


def test_shard_bounds():
    from object_detection_torch2_b200 import parallel
    for n, w in ((1024, 8), (4952, 8), (5, 4), (3, 8)):
        spans = [parallel.shard_bounds(n, w, r) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        sizes = [hi - lo for lo, hi in spans]
        assert max(sizes) - min(sizes) <= 1


def _gloo_worker(rank, world, port, out):
    import torch.distributed as dist
    from object_detection_torch2_b200 import parallel, synth
    from oracle import head
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    torch.set_num_threads(2)
    priors = head.default_boxes()[::37].contiguous()          # 236 priors keep the CPU stand-in fast
    N = 6
    o = synth.make_outputs(N, 7, "D2", num_priors=priors.shape[0])
    t = synth.make_targets(N, 7, 6)
    so, st = parallel.shard_batch(o, t, world, rank)
    # the oracle stands in for the kernels here (CPU test of the HOST-side sharding / reduction logic only)
    local = head.multibox_loss(so, st, priors)
    red = parallel.ScalarAllReducer(width=3, window=2)
    for step in range(3):                                      # 3 steps, window 2: one full and one partial flush
        row = torch.tensor([float(local["loss_per_image"].sum()) / N * (step + 1), so.shape[0], float(local["pos_raw"].sum())],
                           dtype=torch.float64)
        red.push(row)
    rows = red.flush()
    x = o.clone()
    x[:, :, :4] = head.decode_boxes(x, priors)
    x[:, :, 4:] = head.class_scores(x)
    x, _ = head.nms_inplace(x)
    lo, hi = parallel.shard_bounds(N, world, rank)
    tallies, _ = head.eval_batch(x[lo:hi], t[lo:hi])
    parallel.all_reduce_tallies(tallies)
    total = parallel.global_loss(local["loss_per_image"].sum() / N)
    if rank == 0:
        whole = head.multibox_loss(o, t, priors)
        want_t, _ = head.eval_batch(x, t)
        out.put(dict(rows=rows.numpy(), total=float(total), whole=float(whole["loss"]), pos=int(whole["pos_raw"].sum()),
                     tallies=tallies.numpy(), want_tallies=want_t.numpy()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_sharding():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = out.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res["total"] == pytest.approx(res["whole"], rel=1e-6)
    assert res["rows"].shape == (3, 3)
    np.testing.assert_allclose(res["rows"][:, 0], [res["whole"] * k for k in (1, 2, 3)], rtol=1e-6)
    assert res["rows"][:, 1].tolist() == [6, 6, 6] and res["rows"][:, 2].tolist() == [res["pos"]] * 3
    assert np.array_equal(res["tallies"], res["want_tallies"])
