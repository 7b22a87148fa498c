"""Host-side check of the arithmetic behind the dense NMS sweep's division-free verdict (csrc/post.cu): the numpy restatement
in tools/check_nms_band.py must never contradict the reference's fp32 IoU test (src/utils.py:58-77, :108) on pairs placed at the
threshold.  (The kernel itself is checked against the oracle's keep lists in tests/test_gpu_post.py.)"""
import importlib.util
import os

HERE = os.path.dirname(os.path.abspath(__file__))


def test_sweep_band_never_contradicts_the_reference():
    spec = importlib.util.spec_from_file_location("check_nms_band", os.path.join(HERE, "..", "tools", "check_nms_band.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    assert mod.run(n=100_000, seed=7, out=lambda *_: None) == 0
