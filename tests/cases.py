"""Golden case table shared by the generator (tests/golden/make_golden.py) and the tests."""
import importlib.util
import os

_spec = importlib.util.spec_from_file_location(
    "_make_golden", os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "make_golden.py"))


def _load():
    mod = importlib.util.module_from_spec(_spec)
    _spec.loader.exec_module(mod)
    return mod


_mg = _load()
LOSS_CASES = _mg.LOSS_CASES
POST_CASES = _mg.POST_CASES
GRAD_STRIDE = _mg.GRAD_STRIDE


def loss_inputs(case, priors):
    name, N, seed, dist, mb, a, special = case
    _mg._PRIORS = priors
    return _mg.case_inputs(N, seed, dist, mb, special)


def post_inputs(case, priors):
    name, N, seed, dist, thr = case
    _mg._PRIORS = priors
    return _mg.case_inputs(N, seed, dist, 20)
