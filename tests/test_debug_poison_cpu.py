"""PFX_DEBUG_POISON=1 (parallel/debug_poison.py): poisoned buffers, finite checks and the barrier epoch-skew check of the peer-memory
protocols — host-side logic, gloo."""
import torch

from dist_utils import run_distributed
from paddlefleetx_b200.parallel import debug_poison as D


def test_poison_and_check(monkeypatch):
    t = torch.zeros(16)
    D.poison(t)                                    # disabled: untouched
    assert float(t.sum()) == 0.0
    monkeypatch.setenv("PFX_DEBUG_POISON", "1")
    D.poison(t)
    assert torch.isnan(t).all()
    i = torch.zeros(4, dtype=torch.int32)
    D.poison(i)
    assert int(i[0]) == torch.iinfo(torch.int32).max
    t.zero_()
    D.check_finite(t, "clean buffer")
    t[5] = float("nan")
    try:
        D.check_finite(t, "gathered activations", rank=3)
    except RuntimeError as e:
        assert "element 5" in str(e) and "rank 3" in str(e) and "gathered activations" in str(e)
    else:
        raise AssertionError("NaN not reported")


def test_barrier_skew_is_detected_across_ranks():
    run_distributed("dist_fns:barrier_skew_detected", 2)
