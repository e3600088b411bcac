"""Stacked expert parameters (``GroupedExperts``): the grouped-GEMM MoE path keeps per-expert checkpoint names and the loop fallback
computes the same function as a layer with ordinary per-expert parameters (reference layout: moe/moe_layer.py:158-235)."""
import torch

from paddlefleetx_b200.models.language_model.moe.grouped_experts import GroupedExperts, reference_grouped_ffn
from paddlefleetx_b200.models.language_model.moe.moe_layer import ExpertLayer, MoELayer


def _layer(fused, seed=3):
    torch.manual_seed(seed)
    ex = [ExpertLayer(128, 256) for _ in range(3)]
    return MoELayer(128, ex, gate={"type": "naive", "top_k": 2}, fused_p2p=fused)


def test_state_dict_names_and_roundtrip():
    plain, grouped = _layer(False), _layer(True, seed=4)
    assert grouped.grouped is not None and plain.grouped is None
    assert sorted(plain.state_dict().keys()) == sorted(grouped.state_dict().keys())
    grouped.load_state_dict(plain.state_dict())
    for k, v in plain.state_dict().items():
        assert torch.equal(v, grouped.state_dict()[k]), k
    # and back: a grouped checkpoint loads into per-expert parameters
    plain2 = _layer(False, seed=5)
    plain2.load_state_dict(grouped.state_dict())
    assert torch.equal(plain2.experts[2].h4toh.weight, grouped.grouped.w2[2])
    names = [n for n, _ in grouped.named_parameters()]
    assert sum(n.startswith("grouped.") for n in names) == 4 and not any(n.startswith("experts.") for n in names)
    assert all(getattr(p, "is_expert", False) for n, p in grouped.named_parameters() if n.startswith("grouped."))


def test_loop_path_matches_and_grads_land_in_stacks():
    plain, grouped = _layer(False), _layer(True, seed=4)
    grouped.load_state_dict(plain.state_dict())
    x = torch.randn(40, 128)
    y0 = plain(x)
    y1 = grouped(x)
    assert torch.allclose(y0, y1, atol=1e-6)
    y0.square().sum().backward(); y1.square().sum().backward()
    for e in range(3):
        assert torch.allclose(plain.experts[e].htoh4.weight.grad, grouped.grouped.w1.grad[e], atol=1e-5)
        assert torch.allclose(plain.experts[e].h4toh.bias.grad, grouped.grouped.b2.grad[e], atol=1e-5)


def test_views_follow_repointed_storage():
    """Flat-buffer optimizers re-point ``param.data``; the expert modules must see the new storage after ``bind_views``."""
    layer = _layer(True)
    ge: GroupedExperts = layer.grouped
    new = torch.zeros_like(ge.w1.data)
    ge.w1.data = new
    layer(torch.randn(8, 128))          # the loop path re-binds before use
    assert layer.experts[0].htoh4.weight.data_ptr() == new[0].data_ptr()
    seg2 = [0, 3, 3, 0, 3, 5]
    out = reference_grouped_ffn(torch.randn(8, 128), seg2, ge)
    assert out.shape == (8, 128) and torch.isfinite(out).all()


def test_gate_bookkeeping_is_sync_free_and_exact():
    """number_count / prune_gate_by_capacity without boolean indexing or bincount (both force a device->host wait on CUDA): same results
    as the obvious formulation, dropped slots (-1) ignored."""
    from paddlefleetx_b200.models.language_model.moe import utils as U

    torch.manual_seed(0)
    g = torch.randint(-1, 8, (100, 2))
    flat = g.reshape(-1)
    assert torch.equal(U.number_count(g, 8), torch.bincount(flat[flat >= 0], minlength=8))
    cap = torch.tensor([5, 3, 100, 0, 7, 2, 9, 1])
    out = U.prune_gate_by_capacity(g, cap, 8, 1).reshape(-1)
    seen, want = [0] * 8, []
    for v in flat.tolist():
        keep = v >= 0 and seen[v] < int(cap[v])
        if keep:
            seen[v] += 1
        want.append(v if keep else -1)
    assert out.tolist() == want


def test_named_optimizer_state_crosses_the_two_expert_spellings():
    """Universal (named) optimizer checkpoints: state written by an expert-loop run loads into a grouped-GEMM run and back."""
    from paddlefleetx_b200.optims.optimizer import _named_entry

    E, names = 3, {"w1": ("htoh4", "weight", (3, 8, 4)), "b2": ("h4toh", "bias", (3, 4))}
    per = {}
    for e in range(E):
        per[f"layers.2.moe.experts.{e}.htoh4.weight"] = {"moment1": torch.full((8, 4), float(e)), "moment2": torch.full((8, 4), e + 0.5), "master": None}
        per[f"layers.2.moe.experts.{e}.h4toh.bias"] = {"moment1": torch.full((4,), float(e)), "moment2": torch.full((4,), e + 0.5), "master": torch.full((4,), e + 0.25)}
    st = _named_entry(per, "layers.2.moe.grouped.w1")
    assert st["moment1"].shape == (3, 8, 4) and float(st["moment1"][2, 0, 0]) == 2.0 and st["master"] is None
    st = _named_entry(per, "layers.2.moe.grouped.b2")
    assert st["master"].shape == (3, 4) and float(st["master"][1, 0]) == 1.25
    stacked = {"layers.2.moe.grouped.w1": {"moment1": torch.arange(3.0).view(3, 1, 1).expand(3, 8, 4), "moment2": torch.zeros(3, 8, 4), "master": None}}
    st = _named_entry(stacked, "layers.2.moe.experts.1.htoh4.weight")
    assert st["moment1"].shape == (8, 4) and float(st["moment1"][0, 0]) == 1.0
    import pytest
    with pytest.raises(KeyError):
        _named_entry(stacked, "layers.2.moe.experts.1.h4toh.weight")


def test_stacked_expert_biases_stay_out_of_weight_decay():
    from paddlefleetx_b200.optims.optimizer import default_decay_fn

    plain, grouped = _layer(False), _layer(True)
    decayed = lambda layer: sorted(n for n, p in layer.named_parameters() if default_decay_fn(n, p))
    assert all("bias" not in n for n in decayed(plain))
    g = decayed(grouped)
    assert "grouped.w1" in g and "grouped.w2" in g and "grouped.b1" not in g and "grouped.b2" not in g
