"""CPU checks of op-level reference implementations (quantisation recipes)."""


def test_mx_fp8_reference_quantiser_and_linear():
    """OCP MX fp8 recipe (block of 32, power-of-two scales): the reference quantiser stays inside e4m3 precision for blocks of very different
    magnitude, the linear is a straight-through estimator, and the TP options validate the recipe name."""
    import pytest
    import torch

    from paddlefleetx_b200.ops.quant import fp8_linear, quantize_mx_reference
    from paddlefleetx_b200.parallel import tp_layers

    torch.manual_seed(0)
    x = torch.randn(6, 256) * torch.logspace(-3, 3, 8).repeat_interleave(32)          # per-block dynamic range 1e6: a per-row scale would flush blocks
    d = quantize_mx_reference(x)
    rel_block = ((d - x).reshape(6, 8, 32).norm(dim=-1) / x.reshape(6, 8, 32).norm(dim=-1))
    assert float(rel_block.max()) < 0.08                                               # every block keeps e4m3 relative precision
    w, b = torch.randn(40, 256) * 0.05, torch.randn(40)
    xr = x.clone().requires_grad_(True)
    y = fp8_linear(xr, w, b, recipe="mx")
    ref = torch.nn.functional.linear(x, w, b)
    assert float((y - ref).norm() / ref.norm()) < 0.06
    y.sum().backward()
    assert torch.allclose(xr.grad, w.sum(0).expand_as(x), atol=1e-5)                   # straight-through gradient
    assert tp_layers.configure({"fp8_tp_gemm": True})["fp8_recipe"] == "mx"
    assert tp_layers.configure({"fp8_tp_gemm": True, "fp8_recipe": "rowwise"})["fp8_recipe"] == "rowwise"
    with pytest.raises(ValueError):
        tp_layers.configure({"fp8_recipe": "nvfp4"})
    tp_layers.configure({})
