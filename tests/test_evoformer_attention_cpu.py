"""Evoformer gated attention front-end (ops/evoformer_attention.py): the chunked recomputing backward reproduces autograd of the plain
expression (reference semantics: ppfleetx/models/protein_folding/attentions.py:35-180), for every chunking the group count allows."""
import pytest
import torch

from paddlefleetx_b200.ops import evoformer_attention as EA


class _Saved:
    pass


def _run_backward(q, k, v, mask, pair, gate, gpp, scale, go, chunk):
    """Drive _EvoAttnFn.backward on CPU with a context built from the reference forward (the kernel itself needs a GPU)."""
    G, Sq, H, D = q.shape
    logits = torch.einsum("gqhd,gkhd->ghqk", q.float(), k.float()) * scale
    if mask is not None:
        logits = logits + mask.view(G, 1, 1, -1)
    if pair is not None:
        logits = logits + pair.float().repeat_interleave(gpp, 0)
    lse = torch.logsumexp(logits, -1)
    ctx = _Saved()
    e = q.new_empty(0)
    ctx.saved_tensors = (q, k, v, mask if mask is not None else e, pair if pair is not None else e, gate if gate is not None else e, lse)
    ctx.flags = (mask is not None, pair is not None, gate is not None)
    ctx.gpp, ctx.scale = gpp, scale
    ctx.bias_dtypes = (None if mask is None else mask.dtype, None if pair is None else pair.dtype)
    ctx.needs_input_grad = (True, True, True, False, pair is not None, gate is not None, False, False)
    old = EA._CHUNK_ELEMS
    EA._CHUNK_ELEMS = chunk
    try:
        return EA._EvoAttnFn.backward(ctx, go)
    finally:
        EA._CHUNK_ELEMS = old


@pytest.mark.parametrize("chunk", [1, 4 * 12 * 10 * 2, 1 << 30])
@pytest.mark.parametrize("G,gpp", [(6, 3), (4, 4), (5, 1)])
def test_chunked_backward_matches_autograd(G, gpp, chunk):
    torch.manual_seed(G * 10 + gpp)
    Sq, Sk, H, D = 12, 10, 4, 8
    q, k, v = (torch.randn(G, s, H, D) * 0.6 for s in (Sq, Sk, Sk))
    mask = (torch.rand(G, Sk) > 0.2).float()
    mask[:, 0] = 1
    mask = (mask - 1) * 1e9
    pair, gate = torch.randn(G // gpp, H, Sq, Sk), torch.randn(G, Sq, H, D)
    leaves = [t.clone().requires_grad_(True) for t in (q, k, v, pair, gate)]
    ref = EA.reference(leaves[0], leaves[1], leaves[2], mask, leaves[3], leaves[4], gpp, D ** -0.5)
    go = torch.randn_like(ref)
    ref.backward(go)
    got = _run_backward(q, k, v, mask, pair, gate, gpp, D ** -0.5, go, chunk * H * Sq * Sk if chunk < (1 << 30) and chunk > 1 else chunk)
    for name, g_, leaf in zip(("dq", "dk", "dv"), got[:3], leaves[:3]):
        assert torch.allclose(g_, leaf.grad, atol=2e-2, rtol=2e-2), name            # bf16 matmul operands inside the chunk loop
    assert torch.allclose(got[4], leaves[3].grad, atol=2e-2, rtol=2e-2)
    assert torch.allclose(got[5], leaves[4].grad, atol=2e-2, rtol=2e-2)


def test_fallback_matches_sdpa_semantics():
    torch.manual_seed(1)
    G, S, H, D = 3, 9, 2, 8
    q, k, v = (torch.randn(G, S, H, D) for _ in range(3))
    mask = torch.zeros(G, S)
    mask[:, -2:] = -1e9
    out = EA.evoformer_attention(q, k, v, mask, None, None)
    ref = torch.nn.functional.scaled_dot_product_attention(q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2), attn_mask=mask.view(G, 1, 1, S)).transpose(1, 2)
    assert torch.allclose(out, ref, atol=1e-5)
