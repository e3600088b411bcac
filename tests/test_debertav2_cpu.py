"""DeBERTa-V2 pieces beyond the forward pass: masked softmax, replayable dropout, bucketed relative positions, loading a tower from disk and
encoding captions (reference language_model/debertav2/modeling.py)."""
import json
import pickle

import pytest
import torch

from paddlefleetx_b200.models.multimodal_model.debertav2 import modeling as D


def test_xsoftmax_zeroes_masked_positions_and_rows():
    x = torch.randn(2, 3, 5, dtype=torch.float64, requires_grad=True)
    mask = torch.ones(2, 3, 5, dtype=torch.uint8)
    mask[0, :, 3:] = 0
    mask[1, 1] = 0                                   # a fully masked row
    y = D.XSoftmax.apply(x, mask, -1)
    assert float(y[0, :, 3:].detach().abs().max()) == 0 and float(y[1, 1].detach().abs().max()) == 0
    torch.testing.assert_close(y[0, :, :3].sum(-1), torch.ones(3, dtype=torch.float32))
    ref = torch.softmax(x[0, :, :3], -1)
    torch.testing.assert_close(y[0, :, :3].double(), ref, atol=1e-6, rtol=1e-6)
    w = torch.randn(2, 3, 5, dtype=torch.float64)
    (y.double() * w).sum().backward()
    gx = x.grad.clone()
    x.grad = None
    xs = x[0, :, :3]
    (torch.softmax(xs, -1) * w[0, :, :3]).sum().backward()
    torch.testing.assert_close(gx[0, :, :3], x.grad[0, :, :3], atol=1e-6, rtol=1e-5)
    assert float(gx[1, 1].abs().max()) == 0 and float(gx[0, :, 3:].abs().max()) == 0


def test_stable_dropout_replays_its_masks():
    torch.manual_seed(0)
    drop = D.StableDropout(0.5).train()
    x = torch.ones(4, 64)
    drop.init_context()
    a1, b1 = drop(x), drop(x)
    assert not torch.equal(a1, b1) and set(a1.unique().tolist()) <= {0.0, 2.0}
    drop.init_context(reuse_mask=True)
    a2, b2 = drop(x), drop(x)
    assert torch.equal(a1, a2) and torch.equal(b1, b2)        # second pass sees the same masks, in order
    drop.clear_context()
    assert drop.context_stack is None and not torch.equal(drop(x), a1)
    assert torch.equal(drop.eval()(x), x)
    g = torch.ones(4, 64, requires_grad=True)
    out = D.XDropout.apply(g, 0.25)
    out.sum().backward()
    assert torch.equal(g.grad, out.detach())                   # kept elements carry 1 / (1 - p), dropped ones 0
    ctx = D.DropoutContext()
    ctx.dropout = 0.3
    m1, p1 = D.get_mask(x, ctx)
    m2, _ = D.get_mask(x, ctx)
    assert m1 is m2 and p1 == 0.3


def test_relative_position_buckets():
    rel = D.build_relative_position(6, 6)
    assert rel.shape == (1, 6, 6) and int(rel[0, 5, 0]) == 5 and int(rel[0, 0, 5]) == -5
    b = D.build_relative_position(128, 128, bucket_size=16, max_position=128)
    assert int(b.max()) <= 8 + 7 and int(b.min()) >= -(8 + 7)
    small = torch.arange(-7, 8)
    assert torch.equal(D.make_log_bucket_position(small, 16, 128), small)        # short distances keep their value
    far = D.make_log_bucket_position(torch.tensor([50, -50, 127]), 16, 128)
    assert far[0] == -far[1] and far[2] == 15


@pytest.mark.parametrize("paddle_style", [False, True])
def test_tower_from_disk_and_caption_encoding(tmp_path, paddle_style):
    from test_tokenizers_cpu import _train_spm

    from paddlefleetx_b200.data.tokenizers import get_debertav2_tokenizer

    d = tmp_path / "deberta-v2-tiny"
    d.mkdir()
    _train_spm(d / "spm.model", pad_id=0, bos_id=1, eos_id=2, unk_id=3, pad_piece="[PAD]", bos_piece="[CLS]", eos_piece="[SEP]", unk_piece="[UNK]")
    cfg = dict(vocab_size=200, hidden_size=24, num_hidden_layers=2, num_attention_heads=4, intermediate_size=48, position_buckets=8,
               relative_attention=True, norm_rel_ebd="layer_norm", pos_att_type="p2c|c2p", share_att_key=True, conv_kernel_size=3, conv_act="gelu",
               max_position_embeddings=64, model_type="deberta-v2", hidden_dropout_prob=0.1)
    (d / "config.json").write_text(json.dumps(cfg))
    torch.manual_seed(0)
    src = D.DebertaV2Model(**cfg)
    sd = src.state_dict()
    assert not any(k.endswith("position_ids") for k in sd)
    if paddle_style:
        lin = {n + ".weight" for n, m in src.named_modules() if isinstance(m, torch.nn.Linear)}
        arrays = {k: (v.t() if k in lin else v).numpy().copy() for k, v in sd.items()}
        with open(d / "debertav2.pd", "wb") as f:
            pickle.dump({"model": arrays}, f)
    else:
        torch.save({"model": sd}, d / "debertav2.pd")
    assert D.get_debertav2_encoded_dim(str(d)) == 24 and D.get_debertav2_encoded_dim("somewhere/else") == 1536
    tower = D.get_debertav2_model(str(d))
    assert not tower.training and not any(p.requires_grad for p in tower.parameters())
    for k, v in sd.items():
        torch.testing.assert_close(tower.state_dict()[k], v)
    tok = get_debertav2_tokenizer(str(d))
    feats, mask = D.debertav2_encode_text(tower, ["a photo of a small red bird", "dog"], tok, return_attn_mask=True)
    assert feats.shape[:2] == mask.shape and feats.shape[-1] == 24 and mask.dtype == torch.bool
    assert float(feats[~mask].abs().max()) == 0 and float(feats[mask].abs().max()) > 0
    assert D.get_debertav2_model(None) is None
    with pytest.raises(NotImplementedError):
        tower._prune_heads({0: [1]})
    # the Imagen model picks the tower up by directory name
    from paddlefleetx_b200.models.multimodal_model.imagen import modeling as I
    from paddlefleetx_b200.models.multimodal_model.imagen import unet as U

    u = U.Unet(dim=8, text_embed_dim=24, dim_mults=(1, 2), layer_attns=(False, True), layer_cross_attns=(False, True), attn_heads=2, attn_dim_head=4,
               max_text_len=6, attn_pool_num_latents=2, resnet_groups=4)
    m = I.ImagenModel([u], image_sizes=[8], text_encoder_name=str(d), text_embed_dim=None, timesteps=2)
    assert m.text_embed_dim == 24 and isinstance(m.text_encoder, D.DebertaV2Model)
    assert m.sample(texts=["a small bird"]).shape == (1, 3, 8, 8)
