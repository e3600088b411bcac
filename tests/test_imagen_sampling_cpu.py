"""Imagen sampling options of the reference (imagen/modeling.py:541-676): captions through the frozen text tower, inpainting, initial images and
skipped steps, per-stage guidance scales, stand-alone super-resolution, all-stage / PIL outputs."""
import json

import pytest
import torch

from paddlefleetx_b200.models.multimodal_model.imagen import modeling as I
from paddlefleetx_b200.models.multimodal_model.imagen import unet as U


def _unet(**kw):
    return U.Unet(**{**dict(dim=8, text_embed_dim=12, dim_mults=(1, 2), layer_attns=(False, True), layer_cross_attns=(False, True), attn_heads=2,
                            attn_dim_head=4, max_text_len=6, attn_pool_num_latents=2, resnet_groups=4), **kw})


def _emb(b=2):
    g = torch.Generator().manual_seed(0)
    return torch.randn(b, 4, 12, generator=g), torch.ones(b, 4)


def test_inpainting_keeps_the_given_pixels():
    torch.manual_seed(0)
    m = I.ImagenModel([_unet()], image_sizes=[8], text_embed_dim=12, timesteps=3)
    emb, mask = _emb()
    known = torch.rand(2, 3, 8, 8)
    keep = torch.zeros(2, 8, 8, dtype=torch.bool)
    keep[:, :, :4] = True
    out = m.sample(text_embeds=emb, text_masks=mask, inpaint_images=known, inpaint_masks=keep, inpaint_resample_times=2)
    assert out.shape == (2, 3, 8, 8) and float(out.min()) >= 0 and float(out.max()) <= 1
    torch.testing.assert_close(out[:, :, :, :4], known[:, :, :, :4], atol=1e-6, rtol=0)
    assert not torch.allclose(out[:, :, :, 4:], known[:, :, :, 4:])
    with pytest.raises(AssertionError):
        m.sample(text_embeds=emb, text_masks=mask, inpaint_images=known)
    assert m.training                                      # sample() restores the mode it found


def test_cascade_options():
    torch.manual_seed(0)
    m = I.ImagenModel([_unet(), _unet(lowres_cond=True)], image_sizes=[8, 16], text_embed_dim=12, timesteps=2)
    emb, mask = _emb(1)
    outs = m.sample(text_embeds=emb, return_all_unet_outputs=True, cond_scale=(1.0, 1.5), init_images=(torch.rand(1, 3, 8, 8), None),
                    skip_steps=(1, None))
    assert [tuple(o.shape) for o in outs] == [(1, 3, 8, 8), (1, 3, 16, 16)]
    only_first = m.sample(text_embeds=emb, text_masks=mask, stop_at_unet_number=1)
    assert only_first.shape == (1, 3, 8, 8)
    up = m.sample(text_embeds=emb, text_masks=mask, start_at_unet_number=2, start_image_or_video=torch.rand(1, 3, 8, 8))
    assert up.shape == (1, 3, 16, 16)
    with pytest.raises(AssertionError):
        m.sample(text_embeds=emb, start_at_unet_number=2)
    pil = m.sample(text_embeds=emb, return_pil_images=True, stop_at_unet_number=1)
    assert len(pil) == 1 and pil[0].size == (8, 8)
    with pytest.raises(AssertionError):
        m.sample()                                         # text-conditioned model without text
    with pytest.raises(AssertionError):
        m.sample(text_embeds=torch.randn(1, 4, 7))         # wrong embedding width
    # a stand-alone super-resolution stage takes its conditioning image the same way
    sr = I.ImagenModel([_unet(lowres_cond=True)], image_sizes=[16, 8], text_embed_dim=12, timesteps=2)
    assert sr.sample(text_embeds=emb, start_image_or_video=torch.rand(1, 3, 8, 8)).shape == (1, 3, 16, 16)
    uncond = I.ImagenModel([_unet(cond_on_text=False)], image_sizes=[8], condition_on_text=False, timesteps=2)
    assert uncond.sample(batch_size=3).shape == (3, 3, 8, 8)
    with pytest.raises(AssertionError):
        uncond.sample(text_embeds=emb)


def test_renoising_between_two_times_matches_forward_diffusion_statistics():
    sched = I.GaussianDiffusionContinuousTimes("cosine", 10)
    x0 = torch.zeros(20000, 1, 1, 1)
    t_lo, t_hi = torch.full((20000,), 0.3), torch.full((20000,), 0.7)
    x_lo, _, _, _ = sched.q_sample(x0 + 1.0, t_lo)
    x_hi = sched.q_sample_from_to(x_lo, t_lo, t_hi)
    direct, _, a_hi, s_hi = sched.q_sample(x0 + 1.0, t_hi)
    assert abs(float(x_hi.mean()) - float(a_hi.flatten()[0])) < 0.02          # mean alpha(t_hi) * x0
    assert abs(float(x_hi.std()) - float(s_hi.flatten()[0])) < 0.02           # std sigma(t_hi)
    assert abs(float(direct.std()) - float(x_hi.std())) < 0.03


def test_captions_go_through_the_text_tower_on_disk(tmp_path):
    from test_tokenizers_cpu import _train_spm

    from paddlefleetx_b200.models.multimodal_model.t5 import modeling as T

    d = tmp_path / "t5" / "t5-tiny"
    d.mkdir(parents=True)
    _train_spm(d / "spiece.model", pad_id=0, eos_id=1, unk_id=2, bos_id=-1, pad_piece="<pad>", eos_piece="</s>", unk_piece="<unk>")
    shape = dict(vocab_size=200, d_model=12, d_kv=4, d_ff=24, num_layers=1, num_heads=2, relative_attention_num_buckets=8, layer_norm_epsilon=1e-6,
                 feed_forward_proj="relu", dropout_rate=0.0)
    (d / "config.json").write_text(json.dumps(shape))
    torch.manual_seed(0)
    torch.save({"model": T.T5EncoderModel(**shape).state_dict()}, d / "t5.pd")
    m = I.ImagenModel([_unet()], image_sizes=[8], text_encoder_name=str(d), text_embed_dim=None, timesteps=2)
    assert m.text_embed_dim == 12 and isinstance(m.text_encoder, T.T5EncoderModel) and m.tokenizer is not None
    assert not any(p.requires_grad for p in m.text_encoder.parameters())
    emb, mask = m.encode_captions(["a photo of a small red bird", "dog"])
    assert emb.shape[0] == 2 and emb.shape[-1] == 12 and mask.shape == emb.shape[:2]
    feats = T.t5_encode_text(m.text_encoder, ["a photo of a small red bird", "dog"], m.tokenizer)
    torch.testing.assert_close(feats, emb)
    assert m.sample(texts=["a photo of a small red bird", "dog"]).shape == (2, 3, 8, 8)
    none = I.ImagenModel([_unet()], image_sizes=[8], text_encoder_name="t5/t5-11b", text_embed_dim=12, timesteps=2)
    assert none.text_encoder is None                       # directory not on this machine: pre-computed embeddings only
    with pytest.raises(AssertionError):
        none.sample(texts=["x"])
