"""T5 beyond the encoder forward: decoder stacks with caches and cross-attention, head masks / pruning, activations, config, checkpoint reading in
both weight layouts (reference language_model/t5/modeling.py)."""
import json
import os
import pickle

import numpy as np
import pytest
import torch

from paddlefleetx_b200.models.multimodal_model.t5 import modeling as T

KW = dict(d_model=32, num_layers=2, layer_norm_epsilon=1e-6, dropout_rate=0.0, relative_attention_num_buckets=8, feed_forward_proj="gated-gelu", d_kv=8,
          num_heads=4, d_ff=64)


def _encoder(**kw):
    torch.manual_seed(0)
    return T.T5EncoderModel(vocab_size=50, **{**KW, **kw}).eval()


def test_fused_path_equals_explicit_softmax_and_padding_is_ignored():
    m = _encoder()
    ids = torch.randint(1, 50, (2, 9))
    mask = torch.ones(2, 9, dtype=torch.long)
    mask[1, 6:] = 0
    fast = m(ids, mask).last_hidden_state
    slow = m(ids, mask, output_attentions=True)
    torch.testing.assert_close(fast, slow.last_hidden_state, atol=2e-5, rtol=1e-4)
    assert float(slow.attentions[0][1, :, :, 6:].max().detach()) < 1e-3
    ids2 = ids.clone()
    ids2[1, 6:] = 3                                    # content of padded positions must not matter for the kept ones
    torch.testing.assert_close(m(ids2, mask).last_hidden_state[1, :6], fast[1, :6], atol=2e-5, rtol=1e-4)
    # all-ones head mask is the identity, zeroing every head of layer 0 leaves only the residual path there
    ones = m(ids, mask, head_mask=torch.ones(2, 4)).last_hidden_state
    torch.testing.assert_close(ones, fast, atol=2e-5, rtol=1e-4)
    assert not torch.allclose(m(ids, mask, head_mask=torch.tensor([[0.0] * 4, [1.0] * 4])).last_hidden_state, fast)


def test_decoder_stack_cache_matches_full_run_and_is_causal():
    torch.manual_seed(1)
    emb = torch.nn.Embedding(50, 32)
    dec = T.T5Stack(embed_tokens=emb, is_decoder=True, **KW).eval()
    enc_states, enc_mask = torch.randn(2, 5, 32), torch.ones(2, 5)
    enc_mask[0, 3:] = 0
    ids = torch.randint(1, 50, (2, 7))
    full = dec(input_ids=ids, encoder_hidden_states=enc_states, encoder_attention_mask=enc_mask, use_cache=True, output_attentions=True)
    assert len(full.past_key_values) == 2 and len(full.past_key_values[0]) == 4          # self k/v + cross k/v
    assert full.past_key_values[0][0].shape == (2, 4, 7, 8) and full.past_key_values[0][2].shape == (2, 4, 5, 8)
    assert len(full.cross_attentions) == 2 and full.cross_attentions[0].shape == (2, 4, 7, 5)
    assert float(full.attentions[0].triu(1).abs().max()) < 1e-3                          # no attention to the future
    assert float(full.cross_attentions[1][0, :, :, 3:].max()) < 1e-3                     # masked encoder positions
    # incremental decoding: 4 tokens, then 3 more with the cache
    first = dec(input_ids=ids[:, :4], encoder_hidden_states=enc_states, encoder_attention_mask=enc_mask, use_cache=True)
    torch.testing.assert_close(first.last_hidden_state, full.last_hidden_state[:, :4], atol=2e-5, rtol=1e-4)
    out, past = [], first.past_key_values
    for t in range(4, 7):
        step = dec(input_ids=ids[:, t:t + 1], encoder_hidden_states=enc_states, encoder_attention_mask=enc_mask, past_key_values=past, use_cache=True)
        out.append(step.last_hidden_state)
        past = step.past_key_values
    torch.testing.assert_close(torch.cat(out, 1), full.last_hidden_state[:, 4:], atol=3e-5, rtol=1e-4)
    with pytest.raises(AssertionError):
        T.T5Stack(embed_tokens=emb, **KW)(input_ids=ids, use_cache=True)                # encoders have no cache
    with pytest.raises(ValueError):
        dec(input_ids=ids, inputs_embeds=emb(ids))


def test_relative_position_buckets():
    rel = torch.arange(-40, 41)
    bi = T.T5Attention._relative_position_bucket(rel, True, 32, 128)
    assert bi.min() == 0 and bi.max() <= 31 and (bi[rel > 0] >= 16).all() and (bi[rel <= 0] < 16).all()
    assert (bi[(rel <= 0) & (rel > -8)] == -rel[(rel <= 0) & (rel > -8)]).all()          # small offsets get exact buckets
    uni = T.T5Attention._relative_position_bucket(rel, False, 32, 128)
    assert (uni[rel >= 0] == 0).all() and uni.max() <= 31                                 # a decoder only distinguishes the past
    far = T.T5Attention._relative_position_bucket(torch.tensor([-10_000, 10_000]), True, 32, 128)
    assert far.tolist() == [15, 31]


def test_prune_heads_equals_masking_them():
    m = _encoder()
    ids, mask = torch.randint(1, 50, (2, 6)), torch.ones(2, 6, dtype=torch.long)
    hm = torch.ones(2, 4)
    hm[0, 1] = hm[0, 3] = 0
    want = m(ids, mask, head_mask=hm).last_hidden_state
    att = m.encoder.block[0].layer[0].SelfAttention
    m._prune_heads({0: [1, 3]})
    assert att.n_heads == 2 and att.q.weight.shape == (16, 32) and att.o.weight.shape == (32, 16) and att.pruned_heads == {1, 3}
    # the shared relative-position bias keeps all 4 heads: block 1 (unpruned) still gets its rows, block 0 picks heads 0 and 2
    assert att.relative_attention_bias.weight.shape == (8, 4)
    torch.testing.assert_close(m(ids, mask).last_hidden_state, want, atol=2e-5, rtol=1e-4)
    torch.testing.assert_close(m(ids, mask, output_attentions=True).last_hidden_state, want, atol=2e-5, rtol=1e-4)
    heads, index = T.find_pruneable_heads_and_indices([0, 2], 4, 8, {1})          # head 2 sits at slot 1 once head 1 is gone
    assert heads == {0, 2} and index.tolist() == list(range(16, 32))
    lin = torch.nn.Linear(6, 4)
    kept = T.prune_linear_layer(lin, torch.tensor([0, 2]))
    assert torch.equal(kept.weight, lin.weight[[0, 2]]) and torch.equal(kept.bias, lin.bias[[0, 2]])
    kept_in = T.prune_linear_layer(lin, torch.tensor([1, 4, 5]), dim=1)
    assert torch.equal(kept_in.weight, lin.weight[:, [1, 4, 5]]) and torch.equal(kept_in.bias, lin.bias)


def test_activations_and_config():
    x = torch.linspace(-4, 4, 41)
    torch.testing.assert_close(T.get_activation("gelu_new")(x), torch.nn.functional.gelu(x, approximate="tanh"))
    torch.testing.assert_close(T.get_activation("gelu_python")(x), torch.nn.functional.gelu(x))
    torch.testing.assert_close(T.get_activation("gelu_fast")(x), torch.nn.functional.gelu(x, approximate="tanh"), atol=1e-3, rtol=1e-3)
    torch.testing.assert_close(T.get_activation("swish")(x), x * torch.sigmoid(x))
    torch.testing.assert_close(T.MishActivation()._mish_python(x), torch.nn.functional.mish(x))
    assert float(T.get_activation("gelu_10")(torch.tensor(50.0))) == 10.0 and T.get_activation("linear")(x) is x
    with pytest.raises(KeyError):
        T.get_activation("nope")
    with pytest.raises(ValueError):
        T.ClippedGELUActivation(1, -1)
    cfg = T.T5Config(d_model=64, num_heads=4, d_kv=16, d_ff=128, num_layers=1, vocab_size=77, dropout_rate=0.0, layer_norm_epsilon=1e-6,
                     relative_attention_num_buckets=8, feed_forward_proj="relu", return_dict=False, custom_field=3)
    assert cfg.use_return_dict is False and cfg.dense_act_fn == "gelu_new" and cfg.custom_field == 3 and cfg.relative_attention_max_distance == 128
    model = T.T5Model(cfg)
    assert model.shared.weight.shape == (77, 64) and len(model.encoder.block) == 1
    assert T.finfo(torch.float16).max == np.finfo(np.float16).max


@pytest.mark.parametrize("paddle_style", [False, True])
def test_get_t5_model_reads_config_and_weights(tmp_path, paddle_style):
    d = tmp_path / "t5" / "t5-tiny"
    d.mkdir(parents=True)
    shape = dict(vocab_size=50, d_model=32, d_kv=8, d_ff=64, num_layers=2, num_heads=4, relative_attention_num_buckets=8, layer_norm_epsilon=1e-6,
                 feed_forward_proj="relu", dropout_rate=0.1)
    (d / "config.json").write_text(json.dumps(dict(shape, model_type="t5", is_encoder_decoder=False)))
    src = T.T5EncoderModel(**shape)
    sd = src.state_dict()
    assert "encoder.embed_tokens.weight" not in sd
    if paddle_style:          # paddle.save: pickle of numpy arrays, Linear weights [in, out], embedding listed twice
        lin = (".q.weight", ".k.weight", ".v.weight", ".o.weight", ".wi.weight", ".wo.weight")
        arrays = {k: (v.t() if k.endswith(lin) else v).numpy().copy() for k, v in sd.items()}
        arrays["encoder.embed_tokens.weight"] = arrays["shared.weight"]
        with open(d / "t5.pd", "wb") as f:
            pickle.dump({"model": arrays}, f)
    else:
        torch.save({"model": sd}, d / "t5.pd")
    assert T.get_encoded_dim(str(d)) == 32
    m = T.get_t5_model(str(d), pretrained=True)
    assert not m.training and not any(p.requires_grad for p in m.parameters())
    for k, v in sd.items():
        torch.testing.assert_close(m.state_dict()[k], v)
    ids = torch.randint(1, 50, (2, 5))
    torch.testing.assert_close(m(ids).last_hidden_state, src.eval()(ids).last_hidden_state)
    with pytest.raises(FileNotFoundError):
        T.get_t5_model(str(tmp_path / "nothing" / "here"), pretrained=False)
    assert T.get_encoded_dim("t5/t5-11b") == 1024 and len(T.get_t5_model("t5/t5-small", pretrained=False).encoder.block) == 6
