"""Numerical parity of the text towers with the upstream architectures they re-implement: weights are copied BY NAME from a randomly initialised
``transformers`` model (the state-dict keys of our modules are the upstream ones) and the outputs must agree.  ``transformers`` is library code
used here only as an independent oracle; the test is skipped without it."""
import pytest
import torch

transformers = pytest.importorskip("transformers")


def _ids(vocab, b=2, s=11, pad_from=8):
    g = torch.Generator().manual_seed(3)
    ids = torch.randint(1, vocab, (b, s), generator=g)
    mask = torch.ones(b, s, dtype=torch.long)
    mask[0, pad_from:] = 0
    return ids, mask


@pytest.mark.parametrize("ff", ["relu", "gated-gelu"])
def test_t5_encoder_matches_transformers(ff):
    from paddlefleetx_b200.models.multimodal_model.t5 import modeling as T

    shape = dict(vocab_size=120, d_model=48, d_kv=8, d_ff=96, num_layers=3, num_heads=6, relative_attention_num_buckets=16, dropout_rate=0.0,
                 layer_norm_epsilon=1e-6, feed_forward_proj=ff)
    torch.manual_seed(0)
    hf = transformers.T5EncoderModel(transformers.T5Config(**shape, is_encoder_decoder=False, use_cache=False, relative_attention_max_distance=128)).eval()
    mine = T.T5EncoderModel(**shape).eval()
    missing = mine.load_state_dict(hf.state_dict(), strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    ids, mask = _ids(120)
    with torch.no_grad():
        want = hf(input_ids=ids, attention_mask=mask, output_hidden_states=True, output_attentions=True)
        got = mine(input_ids=ids, attention_mask=mask, output_hidden_states=True, output_attentions=True)
        fast = mine(ids, mask).last_hidden_state
    torch.testing.assert_close(got.last_hidden_state, want.last_hidden_state, atol=2e-5, rtol=1e-4)
    torch.testing.assert_close(fast, want.last_hidden_state, atol=2e-5, rtol=1e-4)
    for a, b in zip(got.hidden_states, want.hidden_states):
        torch.testing.assert_close(a, b, atol=2e-5, rtol=1e-4)
    for a, b in zip(got.attentions, want.attentions):
        torch.testing.assert_close(a, b, atol=2e-5, rtol=1e-4)


def test_t5_decoder_stack_matches_transformers():
    from paddlefleetx_b200.models.multimodal_model.t5 import modeling as T

    shape = dict(vocab_size=90, d_model=32, d_kv=8, d_ff=64, num_layers=2, num_heads=4, relative_attention_num_buckets=8, dropout_rate=0.0,
                 layer_norm_epsilon=1e-6, feed_forward_proj="relu")
    torch.manual_seed(1)
    hf = transformers.T5ForConditionalGeneration(transformers.T5Config(**shape, num_decoder_layers=2, decoder_start_token_id=0)).eval()
    emb = torch.nn.Embedding(90, 32)
    dec = T.T5Stack(32, 2, 1e-6, 0.0, 8, "relu", 8, 4, 64, embed_tokens=emb, is_decoder=True).eval()
    sd = {k[len("decoder."):]: v for k, v in hf.state_dict().items() if k.startswith("decoder.") and not k.startswith("decoder.embed_tokens")}
    dec.load_state_dict(sd, strict=True)
    emb.load_state_dict({"weight": hf.state_dict()["shared.weight"]})
    enc_states = torch.randn(2, 5, 32)
    enc_mask = torch.ones(2, 5, dtype=torch.long)
    enc_mask[1, 3:] = 0
    ids, _ = _ids(90, s=6)
    with torch.no_grad():
        want = hf.decoder(input_ids=ids, encoder_hidden_states=enc_states, encoder_attention_mask=enc_mask).last_hidden_state
        got = dec(input_ids=ids, encoder_hidden_states=enc_states, encoder_attention_mask=enc_mask).last_hidden_state
    torch.testing.assert_close(got, want, atol=3e-5, rtol=1e-4)


@pytest.mark.parametrize("conv,share", [(3, True), (0, False)])
def test_debertav2_matches_transformers(conv, share):
    from paddlefleetx_b200.models.multimodal_model.debertav2 import modeling as D

    shape = dict(vocab_size=150, hidden_size=48, num_hidden_layers=3, num_attention_heads=6, intermediate_size=96, max_position_embeddings=64,
                 position_buckets=8, relative_attention=True, norm_rel_ebd="layer_norm", pos_att_type=["p2c", "c2p"], share_att_key=share,
                 conv_kernel_size=conv, conv_act="gelu", hidden_act="gelu", hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0,
                 position_biased_input=False, type_vocab_size=0, layer_norm_eps=1e-7, max_relative_positions=-1, pad_token_id=0)
    torch.manual_seed(0)
    hf = transformers.DebertaV2Model(transformers.DebertaV2Config(**shape)).eval()
    mine = D.DebertaV2Model(**shape).eval()
    sd = {k: v for k, v in hf.state_dict().items() if not k.endswith("position_ids")}
    res = mine.load_state_dict(sd, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    ids, mask = _ids(150)
    with torch.no_grad():
        want = hf(input_ids=ids, attention_mask=mask, output_hidden_states=True, output_attentions=True)
        got = mine(input_ids=ids, attention_mask=mask, output_hidden_states=True, output_attentions=True)
    keep = mask.bool()
    torch.testing.assert_close(got.last_hidden_state[keep], want.last_hidden_state[keep], atol=5e-5, rtol=1e-4)
    for a, b in zip(got.hidden_states, want.hidden_states):
        torch.testing.assert_close(a[keep], b[keep], atol=5e-5, rtol=1e-4)
    for a, b in zip(got.attentions, want.attentions):
        torch.testing.assert_close(a[0, :, :8, :8], b[0, :, :8, :8], atol=5e-5, rtol=1e-4)
        torch.testing.assert_close(a[1], b[1], atol=5e-5, rtol=1e-4)
