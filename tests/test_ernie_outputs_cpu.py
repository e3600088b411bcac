"""ERNIE's inspection interface (reference ernie/dygraph/single_model.py:241-375, layers/model_outputs.py): ``return_dict`` outputs, per-layer
hidden states / attention probabilities, ``inputs_embeds``, incremental ``past_key_values`` — all against the default (flash) path."""
import pickle

import pytest
import torch

from paddlefleetx_b200.models.language_model.ernie.model import ErnieForPretraining, ErnieForSequenceClassification, ErnieModel
from paddlefleetx_b200.models.language_model.ernie.model_outputs import (BaseModelOutputWithPoolingAndCrossAttentions, ErnieForPreTrainingOutput,
                                                                             ModelOutput, SequenceClassifierOutput)


def _model(**kw):
    torch.manual_seed(0)
    m = ErnieModel(vocab_size=97, hidden_size=32, num_hidden_layers=3, num_attention_heads=4, max_position_embeddings=64, hidden_dropout_prob=0.0,
                   attention_probs_dropout_prob=0.0, **kw)
    return m.eval()


def _ids(b=2, s=10):
    g = torch.Generator().manual_seed(1)
    ids = torch.randint(1, 97, (b, s), generator=g)
    ids[0, -3:] = 0            # padding
    return ids


def test_model_output_is_dataclass_mapping_and_tuple():
    a, b = torch.ones(2), torch.zeros(3)
    o = SequenceClassifierOutput(logits=a, attentions=[b, b])
    assert list(o.keys()) == ["logits", "attentions"] and o.loss is None
    assert o["logits"] is a and o[0] is a and o.logits is a
    assert isinstance(o.attentions, tuple) and o.to_tuple()[1] == (b, b)
    assert o[1:] == ((b, b),)
    o.loss = b                                     # a field that becomes non-None joins the mapping in declaration order of insertion
    assert "loss" in o and o["loss"] is b
    o["hidden_states"] = (a,)
    assert o.hidden_states == (a,)
    for bad in (lambda: o.pop("logits"), lambda: o.update({}), lambda: o.setdefault("x", 1), lambda: o.__delitem__("logits")):
        with pytest.raises(Exception):
            bad()
    o2 = pickle.loads(pickle.dumps(SequenceClassifierOutput(logits=a)))
    assert torch.equal(o2.logits, a) and list(o2.keys()) == ["logits"]
    from_pairs = SequenceClassifierOutput([("logits", a), ("loss", b)])
    assert from_pairs.logits is a and from_pairs.loss is b


def test_return_dict_matches_default_path_and_exposes_layers():
    m, ids = _model(), _ids()
    seq, pooled = m(ids)
    out = m(ids, output_hidden_states=True, output_attentions=True, return_dict=True)
    assert isinstance(out, BaseModelOutputWithPoolingAndCrossAttentions) and isinstance(out, ModelOutput)
    torch.testing.assert_close(out.last_hidden_state, seq, atol=2e-5, rtol=1e-4)
    torch.testing.assert_close(out.pooler_output, pooled, atol=2e-5, rtol=1e-4)
    assert len(out.hidden_states) == 4 and len(out.attentions) == 3            # embedding output + one per layer
    assert out.attentions[0].shape == (2, 4, 10, 10)
    torch.testing.assert_close(out.attentions[1].sum(-1), torch.ones(2, 4, 10))
    assert float(out.attentions[2][0, :, :, -3:].max().detach()) < 1e-3                   # padded keys get no probability
    torch.testing.assert_close(out.hidden_states[-1], out.last_hidden_state)
    assert out.past_key_values is None and "past_key_values" not in out
    # tuple call: the reference's encoder drops the extras when return_dict is False
    t = m(ids, output_hidden_states=True)
    assert isinstance(t, tuple) and len(t) == 2


def test_inputs_embeds_equals_ids():
    m, ids = _model(), _ids()
    emb = m.get_input_embeddings()(ids)
    seq_a, _ = m(ids)
    mask = (ids != 0).float()
    seq_b, _ = m(inputs_embeds=emb, attention_mask=mask)
    torch.testing.assert_close(seq_a, seq_b, atol=2e-5, rtol=1e-4)
    with pytest.raises(ValueError):
        m(ids, inputs_embeds=emb)
    with pytest.raises(ValueError):
        m()


def test_incremental_cache_reproduces_full_sequence_keys():
    """Keys / values cached from a prefix plus the new tokens' keys / values equal the keys / values of the whole sequence at layer 0 (deeper
    layers differ by construction: the encoder is bidirectional, so a prefix processed alone has not seen the later tokens)."""
    m = _model()
    ids = _ids(2, 12)
    ids[ids == 0] = 5
    full = m(ids, use_cache=True, return_dict=True)
    assert len(full.past_key_values) == 3 and full.past_key_values[0][0].shape == (2, 4, 12, 8)
    first = m(ids[:, :8], use_cache=True, return_dict=True)
    second = m(ids[:, 8:], past_key_values=first.past_key_values, use_cache=True, return_dict=True)
    assert second.past_key_values[0][0].shape == (2, 4, 12, 8)
    torch.testing.assert_close(second.past_key_values[0][0], full.past_key_values[0][0], atol=2e-5, rtol=1e-4)
    torch.testing.assert_close(second.past_key_values[0][1], full.past_key_values[0][1], atol=2e-5, rtol=1e-4)
    assert second.last_hidden_state.shape == (2, 4, 32)


def test_pretraining_and_classifier_outputs():
    m, ids = _model(), _ids()
    pre = ErnieForPretraining(m).eval()
    scores, rel = pre(ids)
    labels = torch.randint(0, 97, ids.shape)
    nsl = torch.tensor([0, 1])
    out = pre(ids, labels=labels, next_sentence_label=nsl, return_dict=True, output_attentions=True)
    assert isinstance(out, ErnieForPreTrainingOutput)
    torch.testing.assert_close(out.prediction_logits, scores, atol=3e-5, rtol=1e-4)
    want = torch.nn.functional.cross_entropy(scores.reshape(-1, 97), labels.reshape(-1)) + torch.nn.functional.cross_entropy(rel, nsl)
    torch.testing.assert_close(out.loss, want, atol=1e-4, rtol=1e-4)
    tup = pre(ids, labels=labels, next_sentence_label=nsl)
    assert len(tup) == 3 and torch.allclose(tup[0], out.loss)

    cls = ErnieForSequenceClassification(m, num_classes=3, dropout=0.0).eval()
    logits = cls(ids)
    assert logits.shape == (2, 3)
    y = torch.tensor([2, 0])
    loss, logits2 = cls(ids, labels=y)
    torch.testing.assert_close(loss, torch.nn.functional.cross_entropy(logits, y))
    d = cls(ids, labels=torch.rand(2, 3), return_dict=True, output_hidden_states=True)       # float targets: multi-label BCE
    assert isinstance(d, SequenceClassifierOutput) and len(d.hidden_states) == 4 and d.loss.ndim == 0
    reg = ErnieForSequenceClassification(m, num_classes=1, dropout=0.0).eval()
    assert reg(ids, labels=torch.rand(2, 1))[0].ndim == 0


def test_reference_module_paths_resolve():
    import importlib

    mo = importlib.import_module("ppfleetx.models.language_model.ernie.layers.model_outputs")
    assert mo.BaseModelOutputWithPastAndCrossAttentions is not None and mo.ModelOutput is ModelOutput
    ut = importlib.import_module("ppfleetx.models.language_model.ernie.layers.utils")

    class Tracked(torch.nn.Module, metaclass=ut.InitTrackerMeta):
        def __init__(self, a, b=2, *, c=3):
            super().__init__()

    t = Tracked(1, b=5)
    assert t.init_config == {"b": 5, "init_args": (1,), "init_class": "Tracked"}
    assert ut.fn_args_to_dict(lambda a, b=2, c=3: 0, 1, c=9) == {"a": 1, "b": 2, "c": 9}
    t5u = importlib.import_module("ppfleetx.models.language_model.t5.utils")
    lin = torch.nn.Linear(4, 4)
    t5u.constant_init(lin, 0.5, bias=1.0)
    assert float(lin.weight.mean()) == 0.5 and float(lin.bias.mean()) == 1.0
    for name in ("ernie_dataset", "dataset_utils"):
        importlib.import_module(f"ppfleetx.data.data_tools.ernie.preprocess.{name}")
