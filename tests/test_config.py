import os

import pytest

from paddlefleetx_b200.utils import config as C

CFG_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "paddlefleetx_b200", "configs")


def _write(tmp_path, name, text):
    p = tmp_path / name
    p.write_text(text)
    return str(p)


def test_attrdict_setdefault_treats_none_as_missing():
    d = C.AttrDict(a=None, b=2)
    assert d.setdefault("a", 5) == 5 and d.a == 5
    assert d.setdefault("b", 7) == 2
    with pytest.raises(AttributeError):
        d.missing


def test_base_inheritance_and_inherited_false(tmp_path):
    _write(tmp_path, "base.yaml", "A: {x: 1, y: {p: 1, q: 2}}\nB: {k: v}\n")
    child = _write(tmp_path, "child.yaml", "_base_: ./base.yaml\nA: {y: {q: 3}}\nB: {_inherited_: False, z: 9}\n")
    cfg = C.parse_config(child)
    assert cfg.A.x == 1 and cfg.A.y.p == 1 and cfg.A.y.q == 3
    assert dict(cfg.B) == {"z": 9}


def test_override_paths_lists_and_new_keys(capsys):
    cfg = C.AttrDict(Model=C.AttrDict(layers=[C.AttrDict(w=1), C.AttrDict(w=2)], name="a"))
    C.override_config(cfg, ["Model.layers.1.w=5", "Model.name=gpt", "Model.new.sub=1.5e-3", "Flag=True", "N=None"])
    assert cfg.Model.layers[1].w == 5 and cfg.Model.name == "gpt" and cfg.Model.new.sub == 1.5e-3
    assert cfg.Flag is True and cfg.N is None
    assert "new" in capsys.readouterr().out.lower()
    with pytest.raises(IndexError):
        C.override_config(cfg, ["Model.layers.7.w=1"])


def test_derived_degrees_and_batches():
    path = os.path.join(CFG_DIR, "nlp/gpt/pretrain_gpt_6.7B_sharding16.yaml")
    cfg = C.get_config(path, ["Global.device=cpu"], nranks=32)
    d = cfg.Distributed
    assert (d.dp_degree, d.mp_degree, d.pp_degree, d.sharding.sharding_degree) == (2, 1, 1, 16)
    assert cfg.Global.global_batch_size == 8 * 2 * 16
    assert cfg.Engine.accumulate_steps == 1
    assert cfg.Engine.test_iters == cfg.Engine.eval_iters * 10
    assert cfg.Model.hidden_size == 4096 and cfg.Model.num_layers == 32 and cfg.Model.vocab_size == 50304


def test_mismatched_world_raises():
    path = os.path.join(CFG_DIR, "nlp/gpt/pretrain_gpt_6.7B_mp2_pp2_sharding2.yaml")
    with pytest.raises(AssertionError):
        C.get_config(path, ["Global.device=cpu"], nranks=6)
    cfg = C.get_config(path, ["Global.device=cpu"], nranks=8)
    assert cfg.Global.enable_partial_send_recv is False            # sequence_parallel + pp > 1
    assert cfg.Engine.accumulate_steps == 8


def test_overlap_flags_forced_off_for_stage3():
    path = os.path.join(CFG_DIR, "nlp/gpt/pretrain_gpt_6.7B_sharding16.yaml")
    cfg = C.get_config(path, ["Global.device=cpu", "Distributed.sharding.sharding_stage=3"], nranks=16)
    assert cfg.Distributed.sharding.reduce_overlap is False and cfg.Distributed.sharding.broadcast_overlap is False


def test_every_yaml_parses():
    n = 0
    for root, _, files in os.walk(CFG_DIR):
        for f in files:
            if f.endswith(".yaml"):
                cfg = C.parse_config(os.path.join(root, f))
                assert isinstance(cfg, dict)
                n += 1
    assert n >= 10
