"""``tools/auto.py`` / ``tools/auto_export.py`` and ``AutoEngine.tune`` on a tiny CPU model."""
import os
import sys

import pytest

from helpers import ROOT, TINY

sys.path.insert(0, os.path.join(ROOT, "tools"))
AUTO_CFG = os.path.join(ROOT, "paddlefleetx_b200", "configs", "nlp", "gpt", "auto", "pretrain_gpt_345M_single_card.yaml")
SYN = ["Data.Train.dataset.name=SyntheticGPTDataset", "Data.Train.dataset.vocab_size=512", "Data.Train.dataset.input_dir=None", "Data.Train.dataset.split=None",
       "Data.Train.loader.num_workers=0", "Global.local_batch_size=2", "Global.micro_batch_size=2", "Engine.eval_freq=-1", "Engine.save_load.save_steps=-1"]


def _argv(tmp_path, extra=()):
    out = ["-c", AUTO_CFG]
    for o in TINY + SYN + [f"Engine.save_load.output_dir={tmp_path}"] + list(extra):
        out += ["-o", o]
    return out


def test_auto_entry_trains_and_validates_the_mesh(tmp_path):
    import auto

    engine = auto.main(_argv(tmp_path, ["Engine.max_steps=3"]))
    assert type(engine).__name__ == "AutoEngine" and type(engine._module).__name__ == "GPTModuleAuto"
    mesh = engine._configs.Distributed.mesh
    assert list(mesh.shape) == [1, 1, 1]
    engine._configs.Distributed.mesh.shape = [1, 2, 1]
    with pytest.raises(AssertionError, match="mesh"):
        type(engine)(engine._configs, engine._module)


def test_auto_tuning_measures_recompute_candidates(tmp_path, capsys):
    import auto

    engine = auto.main(_argv(tmp_path, ["Tuning.enable=True", "Tuning.tuning_recompute=True", "Tuning.profile_start_step=1", "Tuning.profile_end_step=2",
                                        "Model.use_recompute=True", "Model.recompute_granularity=full"]))
    printed = capsys.readouterr().out
    assert printed.count("'step_s'") == 4                                        # one row per candidate, printed by the tool
    rows = engine.tune(__import__("paddlefleetx_b200.data", fromlist=["build_dataloader"]).build_dataloader(engine._configs.Data, "Train"))
    assert len(rows) == 4 and all(r["status"] == "ok" and r["fits"] and r["step_s"] > 0 for r in rows)
    assert rows == sorted(rows, key=lambda r: r["step_s"])
    kinds = {(r["use_recompute"], r.get("recompute_granularity")) for r in rows}
    assert kinds == {(False, None), (True, "core_attn"), (True, "full_attn"), (True, "full")}
    losses = {round(r["final_loss"], 4) for r in rows}
    assert len(losses) == 1, losses                                              # recompute never changes the numbers, only time / memory
    best = rows[0]
    assert engine._configs.Model.use_recompute == best["use_recompute"]
    if best["use_recompute"]:
        assert engine._configs.Model.recompute_granularity == best["recompute_granularity"]
    # a memory limit nothing satisfies leaves the configuration alone
    before = dict(use=engine._configs.Model.use_recompute)
    engine._configs.Tuning.memory_limit_gb = -1.0
    rows = engine.tune(__import__("paddlefleetx_b200.data", fromlist=["build_dataloader"]).build_dataloader(engine._configs.Data, "Train"))
    assert not any(r["fits"] for r in rows) and engine._configs.Model.use_recompute == before["use"]


def test_auto_tuning_without_recompute_ranks_layouts(tmp_path):
    import auto

    engine = auto.main(_argv(tmp_path, ["Tuning.enable=True", "Tuning.tuning_recompute=False"]))
    rows = engine.tune()
    assert rows and {"dp", "mp", "pp", "sharding", "est_step_s", "est_mem_gb"} <= set(rows[0])


def test_auto_export_writes_an_inference_bundle(tmp_path):
    import auto_export

    gen = os.path.join(ROOT, "paddlefleetx_b200", "configs", "nlp", "gpt", "auto", "generation_gpt_345M_single_card.yaml")
    argv = ["-c", gen]
    for o in TINY + [f"Engine.save_load.output_dir={tmp_path}", "Generation.max_dec_len=4"]:
        argv += ["-o", o]
    engine = auto_export.main(argv)
    assert type(engine).__name__ == "AutoEngine"
    files = [f for _, _, fs in os.walk(tmp_path) for f in fs]
    assert {"model.pdmodel", "model.pdiparams"} <= set(files), files


def test_auto_layout_plans_then_trains(tmp_path, capsys):
    """``Distributed.auto_layout=True``: the planner derives the layout (trivial on one process), records it, and the run trains with it."""
    import auto

    engine = auto.main(_argv(tmp_path, ["Distributed.auto_layout=True", "Engine.max_steps=2"]))
    d = engine._configs.Distributed
    assert d.plan is not None and d.plan.est_mem_gb > 0 and "[auto_layout]" in capsys.readouterr().out
    assert (d.dp_degree, d.mp_degree, d.pp_degree, d.sharding.sharding_degree) == (1, 1, 1, 1) and d.auto_layout is False


def test_auto_tuning_measures_micro_batch_sizes(tmp_path):
    """``Tuning.tuning_micro_batch``: every micro-batch size that divides the local batch is built and timed; gradient accumulation does not change
    the numbers, the fastest candidate is written back together with the matching accumulate_steps."""
    import auto

    extra = ["Tuning.enable=True", "Tuning.tuning_micro_batch=True", "Tuning.profile_start_step=1", "Tuning.profile_end_step=2",
             "Global.local_batch_size=4", "Global.micro_batch_size=4"]
    engine = auto.main(_argv(tmp_path, [o for o in extra]))
    from paddlefleetx_b200.data import build_dataloader

    rows = engine.tune(build_dataloader(engine._configs.Data, "Train"))
    assert sorted(r["micro_batch_size"] for r in rows) == [1, 2, 4] and all(r["status"] == "ok" for r in rows)
    losses = [r["final_loss"] for r in rows]
    assert max(losses) - min(losses) < 1e-4, losses
    best = rows[0]
    assert engine._configs.Global.micro_batch_size == best["micro_batch_size"]
    assert engine._configs.Engine.accumulate_steps == 4 // best["micro_batch_size"]


def test_auto_module_reports_shard_spec_and_rejects_a_wrong_mesh(tmp_path):
    import auto

    engine = auto.main(_argv(tmp_path, ["Engine.max_steps=1"]))
    spec = engine._module.shard_spec()
    assert spec and all(isinstance(v, list) for v in spec.values())
    assert all(set(v) <= {None} for v in spec.values())                  # one process: nothing is split
    name, p = next((n, p) for n, p in engine._module.model.named_parameters() if p.dim() == 2)
    p.tp_sharded, p.split_axis = True, 1                                 # what a column-parallel weight looks like on an mp > 1 mesh
    assert engine._module.shard_spec()[name] == [None, "mp"]
    del p.tp_sharded, p.split_axis
    cfg = engine._configs
    cfg.Distributed.mesh.shape = [2, 1, 1]
    with pytest.raises(ValueError, match="mesh"):
        type(engine._module)(cfg)
