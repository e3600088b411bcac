"""The engine-less ``examples/transformer`` scripts run end to end on CPU with a tiny model (subprocess, like a user would)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TINY = ["Global.device=cpu", "Model.num_layers=2", "Model.hidden_size=64", "Model.num_attention_heads=4", "Model.ffn_hidden_size=128",
        "Model.vocab_size=512", "Data.Train.dataset.max_seq_len=32", "Data.Eval.dataset.max_seq_len=32", "Model.max_position_embeddings=32",
        "Engine.mix_precision.enable=False", "Engine.max_steps=4", "Engine.logging_freq=2", "Data.Train.dataset.name=SyntheticGPTDataset",
        "Data.Train.loader.num_workers=0", "Data.Eval.dataset.name=SyntheticGPTDataset", "Data.Eval.loader.num_workers=0", "Global.local_batch_size=2", "Global.micro_batch_size=2", "Engine.eval_freq=-1"]


def _run(script, cfg, extra=()):
    cmd = [sys.executable, os.path.join(ROOT, script), "-c", os.path.join(ROOT, cfg)]
    for o in TINY + list(extra):
        cmd += ["-o", o]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-3000:]
    return p.stdout + p.stderr


@pytest.mark.parametrize("script,cfg", [
    ("examples/transformer/models/GPT/pretrain/run.py", "examples/transformer/models/GPT/pretrain/configs/pretrain_gpt_345M_single_card.yaml"),
    ("examples/transformer/models/GPT/pretrain_moe/run.py", "examples/transformer/models/GPT/pretrain_moe/configs/pretrain_moe_345M_single_card.yaml"),
])
def test_explicit_loop_examples_train(script, cfg):
    out = _run(script, cfg)
    assert "[train] step: 4/4" in out


def _losses(out):
    import re

    return {int(m.group(1)): m.group(2) for m in re.finditer(r"\[train\] step: (\d+)/\d+, loss: ([0-9.]+)", out)}


@pytest.mark.parametrize("script,cfg", [
    ("examples/transformer/models/GPT/pretrain/run.py", "examples/transformer/models/GPT/pretrain/configs/pretrain_gpt_345M_single_card.yaml"),
    ("examples/transformer/models/GPT/pretrain_moe/run.py", "examples/transformer/models/GPT/pretrain_moe/configs/pretrain_moe_345M_single_card.yaml"),
])
def test_explicit_loop_resume_continues_the_uninterrupted_run(script, cfg, tmp_path):
    """Dropout is on (recipe default): a run resumed from the step-3 checkpoint prints the same losses as the run that never stopped."""
    common = ["Engine.max_steps=6", "Engine.logging_freq=1", "Optimizer.lr.max_lr=1e-2", "Optimizer.lr.warmup_rate=0.0"]
    straight = _losses(_run(script, cfg, common + ["Engine.save_load.save_steps=3", f"Engine.save_load.output_dir={tmp_path}/a"]))
    resumed = _losses(_run(script, cfg, common + ["Engine.save_load.save_steps=-1", f"Engine.save_load.output_dir={tmp_path}/b",
                                                  f"Engine.save_load.ckpt_dir={tmp_path}/a/epoch_0_step_3"]))
    assert sorted(resumed) == [4, 5, 6] and sorted(straight) == [1, 2, 3, 4, 5, 6]
    assert all(resumed[k] == straight[k] for k in resumed), (straight, resumed)


def test_every_project_script_points_at_an_existing_config_and_tool():
    import glob
    import re

    scripts = glob.glob(os.path.join(ROOT, "projects", "**", "*.sh"), recursive=True)
    assert len(scripts) >= 50
    for s in scripts:
        body = open(s).read()
        for cfg in re.findall(r"-c (\S+\.yaml)", body):
            assert os.path.isfile(os.path.join(ROOT, cfg)), (s, cfg)
        for tool in re.findall(r"((?:tools|tasks)/\S+\.py)", body):
            assert os.path.isfile(os.path.join(ROOT, tool)), (s, tool)


def test_offline_eval_example_reports_perplexity_and_cloze_accuracy(tmp_path):
    import json

    wiki = tmp_path / "wiki.valid.tokens"
    wiki.write_text(" = Title = \n\n The quick brown fox jumps over the lazy dog . It was a bright cold day in April , and the clocks were striking thirteen . \n" * 6)
    lamb = tmp_path / "lambada.jsonl"
    lamb.write_text("".join(json.dumps({"text": t}) + "\n" for t in ["the quick brown fox jumps over the lazy dog", "it was a bright cold day in april"] * 3))
    script = "examples/transformer/models/GPT/offline-eval/run.py"
    cfg = "examples/transformer/models/GPT/offline-eval/configs/eval_gpt_345M_single_card.yaml"
    base = ["Global.device=cpu", "Model.num_layers=2", "Model.hidden_size=64", "Model.num_attention_heads=4", "Model.ffn_hidden_size=128", "Model.vocab_size=512",
            "Model.max_position_embeddings=32", "Engine.mix_precision.enable=False", "Offline_Eval.max_seq_len=32", "Offline_Eval.batch_size=2",
            "Offline_Eval.overlapping_eval=8", "Offline_Eval.logging_freq=1"]

    def run(extra):
        cmd = [sys.executable, os.path.join(ROOT, script), "-c", os.path.join(ROOT, cfg)]
        for o in base + extra:
            cmd += ["-o", o]
        p = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=ROOT)
        assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-3000:]
        return p.stdout + p.stderr

    out = run([f"Offline_Eval.eval_path={wiki}", "Offline_Eval.cloze_eval=False"])
    assert "ppl:" in out and "adjusted ppl:" in out and "token ratio:" in out
    out = run([f"Offline_Eval.eval_path={lamb}", "Offline_Eval.cloze_eval=True"])
    assert "number correct:" in out and "avg accuracy:" in out


def test_examples_qat_helper_wraps_and_converts():
    sys.path.insert(0, os.path.join(ROOT, "examples", "transformer"))
    import torch

    from paddlefleetx_b200.utils.config import AttrDict
    from utils import qat

    model = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.GELU(), torch.nn.Linear(16, 4))
    same, quanter = qat.compress_model(AttrDict(), model)
    assert same is model and quanter is None
    cfg = AttrDict(Compress=AttrDict(Quantization=AttrDict(enable=True, weight_quantize_type="channel_wise_abs_max", activation_quantize_type="moving_average_abs_max")))
    qmodel, quanter = qat.compress_model(cfg, model)
    assert callable(quanter) and any("Quant" in type(m).__name__ for m in qmodel.modules())
    y = qmodel(torch.randn(3, 8))
    y.sum().backward()
    assert y.shape == (3, 4)
