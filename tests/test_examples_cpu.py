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
