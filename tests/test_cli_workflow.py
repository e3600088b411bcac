"""The user-facing command line, end to end, as subprocesses: train (with checkpoint) -> resume -> generate -> export -> serve ->
latency tool.  Runs on CPU in the default suite and on the B200 with ``-m gpu`` (bf16, native kernels, CUDA-graph decode)."""
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CFG = os.path.join(ROOT, "paddlefleetx_b200", "configs", "nlp", "gpt")
MODEL = ["Model.num_layers=2", "Model.hidden_size=128", "Model.num_attention_heads=2", "Model.ffn_hidden_size=512", "Model.vocab_size=50304",
         "Model.max_position_embeddings=128", "Model.hidden_dropout_prob=0.0", "Model.attention_probs_dropout_prob=0.0"]


def _run(args, device, timeout=600):
    env = dict(os.environ, PYTHONPATH=ROOT)
    if device == "cpu":
        env["CUDA_VISIBLE_DEVICES"] = ""
    p = subprocess.run([sys.executable] + args, capture_output=True, text=True, timeout=timeout, cwd=ROOT, env=env)
    assert p.returncode == 0, (args, p.stdout[-1500:], p.stderr[-3000:])
    return p.stdout + p.stderr


def _workflow(tmp_path, device):
    out = str(tmp_path / "out")
    dev = ["Global.device=" + device] + (["Engine.mix_precision.enable=False"] if device == "cpu" else [])
    train = MODEL + dev + ["Data.Train.dataset.name=SyntheticGPTDataset", "Data.Eval.dataset.name=SyntheticGPTDataset", "Data.Train.dataset.max_seq_len=64",
                           "Data.Eval.dataset.max_seq_len=64", "Data.Train.loader.num_workers=0", "Data.Eval.loader.num_workers=0",
                           "Global.local_batch_size=4", "Global.micro_batch_size=2", "Engine.logging_freq=1", "Engine.eval_freq=3", "Engine.eval_iters=2",
                           f"Engine.save_load.output_dir={out}"]

    def o(lst):
        return [x for item in lst for x in ("-o", item)]

    log = _run(["tools/train.py", "-c", os.path.join(CFG, "pretrain_gpt_345M_single_card.yaml")] + o(train + ["Engine.max_steps=4", "Engine.save_load.save_steps=4"]), device)
    assert "ips_total:" in log and "[eval]" in log
    ckpt = os.path.join(out, "epoch_0_step_4")
    assert sorted(os.listdir(ckpt)) == ["meta_state.pdopt", "model.pdparams", "model_state.pdopt"]
    log = _run(["tools/train.py", "-c", os.path.join(CFG, "pretrain_gpt_345M_single_card.yaml")] + o(train + ["Engine.max_steps=6", "Engine.save_load.save_steps=-1",
                                                                                                            f"Engine.save_load.ckpt_dir={ckpt}"]), device)
    steps = [int(m) for m in re.findall(r"\[train\].*?batch: \[(\d+)/", log)]
    assert steps and min(steps) >= 4, (steps, log[-1500:])                     # resumed, did not start from scratch
    gen = MODEL + dev + ["Generation.max_dec_len=8", f"Engine.save_load.ckpt_dir={ckpt}", f"Engine.save_load.output_dir={out}"]
    log = _run(["tasks/gpt/generation.py", "-c", os.path.join(CFG, "generation_gpt_345M_single_card.yaml")] + o(gen), device)
    assert "Generation" in log or "generation" in log
    _run(["tools/export.py", "-c", os.path.join(CFG, "generation_gpt_345M_single_card.yaml")] + o(gen), device)
    assert sorted(os.listdir(os.path.join(out, "rank_0"))) == ["model.pdiparams", "model.pdmodel"]
    log = _run(["tasks/gpt/inference.py", "--model_dir", out, "--text", "hello b200"], device)
    assert "Generation:" in log
    conv = str(tmp_path / "conv")
    _run(["tools/reshard.py", "--src", ckpt, "--dst", conv, "--mp", "1"], device)
    assert os.path.isfile(os.path.join(conv, "model.pdparams"))


def test_cli_workflow_cpu(tmp_path):
    _workflow(tmp_path, "cpu")


@pytest.mark.gpu
def test_cli_workflow_gpu(tmp_path):
    _workflow(tmp_path, "gpu")
