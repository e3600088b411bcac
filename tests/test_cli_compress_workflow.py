"""User-level chains through the command-line tools on CPU with tiny models (subprocesses, as a user would run them):
QAT / pruning: train -> offline eval -> export -> serve;  ViT: train -> export -> classify;  ERNIE: export -> serve with the WordPiece vocabulary."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CFG = "paddlefleetx_b200/configs"
CPU = ["Global.device=cpu", "Engine.mix_precision.enable=False"]
TINY_GPT = CPU + ["Model.num_layers=2", "Model.hidden_size=64", "Model.num_attention_heads=4", "Model.ffn_hidden_size=128", "Model.vocab_size=512",
                  "Model.max_position_embeddings=32"]
TRAIN = ["Data.Train.dataset.name=SyntheticGPTDataset", "Data.Train.dataset.max_seq_len=32", "Data.Train.loader.num_workers=0", "Global.local_batch_size=2",
         "Global.micro_batch_size=2", "Engine.max_steps=3", "Engine.eval_freq=-1", "Engine.logging_freq=1", "Engine.save_load.save_steps=3"]


def run(script, cfg=None, opts=(), args=(), timeout=400):
    cmd = [sys.executable, os.path.join(ROOT, script)]
    if cfg:
        cmd += ["-c", os.path.join(ROOT, CFG, cfg)]
    for o in opts:
        cmd += ["-o", o]
    p = subprocess.run(cmd + list(args), capture_output=True, text=True, timeout=timeout, cwd=ROOT)
    assert p.returncode == 0, f"{script} {cfg}\n" + p.stdout[-1500:] + p.stderr[-2500:]
    return p.stdout + p.stderr


@pytest.fixture(scope="module")
def wiki(tmp_path_factory):
    p = tmp_path_factory.mktemp("eval") / "wiki.valid.tokens"
    p.write_text(" = T = \n\n The quick brown fox jumps over the lazy dog . It was a bright cold day in April , and the clocks were striking thirteen . \n" * 6)
    return str(p)


@pytest.mark.parametrize("kind,eval_cfg,gen_cfg", [("qat", "eval_qat_gpt_345M_single_card.yaml", "generation_qat_gpt_345M_single_card.yaml"),
                                                     ("prune", "eval_pruned_gpt_345M_single_card.yaml", "generation_pruned_gpt_345M_single_card.yaml")])
def test_compression_recipe_chain_train_eval_export_serve(tmp_path, wiki, kind, eval_cfg, gen_cfg):
    out = run("tools/train.py", f"nlp/gpt/{kind}_gpt_345M_single_card.yaml", TINY_GPT + TRAIN + [f"Engine.save_load.output_dir={tmp_path}/train"])
    ckpt = os.path.join(tmp_path, "train", "epoch_0_step_3")
    assert os.path.isfile(os.path.join(ckpt, "model.pdparams")), out[-800:]
    ev = run("tools/eval.py", f"nlp/gpt/{eval_cfg}", TINY_GPT + [f"Engine.save_load.ckpt_dir={ckpt}", f"Offline_Eval.eval_path={wiki}", "Offline_Eval.cloze_eval=False",
                                                             "Offline_Eval.max_seq_len=32", "Offline_Eval.batch_size=2", "Offline_Eval.overlapping_eval=8"])
    assert "validation results" in ev and "ppl:" in ev
    ex = run("tools/export.py", f"nlp/gpt/{gen_cfg}", TINY_GPT + [f"Engine.save_load.ckpt_dir={ckpt}", f"Engine.save_load.output_dir={tmp_path}/export", "Generation.max_dec_len=4"])
    assert "exported inference model" in ex
    served = run("projects/gpt/inference.py", args=["--model_dir", f"{tmp_path}/export", "--text", "hello there"])
    assert "Generation:" in served
    if kind == "prune":
        assert "pruned 2 decoder layers" in served          # the serving engine rebuilt the pruned shapes from the recipe


def test_vit_chain_train_export_classify(tmp_path):
    vit = CPU + ["Model.model.img_size=32", "Model.model.patch_size=8", "Model.model.depth=2", "Distributed.dp_degree=1"]
    data = ["Data.Train.dataset.name=SyntheticImageDataset", "Data.Train.dataset.image_size=32", "Data.Train.dataset.num_samples=16",
            "Data.Eval.dataset.name=SyntheticImageDataset", "Data.Eval.dataset.image_size=32", "Data.Eval.dataset.num_samples=8", "Data.Eval.dataset.class_num=10",
            "Global.local_batch_size=4", "Global.micro_batch_size=4", "Data.Train.sampler.batch_size=4", "Data.Eval.sampler.batch_size=4",
            "Data.Train.loader.num_workers=0", "Data.Eval.loader.num_workers=0", "Engine.num_train_epochs=1", "Engine.logging_freq=1"]
    cfg = "vis/vit/ViT_tiny_patch16_224_ci_cifar10_1n8c_dp_fp16o2.yaml"
    out = run("tools/train.py", cfg, vit + data + [f"Engine.save_load.output_dir={tmp_path}/train"])
    assert "images/sec" in out and os.path.isdir(os.path.join(tmp_path, "train", "epoch_0_step_4"))
    run("tools/export.py", cfg, vit + [f"Engine.save_load.ckpt_dir={tmp_path}/train/epoch_0_step_4", f"Engine.save_load.output_dir={tmp_path}/export"])
    served = run("projects/vit/inference.py", args=["--model_dir", f"{tmp_path}/export", "--random", "--size", "32"])
    assert "top-5 classes:" in served


def test_vit_resume_from_end_of_epoch_checkpoint_starts_the_next_epoch(tmp_path):
    import re

    vit = CPU + ["Model.model.img_size=32", "Model.model.patch_size=8", "Model.model.depth=2", "Distributed.dp_degree=1",
                 "Data.Train.dataset.name=SyntheticImageDataset", "Data.Train.dataset.image_size=32", "Data.Train.dataset.num_samples=16",
                 "Data.Eval.dataset.name=SyntheticImageDataset", "Data.Eval.dataset.image_size=32", "Data.Eval.dataset.num_samples=8", "Data.Eval.dataset.class_num=10",
                 "Global.local_batch_size=4", "Global.micro_batch_size=4", "Data.Train.sampler.batch_size=4", "Data.Eval.sampler.batch_size=4",
                 "Data.Train.loader.num_workers=0", "Data.Eval.loader.num_workers=0", "Engine.num_train_epochs=3", "Engine.logging_freq=1", "Engine.save_load.save_epoch=1"]
    cfg = "vis/vit/ViT_tiny_patch16_224_ci_cifar10_1n8c_dp_fp16o2.yaml"

    def losses(out):
        return re.findall(r"\[train\] epoch: (\d+), step: \[(\d+)/4\].*?loss: ([0-9.]+)", out)

    straight = losses(run("tools/train.py", cfg, vit + [f"Engine.save_load.output_dir={tmp_path}/a"]))
    out = run("tools/train.py", cfg, vit + [f"Engine.save_load.output_dir={tmp_path}/b", f"Engine.save_load.ckpt_dir={tmp_path}/a/epoch_0_step_4"])
    assert losses(out) == straight[4:] and len(straight) == 12
    assert "[eval] epoch: 0" not in out and sorted(os.listdir(tmp_path / "b")) == ["epoch_1_step_4", "epoch_2_step_4"]


def test_moco_chain_pretrain_on_image_folder_then_linear_probe(tmp_path):
    """PNG class folders -> MoCo v2 pre-training (two augmented views through the recipe's real transform stack, cosine LR advanced once per EPOCH)
    -> linear probe on the frozen pre-trained backbone with a per-epoch top-1 report."""
    import re

    import numpy as np
    from PIL import Image

    rng = np.random.default_rng(0)
    for split, n in (("train", 8), ("val", 4)):
        for cls, col in (("cat", (200, 60, 60)), ("dog", (60, 60, 200))):
            os.makedirs(tmp_path / "data" / split / cls)
            for i in range(n):
                px = np.clip(np.array(col)[None, None] + rng.normal(0, 30, (48, 56, 3)), 0, 255).astype(np.uint8)
                Image.fromarray(px).save(tmp_path / "data" / split / cls / f"{i}.png")
    base = CPU + ["Distributed.dp_degree=1", "Global.local_batch_size=4", "Global.micro_batch_size=4", "Data.Train.sampler.batch_size=4", "Data.Train.loader.num_workers=0",
                  "Model.model.backbone=resnet18", "Engine.logging_freq=4", f"Data.Train.dataset.root={tmp_path}/data/train", "Data.Train.dataset.transform_ops.1.RandCropImage.size=32"]
    out = run("tools/train.py", "vis/moco/mocov2_pt_in1k_1n8c.yaml", base + ["Model.model.K=16", "Model.model.dim=8", "Engine.num_train_epochs=3", "Engine.save_load.save_epoch=3",
                                                                        f"Engine.save_load.output_dir={tmp_path}/pt"])
    assert re.findall(r"learning rate: ([0-9.]+)", out) == ["0.0300000", "0.0225000", "0.0075000"]
    out = run("tools/train.py", "vis/moco/moco_lincls_in1k_1n8c.yaml", base + [
        "Data.Eval.sampler.batch_size=4", "Data.Eval.loader.num_workers=0", f"Data.Eval.dataset.root={tmp_path}/data/val", "Data.Eval.dataset.transform_ops.1.ResizeImage.resize_short=36",
        "Data.Eval.dataset.transform_ops.2.CenterCropImage.size=32", "Model.model.class_num=2", f"Model.model.pretrained={tmp_path}/pt/epoch_2_step_4/model.pdparams",
        "Optimizer.lr.learning_rate=0.001", "Engine.num_train_epochs=2", f"Engine.save_load.output_dir={tmp_path}/lc"])
    assert len(re.findall(r"\[Eval\] epoch: \d, \{'top1'", out)) == 2 and set(re.findall(r"learning rate: ([0-9.]+)", out)) == {"0.0010000"}


def test_ernie_finetune_cli_learns_a_tsv_task_and_reports_dev_accuracy(tmp_path):
    """WordPiece vocabulary directory + ``text\\tlabel`` TSVs -> tools/train.py on the ERNIE fine-tune recipe: the schedule spans the epochs actually
    run (not the pre-training base recipe's max_steps), the loss falls and every epoch ends with a dev-set accuracy line."""
    import random
    import re

    rnd = random.Random(0)
    words = ["good", "bad", "great", "awful", "movie", "film", "plot", "acting", "the", "was", "is", "very", "not", "i", "liked", "hated", "it"]
    (tmp_path / "vocab").mkdir()
    (tmp_path / "vocab" / "vocab.txt").write_text("\n".join(["[PAD]", "[CLS]", "[SEP]", "[MASK]", "[UNK]"] + words) + "\n")

    def row():
        pos = rnd.random() < 0.5
        w = [rnd.choice(words[4:]) for _ in range(6)] + [rnd.choice(["good", "great", "liked"] if pos else ["bad", "awful", "hated"])]
        rnd.shuffle(w)
        return " ".join(w) + "\t" + str(int(pos))

    (tmp_path / "data").mkdir()
    for name, n in (("train.tsv", 64), ("dev.tsv", 16)):
        (tmp_path / "data" / name).write_text("text_a\tlabel\n" + "\n".join(row() for _ in range(n)) + "\n")
    opts = CPU + ["Model.num_hidden_layers=2", "Model.hidden_size=64", "Model.num_attention_heads=4", "Model.vocab_size=64", "Model.max_position_embeddings=64",
                  "Model.hidden_dropout_prob=0.0", "Model.attention_probs_dropout_prob=0.0", "Global.local_batch_size=8", "Global.micro_batch_size=8",
                  "Engine.num_train_epochs=6", "Engine.logging_freq=8", "Optimizer.lr.learning_rate=3e-3", f"Engine.save_load.output_dir={tmp_path}/out"]
    for split in ("Train", "Eval"):
        opts += [f"Data.{split}.dataset.input_dir={tmp_path}/data", f"Data.{split}.dataset.tokenizer_type={tmp_path}/vocab", f"Data.{split}.dataset.max_seq_len=32",
                 f"Data.{split}.sampler.batch_size=8", f"Data.{split}.loader.num_workers=0"]
    out = run("tools/train.py", "nlp/ernie/finetune_ernie_345M_single_card.yaml", opts)
    losses = [float(x) for x in re.findall(r"\[train\] epoch: \d+, batch: 7, loss: ([0-9.]+)", out)]
    accs = [float(x) for x in re.findall(r"\[Eval\] epoch: \d+, .*accuracy: ([0-9.]+)", out)]
    lrs = [float(x) for x in re.findall(r"learning rate: ([0-9.e+-]+)", out)]
    assert len(losses) == 6 and len(accs) == 6 and losses[-1] < 0.65 < losses[0] and max(accs) >= 0.75, (losses, accs)
    assert lrs[0] > 1e-3 and lrs[-1] == 0.0          # warm-up finished inside epoch 0; linear decay reaches zero at the last step


def test_ernie_chain_export_serve_with_wordpiece_vocab(tmp_path):
    (tmp_path / "vocab.txt").write_text("\n".join(["[PAD]", "[CLS]", "[SEP]", "[MASK]", "[UNK]", "hello", "my", "dog", "is", "cute", ","]) + "\n")
    ernie = CPU + ["Model.num_hidden_layers=2", "Model.hidden_size=64", "Model.num_attention_heads=4", "Model.vocab_size=512", "Model.max_position_embeddings=64"]
    run("tools/export.py", "nlp/ernie/inference_ernie_345M_single_card.yaml", ernie + [f"Engine.save_load.output_dir={tmp_path}/export"])
    served = run("projects/ernie/inference.py", args=["--model_dir", f"{tmp_path}/export", "--vocab_dir", str(tmp_path), "--seq_len", "16", "--text", "Hello, my dog is cute"])
    assert "output_0 (1, 16, 512)" in served and "output_1 (1, 2)" in served


@pytest.mark.parametrize("layout,degree,reshard_args,dirs", [
    ("mp2", "Distributed.mp_degree=2", [], ["mp_00_sharding_00_pp_00", "mp_01_sharding_00_pp_00"]),
    ("pp2", "Distributed.pp_degree=2", ["--to-plain", "--num-layers", "4"], ["mp_00_sharding_00_pp_00", "mp_00_sharding_00_pp_01"]),
])
def test_reshard_cli_checkpoint_resumes_identically_on_one_process(tmp_path, layout, degree, reshard_args, dirs):
    """tools/train.py on 2 gloo ranks (tensor parallel / pipeline) -> tools/reshard.py --mp 1 [--to-plain] -> a single-process resume reproduces the
    losses of resuming on the original layout: weights (the first stage's copy of a tied embedding layer), optimizer moments, LR schedule and
    data position all cross the layout change."""
    import re

    opts = [o for o in TINY_GPT if not o.startswith("Model.num_layers")] + [
        "Model.num_layers=4", "Data.Train.dataset.name=SyntheticGPTDataset", "Data.Train.dataset.max_seq_len=32", "Data.Train.loader.num_workers=0",
        "Global.local_batch_size=4", "Global.micro_batch_size=2", "Engine.eval_freq=-1", "Engine.logging_freq=1", "Model.hidden_dropout_prob=0.0",
        "Model.attention_probs_dropout_prob=0.0", "Optimizer.lr.max_lr=1e-2", "Optimizer.lr.warmup_rate=0.0", "Model.use_flash_attn=False"]
    cfg = os.path.join(ROOT, CFG, "nlp/gpt/pretrain_gpt_345M_single_card.yaml")

    def launch(nproc, port, extra):
        cmd = [sys.executable]
        if nproc > 1:
            cmd += ["-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr=127.0.0.1", f"--master-port={port}"]
        cmd += [os.path.join(ROOT, "tools/train.py"), "-c", cfg]
        for o in opts + extra:
            cmd += ["-o", o]
        p = subprocess.run(cmd, capture_output=True, text=True, timeout=400, cwd=ROOT)
        assert p.returncode == 0, p.stdout[-1500:] + p.stderr[-2500:]
        return sorted(set(re.findall(r"batch: \[(\d+)/\d+\], loss: ([0-9.]+)", p.stdout + p.stderr)))

    base = 29541 if layout == "mp2" else 29551
    launch(2, base, [degree, "Engine.max_steps=3", "Engine.save_load.save_steps=3", f"Engine.save_load.output_dir={tmp_path}/src"])
    src = os.path.join(tmp_path, "src", "epoch_0_step_3")
    assert sorted(os.listdir(src)) == dirs
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools/reshard.py"), "--src", src, "--dst", f"{tmp_path}/plain", "--mp", "1"] + reshard_args,
                       capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and os.path.isfile(os.path.join(tmp_path, "plain", "model.pdparams")), r.stderr[-1500:]
    resume = ["Engine.max_steps=6", "Engine.save_load.save_steps=-1"]
    one = launch(1, 0, resume + [f"Engine.save_load.ckpt_dir={tmp_path}/plain", f"Engine.save_load.output_dir={tmp_path}/o1"])
    two = launch(2, base + 2, [degree] + resume + [f"Engine.save_load.ckpt_dir={src}", f"Engine.save_load.output_dir={tmp_path}/o2"])
    # the two layouts sum the tensor-parallel partial products in a different order: equal up to fp32 rounding, not bit for bit
    assert [s for s, _ in one] == ["3", "4", "5"] and [s for s, _ in two] == ["3", "4", "5"], (one, two)
    assert all(abs(float(a) - float(b)) <= 2e-6 * abs(float(a)) for (_, a), (_, b) in zip(one, two)), (one, two)


def test_periodic_checkpoint_resume_reproduces_the_uninterrupted_run(tmp_path):
    """``Engine.save_load.save_steps=3`` in a 6-step run, then a second process resuming from ``epoch_0_step_3``: the checkpoint is named by the number
    of completed steps, so the resumed run continues with batch 3 and prints exactly the losses of the uninterrupted run."""
    import re

    opts = TINY_GPT + ["Data.Train.dataset.name=SyntheticGPTDataset", "Data.Train.dataset.max_seq_len=32", "Data.Train.loader.num_workers=0", "Global.local_batch_size=2",
                       "Global.micro_batch_size=2", "Engine.eval_freq=-1", "Engine.logging_freq=1", "Engine.max_steps=6", "Optimizer.lr.max_lr=1e-2",
                       "Optimizer.lr.warmup_rate=0.0"]        # dropout stays on: the RNG streams are part of the checkpoint

    def losses(extra):
        out = run("tools/train.py", "nlp/gpt/pretrain_gpt_345M_single_card.yaml", opts + extra)
        return re.findall(r"batch: \[(\d+)/\d+\], loss: ([0-9.]+)", out)

    straight = losses(["Engine.save_load.save_steps=3", f"Engine.save_load.output_dir={tmp_path}/a"])
    assert [s for s, _ in straight] == [str(i) for i in range(6)] and sorted(os.listdir(tmp_path / "a")) == ["epoch_0_step_3", "epoch_0_step_6"]
    resumed = losses(["Engine.save_load.save_steps=-1", f"Engine.save_load.ckpt_dir={tmp_path}/a/epoch_0_step_3", f"Engine.save_load.output_dir={tmp_path}/b"])
    assert resumed == straight[3:], (straight, resumed)


def test_expert_parallel_run_resumes_identically(tmp_path):
    """GPT-MoE on 2 data-parallel gloo ranks (experts spread over the replicas, dropout and random second-expert routing on): the checkpoint
    carries every replica's experts, their optimizer moments and its own random-number streams, so the resumed run repeats the losses."""
    import re

    opts = TINY_GPT + ["Data.Train.dataset.name=SyntheticGPTDataset", "Data.Train.dataset.max_seq_len=32", "Data.Train.dataset.vocab_size=512",
                       "Data.Train.loader.num_workers=0", "Global.local_batch_size=2", "Global.micro_batch_size=2", "Engine.eval_freq=-1", "Engine.logging_freq=1",
                       "Engine.max_steps=6", "Optimizer.lr.max_lr=1e-2", "Optimizer.lr.warmup_rate=0.0", "Distributed.dp_degree=2"]
    cfg = os.path.join(ROOT, CFG, "nlp/moe/pretrain_moe_345M_single_card.yaml")

    def launch(port, extra):
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr=127.0.0.1", f"--master-port={port}",
               os.path.join(ROOT, "tools/train.py"), "-c", cfg]
        for o in opts + extra:
            cmd += ["-o", o]
        p = subprocess.run(cmd, capture_output=True, text=True, timeout=400, cwd=ROOT)
        assert p.returncode == 0, p.stdout[-1500:] + p.stderr[-2500:]
        return sorted(re.findall(r"batch: \[(\d+)/\d+\], loss: ([0-9.]+)", p.stdout + p.stderr))

    straight = launch(29601, ["Engine.save_load.save_steps=3", f"Engine.save_load.output_dir={tmp_path}/a"])
    ck = os.path.join(tmp_path, "a", "epoch_0_step_3", "mp_00_sharding_00_pp_00")
    assert sorted(os.listdir(os.path.join(ck, "dp_01"))) == ["meta_state.pdopt", "model.pdparams", "model_state.pdopt"]
    resumed = launch(29603, ["Engine.save_load.save_steps=-1", f"Engine.save_load.ckpt_dir={tmp_path}/a/epoch_0_step_3", f"Engine.save_load.output_dir={tmp_path}/b"])
    assert len(resumed) == 6 and resumed == [x for x in straight if int(x[0]) >= 3], (straight, resumed)


def test_epoch_mode_resume_reproduces_the_second_epoch(tmp_path):
    """ViT (epoch mode, dropout on, per-epoch shuffling): resuming from the end-of-epoch-0 checkpoint prints the epoch-1 losses of the 2-epoch run."""
    import re

    vit = CPU + ["Model.model.img_size=32", "Model.model.patch_size=8", "Model.model.depth=2", "Distributed.dp_degree=1", "Data.Train.dataset.name=SyntheticImageDataset",
                 "Data.Train.dataset.image_size=32", "Data.Train.dataset.num_samples=16", "Data.Train.dataset.class_num=10", "Global.local_batch_size=4",
                 "Global.micro_batch_size=4", "Data.Train.sampler.batch_size=4", "Data.Train.loader.num_workers=0", "Engine.logging_freq=1", "Engine.eval_freq=-1",
                 "Engine.num_train_epochs=2", "Optimizer.lr.learning_rate=1e-3", "Optimizer.lr.warmup_steps=0"]
    cfg = "vis/vit/ViT_tiny_patch16_224_ci_cifar10_1n8c_dp_fp16o2.yaml"

    def losses(extra):
        return re.findall(r"epoch: (\d+), step: \[(\d+)/\d+\].*?loss: ([0-9.]+)", run("tools/train.py", cfg, vit + extra))

    straight = losses([f"Engine.save_load.output_dir={tmp_path}/a"])
    assert len(straight) == 8 and sorted(os.listdir(tmp_path / "a")) == ["epoch_0_step_4", "epoch_1_step_4"]
    resumed = losses([f"Engine.save_load.ckpt_dir={tmp_path}/a/epoch_0_step_4", f"Engine.save_load.output_dir={tmp_path}/b"])
    assert resumed == [x for x in straight if x[0] == "1"], (straight, resumed)


def test_ernie_pipeline_corpus_tools_to_pretraining(tmp_path):
    """raw jsonl -> create_pretraining_data (WordPiece, sentence-addressable index) -> ErnieDataset (C++ sample mapping, masking with the vocabulary's
    own special ids) -> tools/train.py: the loss is finite and the MLM random replacements stay inside the (tiny) embedding table."""
    import json
    import random
    import re

    random.seed(0)
    words = ["the", "quick", "brown", "fox", "jump", "##s", "over", "lazy", "dog", "cat", "sleep", "run", "fast", "slow", "house", "tree", "river", "cloud"]
    vocab = tmp_path / "vocab"
    vocab.mkdir()
    (vocab / "vocab.txt").write_text("\n".join(["[PAD]", "[CLS]", "[SEP]", "[MASK]", "[UNK]"] + words + [".", ","]) + "\n")
    base = [w for w in words if not w.startswith("##")]
    with open(tmp_path / "corpus.jsonl", "w") as f:
        for _ in range(60):
            sents = [" ".join(random.choice(base) for _ in range(random.randint(5, 12))) + " ." for _ in range(random.randint(3, 6))]
            f.write(json.dumps({"text": " ".join(sents)}) + "\n")
    r = subprocess.run([sys.executable, "-m", "paddlefleetx_b200.data.data_tools.ernie.create_pretraining_data", "--model_name", str(vocab), "--tokenizer_name",
                        "ErnieTokenizer", "--input_path", str(tmp_path / "corpus.jsonl"), "--output_prefix", str(tmp_path / "data" / "c")],
                       capture_output=True, text=True, cwd=ROOT, timeout=120)
    assert r.returncode == 0 and "sentences" in r.stdout, r.stderr[-1500:]
    out = run("tools/train.py", "nlp/ernie/pretrain_ernie_base.yaml", CPU + [
        "Model.num_hidden_layers=2", "Model.hidden_size=64", "Model.num_attention_heads=4", "Model.vocab_size=128", "Model.max_position_embeddings=64",
        f"Data.Train.dataset.input_dir={tmp_path}/data", "Data.Train.dataset.max_seq_length=32", f"Data.Train.dataset.tokenizer_type={vocab}",
        "Data.Train.loader.num_workers=0", "Global.local_batch_size=4", "Global.micro_batch_size=4", "Engine.eval_freq=-1", "Engine.logging_freq=1",
        "Engine.max_steps=4", "Engine.save_load.save_steps=-1", f"Engine.save_load.output_dir={tmp_path}/out"])
    losses = [float(x) for x in re.findall(r"loss: ([0-9.]+)", out)]
    assert len(losses) >= 4 and all(0 < l < 20 for l in losses), losses
