"""Utility subsystems: checkpoint conversion, watchdog / fault injection, profiler, download/file helpers, config checks."""
import json
import os
import time

import pytest
import torch

from dist_utils import run_distributed
from helpers import build_engine, synthetic_batches, tiny_gpt_config


def test_reshard_single_process_roundtrip_and_named_optimizer(tmp_path):
    from paddlefleetx_b200.utils import ckpt_convert as cc

    out = str(tmp_path / "out")
    ov = ["Global.local_batch_size=2", "Global.micro_batch_size=2"]
    cfg = tiny_gpt_config(ov + [f"Engine.save_load.output_dir={out}"])
    eng = build_engine(cfg)
    batches = synthetic_batches(cfg, 5, seed=3)
    for b in batches[:3]:
        eng.train_step(b)
    eng.save(epoch=0, step=3)
    cont = [float(eng.train_step(b)) for b in batches[3:]]

    dst = str(tmp_path / "conv")
    cc.convert(os.path.join(out, "epoch_0_step_3"), dst, mp=1)
    sd = torch.load(os.path.join(dst, "model_state.pdopt"), weights_only=False)
    assert sd["format"] == "named" and sd["step"] == 3
    eng2 = build_engine(tiny_gpt_config(ov + [f"Engine.save_load.ckpt_dir={dst}"]))
    eng2.load()
    resumed = [float(eng2.train_step(b)) for b in batches[3:]]
    assert max(abs(a - b) for a, b in zip(cont, resumed)) < 1e-5


def test_gpt_pipe_plain_key_mapping_is_inverse():
    from paddlefleetx_b200.utils.ckpt_convert import gpt_pipe_to_plain, gpt_plain_to_pipe

    plain = {"gpt.embeddings.word_embeddings.weight": 1, "gpt.decoder.layers.0.norm1.weight": 2, "gpt.decoder.layers.3.linear2.bias": 3,
             "gpt.decoder.norm.weight": 4}
    pipe = gpt_plain_to_pipe(plain, num_layers=4)
    assert set(pipe) == {"shared_layers.embed.word_embeddings.weight", "layers.1.norm1.weight", "layers.4.linear2.bias", "layers.5.norm.weight"}
    assert gpt_pipe_to_plain(pipe, num_layers=4) == plain


def test_reshard_tensor_parallel_checkpoint(tmp_path):
    run_distributed("dist_fns:ckpt_reshard_tp_sharding", 2, str(tmp_path))


def test_heartbeat_detects_stalled_rank(tmp_path):
    from paddlefleetx_b200.utils.watchdog import Heartbeat

    fired = []
    a = Heartbeat(str(tmp_path), rank=0, world=2, timeout_s=0.3, interval_s=0.05, on_stall=fired.append)
    b = Heartbeat(str(tmp_path), rank=1, world=2, timeout_s=0.3, interval_s=0.05, on_stall=lambda late: None)
    b.beat(0)
    a.start()
    for s in range(8):               # rank 0 keeps beating, rank 1 went silent after step 0
        a.beat(s)
        time.sleep(0.08)
    a.stop()
    assert fired and fired[0] == [1]


def test_fault_injector_and_signal_flag():
    from paddlefleetx_b200.utils import watchdog as wd

    inj = wd.FaultInjector("0:3:raise", rank=0)
    inj.maybe_fire(2)
    with pytest.raises(RuntimeError):
        inj.maybe_fire(3)
    nan = wd.FaultInjector("1:0:nan", rank=1).maybe_fire(0, torch.tensor(1.0))
    assert torch.isnan(nan)
    assert not wd.FaultInjector("1:0:nan", rank=0).active
    wd.install_signal_checkpoint()
    import signal

    os.kill(os.getpid(), signal.SIGUSR1)
    time.sleep(0.05)
    assert wd.emergency_requested()
    wd.clear_emergency()


def test_engine_emergency_checkpoint_on_signal(tmp_path):
    from paddlefleetx_b200.utils import watchdog as wd

    out = str(tmp_path / "out")
    cfg = tiny_gpt_config(["Global.local_batch_size=2", "Global.micro_batch_size=2", f"Engine.save_load.output_dir={out}", "Engine.max_steps=50"])
    eng = build_engine(cfg)
    data = synthetic_batches(cfg, 10, seed=1)

    class Loader(list):
        pass

    wd._EMERGENCY.set()
    try:
        eng.fit(epoch=1, train_data_loader=Loader(data))
    finally:
        wd.clear_emergency()
    saved = os.listdir(out)
    assert any(d.startswith("epoch_0_step_1") for d in saved), saved     # stopped and saved after the first step


def test_step_profiler_window(tmp_path):
    from paddlefleetx_b200.utils.config import AttrDict
    from paddlefleetx_b200.utils.profiler import StepProfiler

    prof = StepProfiler(AttrDict(scheduler=[1, 3], profiler_log=str(tmp_path), detailed=False))
    x = torch.randn(64, 64)
    for _ in range(5):
        (x @ x).sum()
        prof.step()
    prof.finish()
    summ = json.load(open(tmp_path / "summary_rank0.json"))
    assert summ["window"] == [1, 3] and summ["events"]
    assert (tmp_path / "trace_rank0.json").exists()


def test_download_and_file_helpers(tmp_path):
    import zipfile

    from paddlefleetx_b200.utils import download, file as pfile

    src = tmp_path / "vocab.json"
    src.write_text("{}")
    got = download.get_path_from_url("file://" + str(src), str(tmp_path / "cache"))
    assert open(got).read() == "{}" and download.cached_path(str(src)) == str(src)
    with pytest.raises(FileNotFoundError):
        download.cached_path(str(tmp_path / "missing"))
    z = tmp_path / "a.zip"
    with zipfile.ZipFile(z, "w") as f:
        f.writestr("d/x.txt", "hi")
    pfile.unzip(str(z), str(tmp_path / "unz"))
    assert (tmp_path / "unz" / "d" / "x.txt").read_text() == "hi"
    (tmp_path / "t.csv").write_text("skip\n1 2\n3 4\n")
    assert pfile.parse_csv(str(tmp_path / "t.csv"), skip_lines=1) == [["1", "2"], ["3", "4"]]


def test_config_checks():
    from paddlefleetx_b200.utils import check

    cfg = tiny_gpt_config()
    assert check.check_config(cfg) and check.check_version()
    cfg.Model.num_attention_heads = 7
    with pytest.raises(ValueError):
        check.check_config(cfg)
    with pytest.raises(ValueError):
        check.check_device("xpu")


def test_tensor_fusion_helper_groups_and_keeps_values():
    from paddlefleetx_b200.utils.tensor_fusion_helper import all_reduce_parameters, fused_parameters

    m = torch.nn.Sequential(torch.nn.Linear(8, 8), torch.nn.LayerNorm(8))
    before = {n: p.detach().clone() for n, p in m.named_parameters()}
    decay, allg = fused_parameters(m.parameters())
    assert len(allg) == 2 and len(decay) == 1 and decay[0].params[0].dim() == 2
    for n, p in m.named_parameters():
        assert torch.equal(p, before[n])
        assert p.data_ptr() >= min(g.param_buf.data_ptr() for g in allg)
    all_reduce_parameters(allg)      # no process group: no-op


def test_data_tools_produce_a_corpus_gpt_dataset_can_read(tmp_path):
    import numpy as np

    from paddlefleetx_b200.data.data_tools.ernie import create_pretraining_data
    from paddlefleetx_b200.data.data_tools.gpt import preprocess_data, raw_trans_to_json

    raw = tmp_path / "a.txt"
    raw.write_text("First document. It has two sentences!\nAnd a second line.\n\nSecond document, also long enough.\n\nshort\n")
    raw_trans_to_json.main(["--input_path", str(raw), "--output_path", str(tmp_path / "corpus")])
    lines = (tmp_path / "corpus.jsonl").read_text().strip().splitlines()
    assert len(lines) == 2 and json.loads(lines[0])["text"].startswith("First")
    preprocess_data.main(["--input_path", str(tmp_path / "corpus.jsonl"), "--output_prefix", str(tmp_path / "c"), "--append_eos",
                          "--tokenizer_name", "ByteTokenizer"])
    ids, idx = np.load(tmp_path / "c_ids.npy"), np.load(tmp_path / "c_idx.npz")
    assert idx["lens"].sum() == ids.size and len(idx["lens"]) == 2
    create_pretraining_data.main(["--input_path", str(tmp_path / "corpus.jsonl"), "--output_prefix", str(tmp_path / "e"),
                                  "--tokenizer_name", "ByteTokenizer"])
    e = np.load(tmp_path / "e_idx.npz")
    assert e["docs"].tolist() == [0, 3, 4] and len(e["lens"]) == 4 and e["lens"].sum() == np.load(tmp_path / "e_ids.npy").size   # lens per sentence


def test_launcher_spawns_ranks_logs_and_restarts(tmp_path):
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    worker = tmp_path / "w.py"
    worker.write_text(
        "import os, sys\n"
        "r, w, n = os.environ['RANK'], os.environ['WORLD_SIZE'], int(os.environ['PFX_RESTART_COUNT'])\n"
        "print(f'hello from {r}/{w} attempt {n}', flush=True)\n"
        "sys.exit(3 if (r == '1' and n == 0) else 0)\n")
    p = subprocess.run([sys.executable, os.path.join(root, "tools", "launch.py"), "--devices", "cpu:2", "--log_dir", str(tmp_path / "log"),
                        "--max_restart", "1", "--master", "127.0.0.1:29731", str(worker)], capture_output=True, text=True, timeout=120)
    assert p.returncode == 0, p.stderr
    log1 = (tmp_path / "log" / "workerlog.1").read_text()
    assert "hello from 1/2 attempt 0" in log1 and "hello from 1/2 attempt 1" in log1
    assert "hello from 0/2" in p.stdout


def test_ppfleetx_alias_resolves_relocated_reference_module_paths():
    """Import paths of the reference layout whose contents live elsewhere here (single/hybrid/auto model files, per-gate files, layer files)."""
    import importlib

    import ppfleetx  # noqa: F401
    import paddlefleetx_b200.models.language_model.gpt.model as gpt_model
    import paddlefleetx_b200.parallel.tp_layers as tp

    cases = {
        "ppfleetx.models.language_model.gpt.dygraph.single_model": ["GPTModel", "GPTForPretraining", "GPTPretrainingCriterion"],
        "ppfleetx.models.language_model.gpt.dygraph.hybrid_model": ["GPTModel", "GPTForPretraining"],
        "ppfleetx.models.language_model.gpt.dygraph.sequence_parallel_utils": ["ColumnSequenceParallelLinear", "RowSequenceParallelLinear", "ScatterOp", "GatherOp",
                                                                               "AllGatherOp", "ReduceScatterOp", "scatter", "all_gather", "reduce_scatter",
                                                                               "mark_as_sequence_parallel_parameter", "register_sequence_parallel_allreduce_hooks"],
        "ppfleetx.models.language_model.gpt.dygraph.processor": ["LogitsProcessorList", "MinLengthLogitsProcessor", "RepetitionPenaltyLogitsProcessor"],
        "ppfleetx.models.language_model.gpt.auto.auto_module": ["GPTModuleAuto", "GPTGenerationModuleAuto"],
        "ppfleetx.models.language_model.ernie.dygraph.single_model": ["ErnieModel", "ErnieForPretraining", "ErniePretrainingCriterion"],
        "ppfleetx.models.language_model.ernie.layers.transformer": ["TransformerEncoder", "TransformerEncoderLayer"],
        "ppfleetx.models.language_model.ernie.auto.auto_module": ["ErnieModuleAuto"],
        "ppfleetx.models.language_model.t5.modeling": ["T5EncoderModel"],
        "ppfleetx.models.language_model.debertav2.modeling": ["DebertaV2Model"],
        "ppfleetx.models.language_model.utils": ["process_configs", "process_model_configs", "process_data_configs", "process_optim_configs", "is_fused_matmul_bias_supported"],
        "ppfleetx.models.language_model.moe.gate.naive_gate": ["NaiveGate"], "ppfleetx.models.language_model.moe.gate.gshard_gate": ["GShardGate"],
        "ppfleetx.models.language_model.moe.gate.switch_gate": ["SwitchGate"], "ppfleetx.models.language_model.moe.gate.base_gate": ["BaseGate"],
        "ppfleetx.models.vision_model.layers.attention": ["ViTAttention"], "ppfleetx.models.vision_model.layers.mlp": ["ViTMLP"],
        "ppfleetx.models.vision_model.layers.droppath": ["DropPath", "drop_path"], "ppfleetx.models.vision_model.layers.embedding": ["ViTPatchEmbed"],
        "ppfleetx.models.vision_model.layers.identity": ["Identity"], "ppfleetx.models.vision_model.layers.initializer": ["xavier_uniform_2d_"],
        "ppfleetx.models.vision_model.loss.cross_entropy": ["CELoss", "ViTCELoss"], "ppfleetx.models.vision_model.metrics.accuracy": ["TopkAcc"],
        "ppfleetx.models.multimodal_model.imagen.utils": ["GaussianDiffusionContinuousTimes", "resize_image_to"],
        "ppfleetx.data.tokenizers.t5_tokenization_utils": ["PreTrainedTokenizer", "Trie"],
        "ppfleetx.data.data_tools.ernie.preprocess.create_pretraining_data": ["main"], "ppfleetx.data.data_tools.ernie.preprocess.words_segmentation": ["main"],
        "ppfleetx.data.utils.batch_collate_fn": ["collate_fn", "gpt_collate_fn", "ErnieCollateData", "DataCollatorWithPadding", "imagen_collate_fn"],
        "ppfleetx.data.transforms.utils": ["transform", "create_preprocess_operators"],
        "ppfleetx.data.dataset.ernie.dataset_utils": ["MMapIndexedDataset", "create_masked_lm_predictions", "get_samples_mapping", "make_indexed_dataset"],
        "ppfleetx.ops.topp_sampling": ["topp_sampling"], "ppfleetx.tools.multiprocess_tool": ["main"],
        "ppfleetx.models.protein_folding.quat_affine": ["QuatAffine"], "ppfleetx.models.protein_folding.template": ["TemplateEmbedding"],
    }
    for name, symbols in cases.items():
        mod = importlib.import_module(name)
        for s in symbols:
            assert hasattr(mod, s), (name, s)
    assert importlib.import_module("ppfleetx.models.language_model.gpt.dygraph.single_model").GPTModel is gpt_model.GPTModel
    assert importlib.import_module("ppfleetx.models.language_model.gpt.dygraph.sequence_parallel_utils").ColumnSequenceParallelLinear is tp.ColumnSequenceParallelLinear
    import pytest

    with pytest.raises(ModuleNotFoundError):
        importlib.import_module("ppfleetx.models.language_model.gpt.dygraph.no_such_module")


def test_multiprocess_tool_runs_commands_and_reports_failures(tmp_path, capsys):
    from paddlefleetx_b200.tools import multiprocess_tool as mt

    cmds = tmp_path / "cmds.txt"
    cmds.write_text(f"# bulk job\ntouch {tmp_path}/a\ntouch {tmp_path}/b && exit 3\necho hi > {tmp_path}/c\n\n")
    assert mt.read_commands(str(cmds)) == [f"touch {tmp_path}/a", f"touch {tmp_path}/b && exit 3", f"echo hi > {tmp_path}/c"]
    rc = mt.main(["--num_proc", "2", "--shell_cmd_list_filename", str(cmds)])
    out = capsys.readouterr().out
    assert rc == 1 and "2 succeeded, 1 failed" in out and "FAILED rc=3" in out
    assert all((tmp_path / n).exists() for n in "abc")
    ok = tmp_path / "ok.txt"
    ok.write_text("true\ntrue\n")
    assert mt.main(["--num_proc", "4", "--shell_cmd_list_filename", str(ok), "--retries", "1"]) == 0
    slow = tmp_path / "slow.txt"
    slow.write_text("sleep 5\n")
    assert mt.main(["--shell_cmd_list_filename", str(slow), "--timeout", "0.2"]) == 1 and "timed out" in capsys.readouterr().out


def test_ernie_corpus_tool_with_wordpiece_vocab(tmp_path):
    """ERNIE pipeline: jsonl -> sentence-split WordPiece ids with document / sentence boundaries (``--tokenizer_name ErnieTokenizer``)."""
    import json

    import numpy as np

    from paddlefleetx_b200.data.data_tools.ernie import create_pretraining_data

    vocab = tmp_path / "vocab"
    vocab.mkdir()
    (vocab / "vocab.txt").write_text("\n".join(["[PAD]", "[CLS]", "[SEP]", "[MASK]", "[UNK]", "the", "quick", "fox", "jump", "##s", ".", "dog", "sleep"]) + "\n")
    corpus = tmp_path / "c.jsonl"
    corpus.write_text(json.dumps({"text": "The quick fox jumps. The dog sleeps."}) + "\n" + json.dumps({"text": "The fox sleeps."}) + "\n")
    create_pretraining_data.main(["--model_name", str(vocab), "--tokenizer_name", "ErnieTokenizer", "--input_path", str(corpus),
                                  "--output_prefix", str(tmp_path / "out"), "--append_eos"])
    ids = np.load(tmp_path / "out_ids.npy")
    idx = np.load(tmp_path / "out_idx.npz")
    assert ids.dtype == np.uint16 and 2 in ids.tolist() and 1 not in ids.tolist()      # [SEP] closes documents, no [CLS] frame in corpus ids
    assert idx["lens"].tolist() == [6, 6, 6] and idx["docs"].tolist() == [0, 2, 3] and idx["lens"].sum() == len(ids)     # per sentence; docs in sentences
    from paddlefleetx_b200.data.dataset.ernie.ernie_dataset import MMapIndexedDataset

    ds = MMapIndexedDataset(str(tmp_path / "out"))
    assert len(ds) == 3 and ds[1].tolist() == [5, 11, 12, 9, 10, 2] and ds.doc_idx.tolist() == [0, 2, 3]
    assert ids[:5].tolist() == [5, 6, 7, 8, 9]                                          # the quick fox jump ##s
