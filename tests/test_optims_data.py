import os

import numpy as np
import pytest
import torch

from paddlefleetx_b200.optims import lr_scheduler as L


def test_cosine_warmup_decay_in_samples():
    s = L.CosineAnnealingWithWarmupDecay(max_lr=1.0, min_lr=0.1, warmup_rate=0.1, decay_steps=1000, use_increments=True)
    assert s() == 0.0
    s.step(epoch=50)
    assert abs(s() - 0.5) < 1e-12
    s.step(epoch=50)
    assert abs(s() - 1.0) < 1e-12
    s.step(epoch=450)           # halfway through the cosine
    assert abs(s() - (0.1 + 0.5 * 0.9)) < 1e-9
    s.step(epoch=10_000)
    assert s() == 0.1
    sd = s.state_dict()
    t = L.CosineAnnealingWithWarmupDecay(1.0, 0.1, 0.1, 1000)
    t.set_state_dict(sd)
    assert t() == s()


def test_fixed_schedulers():
    lin = L.LinearDecayWithWarmup(1.0, step_each_epoch=10, epochs=10, warmup=0.1)
    vals = []
    for _ in range(100):
        vals.append(lin()); lin.step()
    assert vals[0] == 0.0 and abs(vals[10] - 0.9) < 1e-9 and vals[-1] < 0.02      # reference: lr * (1 - t / T_max) after warm-up
    cd = L.CosineDecay(1.0, step_each_epoch=5, epochs=4, update_unit="step", warmups=1)
    assert cd() == pytest.approx(1 / 5)
    ms = L.MultiStepDecay(1.0, [2, 4], 0.1)
    seq = []
    for _ in range(5):
        seq.append(ms()); ms.step()
    assert seq == pytest.approx([1.0, 1.0, 0.1, 0.1, 0.01])
    v = L.ViTLRScheduler(1.0, step_each_epoch=10, epochs=1, decay_type="cosine", warmup_steps=2)
    assert v() == 0.0


def test_fused_adamw_matches_torch_adamw_cpu():
    from paddlefleetx_b200.optims import ClipGradByGlobalNorm, FusedAdamW

    torch.manual_seed(0)
    m1 = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.LayerNorm(16), torch.nn.Linear(16, 4))
    m2 = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.LayerNorm(16), torch.nn.Linear(16, 4))
    m2.load_state_dict(m1.state_dict())
    opt = FusedAdamW(1e-2, named_parameters=list(m1.named_parameters()), weight_decay=0.1, beta2=0.95, grad_clip=ClipGradByGlobalNorm(0.5))
    decay = [p for n, p in m2.named_parameters() if p.ndim > 1]
    no_decay = [p for n, p in m2.named_parameters() if p.ndim <= 1]
    ref = torch.optim.AdamW([{"params": decay, "weight_decay": 0.1}, {"params": no_decay, "weight_decay": 0.0}], lr=1e-2, betas=(0.9, 0.95))
    x = torch.randn(5, 8)
    for _ in range(4):
        m1(x).pow(2).sum().backward()
        m2(x).pow(2).sum().backward()
        torch.nn.utils.clip_grad_norm_(m2.parameters(), 0.5)
        opt.step(); ref.step()
        opt.clear_grad(); ref.zero_grad()
    for (n, a), (_, b) in zip(m1.named_parameters(), m2.named_parameters()):
        assert torch.allclose(a, b, atol=2e-5, rtol=1e-4), n
    sd = opt.state_dict()
    opt.set_state_dict(sd)


def _toy_corpus(tmp_path, n_docs=40, seed=0):
    rng = np.random.RandomState(seed)
    lens = rng.randint(5, 60, size=n_docs).astype(np.int32)
    ids = rng.randint(0, 1000, size=int(lens.sum())).astype(np.int32)
    np.save(tmp_path / "toy_ids.npy", ids)
    np.savez(tmp_path / "toy_idx.npz", lens=lens)
    return ids, lens


def test_cpp_sample_idx_matches_python_oracle(tmp_path):
    from paddlefleetx_b200.data.dataset import gpt_dataset as G

    _, lens = _toy_corpus(tmp_path)
    docs = np.arange(len(lens))
    rng = np.random.RandomState(1)
    for seq_len, epochs in ((16, 1), (33, 3), (7, 2)):
        doc_idx = G._doc_order(docs, epochs, rng, False)
        tpe = int(lens.sum())
        a = G._helpers().build_sample_idx(lens, doc_idx, seq_len, epochs, tpe)
        b = G.python_sample_idx(lens, doc_idx, seq_len, epochs, tpe)
        assert a.dtype == np.int64 and np.array_equal(a, b)


def test_gpt_dataset_samples_are_contiguous_token_windows(tmp_path):
    from paddlefleetx_b200.data.dataset.gpt_dataset import GPTDataset

    ids, lens = _toy_corpus(tmp_path)
    ds = GPTDataset(str(tmp_path), [8, 1, 1], 16, num_samples=50, mode="Train", seed=7)
    assert len(ds) >= 50
    tok, pos, lab, mask = ds[3]
    assert tok.shape == (16,) and lab.shape == (16,) and np.array_equal(tok[1:], lab[:-1])
    assert mask.sum() == 16 and np.array_equal(pos, np.arange(16))
    # cache files are reused and named like the reference
    names = sorted(os.listdir(tmp_path))
    assert any(n.endswith("_gpt_Train_indexmap_50ns_16sl_sample_idx.npy") for n in names)
    ds2 = GPTDataset(str(tmp_path), [8, 1, 1], 16, num_samples=50, mode="Train", seed=7)
    assert np.array_equal(ds2[3][0], tok)
    ds3 = GPTDataset(str(tmp_path), [8, 1, 1], 16, num_samples=50, mode="Train", seed=7, mask_eos=True, eos_id=int(tok[2]))
    assert ds3[3][3][2] == 0.0


def test_batch_sampler_rank_slices_are_disjoint_and_resumable():
    from paddlefleetx_b200.data.sampler.batch_sampler import GPTBatchSampler

    data = list(range(64))
    per_rank = [list(GPTBatchSampler(data, 4, num_replicas=2, rank=r)) for r in range(2)]
    assert per_rank[0][0] == [0, 1, 2, 3] and per_rank[1][0] == [4, 5, 6, 7] and per_rank[0][1] == [8, 9, 10, 11]
    assert len(per_rank[0]) == 8
    resumed = list(GPTBatchSampler(data, 4, num_replicas=2, rank=1, consumed_samples=16))
    assert resumed[0] == per_rank[1][2]


def test_mapping_helpers_shapes_and_determinism():
    from paddlefleetx_b200.data.dataset.gpt_dataset import _helpers

    h = _helpers()
    rng = np.random.RandomState(0)
    sizes = rng.randint(3, 40, size=200).astype(np.int32)
    docs = np.concatenate([[0], np.cumsum(rng.randint(1, 8, size=40))]).astype(np.int64)
    docs = docs[docs <= 200]
    a = h.build_mapping(docs, sizes, 2, 10_000, 64, 0.1, 1234, False, 2)
    b = h.build_mapping(docs, sizes, 2, 10_000, 64, 0.1, 1234, False, 2)
    assert a.shape[1] == 3 and np.array_equal(a, b) and a.dtype == np.uint32
    assert (a[:, 1] > a[:, 0]).all() and (a[:, 2] <= 64).all()
    titles = np.full(len(docs) - 1, 4, dtype=np.int32)
    blk = h.build_blocks_mapping(docs, sizes, titles, 1, 10_000, 64, 7, False, False)
    assert blk.shape[1] == 4
    w = np.array([0.7, 0.2, 0.1])
    di, si = np.zeros(1000, np.uint8), np.zeros(1000, np.int64)
    h.build_blending_indices(di, si, w, 3, 1000, False)
    assert abs((di == 0).mean() - 0.7) < 0.01 and si[di == 1].max() == (di == 1).sum() - 1


def test_forward_hooks_cover_parameters_read_through_a_child(monkeypatch):
    """Side-stream work (ZeRO all-gather / overlapped AdamW) is awaited per bucket by module pre-hooks.  Fused call sites read a child's
    tensors without calling the child, so the parent must wait for its children's buckets too."""
    import torch
    import torch.nn as nn
    import torch.nn.functional as F

    from paddlefleetx_b200.optims import FusedAdamW

    class Fused(nn.Module):                      # reads lin's parameters directly, never calls lin()
        def __init__(self):
            super().__init__()
            self.lin = nn.Linear(8, 8)

        def forward(self, x):
            return F.linear(x, self.lin.weight, self.lin.bias)

    class Net(nn.Module):
        def __init__(self):
            super().__init__()
            self.emb = nn.Linear(8, 8)
            self.blocks = nn.ModuleList([Fused(), Fused()])

        def forward(self, x):
            x = F.linear(x, self.emb.weight, self.emb.bias)
            for b in self.blocks:
                x = b(x)
            return x

    net = Net()
    opt = FusedAdamW(1e-3, named_parameters=list(net.named_parameters()), step_overlap=True)
    assert opt.step_overlap is False                       # CPU tensors: the flag is inert ...
    opt.step_overlap, opt._comm_stream = True, object()    # ... force the hook installation path
    waited = []

    class _Stream:
        def wait_event(self, ev):
            waited.append(ev)

    monkeypatch.setattr(torch.cuda, "current_stream", lambda *a, **k: _Stream())
    opt.install_forward_hooks(net)
    assert opt._fwd_hooks_installed
    opt._ag_events = {id(g): f"ev{i}" for i, g in enumerate(opt.groups)}
    expected = set(opt._ag_events.values())
    net(torch.randn(2, 8))
    assert set(waited) == expected and not opt._ag_events   # every bucket was awaited although lin.forward never ran
    waited.clear()
    net(torch.randn(2, 8))
    assert waited == []                                     # nothing in flight: hooks are free


def test_glue_split_names_select_their_files(tmp_path):
    """``split`` values the recipes / scripts use: train, dev, test, dev_matched, dev_mismatched; CoLA also in its public ``raw/in_domain_*`` layout."""
    from paddlefleetx_b200.data.dataset import glue_dataset as G

    d = tmp_path / "MNLI"
    d.mkdir()
    hdr = "\t".join(f"c{i}" for i in range(12))
    row = lambda lab, t: "\t".join(["x"] * 8 + [f"premise {t}", f"hypothesis {t}", "x", lab])  # noqa: E731
    for fn, t in (("train.tsv", "tr"), ("dev_matched.tsv", "m"), ("dev_mismatched.tsv", "mm")):
        (d / fn).write_text(hdr + "\n" + row("entailment", t) + "\n" + row("neutral", t) + "\n")
    got = {split: G.MNLI(root=str(tmp_path), split=split).samples[0][0] for split in ("train", "dev", "dev_matched", "dev_mismatched")}
    assert got == {"train": "premise tr", "dev": "premise m", "dev_matched": "premise m", "dev_mismatched": "premise mm"}
    raw = tmp_path / "cola_public" / "raw"
    raw.mkdir(parents=True)
    (raw / "in_domain_train.tsv").write_text("s\t1\t*\tgood sentence\ns\t0\t*\tbad sentence\n")
    (raw / "in_domain_dev.tsv").write_text("s\t1\t*\tdev sentence\n")
    assert len(G.CoLA(root=str(tmp_path / "cola_public"), split="train")) == 2 and len(G.CoLA(root=str(tmp_path / "cola_public"), split="dev")) == 1


def test_samplers_skip_consumed_batches_once():
    """Resume support: ``skip_batches`` makes the next pass start that many local batches in (nothing before it is ever indexed); the pass after is full again."""
    from paddlefleetx_b200.data.sampler.batch_sampler import DistributedBatchSampler, GPTBatchSampler

    data = list(range(40))
    for make in (lambda: GPTBatchSampler(data, batch_size=3, num_replicas=2, rank=1),
                 lambda: DistributedBatchSampler(data, batch_size=3, num_replicas=2, rank=1, shuffle=True, seed=5)):
        s = make()
        full = list(s)
        s.skip_batches = 2
        assert list(s) == full[2:]
        assert list(s) == full and len(s) == len(full)
        s.skip_batches = len(full) + 3
        assert list(s) == []
