import os

import torch

from helpers import build_engine, synthetic_batches, tiny_gpt_config


def _run(engine, batches):
    losses = []
    for b in batches:
        losses.append(float(engine.train_step(b)))
    return losses


def test_train_loss_decreases_and_grad_accumulation_is_equivalent():
    cfg1 = tiny_gpt_config(["Global.local_batch_size=4", "Global.micro_batch_size=4", "Engine.max_steps=8"])
    e1 = build_engine(cfg1)
    batches = synthetic_batches(cfg1, 1) * 8           # overfit one batch
    l1 = _run(e1, batches)
    assert l1[-1] < l1[0] - 0.5, l1
    cfg2 = tiny_gpt_config(["Global.local_batch_size=4", "Global.micro_batch_size=1", "Engine.max_steps=8"])
    assert cfg2.Engine.accumulate_steps == 4
    e2 = build_engine(cfg2)
    l2 = _run(e2, batches)
    assert max(abs(a - b) for a, b in zip(l1, l2)) < 2e-3, (l1, l2)


def test_recompute_granularities_match_plain_run():
    base = ["Global.local_batch_size=2", "Global.micro_batch_size=2", "Model.hidden_dropout_prob=0.1"]
    ref = None
    for gran in (None, "full", "full_attn", "core_attn"):
        ov = base + ([f"Model.use_recompute=True", f"Model.recompute_granularity={gran}", "Model.use_flash_attn=False"] if gran
                     else ["Model.use_flash_attn=False"])
        cfg = tiny_gpt_config(ov)
        eng = build_engine(cfg)
        losses = _run(eng, synthetic_batches(cfg, 3, seed=3))
        if ref is None:
            ref = losses
        else:
            assert max(abs(a - b) for a, b in zip(ref, losses)) < 1e-4, (gran, ref, losses)


def test_checkpoint_layout_roundtrip_and_resume(tmp_path):
    out = str(tmp_path / "out")
    cfg = tiny_gpt_config(["Global.local_batch_size=2", "Global.micro_batch_size=2", f"Engine.save_load.output_dir={out}"])
    eng = build_engine(cfg)
    batches = synthetic_batches(cfg, 6, seed=5)
    _run(eng, batches[:3])
    eng.save(epoch=0, step=3)
    d = os.path.join(out, "epoch_0_step_3")
    assert sorted(os.listdir(d)) == ["meta_state.pdopt", "model.pdparams", "model_state.pdopt"]
    keys = torch.load(os.path.join(d, "model.pdparams"), weights_only=False).keys()
    assert "gpt.decoder.layers.0.self_attn.qkv_proj.weight" in keys and "gpt.embeddings.word_embeddings.weight" in keys
    assert "gpt.decoder.norm.bias" in keys and "gpt.decoder.layers.1.linear2.weight" in keys
    cont = _run(eng, batches[3:])

    cfg2 = tiny_gpt_config(["Global.local_batch_size=2", "Global.micro_batch_size=2", f"Engine.save_load.ckpt_dir={d}"])
    eng2 = build_engine(cfg2)
    eng2.load()
    assert eng2._load_recovery["step"] == 3
    resumed = _run(eng2, batches[3:])
    assert max(abs(a - b) for a, b in zip(cont, resumed)) < 1e-5, (cont, resumed)


def test_fit_loop_logs_and_respects_max_steps(caplog):
    from paddlefleetx_b200.data import build_dataloader

    cfg = tiny_gpt_config(["Global.local_batch_size=2", "Global.micro_batch_size=2", "Engine.max_steps=4"])
    eng = build_engine(cfg)
    loader = build_dataloader(cfg.Data, "Train")
    assert len(loader) == 4
    import logging

    with caplog.at_level(logging.INFO, logger="paddlefleetx_b200"):
        eng.fit(train_data_loader=loader, epoch=1)
    out = caplog.text
    assert out.count("[train] epoch: [0/1]") == 4 and "ips_total:" in out and "tokens/s" in out


def test_direct_grad_layout_matches_classic_layout_with_accumulation():
    """``direct_grad`` (autograd never owns flat-grad views; first touch overwrites, no memset) trains identically to the classic
    p.grad-view layout, including gradient accumulation and parameters that get no gradient in a step."""
    import torch

    ov = ["Global.local_batch_size=4", "Global.micro_batch_size=2"]
    cfg_a = tiny_gpt_config(ov)
    cfg_b = tiny_gpt_config(ov + ["Optimizer.direct_grad=True"])
    ea, eb = build_engine(cfg_a), build_engine(cfg_b)
    eb._module.model.load_state_dict(ea._module.model.state_dict())
    assert eb.optimizer.direct_grad and not ea.optimizer.direct_grad
    batches = synthetic_batches(cfg_a, 4, seed=9)
    la = [float(ea.train_step(b)) for b in batches]
    lb = [float(eb.train_step(b)) for b in batches]
    assert max(abs(x - y) for x, y in zip(la, lb)) < 1e-6, (la, lb)
    for (n, p), (_, q) in zip(ea._module.model.named_parameters(), eb._module.model.named_parameters()):
        assert torch.allclose(p, q, atol=1e-6), n
    # a parameter without gradient this step must see zeros, not last step's values
    opt = eb.optimizer
    opt.clear_grad()
    g0 = opt.groups[0]
    g0.params[0].main_grad.fill_(7.0)
    opt._finalize_fresh(g0)
    assert float(g0.params[0].main_grad.abs().sum()) == 0.0


def test_optimizer_state_offload_matches_resident_state():
    """``sharding_offload``: master weights and moments on the host, update on the host, weights copied back — same trajectory."""
    import torch

    ov = ["Global.local_batch_size=2", "Global.micro_batch_size=2"]
    ea = build_engine(tiny_gpt_config(ov))
    eb = build_engine(tiny_gpt_config(ov + ["Optimizer.offload=True"]))
    eb._module.model.load_state_dict(ea._module.model.state_dict())
    assert eb.optimizer.offload and all(g.meta["m"].device.type == "cpu" for g in eb.optimizer.groups)
    batches = synthetic_batches(tiny_gpt_config(ov), 3, seed=4)
    la = [float(ea.train_step(b)) for b in batches]
    lb = [float(eb.train_step(b)) for b in batches]
    assert max(abs(x - y) for x, y in zip(la, lb)) < 1e-6, (la, lb)
    sd = eb.optimizer.state_dict()
    eb.optimizer.set_state_dict(sd)


def test_save_auto_inference_flag_writes_layout_annotated_weights(tmp_path):
    from paddlefleetx_b200.distributed.apis import io

    cfg = tiny_gpt_config([f"Engine.save_load.output_dir={tmp_path}", "Engine.save_load.save_auto_inference=True"])
    eng = build_engine(cfg)
    eng.train_step(synthetic_batches(cfg, 1)[0])
    eng.save(epoch=0, step=1)
    prefix = os.path.join(tmp_path, "auto_infer", "auto")
    assert os.path.isfile(prefix + "_dist0.pdparams") and os.path.isfile(prefix + "_dist0.pdattr")
    attr = torch.load(prefix + "_dist0.pdattr", weights_only=False)
    assert attr["mesh"] == [1, 1] and all(set(a["dims_mapping"]) == {-1} for a in attr["tensors"].values())
    other = build_engine(tiny_gpt_config(["Global.seed=77"]))._module.model
    io.load_auto_inference(prefix, other)
    for (n, a), (_, b) in zip(eng._module.model.state_dict().items(), other.state_dict().items()):
        assert torch.equal(a, b), n
