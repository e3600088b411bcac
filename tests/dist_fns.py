"""Rank functions for the multi-process gloo tests (imported by name inside the spawned children)."""

import torch
import torch.distributed as dist

from helpers import build_engine, synthetic_batches, tiny_gpt_config


def _slice(batch, rank, world):
    n = batch[0].shape[0] // world
    return [t[rank * n:(rank + 1) * n].contiguous() for t in batch]


def _reference_losses_and_state(overrides, n_steps, seed=11, batch_hook=None):
    """Single-process run of the same global batch, executed redundantly on every rank (world-1 topology)."""
    from paddlefleetx_b200.distributed.apis import env
    from paddlefleetx_b200.parallel.topology import HybridCommunicateGroup

    cfg = tiny_gpt_config(overrides, nranks=1)
    env.set_hcg(HybridCommunicateGroup(world_size=1, rank=0, build_groups=False))
    real_ws = env.world_size
    env.world_size = lambda: 1                    # build as if single process
    try:
        from paddlefleetx_b200.core import EagerEngine
        from paddlefleetx_b200.models import build_module

        env.set_seed(cfg.Global.seed)
        module = build_module(cfg)
        init = {k: v.detach().clone() for k, v in module.model.state_dict().items()}
        eng = EagerEngine(configs=cfg, module=module)
        batches = synthetic_batches(cfg, n_steps, seed=seed)
        if batch_hook is not None:
            batches = [batch_hook(b) for b in batches]
        losses = [float(eng.train_step(b)) for b in batches]
        state = {k: v.detach().clone() for k, v in module.model.state_dict().items()}
    finally:
        env.world_size = real_ws
        env.set_hcg(None)
    return cfg, batches, losses, state, init


def dp_sharding_matches_single(rank, world, dp, sharding, stage, extra=()):
    gb = 4
    base = [f"Global.global_batch_size={gb}", "Global.local_batch_size=None", "Global.micro_batch_size=1"]
    _, batches, ref_losses, ref_state, init = _reference_losses_and_state(
        ["Global.global_batch_size=None", f"Global.local_batch_size={gb}", f"Global.micro_batch_size={gb}"], 4)
    cfg = tiny_gpt_config(base + [f"Distributed.dp_degree={dp}", f"Distributed.sharding.sharding_degree={sharding}",
                                  f"Distributed.sharding.sharding_stage={stage}"] + list(extra), nranks=world)
    from paddlefleetx_b200.core import EagerEngine
    from paddlefleetx_b200.distributed.apis import env
    from paddlefleetx_b200.models import build_module

    env.init_dist_env(cfg)
    env.set_seed(cfg.Global.seed)
    module = build_module(cfg)
    module.model.load_state_dict(init)
    eng = EagerEngine(configs=cfg, module=module)
    dr, dw = env.get_data_world_rank(), env.get_data_world_size()
    assert dw == dp * sharding
    losses = []
    for b in batches:
        l = eng.train_step(_slice(b, dr, dw)).detach().clone()
        dist.all_reduce(l)
        losses.append(float(l) / world)
    assert max(abs(a - b) for a, b in zip(losses, ref_losses)) < 2e-4, (losses, ref_losses)
    for k, v in eng.module.model.state_dict().items():
        assert torch.allclose(v, ref_state[k], atol=5e-5, rtol=1e-4), k
    if sharding > 1 and stage < 3:        # optimizer state really is sharded
        g = eng.optimizer.groups[0]
        assert g.meta["m"].numel() == g.numel // sharding


def zero2_ring_matches_single(rank, world, micro):
    """ZeRO stage 2 with the gradient ring (many small buckets sharing 2 slots), with and without gradient accumulation: same losses and
    weights as one process, and the full-size gradient really is gone (buckets alias the ring, each keeps a 1/N shard)."""
    import os

    os.environ["PFX_ZERO2_SLOTS"] = "2"
    gb = 8
    _, batches, ref_losses, ref_state, init = _reference_losses_and_state(
        ["Global.global_batch_size=None", f"Global.local_batch_size={gb}", f"Global.micro_batch_size={gb}", "Model.num_layers=4"], 3)
    cfg = tiny_gpt_config(["Model.num_layers=4", f"Global.global_batch_size={gb}", "Global.local_batch_size=None", f"Global.micro_batch_size={micro}",
                           "Distributed.sharding.sharding_degree=2", "Distributed.sharding.sharding_stage=2", "Distributed.sharding.bucket_mb=0.02",
                           "Distributed.sharding.reduce_overlap=True", "Optimizer.direct_grad=True"], nranks=world)
    from paddlefleetx_b200.core import EagerEngine
    from paddlefleetx_b200.distributed.apis import env
    from paddlefleetx_b200.models import build_module

    env.init_dist_env(cfg)
    env.set_seed(cfg.Global.seed)
    module = build_module(cfg)
    module.model.load_state_dict(init)
    eng = EagerEngine(configs=cfg, module=module)
    opt = eng.optimizer
    ring = [g for g in opt.groups if "ring_idx" in g.meta]
    assert opt.grad_ring and len(ring) >= 4, (opt.grad_ring, len(ring), len(opt.groups))
    assert ring[0].grad_buf.data_ptr() == ring[2].grad_buf.data_ptr() != ring[1].grad_buf.data_ptr()
    assert all(g.meta["grad_shard"].numel() == g.numel // 2 for g in ring)
    dr, dw = env.get_data_world_rank(), env.get_data_world_size()
    losses = []
    for b in batches:
        l = eng.train_step(_slice(b, dr, dw)).detach().clone()
        dist.all_reduce(l)
        losses.append(float(l) / world)
    assert max(abs(a - b) for a, b in zip(losses, ref_losses)) < 2e-4, (losses, ref_losses)
    for k, v in eng.module.model.state_dict().items():
        assert torch.allclose(v, ref_state[k], atol=5e-5, rtol=1e-4), k


def _shard_like(full: torch.Tensor, p: torch.nn.Parameter, mp_rank: int, mp: int) -> torch.Tensor:
    if getattr(p, "tp_sharded", False):
        return full.chunk(mp, dim=p.split_axis)[mp_rank].clone()
    return full.clone()


def tp_matches_single(rank, world, sequence_parallel):
    mp = world
    gb = 2
    single_ov = ["Global.global_batch_size=None", f"Global.local_batch_size={gb}", f"Global.micro_batch_size={gb}"]
    from paddlefleetx_b200.distributed.apis import env
    from paddlefleetx_b200.parallel.topology import HybridCommunicateGroup

    # reference model: remember its *initial* weights
    cfg1 = tiny_gpt_config(single_ov, nranks=1)
    env.set_hcg(HybridCommunicateGroup(world_size=1, rank=0, build_groups=False))
    real_ws = env.world_size
    env.world_size = lambda: 1
    try:
        from paddlefleetx_b200.core import EagerEngine
        from paddlefleetx_b200.models import build_module

        env.set_seed(cfg1.Global.seed)
        module1 = build_module(cfg1)
        init = {k: v.detach().clone() for k, v in module1.model.state_dict().items()}
        eng1 = EagerEngine(configs=cfg1, module=module1)
        batches = synthetic_batches(cfg1, 3, seed=21)
        ref_losses = [float(eng1.train_step(b)) for b in batches]
        ref_state = {k: v.detach().clone() for k, v in module1.model.state_dict().items()}
    finally:
        env.world_size = real_ws
        env.set_hcg(None)

    cfg = tiny_gpt_config(single_ov + [f"Distributed.mp_degree={mp}", f"Model.sequence_parallel={sequence_parallel}"], nranks=world)
    env.init_dist_env(cfg)
    env.set_seed(cfg.Global.seed)
    module = build_module(cfg)
    hcg = env.get_hcg()
    with torch.no_grad():
        for k, p in module.model.named_parameters():
            p.copy_(_shard_like(init[k], p, hcg.get_model_parallel_rank(), mp))
    eng = EagerEngine(configs=cfg, module=module)
    losses = [float(eng.train_step(b)) for b in batches]
    assert max(abs(a - b) for a, b in zip(losses, ref_losses)) < 3e-4, (losses, ref_losses)
    for k, p in module.model.named_parameters():
        want = _shard_like(ref_state[k], p, hcg.get_model_parallel_rank(), mp)
        assert torch.allclose(p.detach(), want, atol=1e-4, rtol=1e-3), (k, (p.detach() - want).abs().max())


def tp_shards_differ_across_ranks(rank, world):
    """From-scratch tensor-parallel init: every mp rank must draw a DIFFERENT slice (identical shards = duplicated neurons)."""
    from paddlefleetx_b200.distributed.apis import env
    from paddlefleetx_b200.models import build_module

    cfg = tiny_gpt_config(["Global.global_batch_size=None", "Global.local_batch_size=2", "Global.micro_batch_size=2",
                           f"Distributed.mp_degree={world}"], nranks=world)
    env.init_dist_env(cfg)
    env.set_seed(cfg.Global.seed)
    module = build_module(cfg)
    checked = 0
    for n, p in module.model.named_parameters():
        if not getattr(p, "tp_sharded", False) or p.ndim < 2:
            continue
        parts = [torch.empty_like(p.detach()) for _ in range(world)]
        dist.all_gather(parts, p.detach().contiguous())
        assert not torch.equal(parts[0], parts[1]), f"{n}: identical shards on mp ranks 0 and 1"
        checked += 1
    assert checked >= 4, checked
    # replicated tensors still agree inside the mp group
    for n, p in module.model.named_parameters():
        if getattr(p, "tp_sharded", False):
            continue
        parts = [torch.empty_like(p.detach()) for _ in range(world)]
        dist.all_gather(parts, p.detach().contiguous())
        assert torch.equal(parts[0], parts[1]), n


def pipeline_matches_single(rank, world, pp, mp, vpp, acc, sp=False, cp=1, cp_mode="ulysses", extra=()):
    """pp (x mp, optionally with Megatron sequence parallelism) pipeline with tied embeddings reproduces the single-process loss curve
    AND the single-process weights (sequence-partial LayerNorm / bias gradients must be summed over the mp group on every stage)."""
    from paddlefleetx_b200.core import EagerEngine
    from paddlefleetx_b200.distributed.apis import env
    from paddlefleetx_b200.models import build_module

    gb = acc
    L = 4
    common = [f"Model.num_layers={L}", "Model.use_flash_attn=False"] + list(extra)
    _, batches, ref_losses, ref_state, init = _reference_losses_and_state(
        # with an uneven mask the mean of per-micro-batch means is not the mean over the batch: the reference accumulates the same micro-batches
        common + ["Global.global_batch_size=None", f"Global.local_batch_size={gb}", f"Global.micro_batch_size={1 if cp > 1 else gb}"], 3, seed=31,
        batch_hook=_mask_out_a_ragged_tail if cp > 1 else None)
    ov = common + ["Global.global_batch_size=None", f"Global.local_batch_size={gb}", "Global.micro_batch_size=1",
                   f"Distributed.pp_degree={pp}", f"Distributed.mp_degree={mp}"]
    if cp > 1:       # the data axis (world / (pp x mp) ranks, ZeRO-1) is one context-parallel group: every rank sees the whole batch
        ov += [f"Distributed.sharding.sharding_degree={cp}", "Distributed.sharding.sharding_stage=1", f"Distributed.cp_degree={cp}", f"Distributed.cp_mode={cp_mode}"]
    if vpp > 1:
        ov.append(f"Model.virtual_pp_degree={vpp}")
    if sp:
        ov.append("Model.sequence_parallel=True")
    cfg = tiny_gpt_config(ov, nranks=world)
    env.init_dist_env(cfg)
    env.set_seed(cfg.Global.seed)
    module = build_module(cfg)
    hcg = env.get_hcg()
    pipe = module.model
    # map the single-model initial weights onto this stage's layers through the converter's plain -> pipe key map
    # (pipeline keys carry the global layer index, so the same map works for every pp / virtual-pp layout)
    import re

    from paddlefleetx_b200.utils.ckpt_convert import gpt_plain_to_pipe

    pipe_init = gpt_plain_to_pipe(init, L)
    with torch.no_grad():
        seen = 0
        for n, p in pipe.named_parameters():
            key = re.sub(r"^_model_chunks\.\d+\.", "layers.", n)
            p.copy_(_shard_like(pipe_init[key], p, hcg.get_model_parallel_rank(), mp))
            seen += 1
        assert seen > 0
    eng = EagerEngine(configs=cfg, module=module)
    losses = []
    for b in batches:
        l = eng.train_step(b).detach().float().clone()
        if cp > 1:                                   # the cp ranks' losses average to the loss over whole sequences
            dist.all_reduce(l)
            l = l / world
        losses.append(float(l))
    assert max(abs(a - b) for a, b in zip(losses, ref_losses)) < 3e-4, (rank, losses, ref_losses)
    # with tensor parallelism every stage-boundary tensor travels as a 1/mp slice + an all-gather on the receiving side
    p2p = eng._module.model._p2p if hasattr(eng._module.model, "_p2p") else eng._dist_model._p2p
    if not sp:       # SP activations are already sequence shards: partial send/recv is off (reference utils/config.py:112-119)
        assert p2p.partial == (mp > 1) and (p2p.partial_transfers > 0) == (mp > 1), (mp, p2p.partial, p2p.partial_transfers)
    # every parameter of this stage (replicated LayerNorm / bias included) followed the single-process trajectory
    ref_pipe = gpt_plain_to_pipe(ref_state, L)
    for n, p in pipe.named_parameters():
        key = re.sub(r"^_model_chunks\.\d+\.", "layers.", n)
        if n.startswith("shared_layers.") and "word_embeddings" not in n and hcg.get_stage_id() != 0:
            continue      # the last stage's copy of the embedding block only lends its (tied) word-embedding matrix to the LM head
        want, got = _shard_like(ref_pipe[key], p, hcg.get_model_parallel_rank(), mp), p.detach()
        if cp > 1 and n.endswith("qkv_proj.bias"):      # K-bias: zero gradient in theory, Adam-normalised rounding noise in practice
            hl = cfg.Model.num_attention_heads // mp
            want, got = want.view(hl, 3, -1)[:, [0, 2]], got.view(hl, 3, -1)[:, [0, 2]]
        bad = (got - want).abs() > 1e-4 + 1e-3 * want.abs()
        assert int(bad.sum()) <= (max(1, bad.numel() // 1000) if cp > 1 else 0) and float((got - want).abs().max()) < 1e-3, (n, (got - want).abs().max())
    # tied embedding stays identical on first and last stage
    if "embed" in pipe.shared_layers:
        w = pipe.shared_layers["embed"].word_embeddings.weight.detach()
        want = _shard_like(ref_state["gpt.embeddings.word_embeddings.weight"], pipe.shared_layers["embed"].word_embeddings.weight,
                           hcg.get_model_parallel_rank(), mp)
        assert torch.allclose(w, want, atol=1e-4, rtol=1e-3), (w - want).abs().max()


def moe_ep_matches_single(rank, world, gate_type):
    """Expert-parallel MoELayer (2 ranks x 2 experts) == single-process layer with 4 experts on the concatenated batch."""
    from paddlefleetx_b200.models.language_model.moe.moe_layer import ExpertLayer, MoELayer
    from paddlefleetx_b200.parallel.topology import HybridCommunicateGroup

    hcg = HybridCommunicateGroup(dp=world)
    grp = hcg.get_moe_group()
    d, f, e_local, n_tok = 16, 32, 2, 24
    torch.manual_seed(5)
    full = MoELayer(d, [ExpertLayer(d, f) for _ in range(e_local * world)], gate={"type": gate_type, "top_k": 2}, moe_group=None)
    x_all = torch.randn(world * n_tok, d)
    torch.manual_seed(99)
    ep = MoELayer(d, [ExpertLayer(d, f) for _ in range(e_local)], gate={"type": gate_type, "top_k": 2}, moe_group=grp)
    with torch.no_grad():
        ep.gate.gate.weight.copy_(full.gate.gate.weight); ep.gate.gate.bias.copy_(full.gate.gate.bias)
        for i in range(e_local):
            ep.experts[i].load_state_dict(full.experts[rank * e_local + i].state_dict())
    full.eval(); ep.eval()          # eval: no random routing
    if gate_type == "gshard":
        full.gate.random_routing = ep.gate.random_routing = False
    x = x_all[rank * n_tok:(rank + 1) * n_tok].clone().requires_grad_(True)
    xa = x_all.clone().requires_grad_(True)
    y = ep(x)
    ya = full(xa)
    assert torch.allclose(y, ya[rank * n_tok:(rank + 1) * n_tok], atol=1e-5), (y - ya[rank * n_tok:(rank + 1) * n_tok]).abs().max()
    g_all = torch.randn(world * n_tok, d, generator=torch.Generator().manual_seed(1))
    y.backward(g_all[rank * n_tok:(rank + 1) * n_tok])
    ya.backward(g_all)
    assert torch.allclose(x.grad, xa.grad[rank * n_tok:(rank + 1) * n_tok], atol=1e-5)
    for i in range(e_local):
        for (n1, p1), (n2, p2) in zip(ep.experts[i].named_parameters(), full.experts[rank * e_local + i].named_parameters()):
            assert torch.allclose(p1.grad, p2.grad, atol=1e-5), (i, n1)


def moe_module_trains(rank, world):
    cfg = tiny_gpt_config(["Model.module=MoEModule", "Model.moe_configs={'expert_mode':True,'num_experts':2,'gate':'gshard','top_k':2}",
                           "Global.local_batch_size=2", "Global.micro_batch_size=2", "Optimizer.grad_clip.name=ClipGradForMOEByGlobalNorm",
                           "Distributed.hcg=HybridCommGroupForMoE", f"Distributed.dp_degree={world}"], nranks=world)
    eng = build_engine(cfg)
    b = synthetic_batches(cfg, 1, seed=rank)
    losses = [float(eng.train_step([t[rank * 2:(rank + 1) * 2] for t in b[0]])) for _ in range(6)]
    assert losses[-1] < losses[0] - 0.5, losses
    # dense params stay in sync across dp, expert params differ
    for n, p in eng.module.model.named_parameters():
        t = p.detach().clone()
        dist.broadcast(t, src=0)
        same = torch.allclose(t, p.detach())
        if rank != 0:
            assert same != bool(getattr(p, "is_expert", False)), (n, same)


def ckpt_reshard_tp_sharding(rank, world, tmpdir):
    """mp2 (world 2) or sharding2 run → save → offline merge equals the in-memory full state; the merged optimizer state,
    re-split, resumes bit-compatibly on the same layout through the ``format: named`` path."""
    import os

    from paddlefleetx_b200.core import EagerEngine
    from paddlefleetx_b200.distributed.apis import env
    from paddlefleetx_b200.models import build_module
    from paddlefleetx_b200.utils import ckpt_convert as cc

    out = os.path.join(tmpdir, "out")
    ov = ["Global.global_batch_size=None", "Global.local_batch_size=2", "Global.micro_batch_size=2", f"Distributed.mp_degree={world}",
          f"Engine.save_load.output_dir={out}"]
    cfg = tiny_gpt_config(ov, nranks=world)
    env.init_dist_env(cfg)
    env.set_seed(cfg.Global.seed)
    module = build_module(cfg)
    eng = EagerEngine(configs=cfg, module=module)
    batches = synthetic_batches(cfg, 4, seed=5)
    for b in batches[:2]:
        eng.train_step(b)
    eng.save(epoch=0, step=2)
    dist.barrier()
    cont = [float(eng.train_step(b)) for b in batches[2:]]

    src = os.path.join(out, "epoch_0_step_2")
    dst = os.path.join(tmpdir, "resplit")
    if rank == 0:
        model, optim, meta = cc.merge_checkpoint(src)
        assert model["gpt.decoder.layers.0.self_attn.qkv_proj.weight"].shape[0] == 3 * cfg.Model.hidden_size
        assert optim["format"] == "named" and "gpt.decoder.layers.0.linear2.weight" in optim["state"]
        cc.split_checkpoint(model, optim, meta, dst, mp=world)
    dist.barrier()
    cfg2 = tiny_gpt_config(ov + [f"Engine.save_load.ckpt_dir={dst}"], nranks=world)
    env.set_seed(cfg2.Global.seed)
    module2 = build_module(cfg2)
    eng2 = EagerEngine(configs=cfg2, module=module2)
    eng2.load()
    resumed = [float(eng2.train_step(b)) for b in batches[2:]]
    assert max(abs(a - b) for a, b in zip(cont, resumed)) < 2e-4, (cont, resumed)


def moe_exp_ep_matches_single(rank, world):
    """GShard-style MoE: experts sharded over ep=world (live all-to-all) reproduce the ep=1 layer when every rank feeds the same
    tokens (outputs and expert gradients of the locally-owned experts)."""
    from paddlefleetx_b200.models.language_model.moe_exp import MoE
    from paddlefleetx_b200.parallel.topology import HybridCommunicateGroup

    hcg = HybridCommunicateGroup(dp=world)
    grp = hcg.get_moe_group()
    torch.manual_seed(3)
    expert = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.GELU(), torch.nn.Linear(16, 8))
    E = 2 * world
    full = MoE(8, expert, num_experts=E, ep_size=1, k=1, capacity_factor=4.0, use_rts=False)
    for e in full.fleetx_moe.experts.experts:
        for p in e.parameters():
            torch.nn.init.normal_(p, std=0.3)
    sharded = MoE(8, expert, num_experts=E, ep_size=world, k=1, capacity_factor=4.0, use_rts=False, ep_group=grp)
    sharded.gate.load_state_dict(full.gate.state_dict())
    for i, e in enumerate(sharded.fleetx_moe.experts.experts):
        e.load_state_dict(full.fleetx_moe.experts.experts[rank * 2 + i].state_dict())
    x = torch.randn(24, 8)
    y_full, _, _ = full(x)
    y_sh, _, _ = sharded(x)
    assert torch.allclose(y_full, y_sh, atol=1e-5), (y_full - y_sh).abs().max()
    y_full.pow(2).sum().backward()
    y_sh.pow(2).sum().backward()
    for i, e in enumerate(sharded.fleetx_moe.experts.experts):
        ref = full.fleetx_moe.experts.experts[rank * 2 + i]
        for p, q in zip(e.parameters(), ref.parameters()):
            # every rank contributed the same tokens -> the owner sees `world` copies of each routed token
            assert torch.allclose(p.grad, q.grad * world, atol=1e-4), (p.grad - q.grad * world).abs().max()


# ------------------------------------------------------------------------------------------------ protein folding: DAP / BP
def _fold_model_and_batch(seed=0, outer="origin"):
    from paddlefleetx_b200.models.protein_folding import EmbeddingsAndEvoformer

    torch.manual_seed(seed)
    b, S, R, T, E = 1, 4, 8, 2, 6
    evo = EmbeddingsAndEvoformer(msa_feat_dim=9, target_feat_dim=6, c_m=16, c_z=8, c_s=12, num_blocks=2, max_relative_feature=4, msa_heads=2,
                                 pair_heads=2, extra_msa_channel=8, extra_msa_blocks=1, outer_product_mean_position=outer,
                                 template=dict(enabled=True, c_t=8, num_block=1, num_head=2, attn_key_dim=8, embed_torsion_angles=True,
                                               use_template_unit_vector=True)).double()
    for p in evo.parameters():            # zero-initialised output projections would hide layout bugs behind zeros
        if p.abs().sum() == 0:
            torch.nn.init.normal_(p, std=0.1)
    evo.eval()                            # no dropout: outputs must agree exactly across layouts
    g = torch.Generator().manual_seed(seed + 1)
    r = lambda *s: torch.randn(*s, generator=g, dtype=torch.float64)  # noqa: E731
    batch = dict(target_feat=r(b, R, 6), msa_feat=r(b, S, R, 9), residue_index=torch.arange(R)[None], aatype=torch.randint(0, 20, (b, R), generator=g),
                 msa_mask=(torch.rand(b, S, R, generator=g) > 0.1).double(), seq_mask=torch.ones(b, R, dtype=torch.float64),
                 extra_msa=torch.randint(0, 23, (b, E, R), generator=g), extra_has_deletion=torch.rand(b, E, R, generator=g).double(),
                 extra_deletion_value=torch.rand(b, E, R, generator=g).double(), extra_msa_mask=(torch.rand(b, E, R, generator=g) > 0.1).double(),
                 template_mask=torch.ones(b, T, dtype=torch.float64), template_aatype=torch.randint(0, 20, (b, T, R), generator=g),
                 template_pseudo_beta=r(b, T, R, 3) * 5, template_pseudo_beta_mask=torch.ones(b, T, R, dtype=torch.float64),
                 template_all_atom_positions=r(b, T, R, 37, 3) * 3, template_all_atom_masks=torch.ones(b, T, R, 37, dtype=torch.float64))
    prev = dict(prev_pos=r(b, R, 37, 3), prev_msa_first_row=r(b, R, 16), prev_pair=r(b, R, R, 8))
    return evo, batch, prev


def _fold_loss(out):
    return (out["single"] ** 2).sum() + (out["pair"] ** 2).sum() + (out["msa"] ** 2).sum()


def evoformer_parallel_matches_single(rank, world, mode):
    """DAP (activations sharded over 2 ranks) or BP (MSA branch on rank 0, pair branch on rank 1) vs the unsharded model: same outputs,
    same parameter gradients after the group's gradient synchronisation."""
    from paddlefleetx_b200.distributed.protein_folding.scg import scg

    evo, batch, prev = _fold_model_and_batch(outer="origin" if mode == "dap" else "end")
    ref_out = evo(batch, prev)                      # scg not initialised yet: every collective is the identity
    _fold_loss(ref_out).backward()
    ref_grads = {n: p.grad.clone() for n, p in evo.named_parameters()}
    evo.zero_grad()
    scg.init_process_group([("dp", None), ("dap", 2 if mode == "dap" else 1), ("bp", 2 if mode == "bp" else 1)])
    assert (scg.get_dap_world_size(), scg.get_bp_world_size()) == ((2, 1) if mode == "dap" else (1, 2))
    out = evo(batch, prev)
    for k in ("single", "pair", "msa"):
        torch.testing.assert_close(out[k], ref_out[k], rtol=1e-8, atol=1e-9, msg=lambda m, k=k: f"{mode} {k}: {m}")
    _fold_loss(out).backward()
    evo.sync_gradients()
    for n, p in evo.named_parameters():
        torch.testing.assert_close(p.grad, ref_grads[n], rtol=1e-6, atol=1e-8, msg=lambda m, n=n: f"{mode} grad {n}: {m}")


def auto_inference_weights_roundtrip(rank, world, tmpdir):
    """mp2 training model -> ``save_for_auto_inference`` -> (a) merged full tensors equal the all-gathered shards, (b) a fresh mp2 model and
    (c) a single-process (mp1) model load them and produce the same logits."""
    import os

    from paddlefleetx_b200.distributed.apis import env, io
    from paddlefleetx_b200.models import build_module
    from paddlefleetx_b200.parallel.topology import HybridCommunicateGroup

    ov = ["Global.global_batch_size=None", "Global.local_batch_size=2", "Global.micro_batch_size=2", f"Distributed.mp_degree={world}"]
    cfg = tiny_gpt_config(ov, nranks=world)
    env.init_dist_env(cfg)
    env.set_seed(cfg.Global.seed)
    model = build_module(cfg).model
    prefix = os.path.join(tmpdir, "auto_infer", "auto")
    io.save_for_auto_inference(prefix, model)
    dist.barrier()
    full = io.merge_auto_inference(prefix)
    mp_group = env.get_hcg().get_model_parallel_group().process_group
    for name, p in model.named_parameters():
        if getattr(p, "tp_sharded", False):
            parts = [torch.empty_like(p.data) for _ in range(world)]
            dist.all_gather(parts, p.data.contiguous(), group=mp_group)
            assert torch.equal(full[name], torch.cat(parts, dim=p.split_axis)), name
        else:
            assert torch.equal(full[name], p.data), name
    batch = synthetic_batches(cfg, 1, seed=3)[0]
    model.eval()
    with torch.no_grad():
        ref = model(batch[0], batch[1])
    env.set_seed(cfg.Global.seed + 7)                 # different initial weights, same layout
    again = build_module(cfg).model.eval()
    io.load_auto_inference(prefix, again)
    with torch.no_grad():
        torch.testing.assert_close(again(batch[0], batch[1]), ref)
    # single-process layout: rebuild the model as if world == 1 and load the same files
    hcg = env.get_hcg()
    real_ws = env.world_size
    env.set_hcg(HybridCommunicateGroup(world_size=1, rank=0, build_groups=False))
    env.world_size = lambda: 1
    try:
        single = build_module(tiny_gpt_config(["Global.global_batch_size=None", "Global.local_batch_size=2", "Global.micro_batch_size=2"], nranks=1)).model.eval()
        io.load_auto_inference(prefix, single)
        with torch.no_grad():
            logits = single(batch[0], batch[1])
    finally:
        env.world_size = real_ws
        env.set_hcg(hcg)
    # the mp model returns vocabulary-parallel logits: compare this rank's slice
    v = ref.shape[-1]
    torch.testing.assert_close(logits[..., rank * v:(rank + 1) * v], ref, rtol=1e-4, atol=1e-5)


def ernie_tp_matches_single(rank, world, sequence_parallel):
    """ERNIE encoder under tensor parallelism (optionally with sequence parallelism): same MLM / SOP scores, same parameter gradients
    (after the sequence-parallel gradient all-reduce) as the single-process model."""
    from paddlefleetx_b200.distributed.apis import env
    from paddlefleetx_b200.models.language_model.ernie import model as E
    from paddlefleetx_b200.parallel.tp_layers import allreduce_sequence_parallel_grads, register_sequence_parallel_allreduce_hooks

    mp = world
    cfg = tiny_gpt_config(["Global.global_batch_size=None", "Global.local_batch_size=2", "Global.micro_batch_size=2", f"Distributed.mp_degree={mp}"], nranks=world)
    env.init_dist_env(cfg)
    hcg = env.get_hcg()
    group = hcg.get_model_parallel_group()
    kw = dict(vocab_size=256, hidden_size=32, num_hidden_layers=2, num_attention_heads=4, ffn_hidden_size=64, max_position_embeddings=32,
              hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0, use_flash_attn=False, dtype=torch.float64)
    torch.manual_seed(5)
    single = E.ErnieForPretraining(E.ErnieModel(**kw), 256).double()
    init = {k: v.detach().clone() for k, v in single.state_dict().items()}
    par = E.ErnieForPretraining(E.ErnieModel(mp_group=group, sequence_parallel=sequence_parallel, **kw), 256).double()
    assert par.ernie.sequence_parallel == bool(sequence_parallel)
    with torch.no_grad():
        for k, p in par.named_parameters():
            p.copy_(_shard_like(init[k], p, hcg.get_model_parallel_rank(), mp))
    if sequence_parallel:
        sp_params = register_sequence_parallel_allreduce_hooks(par, 1, False, group)
        assert len(sp_params) == 2 * 2 * 2 + 2 * 2          # two LayerNorms (w, b) per layer + the two row-linear biases per layer
    g = torch.Generator().manual_seed(9)
    ids = torch.randint(1, 256, (2, 16), generator=g)
    seg = torch.randint(0, 2, (2, 16), generator=g)
    masked = torch.tensor([1, 5, 9, 17, 20, 30])
    labels = torch.randint(0, 256, (6,), generator=g)
    single.train(); par.train()

    def loss_of(model, vocab_parallel):
        scores, rel = model(ids, seg, None, None, masked)
        if vocab_parallel:
            v = scores.shape[-1]
            lo = hcg.get_model_parallel_rank() * v
            # squared error against a one-hot target restricted to this rank's vocabulary slice: sums to the full-vocabulary loss over ranks
            tgt = torch.zeros_like(scores)
            for i, l in enumerate(labels.tolist()):
                if lo <= l < lo + v:
                    tgt[i, l - lo] = 1.0
            part = ((scores - tgt) ** 2).sum()
            dist.all_reduce(part, group=group.process_group)        # value only; each rank back-propagates its own slice
            # the sentence-order branch is replicated: every rank back-propagates it in full (the TP layers' own backward collectives
            # combine the per-rank partial input gradients), exactly like a replicated scalar loss in training
            return scores, rel, ((scores - tgt) ** 2).sum() + (rel ** 2).sum(), part + (rel ** 2).sum().detach()
        tgt = torch.nn.functional.one_hot(labels, scores.shape[-1]).to(scores.dtype)
        val = ((scores - tgt) ** 2).sum() + (rel ** 2).sum()
        return scores, rel, val, val.detach()

    s1, r1, l1, v1 = loss_of(single, False)
    l1.backward()
    s2, r2, l2, v2 = loss_of(par, True)
    l2.backward()
    if sequence_parallel:
        allreduce_sequence_parallel_grads(par)
    v = s2.shape[-1]
    torch.testing.assert_close(s2, s1[:, hcg.get_model_parallel_rank() * v:(hcg.get_model_parallel_rank() + 1) * v], rtol=1e-8, atol=1e-9)
    torch.testing.assert_close(r2, r1, rtol=1e-8, atol=1e-9)
    torch.testing.assert_close(v2, v1, rtol=1e-8, atol=1e-9)
    ref_grads = {k: p.grad for k, p in single.named_parameters()}
    for k, p in par.named_parameters():
        want = _shard_like(ref_grads[k], p, hcg.get_model_parallel_rank(), mp)
        torch.testing.assert_close(p.grad, want, rtol=1e-6, atol=1e-8, msg=lambda m, k=k: f"{k}: {m}")


def resume_matches_uninterrupted(rank, world, layout):
    """3 steps -> save -> 2 more steps, against: fresh engine -> load -> the same 2 steps.  Bit-identical losses for data parallel, ZeRO-2 and
    ZeRO-3 (whose resident unit — embeddings, final norm — must come back from the checkpoint like every other unit)."""
    import tempfile

    from paddlefleetx_b200.core import EagerEngine
    from paddlefleetx_b200.distributed.apis import env
    from paddlefleetx_b200.models import build_module

    out = [tempfile.mkdtemp() if rank == 0 else None]
    dist.broadcast_object_list(out, src=0)
    lay = {"dp2": ["Distributed.dp_degree=2"], "zero2": ["Distributed.sharding.sharding_degree=2", "Distributed.sharding.sharding_stage=2"],
           "zero3": ["Distributed.sharding.sharding_degree=2", "Distributed.sharding.sharding_stage=3"]}[layout]
    ov = ["Global.global_batch_size=None", "Global.local_batch_size=2", "Global.micro_batch_size=2", f"Engine.save_load.output_dir={out[0]}"] + lay
    cfg = tiny_gpt_config(ov, nranks=world)
    env.init_dist_env(cfg)
    env.set_seed(cfg.Global.seed)
    eng = EagerEngine(configs=cfg, module=build_module(cfg))
    batches = [_slice(b, rank, world) for b in synthetic_batches(cfg, 5, seed=3)]
    for b in batches[:3]:
        eng.train_step(b)
    eng.save(epoch=0, step=3)
    dist.barrier()
    cont = [float(eng.train_step(b)) for b in batches[3:]]
    cfg2 = tiny_gpt_config(ov + [f"Engine.save_load.ckpt_dir={out[0]}/epoch_0_step_3"], nranks=world)
    env.set_seed(cfg2.Global.seed + 17)                      # different initial weights: everything must come from the checkpoint
    eng2 = EagerEngine(configs=cfg2, module=build_module(cfg2))
    eng2.load()
    resumed = [float(eng2.train_step(b)) for b in batches[3:]]
    assert cont == resumed, (layout, cont, resumed)


def barrier_skew_detected(rank, world):
    """PFX_DEBUG_POISON: equal barrier counts pass, a rank that issued one barrier more on a channel is reported on every rank."""
    import os

    from paddlefleetx_b200.parallel import debug_poison as D

    os.environ["PFX_DEBUG_POISON"] = "1"
    D.barrier_skew_check({0: 5, 2: 1})
    try:
        D.barrier_skew_check({0: 5, 2: 1 + (rank == 1)})
    except RuntimeError as e:
        assert "channel 2" in str(e) and "[1, 2]" in str(e), str(e)
    else:
        raise AssertionError("skew not detected")


def _mask_out_a_ragged_tail(batch):
    """Loss mask with a different number of live positions in every quarter of the sequence (and per sample)."""
    tokens, pos, labels, mask = batch
    mask = mask.clone()
    s = mask.shape[1]
    for i in range(mask.shape[0]):
        mask[i, s - 1 - (i % 3) - s // 4:] = 0          # the tail (mostly the last cp shard) is dead
        mask[i, 1:3 + i % 2] = 0
    return tokens, pos, labels, mask


def context_parallel_matches_single(rank, world, dp, sharding, cp, extra=(), ragged_mask=False):
    """Ulysses context parallelism (Distributed.cp_degree): the ranks of a cp group take the same batch, each a slice of the sequence; losses
    and weights must equal the single-process run.  Heads 4, seq 16 in the tiny recipe: cp 2 leaves 2 heads x 16 positions per rank."""
    data = dp * sharding // cp
    gb = 4
    _, batches, ref_losses, ref_state, init = _reference_losses_and_state(
        ["Global.global_batch_size=None", f"Global.local_batch_size={gb}", f"Global.micro_batch_size={gb}"] + list(extra), 4,
        batch_hook=_mask_out_a_ragged_tail if ragged_mask else None)
    cfg = tiny_gpt_config([f"Global.global_batch_size={gb}", "Global.local_batch_size=None", f"Global.micro_batch_size={gb // data}",
                           f"Distributed.dp_degree={dp}", f"Distributed.sharding.sharding_degree={sharding}", "Distributed.sharding.sharding_stage=1",
                           f"Distributed.cp_degree={cp}"] + list(extra), nranks=world)
    assert cfg.Global.local_batch_size == gb // data
    from paddlefleetx_b200.core import EagerEngine
    from paddlefleetx_b200.distributed.apis import env
    from paddlefleetx_b200.models import build_module

    hcg = env.init_dist_env(cfg)
    assert hcg.get_context_parallel_world_size() == cp and len(hcg.get_context_parallel_group().ranks) == cp
    env.set_seed(cfg.Global.seed)
    module = build_module(cfg)
    module.model.load_state_dict(init)
    eng = EagerEngine(configs=cfg, module=module)
    dr, dw = env.get_data_world_rank(), env.get_data_world_size()
    assert dw == data
    losses = []
    for b in batches:
        l = eng.train_step(_slice(b, dr, dw)).detach().clone()        # the module keeps this rank's slice of the sequence
        if data == 1:
            mine = float(l)                                           # every rank of a cp group reports the loss of the whole sequences
        dist.all_reduce(l)
        losses.append(float(l) / world)
        assert data > 1 or abs(mine - losses[-1]) < 1e-5, (mine, losses[-1])
    assert max(abs(a - b) for a, b in zip(losses, ref_losses)) < 2e-4, (losses, ref_losses)
    heads = cfg.Model.num_attention_heads
    for k, v in eng.module.model.state_dict().items():
        ref = ref_state[k]
        if k.endswith("qkv_proj.bias"):
            # softmax is invariant to a shift of the keys: the K-bias gradient is exactly zero in theory and rounding noise in practice, which
            # Adam normalises to full-size steps — those entries follow the summation order of the attention implementation, not the layout
            v, ref = v.view(heads, 3, -1), ref.view(heads, 3, -1)
            assert torch.allclose(v[:, 1], ref[:, 1], atol=1e-3), k
            v, ref = v[:, [0, 2]], ref[:, [0, 2]]
        assert torch.allclose(v, ref, atol=5e-5, rtol=1e-4), (k, float((v - ref).abs().max()))


def dap_split_phase_pairs(rank, world):
    """Reference call style of the DAP collectives: ``y = all_gather(x, axis)`` ... ``z = all_gather_opp(y, axis)`` (and the all-to-all pair),
    in blocking and asynchronous mode, equal the one-shot ops in value and gradient; the class forms follow the reference signatures."""
    from paddlefleetx_b200.distributed.protein_folding import dap
    from paddlefleetx_b200.distributed.protein_folding.scg import scg

    scg.init_process_group([("dp", None), ("dap", world)])
    assert dap.get_world_size() == world and dap.get_rank_in_group() == rank
    torch.manual_seed(10 + rank)
    base = torch.randn(2, 4, 6, 3, dtype=torch.float64)
    w = torch.randn(2, 4 * world, 6, 3, dtype=torch.float64)          # rank-specific weight: every rank's loss differs
    for sync in (True, False):
        dap.set_dap_sync_op(sync)
        assert dap.get_dap_sync_op() is sync
        for axis in (0, 1, 2):
            shape = list(base.shape)
            shape[axis] *= world
            wa = torch.randn(shape, dtype=torch.float64, generator=torch.Generator().manual_seed(rank * 7 + axis))
            x1 = base.clone().requires_grad_(True)
            y = dap.all_gather(x1, axis=axis)
            assert y.shape[0] == base.shape[0] * world                 # the rank-major stack, whatever the axis
            side = (x1 * 2).sum()                                      # independent work between the halves
            z = dap.all_gather_opp(y, axis=axis)
            ((z * wa).sum() + side).backward()
            x2 = base.clone().requires_grad_(True)
            z2 = dap.gather_full(x2, axis)
            ((z2 * wa).sum() + (x2 * 2).sum()).backward()
            torch.testing.assert_close(z, z2)
            torch.testing.assert_close(x1.grad, x2.grad)
        # all-to-all pair: rows sharded -> columns sharded
        x1 = base.clone().requires_grad_(True)
        y = dap.all_to_all(x1, in_axis=2, out_axis=1)
        z = dap.all_to_all_opp(y, in_axis=2, out_axis=1)
        wz = torch.randn(z.shape, dtype=torch.float64, generator=torch.Generator().manual_seed(rank + 99))
        (z * wz).sum().backward()
        x2 = base.clone().requires_grad_(True)
        z2 = dap.row_to_col(x2)
        (z2 * wz).sum().backward()
        torch.testing.assert_close(z, z2)
        torch.testing.assert_close(x1.grad, x2.grad)
        # without gradients the pair is pure data movement
        with torch.no_grad():
            torch.testing.assert_close(dap.all_gather_opp(dap.all_gather(base, axis=1), axis=1), dap.gather_full(base, 1))
    dap.set_dap_sync_op(True)
    # class forms (reference signatures: no group argument, the dap group is implied)
    x = base.clone().requires_grad_(True)
    full = dap.Gather.apply(x, 1)
    assert full.shape[1] == base.shape[1] * world
    back = dap.Scatter.apply(full, 1)
    torch.testing.assert_close(back, x)
    (back * w[:, :4]).sum().backward()
    torch.testing.assert_close(x.grad, w[:, :4])
    stack = torch.cat((base * (rank + 1)).chunk(world, dim=2), dim=0)
    out = dap.All2All.apply(stack, 2, 1)
    torch.testing.assert_close(torch.cat(out.chunk(world, 0), dim=1), dap.row_to_col(base * (rank + 1)))
    # reference-style optimizer groups in grad_sync
    p_rep, p_skip = torch.nn.Parameter(torch.ones(3)), torch.nn.Parameter(torch.ones(3))
    p_rep.grad, p_skip.grad = torch.full((3,), float(rank + 1)), torch.full((3,), float(rank + 1))
    dap.grad_sync([{"params": [p_rep], "dap": True}, {"params": [p_skip]}])
    assert float(p_rep.grad[0]) == sum(range(1, world + 1)) and float(p_skip.grad[0]) == rank + 1


def ring_attention_matches_full(rank, world):
    """parallel/ring_attention.py: zigzag shards + K / V blocks on the ring reproduce attention over the whole sequence — outputs and the
    gradients of q, k and v — causal and unmasked, and with the kernels' counter-hash dropout the forward / backward pair stays consistent
    (the gradient of a linear functional equals its finite difference)."""
    from paddlefleetx_b200.parallel import ring_attention as R
    from paddlefleetx_b200.parallel.topology import HybridCommunicateGroup

    hcg = HybridCommunicateGroup(dp=world, cp=world, cp_mode="ring")
    group = hcg.get_context_parallel_group()
    assert group.nranks == world and group.rank == rank
    b, s, h, d = 2, 8 * world, 3, 4                                  # 3 heads on 2 / 4 ranks: no Ulysses split exists for this shape
    gen = torch.Generator().manual_seed(5)
    full = [torch.randn(b, s, h, d, dtype=torch.float64, generator=gen) for _ in range(3)]
    w = torch.randn(b, s, h, d, dtype=torch.float64, generator=gen)
    assert torch.equal(R.zigzag_merge([R.zigzag_slice(w, world, r) for r in range(world)]), w)
    for causal in (True, False):
        ref_in = [t.clone().requires_grad_(True) for t in full]
        qt, kt, vt = (t.transpose(1, 2) for t in ref_in)
        ref = torch.nn.functional.scaled_dot_product_attention(qt, kt, vt, is_causal=causal).transpose(1, 2)
        (ref * w).sum().backward()
        loc = [R.zigzag_slice(t, world, rank).clone().requires_grad_(True) for t in full]
        out = R.ring_attention(*loc, group, causal=causal)
        torch.testing.assert_close(out, R.zigzag_slice(ref.detach(), world, rank), atol=1e-10, rtol=1e-8)
        (out * R.zigzag_slice(w, world, rank)).sum().backward()
        for a, r_ in zip(loc, ref_in):
            torch.testing.assert_close(a.grad, R.zigzag_slice(r_.grad, world, rank), atol=1e-10, rtol=1e-8)
    # dropout: same pattern in forward and backward of every block -> directional derivative matches
    from paddlefleetx_b200.parallel.rng import get_rng_state_tracker

    tr = get_rng_state_tracker()
    if not tr.has("local_seed"):
        tr.add("local_seed", 1234 + rank)
    loc = [R.zigzag_slice(t, world, rank).clone().requires_grad_(True) for t in full]
    wl = R.zigzag_slice(w, world, rank)

    def run(q, k, v):
        state = tr.get_states_tracker()
        with tr.rng_state("local_seed"):
            o = R.ring_attention(q, k, v, group, causal=True, dropout_p=0.25)
        tr.set_states_tracker(state)                                 # replay the same dropout pattern on the next call
        return (o * wl).sum()

    run(*loc).backward()
    dirs = [torch.randn(t.shape, dtype=torch.float64, generator=gen) for t in loc]
    # a perturbation of THIS rank's shards changes every rank's loss: sum the losses over the group to differentiate the same function
    eps = 1e-6
    total = []
    for sign in (1, -1):
        l = run(*(t.detach() + sign * eps * dd for t, dd in zip(loc, dirs))).detach().clone()
        dist.all_reduce(l)
        total.append(float(l))
    fd = (total[0] - total[1]) / (2 * eps)
    an = torch.stack([(t.grad * dd).sum() for t, dd in zip(loc, dirs)]).sum()
    dist.all_reduce(an)
    assert abs(fd - float(an)) < 1e-5 * max(1.0, abs(fd)), (fd, float(an))


def context_parallel_with_tp_matches_single(rank, world, mp, cp, sequence_parallel, mode):
    """cp x mp on one job: tensor parallel ranks (with or without Megatron sequence parallelism, which shards the cp-local sequence once more)
    inside context-parallel groups; weights are the single-process model's, sharded per mp rank."""
    gb = 2
    single_ov = ["Global.global_batch_size=None", f"Global.local_batch_size={gb}", f"Global.micro_batch_size={gb}"]
    _, batches, ref_losses, ref_state, init = _reference_losses_and_state(single_ov, 3, seed=21, batch_hook=_mask_out_a_ragged_tail)
    from paddlefleetx_b200.core import EagerEngine
    from paddlefleetx_b200.distributed.apis import env
    from paddlefleetx_b200.models import build_module

    data = world // mp
    assert data == cp, "one replica: the whole data axis is one context-parallel group"
    cfg = tiny_gpt_config(single_ov + [f"Distributed.mp_degree={mp}", f"Model.sequence_parallel={sequence_parallel}", f"Distributed.sharding.sharding_degree={data}",
                                       "Distributed.sharding.sharding_stage=1", f"Distributed.cp_degree={cp}", f"Distributed.cp_mode={mode}"], nranks=world)
    hcg = env.init_dist_env(cfg)
    env.set_seed(cfg.Global.seed)
    module = build_module(cfg)
    with torch.no_grad():
        for k, p in module.model.named_parameters():
            p.copy_(_shard_like(init[k], p, hcg.get_model_parallel_rank(), mp))
    eng = EagerEngine(configs=cfg, module=module)
    losses = []
    for b in batches:
        l = eng.train_step(b).detach().clone()
        dist.all_reduce(l)
        losses.append(float(l) / world)
    assert max(abs(a - b) for a, b in zip(losses, ref_losses)) < 3e-4, (losses, ref_losses)
    heads = cfg.Model.num_attention_heads // mp
    for k, p in module.model.named_parameters():
        want, got = _shard_like(ref_state[k], p, hcg.get_model_parallel_rank(), mp), p.detach()
        if k.endswith("qkv_proj.bias"):          # K-bias entries: zero gradient in theory, Adam-normalised rounding noise in practice
            want, got = want.view(heads, 3, -1)[:, [0, 2]], got.view(heads, 3, -1)[:, [0, 2]]
        # Adam's first steps move an entry by ~lr whatever the size of its gradient: an entry whose gradient is rounding noise can land a few
        # 1e-4 away when the summation order changes (4 ranks' partial sums vs one) — tolerate isolated entries, not a pattern
        bad = ((got - want).abs() > 1e-4 + 1e-3 * want.abs())
        assert int(bad.sum()) <= max(1, bad.numel() // 1000) and float((got - want).abs().max()) < 1e-3, \
            (k, float((got - want).abs().max()), int(bad.sum()), bad.numel())
