"""Multi-process correctness on CPU/gloo: layouts must reproduce the single-process loss curve and weights."""
import pytest

from dist_utils import run_distributed


def test_dp2_sharding_stage1_matches_single():           # BASELINE.json config #1 layout (plumbing on gloo)
    run_distributed("dist_fns:dp_sharding_matches_single", 2, 2, 1, 1)


def test_sharding2_stage1_matches_single():
    run_distributed("dist_fns:dp_sharding_matches_single", 2, 1, 2, 1)


def test_sharding2_stage2_matches_single():
    run_distributed("dist_fns:dp_sharding_matches_single", 2, 1, 2, 2)


def test_sharding2_stage2_direct_grad_overlap_matches_single():
    # direct gradient writes (no autograd-owned flat views, lazy zero) + overlapped bucket reduction
    run_distributed("dist_fns:dp_sharding_matches_single", 2, 1, 2, 2,
                    ("Optimizer.direct_grad=True", "Distributed.sharding.reduce_overlap=True"))


@pytest.mark.parametrize("micro", [4, 2])
def test_sharding2_stage2_gradient_ring_matches_single(micro):
    run_distributed("dist_fns:zero2_ring_matches_single", 2, micro)


def test_dp2_x_sharding2_matches_single():
    run_distributed("dist_fns:dp_sharding_matches_single", 4, 2, 2, 1)


def test_tensor_parallel_matches_single():
    run_distributed("dist_fns:tp_matches_single", 2, False)


def test_sequence_parallel_matches_single():
    run_distributed("dist_fns:tp_matches_single", 2, True)


def test_pipeline_1f1b_pp2_matches_single():
    run_distributed("dist_fns:pipeline_matches_single", 2, 2, 1, 1, 4)


def test_pipeline_pp2_mp2_matches_single():
    run_distributed("dist_fns:pipeline_matches_single", 4, 2, 2, 1, 4)


def test_pipeline_pp2_mp2_sequence_parallel_matches_single():
    run_distributed("dist_fns:pipeline_matches_single", 4, 2, 2, 1, 4, True)


def test_tensor_parallel_init_differs_across_mp_ranks():
    run_distributed("dist_fns:tp_shards_differ_across_ranks", 2)


def test_pipeline_interleaved_pp2_vpp2_matches_single():
    run_distributed("dist_fns:pipeline_matches_single", 2, 2, 1, 2, 4)


@pytest.mark.parametrize("gate", ["naive", "gshard"])
def test_moe_expert_parallel_matches_single(gate):
    run_distributed("dist_fns:moe_ep_matches_single", 2, gate)


def test_moe_module_trains_with_expert_parallel():
    run_distributed("dist_fns:moe_module_trains", 2)


def test_sharding2_stage3_matches_single():
    run_distributed("dist_fns:dp_sharding_matches_single", 2, 1, 2, 3)


def test_stage3_with_recompute_and_dp():
    run_distributed("dist_fns:dp_sharding_matches_single", 4, 2, 2, 3)


def test_moe_exp_expert_parallel_all_to_all_matches_single():
    run_distributed("dist_fns:moe_exp_ep_matches_single", 2)


def test_auto_inference_weights_reload_on_another_tensor_parallel_degree(tmp_path):
    run_distributed("dist_fns:auto_inference_weights_roundtrip", 2, str(tmp_path))


@pytest.mark.parametrize("sp", [False, True])
def test_ernie_tensor_and_sequence_parallel_match_single(sp):
    run_distributed("dist_fns:ernie_tp_matches_single", 2, sp)


@pytest.mark.parametrize("layout", ["dp2", "zero2", "zero3"])
def test_resume_from_checkpoint_matches_uninterrupted_run(layout):
    run_distributed("dist_fns:resume_matches_uninterrupted", 2, layout)


def test_context_parallel_cp2_matches_single():          # beyond the reference: Ulysses sequence sharding inside the data-parallel ranks
    run_distributed("dist_fns:context_parallel_matches_single", 2, 2, 1, 2)


def test_context_parallel_cp2_with_zero_sharding_and_rope_matches_single():
    run_distributed("dist_fns:context_parallel_matches_single", 4, 1, 4, 2, ["Model.use_rope=True"])


@pytest.mark.parametrize("world", [2, 4])
def test_ring_attention_matches_full_attention(world):          # beyond the reference: zigzag ring attention (Distributed.cp_mode: ring)
    run_distributed("dist_fns:ring_attention_matches_full", world)


def test_context_parallel_ring_cp2_matches_single():
    run_distributed("dist_fns:context_parallel_matches_single", 2, 2, 1, 2, ["Distributed.cp_mode=ring"])


def test_context_parallel_ring_cp4_more_ranks_than_ulysses_allows_with_zero_sharding_and_rope():
    # 4 heads on mp 1 would allow Ulysses cp 4 too; 3 heads (hidden 48) do not: only the ring runs this layout
    run_distributed("dist_fns:context_parallel_matches_single", 4, 1, 4, 4,
                    ["Distributed.cp_mode=ring", "Model.use_rope=True", "Model.num_attention_heads=3", "Model.hidden_size=48", "Model.ffn_hidden_size=96"])


@pytest.mark.parametrize("mode", ["ulysses", "ring"])
def test_context_parallel_with_an_uneven_loss_mask_matches_single(mode):
    # one replica of 2 cp ranks whose sequence shards hold different numbers of live positions: the loss is still sum(CE x mask) / sum(mask)
    run_distributed("dist_fns:context_parallel_matches_single", 2, 2, 1, 2, [f"Distributed.cp_mode={mode}"], True)


@pytest.mark.parametrize("sp,mode", [(False, "ulysses"), (True, "ulysses"), (True, "ring")])
def test_context_parallel_inside_tensor_parallel_matches_single(sp, mode):
    run_distributed("dist_fns:context_parallel_with_tp_matches_single", 4, 2, 2, sp, mode)


@pytest.mark.parametrize("mode", ["ulysses", "ring"])
def test_context_parallel_under_the_pipeline_schedule_matches_single(mode):
    # pp2 x (sharding2 = cp2): sequence shards travel through the 1F1B schedule, the last stage normalises by the group's live-token count
    run_distributed("dist_fns:pipeline_matches_single", 4, 2, 1, 1, 4, False, 2, mode)


@pytest.mark.parametrize("mode", ["ulysses", "ring"])
def test_context_parallel_under_the_pipeline_schedule_with_rope_matches_single(mode):
    # stages > 0 receive activations only: the rotary positions of a sequence shard are rebuilt from the cp layout
    run_distributed("dist_fns:pipeline_matches_single", 4, 2, 1, 1, 4, False, 2, mode, ["Model.use_rope=True"])


@pytest.mark.parametrize("stage", [2, 3])
def test_context_parallel_ring_with_zero_stage_2_and_3_matches_single(stage):
    run_distributed("dist_fns:context_parallel_matches_single", 2, 1, 2, 2, ["Distributed.cp_mode=ring", f"Distributed.sharding.sharding_stage={stage}"], True)
