import itertools

import pytest

from paddlefleetx_b200.parallel.topology import CommunicateTopology, HybridCommunicateGroup, all_axis_products


@pytest.mark.parametrize("world", [1, 2, 4, 8])
def test_groups_partition_world(world):
    for dp, mp, pp, sd in all_axis_products(world):
        topo = CommunicateTopology({"dp": dp, "mp": mp, "pp": pp, "sharding": sd})
        assert topo.world_size() == world
        for axes in (("dp",), ("mp",), ("pp",), ("sharding",), ("mp", "pp"), ("dp", "mp")):
            groups = topo.groups_along(*axes)
            flat = sorted(itertools.chain.from_iterable(groups))
            assert flat == list(range(world)), (axes, groups)
        for r in range(world):
            assert topo.rank_of(**topo.coord_of(r)) == r


def test_mp_is_fastest_axis():
    topo = CommunicateTopology({"dp": 2, "mp": 2, "pp": 2, "sharding": 1})
    assert topo.groups_along("mp")[0] == [0, 1]
    assert topo.groups_along("pp")[0] == [0, 2]
    assert topo.groups_along("dp")[0] == [0, 4]


def test_hcg_accessors_without_process_group():
    h = HybridCommunicateGroup(dp=2, mp=2, pp=2, sharding=1, rank=5, world_size=8, build_groups=False)
    assert h.get_model_parallel_rank() == 1 and h.get_stage_id() == 0 and h.get_data_parallel_rank() == 1
    assert h.get_model_parallel_group().ranks == [4, 5]
    assert h.get_pipe_parallel_group().ranks == [5, 7]
    assert h.get_check_parallel_group().ranks == [4, 5, 6, 7]
    assert h.is_first_stage and not h.is_last_stage and h.next_rank == 7
    assert h.get_rank_from_stage(1) == 7
    assert h.get_parallel_mode() == "pipeline"


def test_context_parallel_groups_and_modes_without_process_group():
    """cp consecutive DATA ranks (dp_rank * sharding + sharding_rank) share a batch; the mode is validated where the topology is built and
    where the YAML is read."""
    from helpers import tiny_gpt_config

    for rank in range(8):
        h = HybridCommunicateGroup(dp=2, sharding=4, cp=4, cp_mode="ring", rank=rank, world_size=8, build_groups=False)
        ranks = h.get_context_parallel_group().ranks
        assert len(ranks) == 4 and rank in ranks and h.cp_mode == "ring" and h.get_context_parallel_rank() == ranks.index(rank)
    with pytest.raises(ValueError):
        HybridCommunicateGroup(dp=2, sharding=4, cp=3, rank=0, world_size=8, build_groups=False)
    with pytest.raises(ValueError):
        HybridCommunicateGroup(dp=8, cp=2, cp_mode="striped", rank=0, world_size=8, build_groups=False)
    cfg = tiny_gpt_config(["Distributed.dp_degree=4", "Distributed.cp_degree=2", "Distributed.cp_mode=Ring"], nranks=4)
    assert cfg.Distributed.cp_mode == "ring" and cfg.Distributed.cp_degree == 2
    assert tiny_gpt_config([], nranks=1).Distributed.cp_mode == "ulysses"
    with pytest.raises(AssertionError):
        tiny_gpt_config(["Distributed.dp_degree=4", "Distributed.cp_degree=2", "Distributed.cp_mode=blockwise"], nranks=4)
