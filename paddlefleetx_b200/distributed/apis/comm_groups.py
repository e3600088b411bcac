"""``create_hcg`` — build the hybrid communicator set from a strategy.

Reference: ppfleetx/distributed/apis/comm_groups.py:27-153 offers two classes selected by the YAML
key ``Distributed.hcg``: Paddle's ``HybridCommunicateGroup`` and an in-tree
``HybridCommGroupForMoE`` that adds the fused dp x mp ``moe`` group.  Here both are one
implementation (parallel/topology.py) — the fused groups always exist.
"""
from __future__ import annotations

import torch.distributed as dist

from ...parallel.topology import HybridCommGroupForMoE, HybridCommunicateGroup, _Group


class MoEGroup:
    """Expert-parallel communicator from an explicit partition of the ranks, e.g. ``[[0, 1, 2, 3], [4, 5, 6, 7]]``: every rank calls this
    with the same lists (communicator creation is collective) and keeps the group it belongs to (reference comm_groups.py:38-52)."""

    def __init__(self, list_of_ranks):
        self.list_of_ranks = [list(r) for r in list_of_ranks]
        initialised = dist.is_available() and dist.is_initialized()
        self._rank = dist.get_rank() if initialised else 0
        mine = [r for r in self.list_of_ranks if self._rank in r]
        assert len(mine) == 1, f"Rank {self._rank} belongs to {'multi' if mine else 'no'} moe groups"
        self.group = None
        for ranks in self.list_of_ranks:
            pg = dist.new_group(ranks) if (initialised and len(ranks) > 1) else None
            if ranks is mine[0]:
                self.group = _Group(ranks, pg, self._rank)

    @property
    def world_size(self) -> int:
        return self.group.nranks

    @property
    def ranks(self):
        return self.group.ranks

    @property
    def rank_in_group(self) -> int:
        return self.group.rank


class Hybrid4DCommGroup(HybridCommunicateGroup):
    """The reference's orthogonal-strategy constructor form (comm_groups.py:55-122): ``Hybrid4DCommGroup([("dp", 2, _), ("mp", 4, _), ...],
    {"moe": ["dp", "mp"]})`` — a list of ``(axis, degree, group class)`` triples (the class is ignored: one communicator type serves every
    axis here) plus named fused groups spanning several axes.  Degrees may also come as keywords, like the base class."""

    def __init__(self, list_of_strategy=None, fused_strategy_dict=None, **degrees):
        if list_of_strategy is not None and hasattr(list_of_strategy, "hybrid_configs"):      # a DistributedStrategy
            hc = list_of_strategy.hybrid_configs
            degrees = dict(dp=hc.get("dp_degree", 1), mp=hc.get("mp_degree", 1), pp=hc.get("pp_degree", 1), sharding=hc.get("sharding_degree", 1),
                           **degrees)
        elif list_of_strategy is not None:
            for name, degree, *_ in list_of_strategy:
                degrees[name] = int(degree)
        super().__init__(**degrees)
        can_build = dist.is_available() and dist.is_initialized() and self.nranks > 1
        self._fused = {"check": self._groups["check"], "moe": self._groups["moe"]}
        for name, axes in (fused_strategy_dict or {}).items():
            if name not in self._fused:
                self._fused[name] = self._make(tuple(axes), can_build)

    def strategy_group(self, name: str):
        return self._groups[name]

    def fused_strategy_group(self, name: str):
        return self._fused[name]

    def rank_in_strategy(self, name: str) -> int:
        return self._coord[name]


_REGISTRY = {
    "HybridCommunicateGroup": HybridCommunicateGroup,
    "HybridCommGroupForMoE": HybridCommGroupForMoE,
    "Hybrid4DCommGroup": Hybrid4DCommGroup,
}


def create_hcg(strategy, hcg_name: str = "HybridCommunicateGroup"):
    if hcg_name not in _REGISTRY:
        raise ValueError(f"unknown hcg {hcg_name}; choose from {sorted(_REGISTRY)}")
    hc = strategy.hybrid_configs
    return _REGISTRY[hcg_name](dp=hc.get("dp_degree", 1), mp=hc.get("mp_degree", 1),
                               pp=hc.get("pp_degree", 1), sharding=hc.get("sharding_degree", 1), cp=hc.get("cp_degree", 1), cp_mode=hc.get("cp_mode", "ulysses"))
