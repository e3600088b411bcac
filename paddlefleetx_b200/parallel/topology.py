"""N-D orthogonal process-group topology on ``torch.distributed`` (NCCL on B200, gloo on CPU).

This supplies what the reference obtains from ``fleet.init`` /
``get_hybrid_communicate_group()`` (contract listed in SURVEY §2.5, call sites
ppfleetx/distributed/apis/comm_groups.py:27-153, env.py:41-55): one communicator per axis
(dp / mp / pp / sharding), fused groups (``check`` = mp x pp, ``moe`` = dp x mp), p2p
neighbour groups for the pipeline and the HCG accessor names the model zoo calls.

Rank order: ``mp`` varies fastest, then ``pp``, ``dp``, ``sharding`` — the order implied by the
reference's seed formula (env.py:74-84).  On an NVSwitch box all peers are equidistant, so the
order is a convention, not a locality optimisation.
"""
from __future__ import annotations

import itertools
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch.distributed as dist

AXES_FAST_TO_SLOW = ("mp", "pp", "dp", "sharding")


class CommunicateTopology:
    """Pure rank arithmetic (no communicators) so it is unit-testable without a process group."""

    def __init__(self, dims: Dict[str, int], order: Sequence[str] = AXES_FAST_TO_SLOW):
        self.order = tuple(order)  # fastest-varying first
        self.dims = {k: int(dims.get(k, 1)) for k in self.order}
        shape_slow_first = [self.dims[a] for a in reversed(self.order)]
        self._grid = np.arange(int(np.prod(shape_slow_first))).reshape(shape_slow_first)

    # -- basic queries ---------------------------------------------------------------
    def world_size(self) -> int:
        return int(self._grid.size)

    def get_dim_size(self, axis: str) -> int:
        return self.dims[_canon(axis)]

    def _axis_index(self, axis: str) -> int:
        return len(self.order) - 1 - self.order.index(_canon(axis))

    def coord_of(self, rank: int) -> Dict[str, int]:
        idx = np.unravel_index(rank, self._grid.shape)
        return {a: int(idx[self._axis_index(a)]) for a in self.order}

    def rank_of(self, **coord: int) -> int:
        idx = [0] * len(self.order)
        for a, v in coord.items():
            idx[self._axis_index(a)] = v
        return int(self._grid[tuple(idx)])

    def groups_along(self, *axes: str) -> List[List[int]]:
        """All rank lists obtained by varying ``axes`` and fixing every other axis."""
        axes_idx = sorted(self._axis_index(a) for a in axes)
        other = [i for i in range(self._grid.ndim) if i not in axes_idx]
        moved = np.transpose(self._grid, other + axes_idx)
        n = int(np.prod([self._grid.shape[i] for i in axes_idx])) if axes_idx else 1
        return [list(map(int, row)) for row in moved.reshape(-1, n)]

    def group_of(self, rank: int, *axes: str) -> List[int]:
        for g in self.groups_along(*axes):
            if rank in g:
                return g
        raise ValueError(rank)


def _canon(axis: str) -> str:
    return {"data": "dp", "model": "mp", "pipe": "pp", "sharding": "sharding"}.get(axis, axis)


class _Group:
    """A communicator plus its rank list; ``process_group`` is None for a size-1 axis."""

    def __init__(self, ranks: List[int], pg, my_rank: int):
        self.ranks = ranks
        self.process_group = pg
        self.nranks = len(ranks)
        self.rank = ranks.index(my_rank) if my_rank in ranks else -1

    @property
    def world_size(self) -> int:
        return self.nranks

    def __repr__(self) -> str:
        return f"Group(ranks={self.ranks}, rank={self.rank})"


class HybridCommunicateGroup:
    """4-D (+fused) communicator set.  Works with world_size==1 and without an initialised
    process group (all groups degenerate to size 1), which is what single-card runs use."""

    def __init__(self, dp: int = 1, mp: int = 1, pp: int = 1, sharding: int = 1,
                 rank: Optional[int] = None, world_size: Optional[int] = None, build_groups: bool = True, cp: int = 1, cp_mode: str = "ulysses"):
        initialised = dist.is_available() and dist.is_initialized()
        self.global_rank = rank if rank is not None else (dist.get_rank() if initialised else 0)
        self.nranks = world_size if world_size is not None else (dist.get_world_size() if initialised else 1)
        self._topo = CommunicateTopology({"dp": dp, "mp": mp, "pp": pp, "sharding": sharding})
        if self._topo.world_size() != self.nranks:
            raise ValueError(f"dp{dp} x mp{mp} x pp{pp} x sharding{sharding} != world {self.nranks}")
        self._coord = self._topo.coord_of(self.global_rank)
        self._groups: Dict[str, _Group] = {}
        can_build = build_groups and initialised and self.nranks > 1
        for name, axes in (("dp", ("dp",)), ("mp", ("mp",)), ("pp", ("pp",)), ("sharding", ("sharding",)),
                           ("check", ("mp", "pp")), ("moe", ("dp", "mp")),
                           ("dp_sharding", ("dp", "sharding"))):
            self._groups[name] = self._make(axes, can_build)
        # context parallelism (beyond the reference): ``cp`` consecutive DATA ranks (data index = dp_rank * sharding + sharding_rank) form a
        # group that works on one batch, each rank on its slice of the sequence; to the samplers the group is one data replica, to the
        # gradient reduction its members are ordinary data ranks (different tokens, same parameters)
        self.cp = int(cp)
        self.cp_mode = str(cp_mode or "ulysses").lower()      # "ulysses": heads <-> sequence all-to-all; "ring": zigzag shards, K / V blocks on a ring
        if self.cp_mode not in ("ulysses", "ring"):
            raise ValueError(f"cp_mode {cp_mode!r}: 'ulysses' or 'ring'")
        data = dp * sharding
        if self.cp < 1 or data % self.cp:
            raise ValueError(f"cp_degree {cp} must divide dp x sharding = {data}")
        self._groups["cp"] = _Group([self.global_rank], None, self.global_rank)
        if self.cp > 1:
            for ranks in self._topo.groups_along("dp", "sharding"):
                by_data = sorted(ranks, key=lambda r: self._topo.coord_of(r)["dp"] * sharding + self._topo.coord_of(r)["sharding"])
                for i in range(0, len(by_data), self.cp):
                    chunk = by_data[i:i + self.cp]
                    pg = dist.new_group(chunk) if can_build else None
                    if self.global_rank in chunk:
                        self._groups["cp"] = _Group(chunk, pg, self.global_rank)
        # pipeline neighbours
        pp_ranks = self._groups["pp"].ranks
        s = self._coord["pp"]
        self.is_first_stage = s == 0
        self.is_last_stage = s == pp - 1
        self.prev_rank = pp_ranks[(s - 1) % pp]
        self.next_rank = pp_ranks[(s + 1) % pp]
        # first<->last stage pair for tied-embedding grad all-reduce (SharedLayerDesc)
        self._embed_group = None
        if pp > 1:
            pairs = [[g[0], g[-1]] for g in self._topo.groups_along("pp")]
            for pr in pairs:
                pg = dist.new_group(pr) if can_build else None
                if self.global_rank in pr:
                    self._embed_group = _Group(pr, pg, self.global_rank)

    def _make(self, axes: Tuple[str, ...], can_build: bool) -> _Group:
        mine = None
        for ranks in self._topo.groups_along(*axes):
            pg = dist.new_group(ranks) if (can_build and len(ranks) > 1) else None
            if self.global_rank in ranks:
                mine = _Group(ranks, pg, self.global_rank)
        assert mine is not None
        return mine

    # -- HCG accessor surface (names follow the call sites in SURVEY §2.5) ------------
    def topology(self) -> CommunicateTopology:
        return self._topo

    def get_global_rank(self) -> int:
        return self.global_rank

    def get_parallel_mode(self) -> str:
        d = self._topo.dims
        if d["pp"] > 1:
            return "pipeline"
        if d["mp"] > 1:
            return "tensor"
        if d["sharding"] > 1:
            return "sharding"
        return "data"

    def get_data_parallel_rank(self): return self._coord["dp"]
    def get_data_parallel_world_size(self): return self._topo.dims["dp"]
    def get_data_parallel_group(self): return self._groups["dp"]
    def get_data_parallel_group_src_rank(self): return self._groups["dp"].ranks[0]

    def get_model_parallel_rank(self): return self._coord["mp"]
    def get_model_parallel_world_size(self): return self._topo.dims["mp"]
    def get_model_parallel_group(self): return self._groups["mp"]
    def get_model_parallel_group_src_rank(self): return self._groups["mp"].ranks[0]

    def get_stage_id(self): return self._coord["pp"]
    def get_pipe_parallel_rank(self): return self._coord["pp"]
    def get_pipe_parallel_world_size(self): return self._topo.dims["pp"]
    def get_pipe_parallel_group(self): return self._groups["pp"]

    def get_sharding_parallel_rank(self): return self._coord["sharding"]
    def get_sharding_parallel_world_size(self): return self._topo.dims["sharding"]
    def get_sharding_parallel_group(self): return self._groups["sharding"]
    def get_sharding_parallel_group_src_rank(self): return self._groups["sharding"].ranks[0]

    def get_context_parallel_group(self): return self._groups["cp"]
    def get_context_parallel_world_size(self): return self.cp
    def get_context_parallel_rank(self): return self._groups["cp"].ranks.index(self.global_rank)

    def get_check_parallel_group(self): return self._groups["check"]
    def get_moe_group(self): return self._groups["moe"]
    def get_expert_parallel_group(self): return self._groups["moe"]
    def get_expert_parallel_world_size(self): return self._groups["moe"].nranks
    def get_dp_sharding_group(self): return self._groups["dp_sharding"]
    def get_embedding_group(self): return self._embed_group

    def get_p2p_groups(self):
        return self.prev_rank, self.next_rank

    def get_rank_from_stage(self, stage_id: int) -> int:
        c = dict(self._coord)
        c["pp"] = stage_id
        return self._topo.rank_of(**c)

    def __repr__(self) -> str:
        return f"HybridCommunicateGroup(rank={self.global_rank}, coord={self._coord})"


class HybridCommGroupForMoE(HybridCommunicateGroup):
    """MoE variant (reference: comm_groups.py:125-153): identical axes, and the fused
    ``moe`` = dp x mp group is the expert-parallel world."""


def all_axis_products(world: int) -> List[Tuple[int, int, int, int]]:
    """Utility for tests: every (dp, mp, pp, sharding) factorisation of ``world``."""
    out = []
    for dp, mp, pp, sd in itertools.product(range(1, world + 1), repeat=4):
        if dp * mp * pp * sd == world:
            out.append((dp, mp, pp, sd))
    return out
