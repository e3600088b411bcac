"""``SingletonCommunicationGroup`` — generic N-D orthogonal process-group builder (reference
distributed/protein_folding/scg.py:28-224): ``scg.init_process_group([('dp', None), ('dap', d), ('bp', b)])`` builds one
communicator per named axis (``None`` = whatever is left of the world), supports shared-group aliases and custom rank lists."""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch.distributed as dist

from ...parallel.topology import _Group


def ensure_divisibility(numerator: int, denominator: int) -> None:
    assert numerator % denominator == 0, f"{numerator} is not divisible by {denominator}"


class SingletonCommunicationGroup:
    def __init__(self):
        self._groups: Dict[str, _Group] = {}
        self._dims: Dict[str, int] = {}
        self.world_size, self.rank, self._initialised = 1, 0, False

    def init_process_group(self, parallel_degree: Sequence[Tuple[str, Optional[int]]] = (("dp", None),), custom_ranks: Optional[Dict[str, List[List[int]]]] = None,
                           shared: Optional[Dict[str, str]] = None) -> None:
        live = dist.is_available() and dist.is_initialized()
        self.world_size = dist.get_world_size() if live else 1
        self.rank = dist.get_rank() if live else 0
        known = int(np.prod([d for _, d in parallel_degree if d]))
        assert self.world_size % known == 0, f"world {self.world_size} not divisible by {known}"
        dims = [(n, d if d else self.world_size // known) for n, d in parallel_degree]
        assert int(np.prod([d for _, d in dims])) == self.world_size, "parallel degrees do not multiply to the world size"
        grid = np.arange(self.world_size).reshape([d for _, d in dims])        # first axis slowest
        for i, (name, d) in enumerate(dims):
            self._dims[name] = d
            moved = np.moveaxis(grid, i, -1).reshape(-1, d)
            lists = custom_ranks[name] if custom_ranks and name in custom_ranks else [list(map(int, r)) for r in moved]
            mine = None
            for ranks in lists:
                pg = dist.new_group(ranks) if (live and len(ranks) > 1) else None
                if self.rank in ranks:
                    mine = _Group(list(ranks), pg, self.rank)
            self._groups[name] = mine
        for alias, target in (shared or {}).items():
            self._groups[alias], self._dims[alias] = self._groups[target], self._dims[target]
        self._initialised = True

    def get_group(self, name: str) -> Optional[_Group]:
        return self._groups.get(name)

    def get_world_size(self, name: str) -> int:
        return self._dims.get(name, 1)

    def get_rank_in_group(self, name: str) -> int:
        g = self._groups.get(name)
        return 0 if g is None else max(g.rank, 0)

    @property
    def initialized(self) -> bool:
        return self._initialised

    def __getattr__(self, name: str):
        """The reference attaches per-axis attributes at ``init_process_group`` time (scg.py:165-193): ``scg.<axis>_group``,
        ``scg.get_rank_in_<axis>_group()`` and ``scg.get_<axis>_world_size()`` for whatever axis names the caller chose.  They resolve here
        for every axis that was built, and — like the reference, where user code probes them with ``hasattr`` — do not exist before."""
        groups = self.__dict__.get("_groups", {})
        if name.endswith("_group") and not name.startswith("get_") and name[:-len("_group")] in groups:
            return groups[name[:-len("_group")]]
        if name.startswith("get_rank_in_") and name.endswith("_group") and name[len("get_rank_in_"):-len("_group")] in groups:
            axis = name[len("get_rank_in_"):-len("_group")]
            return lambda: self.get_rank_in_group(axis)
        if name.startswith("get_") and name.endswith("_world_size") and name[len("get_"):-len("_world_size")] in groups:
            axis = name[len("get_"):-len("_world_size")]
            return lambda: self.get_world_size(axis)
        raise AttributeError(name)

    # convenience accessors used by the model code
    def get_dp_group(self): return self.get_group("dp")
    def get_dap_group(self): return self.get_group("dap")
    def get_bp_group(self): return self.get_group("bp")
    def get_dp_world_size(self): return self.get_world_size("dp")
    def get_dap_world_size(self): return self.get_world_size("dap")
    def get_bp_world_size(self): return self.get_world_size("bp")
    def get_dap_rank(self): return self.get_rank_in_group("dap")
    def get_bp_rank(self): return self.get_rank_in_group("bp")


scg = SingletonCommunicationGroup()
