"""``create_hcg`` — build the hybrid communicator set from a strategy.

Reference: ppfleetx/distributed/apis/comm_groups.py:27-153 offers two classes selected by the YAML
key ``Distributed.hcg``: Paddle's ``HybridCommunicateGroup`` and an in-tree
``HybridCommGroupForMoE`` that adds the fused dp x mp ``moe`` group.  Here both are one
implementation (parallel/topology.py) — the fused groups always exist.
"""
from __future__ import annotations

from ...parallel.topology import HybridCommGroupForMoE, HybridCommunicateGroup

_REGISTRY = {
    "HybridCommunicateGroup": HybridCommunicateGroup,
    "HybridCommGroupForMoE": HybridCommGroupForMoE,
    "Hybrid4DCommGroup": HybridCommunicateGroup,
}


def create_hcg(strategy, hcg_name: str = "HybridCommunicateGroup"):
    if hcg_name not in _REGISTRY:
        raise ValueError(f"unknown hcg {hcg_name}; choose from {sorted(_REGISTRY)}")
    hc = strategy.hybrid_configs
    return _REGISTRY[hcg_name](dp=hc.get("dp_degree", 1), mp=hc.get("mp_degree", 1),
                               pp=hc.get("pp_degree", 1), sharding=hc.get("sharding_degree", 1), cp=hc.get("cp_degree", 1))
