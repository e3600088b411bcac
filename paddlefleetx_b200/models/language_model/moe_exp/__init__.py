"""GShard / DeepSpeed-style capacity-based MoE ("moe_exp" in the reference: ppfleetx/models/language_model/moe_exp/)."""
from .experts import Experts  # noqa: F401
from .layer import MoE  # noqa: F401
from .sharded_moe import MOELayer, TopKGate, top1gating, top2gating  # noqa: F401
