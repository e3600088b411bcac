from .moe_layer import ExpertLayer, MoELayer, build_moe_layer  # noqa: F401
