from .gates import BaseGate, GShardGate, NaiveGate, SwitchGate  # noqa: F401
