from . import bp, dap, dp  # noqa: F401
from .scg import scg  # noqa: F401
