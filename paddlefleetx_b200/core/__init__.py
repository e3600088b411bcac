from .engine import BasicEngine, EagerEngine  # noqa: F401
from .module import BasicModule  # noqa: F401
