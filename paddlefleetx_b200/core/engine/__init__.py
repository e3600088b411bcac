from .basic_engine import BasicEngine  # noqa: F401
from .eager_engine import EagerEngine  # noqa: F401
