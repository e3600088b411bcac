from .basic_module import BasicModule  # noqa: F401
