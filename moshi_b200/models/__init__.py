"""Drop-in counterparts of ``moshi.models`` for the hot path (``moshi/moshi/models/__init__.py``)."""
from . import loaders  # noqa: F401
from .compression import MimiModel  # noqa: F401
from .lm import LMGen, LMModel  # noqa: F401
