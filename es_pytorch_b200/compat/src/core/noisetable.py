"""`src.core.noisetable` -> `es_pytorch_b200.core.noisetable` (same module object)."""
import sys as _sys
from es_pytorch_b200.core import noisetable as _impl
_sys.modules[__name__] = _impl
