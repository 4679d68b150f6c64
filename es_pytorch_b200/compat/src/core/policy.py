"""`src.core.policy` -> `es_pytorch_b200.core.policy` (same module object)."""
import sys as _sys
from es_pytorch_b200.core import policy as _impl
_sys.modules[__name__] = _impl
