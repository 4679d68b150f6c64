"""`src.nn.obstat` -> `es_pytorch_b200.nn.obstat` (same module object)."""
import sys as _sys
from es_pytorch_b200.nn import obstat as _impl
_sys.modules[__name__] = _impl
