"""`src.nn.nn` -> `es_pytorch_b200.nn.nn` (same module object)."""
import sys as _sys
from es_pytorch_b200.nn import nn as _impl
_sys.modules[__name__] = _impl
