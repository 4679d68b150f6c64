"""`src.nn.optimizers` -> `es_pytorch_b200.nn.optimizers` (same module object)."""
import sys as _sys
from es_pytorch_b200.nn import optimizers as _impl
_sys.modules[__name__] = _impl
