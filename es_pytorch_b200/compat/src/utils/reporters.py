"""`src.utils.reporters` -> `es_pytorch_b200.utils.reporters` (same module object)."""
import sys as _sys
from es_pytorch_b200.utils import reporters as _impl
_sys.modules[__name__] = _impl
