"""`src.utils.utils` -> `es_pytorch_b200.utils.utils` (same module object)."""
import sys as _sys
from es_pytorch_b200.utils import utils as _impl
_sys.modules[__name__] = _impl
