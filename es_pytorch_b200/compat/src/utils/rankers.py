"""`src.utils.rankers` -> `es_pytorch_b200.utils.rankers` (same module object)."""
import sys as _sys
from es_pytorch_b200.utils import rankers as _impl
_sys.modules[__name__] = _impl
