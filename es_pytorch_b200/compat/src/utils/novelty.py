"""`src.utils.novelty` -> `es_pytorch_b200.utils.novelty` (same module object)."""
import sys as _sys
from es_pytorch_b200.utils import novelty as _impl
_sys.modules[__name__] = _impl
