"""`src.core.es` -> `es_pytorch_b200.core.es` (same module object)."""
import sys as _sys
from es_pytorch_b200.core import es as _impl
_sys.modules[__name__] = _impl
