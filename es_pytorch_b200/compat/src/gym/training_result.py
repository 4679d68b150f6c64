"""`src.gym.training_result` -> `es_pytorch_b200.gym.training_result` (same module object)."""
import sys as _sys
from es_pytorch_b200.gym import training_result as _impl
_sys.modules[__name__] = _impl
