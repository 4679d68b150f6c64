"""`src.gym.gym_runner` -> `es_pytorch_b200.gym.gym_runner` (same module object)."""
import sys as _sys
from es_pytorch_b200.gym import gym_runner as _impl
_sys.modules[__name__] = _impl
