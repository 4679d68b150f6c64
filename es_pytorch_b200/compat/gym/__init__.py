"""Minimal ``gym`` stand-in: ``make`` returns the synthetic open-loop env of matching shape."""
from es_pytorch_b200.gym.synthetic_env import SyntheticEnv as Env, make, Box  # noqa: F401
from . import spaces, utils, logger  # noqa: F401
