from es_pytorch_b200.gym.synthetic_env import Box  # noqa: F401
