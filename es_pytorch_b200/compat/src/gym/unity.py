"""`src.gym.unity`: the reference wraps a Unity ML-Agents executable (mlagents_envs) here.  Not part of the ES hot path and
not available offline: importing works (so that multi_agent.py resolves its imports), constructing the wrapper says why it
cannot run."""


class UnityGymWrapper:
    def __init__(self, *args, **kwargs):
        raise NotImplementedError('UnityGymWrapper needs mlagents_envs and a Unity build; es_pytorch_b200 covers the '
                                  'single-agent ES generation path (SURVEY.md section 8)')
