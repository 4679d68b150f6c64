"""Fitness adaptors (mirror of src/gym/training_result.py): how a rollout's raw record
becomes the per-objective result list that es.test_params shares."""
from __future__ import annotations

from abc import ABC, abstractmethod
from typing import List, Tuple

import numpy as np


class TrainingResult(ABC):
    """One rollout's record (training_result.py:9-30)."""

    def __init__(self, rewards: List[float], positions: List[float], obs: np.ndarray, steps: int, *args, **kwargs):
        self.rewards = rewards
        self.positions = positions
        self.obs = obs
        self.steps = steps

    @property
    def ob_sum_sq_cnt(self) -> Tuple[np.ndarray, np.ndarray, int]:
        cnt = len(self.obs) if np.any(self.obs) else 0
        return self.obs.sum(axis=0), np.square(self.obs).sum(axis=0), cnt

    @abstractmethod
    def get_result(self) -> List[float]:
        pass

    result = property(lambda self: self.get_result())
    reward = property(lambda self: sum(self.rewards))
    behaviour = property(lambda self: self.positions[-3:-1])


def _adaptor(name: str, base, fn, doc: str):
    """A TrainingResult subclass whose ``get_result`` is ``fn`` (module-level classes: picklable under their own names)."""
    return type(name, (base,), {'get_result': fn, '__doc__': doc, '__module__': __name__})


# single-objective adaptors: (training_result.py:62-79)
RewardResult = _adaptor('RewardResult', TrainingResult, lambda self: [self.reward],
                        'Fitness = the total reward of the episode (python sum of the per-step rewards).')
MeanRewardResult = _adaptor('MeanRewardResult', TrainingResult, lambda self: [self.reward / self.steps],
                            'Reward per step; ``steps`` is the last loop index run_model returns.')
DistResult = _adaptor('DistResult', TrainingResult, lambda self: [np.linalg.norm(self.positions[-3:-1])],
                      'Distance of the final (x, y) from the origin.')
XDistResult = _adaptor('XDistResult', DistResult, lambda self: [self.positions[-3]], 'Final x.')


class MultiAgentTrainingResult(TrainingResult):
    """Record of a multi-agent episode: ``rewards`` / ``obs`` carry one column per agent (training_result.py:33-59).
    Host-side bookkeeping only: the multi-agent (Unity) runner is outside the synthetic-env hot path."""

    def get_result(self):
        return self.reward

    @property
    def ob_sum_sq_cnt(self):
        per_agent = []
        for i in range(self.obs.shape[1]):
            o = self.obs[:, i]
            per_agent.append((o.sum(axis=0), np.square(o).sum(axis=0), len(o) if np.any(o) else 0))
        return per_agent

    def trainingresults(self, tr_type):
        """One single-agent result of class ``tr_type`` per agent."""
        rews, obs = np.array(self.rewards), np.array(self.obs)
        return [tr_type(rews[:, i], self.positions, obs[:, i], self.steps) for i in range(rews.shape[1])]

    reward = property(lambda self: np.sum(self.rewards, axis=0).tolist())


class NSResult(TrainingResult):
    def __init__(self, rewards, positions, obs, steps, archive: np.ndarray, k: int):
        super().__init__(rewards, positions, obs, steps)
        self.archive = archive
        self.k = k

    @property
    def novelty(self):
        from ..utils.novelty import novelty
        return novelty(np.array(self.behaviour), self.archive, self.k)

    def get_result(self) -> List[float]:
        return [self.novelty]


class NSRResult(NSResult):
    def get_result(self) -> List[float]:
        return [sum(self.rewards), self.novelty]
