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


class RewardResult(TrainingResult):
    def get_result(self) -> List[float]:
        return [self.reward]


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
