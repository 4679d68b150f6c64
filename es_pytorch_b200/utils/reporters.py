"""Reporters (import-compatible subset of src/utils/reporters.py).

Observability is outside the hot-path scope; these classes keep
the names, constructor signatures and the ``start_gen / log_gen / end_gen / print / log``
protocol that ``es.step`` and the experiment scripts call, with rank-0 gating.
"""
from __future__ import annotations

import logging
import os
import time
from datetime import datetime
from typing import Dict, Tuple

import numpy as np


def calc_dist_rew(tr) -> Tuple[float, float]:
    return float(np.linalg.norm(np.array(tr.positions[-3:-1]))), float(np.sum(tr.rewards))


class Reporter:
    def start_gen(self): ...
    def log_gen(self, fits, noiseless_tr, policy, steps): ...
    def end_gen(self): ...
    def print(self, s: str): ...
    def log(self, d: Dict[str, float]): ...


class ReporterSet(Reporter):
    def __init__(self, *reporters):
        self.reporters = [r for r in reporters if r is not None]

    def _each(self, method, *args):
        for r in self.reporters:
            getattr(r, method)(*args)

    def start_gen(self): self._each('start_gen')
    def log_gen(self, fits, noiseless_tr, policy, steps): self._each('log_gen', fits, noiseless_tr, policy, steps)
    def end_gen(self): self._each('end_gen')
    def print(self, s): self._each('print', s)
    def log(self, d): self._each('log', d)


class MpiReporter(Reporter):
    """Only rank ``MAIN`` emits; subclasses implement the underscore hooks."""
    MAIN = 0

    def __init__(self, comm):
        self.comm = comm

    def _main(self) -> bool:
        return getattr(self.comm, 'rank', 0) == MpiReporter.MAIN

    def start_gen(self):
        if self._main(): self._start_gen()

    def log_gen(self, fits, noiseless_tr, policy, steps):
        if self._main(): self._log_gen(fits, noiseless_tr, policy, steps)

    def end_gen(self):
        if self._main(): self._end_gen()

    def print(self, s):
        if self._main(): self._print(s)

    def log(self, d):
        if self._main(): self._log(d)

    def _start_gen(self): ...
    def _log_gen(self, fits, noiseless_tr, policy, steps): ...
    def _end_gen(self): ...
    def _print(self, s): ...
    def _log(self, d): ...


class DefaultMpiReporter(MpiReporter):
    """Per-generation summary: avg/max per objective, dist, rew, steps, wall time."""

    def __init__(self, comm):
        super().__init__(comm)
        self.gen = 0
        self.cum_steps = 0
        self.gen_start_time = 0

    def _start_gen(self):
        self.gen_start_time = time.time()
        self.print('\n\n----------------------------------------')
        self.log({'gen': self.gen})

    def _summary(self, fits, noiseless_tr, steps) -> Dict[str, float]:
        """One flat record per generation (the metric names are what the reference's logs / mlflow runs use)."""
        cols = np.asarray(fits, dtype=np.float64).reshape(len(fits), -1)
        rec = {}
        for i in range(cols.shape[1]):
            rec[f'avg-{i}'] = float(np.round(cols[:, i].mean(), 2))
            rec[f'max-{i}'] = float(np.round(cols[:, i].max(), 2))
        rec['dist'], rec['rew'] = calc_dist_rew(noiseless_tr)
        rec['steps'], rec['cum steps'], rec['n fits ranked'] = steps, self.cum_steps, len(fits)
        return rec

    def _log_gen(self, fits, noiseless_tr, policy, steps):
        self.cum_steps += steps
        for key, value in self._summary(fits, noiseless_tr, steps).items():
            self.log({key: value})

    def _end_gen(self):
        self.log({'time': round(time.time() - self.gen_start_time, 2)})
        self.gen += 1


class DefaultMpiReporterSet(DefaultMpiReporter):
    def __init__(self, comm, run_name, *reporters):
        super().__init__(comm)
        self.fit_folder = os.path.join('saved', run_name, 'fits')
        self.policy_folder = os.path.join('saved', run_name, 'weights')
        if self._main():
            os.makedirs(self.fit_folder, exist_ok=True)
            os.makedirs(self.policy_folder, exist_ok=True)
        self.reporters = [r for r in reporters if r is not None]
        self.best_rew = 0
        self.best_dist = 0

    def _log_gen(self, fits, noiseless_tr, policy, steps):
        super()._log_gen(fits, noiseless_tr, policy, steps)
        dist_, rew = calc_dist_rew(noiseless_tr)
        if rew > self.best_rew or dist_ > self.best_dist:             # a new best in either measure: keep the policy
            policy.save(self.policy_folder, str(self.gen))
            self.print(f'saving policy with rew:{rew:0.2f} and dist:{dist_:0.2f}')
        self.best_rew, self.best_dist = max(rew, self.best_rew), max(dist_, self.best_dist)
        np.save(os.path.join(self.fit_folder, f'{self.gen}.np'), fits)   # every generation's fitness matrix

    def _log(self, d):
        for r in self.reporters:
            r.log(d)

    def _print(self, s):
        for r in self.reporters:
            r.print(s)


class StdoutReporter(DefaultMpiReporter):
    def _print(self, s):
        print(s)

    def _log(self, d):
        for k, v in d.items():
            print(f'{k}:{v}')


class LoggerReporter(DefaultMpiReporter):
    def __init__(self, comm, log_folder=None):
        super().__init__(comm)
        if self._main():
            log_folder = log_folder or datetime.now().strftime('es__%d_%m_%y__%H_%M_%S')
            os.makedirs(os.path.join('saved', log_folder), exist_ok=True)
            logging.basicConfig(filename=os.path.join('saved', log_folder, 'es.log'), level=logging.DEBUG)
            logging.info('initialized logger')

    def _print(self, s):
        logging.info(s)

    def _log(self, d):
        for k, v in d.items():
            logging.info(f'{k}:{v}')


class MLFlowReporter(DefaultMpiReporter):
    """Forwards metrics to mlflow when it is importable, otherwise records nothing."""

    def __init__(self, comm, cfg):
        super().__init__(comm)
        self.active_run = None
        self.gens = [0] * int(getattr(getattr(cfg, 'general', None), 'n_policies', 1) or 1)
        try:
            import mlflow
            self._mlflow = mlflow if hasattr(mlflow, 'log_metrics') and self._main() else None
        except Exception:
            self._mlflow = None

    def set_active_run(self, i: int):
        if self._main():
            self.active_run = i

    def _start_gen(self):
        pass

    def _end_gen(self):
        if self.active_run is not None:
            self.gens[self.active_run] += 1
        self.active_run = None

    def _print(self, s):
        pass

    def _log(self, d):
        if self._mlflow is not None and self.active_run is not None:
            try:
                self._mlflow.log_metrics(d, self.gens[self.active_run])
            except Exception:
                pass
