"""No-op ``mlflow`` stand-in (the real package is not installed; tracking is out of scope)."""
import contextlib


class _Run:
    class info:
        run_id = 'none'


def set_experiment(*a, **k): pass
def log_params(*a, **k): pass
def log_param(*a, **k): pass
def end_run(*a, **k): pass


@contextlib.contextmanager
def start_run(*a, **k):
    yield _Run()
