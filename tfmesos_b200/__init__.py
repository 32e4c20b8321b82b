"""tfmesos_b200 -- the tfmesos API surface on a single box of B200s.

    from tfmesos_b200 import cluster
    with cluster([{'name': 'ps', 'num': 2}, {'name': 'worker', 'num': 2}]) as c:
        c.targets['/job:worker/task:0']

``cluster`` accepts what the reference's does (tfmesos/__init__.py:7-22): one
job or a list of jobs, each a ``Job`` or a dict of ``Job`` arguments; the
scheduler is started on entry and ALWAYS stopped on exit.
"""
from .scheduler import Job, TFMesosScheduler

__VERSION__ = '0.1.0'


def _as_jobs(spec):
    items = spec if isinstance(spec, (list, tuple)) else [spec]
    return [item if isinstance(item, Job) else Job(**item) for item in items]


class cluster(object):
    """Context manager yielding the started scheduler."""

    def __init__(self, jobs, **kw):
        self._jobs = _as_jobs(jobs)
        self._kw = kw
        self._scheduler = None

    def __enter__(self):
        # looked up at call time so tests may substitute the scheduler class
        self._scheduler = globals()['TFMesosScheduler'](self._jobs, **self._kw)
        try:
            self._scheduler.start()
        except BaseException:
            self._scheduler.stop()
            raise
        return self._scheduler

    def __exit__(self, exc_type, exc, tb):
        self._scheduler.stop()
        return False
