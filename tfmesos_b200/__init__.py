"""tfmesos_b200 -- the tfmesos API surface on a single box of B200s.

    from tfmesos_b200 import cluster          # tfmesos/__init__.py:7-22
    with cluster([{'name': 'ps', 'num': 2}, {'name': 'worker', 'num': 2}]) as c:
        c.targets['/job:worker/task:0']
"""
from contextlib import contextmanager

from .scheduler import Job, TFMesosScheduler

__VERSION__ = '0.1.0'


@contextmanager
def cluster(jobs, **kw):
    """dict | Job | list of either -> [Job]; start on entry, always stop on exit."""
    if isinstance(jobs, (dict, Job)):
        jobs = [jobs]
    jobs = [j if isinstance(j, Job) else Job(**j) for j in jobs]
    s = TFMesosScheduler(jobs, **kw)
    try:
        s.start()
        yield s
    finally:
        s.stop()
