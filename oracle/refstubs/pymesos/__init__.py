"""Stub of pymesos: just the two names tfmesos/scheduler.py:12 imports, so the
UNMODIFIED reference control plane can be imported from /root/reference as the
placement oracle.  Test infrastructure only."""


class Scheduler(object):
    pass


class MesosSchedulerDriver(object):
    """Replaced per test by a fake driver; this default refuses to run."""

    def __init__(self, sched, framework, master, use_addict=False):
        raise RuntimeError("no Mesos here: install a fake driver "
                           "(see tests/golden/make_golden.py)")
