"""Drop-in alias: ``from tfmesos import cluster`` keeps working and resolves to
the B200 engine (tfmesos_b200)."""
import sys

import tfmesos_b200
from tfmesos_b200 import Job, TFMesosScheduler, cluster, __VERSION__  # noqa: F401
from tfmesos_b200 import scheduler, server, utils  # noqa: F401

sys.modules[__name__ + '.scheduler'] = scheduler
sys.modules[__name__ + '.server'] = server
sys.modules[__name__ + '.utils'] = utils
