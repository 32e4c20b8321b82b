"""Single-box scheduler: same public surface and same placement decisions as the
reference's Mesos scheduler, but tasks are local processes, one per B200.

Mirrors (behaviour, not code) tfmesos/scheduler.py:
    Job                  :21-31    plain record
    Task / TaskInfo      :34-177   name, resources, bootstrap command, env
    TFMesosScheduler     :180-481  task enumeration :201-217, first-fit offer
                         matching :223-277, targets :279-286, cluster_def and
                         config broadcast :288-318, rendez-vous loop :320-369,
                         failure policy :384-434, finished :474-477
Parity with the reference on all of these is pinned by
tests/test_control_plane_parity.py against tests/golden/control_plane.json,
which was produced by running the unmodified reference.

What replaces Mesos: ``LocalSchedulerDriver`` makes ONE offer describing this
box (host cores, memory, the visible GPUs as a SET) and launches each accepted
TaskInfo as a child process; a reaper thread turns child exits into
``statusUpdate`` calls, so the reference's state machine is unchanged.
Documented deviations: (1) ``master`` is optional -- there is no Mesos master;
(2) TaskInfos carry ``CUDA_VISIBLE_DEVICES`` built from the GPU slice, because
no Mesos GPU isolator pins devices here;
(3) an error raised inside a driver callback is kept and re-raised from
``finished()`` instead of dying silently in the driver thread.
"""
import getpass
import logging
import math
import os
import select
import socket
import subprocess
import sys
import textwrap
import threading
import time
import uuid

from .utils import AttrDict, bind_advertised, local_hostname, recv, send, setup_logger

FOREVER = 0xFFFFFFFF
TERMINAL_STATES = ('TASK_FINISHED', 'TASK_FAILED', 'TASK_KILLED', 'TASK_ERROR')
logger = logging.getLogger(__name__)


class Job(object):

    def __init__(self, name, num, cpus=1.0, mem=1024.0, gpus=0, cmd=None, start=0):
        self.name, self.num = name, num
        self.cpus, self.mem, self.gpus = cpus, mem, gpus
        self.cmd, self.start = cmd, start


class Task(object):

    def __init__(self, mesos_task_id, job_name, task_index, cpus=1.0, mem=1024.0,
                 gpus=0, cmd=None, volumes=None, env=None):
        self.mesos_task_id = mesos_task_id
        self.job_name, self.task_index = job_name, task_index
        self.cpus, self.mem, self.gpus, self.cmd = cpus, mem, gpus, cmd
        self.volumes = volumes or {}
        self.env = env or {}
        self.offered = False
        self.addr = None
        self.connection = None
        self.initalized = False          # (sic) external code reads this spelling
        self.gpu_slice = []

    def __str__(self):
        return textwrap.dedent('''
        <Task
          mesos_task_id=%s
          addr=%s
        >''' % (self.mesos_task_id, self.addr))

    def to_task_info(self, offer, master_addr, gpu_uuids=(), gpu_resource_type=None,
                     containerizer_type=None, force_pull_image=False):
        """TaskInfo for this task.  Containers / volumes / nvidia-docker of the
        reference (scheduler.py:82-146) have no meaning on one box and are not
        built; name, resources, command and environment are the reference's."""
        gpu_uuids = list(gpu_uuids)
        self.gpu_slice = gpu_uuids
        ti = AttrDict()
        ti.task_id.value = str(self.mesos_task_id)
        ti.agent_id.value = offer.agent_id.value
        ti.name = '/job:%s/task:%s' % (self.job_name, self.task_index)

        def scalar(name, value):
            r = AttrDict(name=name, type='SCALAR')
            r.scalar.value = value
            return r

        ti.resources = [scalar('cpus', self.cpus), scalar('mem', self.mem)]
        if self.gpus and gpu_uuids and gpu_resource_type is not None:
            if gpu_resource_type == 'SET':
                r = AttrDict(name='gpus', type='SET')
                r.set.item = gpu_uuids
                ti.resources.append(r)
            else:
                ti.resources.append(scalar('gpus', len(gpu_uuids)))

        ti.command.shell = True
        ti.command.value = ' '.join([
            sys.executable, '-m', '%s.server' % __package__,
            str(self.mesos_task_id), master_addr])
        variables = [AttrDict(name=k, value=v) for k, v in self.env.items()
                     if k != 'PYTHONPATH']
        variables.append(AttrDict(name='PYTHONPATH', value=':'.join(sys.path)))
        if self.gpus and gpu_uuids:      # deviation (2): pin the slice ourselves
            variables.append(AttrDict(name='CUDA_VISIBLE_DEVICES',
                                      value=','.join(str(g) for g in gpu_uuids)))
        ti.command.environment.variables = variables
        return ti


class TFMesosScheduler(object):
    MAX_FAILURE_COUNT = 3

    def __init__(self, task_spec, role=None, master=None, name=None, quiet=False,
                 volumes=None, containerizer_type=None, force_pull_image=False,
                 forward_addresses=None, protocol='grpc', env=None, extra_config=None):
        self.started = False
        self.master = master or os.environ.get('MESOS_MASTER') or 'local'
        self.name = name or '[tensorflow] %s %s' % (
            os.path.abspath(sys.argv[0]), ' '.join(sys.argv[1:]))
        self.task_spec = task_spec
        self.containerizer_type = containerizer_type
        self.force_pull_image = force_pull_image
        self.protocol = protocol
        self.extra_config = {} if extra_config is None else extra_config
        self.forward_addresses = forward_addresses
        self.role = role or '*'
        self.tasks = {}
        self.task_failure_count = {}
        self.job_finished = {}
        self.callback_error = None
        for job in task_spec:
            self.job_finished[job.name] = 0
            for index in range(job.start, job.num):
                task = Task(str(uuid.uuid4()), job.name, index, cpus=job.cpus,
                            mem=job.mem, gpus=job.gpus, cmd=job.cmd,
                            volumes=volumes, env=env)
                self.tasks[task.mesos_task_id] = task
                self.task_failure_count[self.decorated_task_index(task)] = 0
        if not quiet:
            setup_logger(logger)

    # ------------------------------------------------------------ offers ----
    @staticmethod
    def _read_offer(offer):
        cpus = mem = 0.0
        gpus, gpu_type = [], None
        for res in offer.resources:
            if res.name == 'cpus':
                cpus = res.scalar.value
            elif res.name == 'mem':
                mem = res.scalar.value
            elif res.name == 'gpus':
                gpu_type = res.type
                if gpu_type == 'SET':
                    gpus = res.set.item
                else:
                    gpus = list(range(int(res.scalar.value)))
        return cpus, mem, gpus, gpu_type

    def resourceOffers(self, driver, offers):
        """First fit in task order: a task takes the head of what is left of
        the offer (scheduler.py:252-275); once nothing is pending, offers are
        suppressed and declined for good (scheduler.py:229-232)."""
        for offer in offers:
            if all(t.offered for t in self.tasks.values()):
                self.driver.suppressOffers()
                driver.declineOffer(offer.id, AttrDict(refuse_seconds=FOREVER))
                continue
            cpus, mem, gpus, gpu_type = self._read_offer(offer)
            launch = []
            for task in self.tasks.values():
                if task.offered:
                    continue
                fits = task.cpus <= cpus and task.mem <= mem and task.gpus <= len(gpus)
                if not fits:
                    continue
                cpus -= task.cpus
                mem -= task.mem
                take = int(math.ceil(task.gpus))
                mine, gpus = gpus[:take], gpus[take:]
                task.offered = True
                launch.append(task.to_task_info(
                    offer, self.addr, gpu_uuids=mine, gpu_resource_type=gpu_type,
                    containerizer_type=self.containerizer_type,
                    force_pull_image=self.force_pull_image))
            driver.launchTasks(offer.id, launch)

    # ------------------------------------------------------- cluster view ----
    @property
    def targets(self):
        return {'/job:%s/task:%s' % (t.job_name, t.task_index): 'grpc://%s' % t.addr
                for t in self.tasks.values()}

    def _cluster_def(self):
        cluster_def = {}
        for task in sorted(self.tasks.values(), key=lambda t: t.task_index):
            cluster_def.setdefault(task.job_name, []).append(task.addr)
        return cluster_def

    def _start_tf_cluster(self):
        cluster_def = self._cluster_def()
        for task in self.tasks.values():
            send(task.connection, {
                'job_name': task.job_name,
                'task_index': task.task_index,
                'cpus': task.cpus,
                'mem': task.mem,
                'gpus': task.gpus,
                'cmd': task.cmd,
                'cwd': os.getcwd(),
                'cluster_def': cluster_def,
                'forward_addresses': self.forward_addresses,
                'extra_config': self.extra_config,
                'protocol': self.protocol,
            })
            assert recv(task.connection) == 'ok'
            logger.info('Device /job:%s/task:%s activated @ grpc://%s ',
                        task.job_name, task.task_index, task.addr)
            task.connection.close()

    def start(self):
        listener = socket.socket()
        try:
            listener.setsockopt(socket.SOL_SOCKET, socket.SO_REUSEADDR, 1)
            bind_advertised(listener)
            self.addr = '%s:%s' % (local_hostname(), listener.getsockname()[1])
            listener.listen(64)
            framework = AttrDict(user=getpass.getuser(), name=self.name,
                                 hostname=local_hostname(), role=self.role)
            self.driver = MesosSchedulerDriver(self, framework, self.master,
                                               use_addict=True)
            self.driver.start()
            registered = 0
            while any(not t.initalized for t in list(self.tasks.values())):
                if self.callback_error is not None:
                    raise self.callback_error
                if not select.select([listener], [], [], 0.1)[0]:
                    continue
                conn, _ = listener.accept()
                if not select.select([conn], [], [], 0.1)[0]:
                    conn.close()
                    continue
                task_id, addr = recv(conn)
                task = self.tasks[task_id]
                task.addr, task.connection, task.initalized = addr, conn, True
                registered += 1
                logger.info('Task %s with mesos_task_id %s has registered',
                            '%s:%s' % (task.job_name, task.task_index), task_id)
                logger.info('Out of %d tasks %d tasks have been registered',
                            len(self.tasks), registered)
            self.started = True
            self._start_tf_cluster()
        except Exception:
            self.stop()
            raise
        finally:
            listener.close()

    # --------------------------------------------------------- callbacks ----
    def registered(self, driver, framework_id, master_info):
        logger.info('Tensorflow cluster registered (framework %s, single-box driver)',
                    framework_id.value)
        if self.containerizer_type is None:
            version = tuple(int(x) for x in driver.version.split('.'))
            self.containerizer_type = 'MESOS' if version >= (1, 0, 0) else 'DOCKER'

    def statusUpdate(self, driver, update):
        """Before the cluster is up a dead task is relaunched (at most
        MAX_FAILURE_COUNT failures per task); afterwards anything but
        TASK_FINISHED is fatal (scheduler.py:384-420)."""
        logger.debug('Received status update %s', str(update.state))
        if update.state not in TERMINAL_STATES:
            return
        task_id = update.task_id.value
        task = self.tasks.get(task_id)
        if task is None:
            logger.info('Task not found for mesos task id %s', task_id)
            return
        if self.started:
            if update.state == 'TASK_FINISHED':
                self.job_finished[task.job_name] += 1
                return
            logger.error('Task failed: %s, %s with state %s', task, update.message,
                         update.state)
            raise RuntimeError('Task %s failed! %s with state %s'
                               % (task, update.message, update.state))
        logger.warning('Task failed while launching the server: %s, %s with state %s',
                       task, update.message, update.state)
        if task.connection:
            task.connection.close()
        key = self.decorated_task_index(task)
        self.task_failure_count[key] += 1
        if self.task_failure_count[key] < self.MAX_FAILURE_COUNT:
            self.revive_task(driver, task_id, task)
        else:
            raise RuntimeError('Task %s failed %s with state %s and retries=%s'
                               % (task, update.message, update.state,
                                  self.MAX_FAILURE_COUNT))

    def revive_task(self, driver, mesos_task_id, task):
        logger.info('Going to revive task %s ', task.task_index)
        del self.tasks[mesos_task_id]
        task.offered, task.addr, task.connection = False, None, None
        task.mesos_task_id = str(uuid.uuid4())
        self.tasks[task.mesos_task_id] = task     # re-queued at the END of task order
        driver.reviveOffers()

    @staticmethod
    def decorated_task_index(task):
        return '%s.%s' % (task.job_name, task.task_index)

    def slaveLost(self, driver, agent_id):
        if self.started:
            logger.error('Slave %s lost:', agent_id.value)
            raise RuntimeError('Slave %s lost' % agent_id)

    def executorLost(self, driver, executor_id, agent_id, status):
        if self.started:
            logger.error('Executor %s lost:', executor_id.value)
            raise RuntimeError('Executor %s@%s lost' % (executor_id, agent_id))

    def error(self, driver, message):
        logger.error('Mesos error: %s', message)
        raise RuntimeError('Error ' + message)

    def processHeartBeat(self):
        pass

    # ----------------------------------------------------------- lifecycle ---
    def stop(self):
        logger.debug('exit')
        if hasattr(self, 'tasks'):
            for task in self.tasks.values():
                if task.connection:
                    task.connection.close()
            del self.tasks
        if hasattr(self, 'driver'):
            self.driver.stop()
            self.driver.join()
            del self.driver

    def finished(self):
        """True as soon as ONE whole job has finished -- compared with job.num,
        not num-start (scheduler.py:474-477).  Deviation (3): re-raises an error
        a driver callback hit, instead of looping for ever."""
        if self.callback_error is not None:
            raise self.callback_error
        return any(self.job_finished[job.name] >= job.num for job in self.task_spec)


# ---------------------------------------------------------------------------
def visible_gpus():
    """GPU ordinals this box offers, as strings (a Mesos SET resource)."""
    env = os.environ.get('CUDA_VISIBLE_DEVICES')
    if env is not None:
        return [g for g in env.split(',') if g.strip() != '']
    try:
        out = subprocess.check_output(['nvidia-smi', '-L'], stderr=subprocess.DEVNULL,
                                      timeout=30)
        return [str(i) for i, l in enumerate(out.decode().splitlines())
                if l.startswith('GPU ')]
    except Exception:
        return []


def host_memory_mb():
    try:
        return os.sysconf('SC_PAGE_SIZE') * os.sysconf('SC_PHYS_PAGES') / (1024.0 * 1024.0)
    except (ValueError, OSError):
        return 65536.0


class LocalSchedulerDriver(object):
    """Stands where pymesos.MesosSchedulerDriver stood (scheduler.py:336-339):
    offers this box once, runs accepted tasks as child processes."""
    version = '1.0.0'

    def __init__(self, sched, framework, master, use_addict=True):
        self.sched = sched
        self.framework = framework
        self.children = {}                 # task id -> Popen
        self.held = {}                     # task id -> resources taken from the offer
        self.lock = threading.Lock()
        self.stopping = False
        self.suppressed = False
        self.free = None
        self.reaper = None

    # -- what a Mesos master would send ----------------------------------
    def _offer(self):
        offer = AttrDict()
        offer.id.value = 'local-%s' % uuid.uuid4()
        offer.agent_id.value = 'local'
        offer.hostname = local_hostname()
        cpus = AttrDict(name='cpus', type='SCALAR')
        cpus.scalar.value = self.free['cpus']
        mem = AttrDict(name='mem', type='SCALAR')
        mem.scalar.value = self.free['mem']
        gpus = AttrDict(name='gpus', type='SET')
        gpus.set.item = list(self.free['gpus'])
        offer.resources = [cpus, mem, gpus]
        return offer

    def _call(self, fn, *args):
        try:
            fn(*args)
        except Exception as exc:          # surfaced by sched.start()/finished()
            self.sched.callback_error = exc

    def start(self):
        self.free = {'cpus': float(os.cpu_count() or 1), 'mem': host_memory_mb(),
                     'gpus': visible_gpus()}
        self._call(self.sched.registered, self, AttrDict(value='local-framework'),
                   AttrDict(hostname=local_hostname(), port=0))
        self.reaper = threading.Thread(target=self._reap, name='tfmesos-reaper')
        self.reaper.daemon = True
        self.reaper.start()
        self._call(self.sched.resourceOffers, self, [self._offer()])
        self._check_placeable()

    def _check_placeable(self):
        """A Mesos cluster may offer more later; this box never will.  Tasks that
        did not fit the one local offer would leave start() waiting for ever, so
        fail loudly instead (deviation from the reference, which keeps waiting)."""
        tasks = getattr(self.sched, 'tasks', {})
        left = ['%s:%s' % (t.job_name, t.task_index) for t in tasks.values() if not t.offered]
        if left and self.sched.callback_error is None:
            self.sched.callback_error = RuntimeError(
                'this box cannot place %s: it offers %.0f cpus, %.0f MB, gpus %s -- '
                'lower -Cw/-Cs/-Mw/-Ms/-Gw/-Gs or the task counts'
                % (', '.join(left), float(os.cpu_count() or 1), host_memory_mb(),
                   visible_gpus() or 'none'))

    # -- what the scheduler asks of the driver -----------------------------
    def launchTasks(self, offer_id, infos):
        for ti in infos:
            env = dict(os.environ)
            for var in ti.command.environment.variables:
                env[var.name] = var.value
            held = {'cpus': 0.0, 'mem': 0.0, 'gpus': []}
            for res in ti.resources:
                if res.name == 'cpus':
                    held['cpus'] = res.scalar.value
                elif res.name == 'mem':
                    held['mem'] = res.scalar.value
                elif res.name == 'gpus':
                    held['gpus'] = list(res.set.item) if res.type == 'SET' else []
            self.free['cpus'] -= held['cpus']
            self.free['mem'] -= held['mem']
            self.free['gpus'] = [g for g in self.free['gpus'] if g not in held['gpus']]
            self.held[ti.task_id.value] = held
            proc = subprocess.Popen(ti.command.value, shell=True, env=env,
                                    start_new_session=True)
            with self.lock:
                self.children[ti.task_id.value] = proc

    def declineOffer(self, offer_id, filters=None):
        pass

    def suppressOffers(self):
        self.suppressed = True

    def reviveOffers(self):
        self.suppressed = False
        self._call(self.sched.resourceOffers, self, [self._offer()])
        self._check_placeable()

    def _reap(self):
        while not self.stopping:
            done = []
            with self.lock:
                for task_id, proc in list(self.children.items()):
                    rc = proc.poll()
                    if rc is not None:
                        done.append((task_id, rc))
                        del self.children[task_id]
                        held = self.held.pop(task_id, None)
                        if held:           # a finished task gives its slice back
                            self.free['cpus'] += held['cpus']
                            self.free['mem'] += held['mem']
                            self.free['gpus'] = self.free['gpus'] + held['gpus']
            for task_id, rc in done:
                update = AttrDict(state='TASK_FINISHED' if rc == 0 else 'TASK_FAILED',
                                  message='exit status %s' % rc)
                update.task_id.value = task_id
                self._call(self.sched.statusUpdate, self, update)
            time.sleep(0.05)

    def stop(self):
        self.stopping = True
        with self.lock:
            procs = list(self.children.values())
            self.children.clear()
        for proc in procs:                # like driver.stop(): tasks are killed
            if proc.poll() is None:
                try:
                    os.killpg(proc.pid, 15)
                except OSError:
                    pass
        deadline = time.time() + 5
        for proc in procs:
            try:
                proc.wait(max(0.1, deadline - time.time()))
            except subprocess.TimeoutExpired:
                try:
                    os.killpg(proc.pid, 9)
                except OSError:
                    pass

    def join(self):
        if self.reaper is not None and self.reaper.is_alive():
            self.reaper.join(2)


# scheduler.start() constructs whatever this name is bound to, exactly where the
# reference constructs pymesos' driver (scheduler.py:336); tests rebind it.
MesosSchedulerDriver = LocalSchedulerDriver
