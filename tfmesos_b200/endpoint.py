"""The per-task endpoint: what a task process becomes where the reference turns
it into ``tf.train.Server(server_def)`` and joins (tfmesos/server.py:51-66, and
``server.join()`` for job 'ps' at examples/mnist/mnist_replica.py:93-95).

Control plane only: requests are pickle frames (tfmesos_b200.utils, the
reference's framing) on the task's reserved port and ask the endpoint to create
/ export / register / apply shards on ITS GPU.  The data plane never touches
this socket: gradients and parameters move GPU<->GPU through libpsx.so kernels
on the IPC-mapped shard memory.

PS-task GPU policy: a task launched with GPUs sees them through
CUDA_VISIBLE_DEVICES and uses the first; a task launched without (the
reference's default ``-Gs 0``, script/tfrun:25) pins its shards to GPU
``task_index mod n_gpus`` (north star: "PS shards pinned on designated GPUs").
"""
import logging
import os
import socket
import threading
import traceback

from .utils import recv, send

logger = logging.getLogger(__name__)


_channels = threading.local()


def _channel(addr):
    """One kept-alive connection per (thread, endpoint): a training step makes a
    request or two per PS task, and a TCP connect per request would dominate."""
    pool = getattr(_channels, 'pool', None)
    if pool is None:
        pool = _channels.pool = {}
    conn = pool.get(addr)
    if conn is None:
        host, port = addr.rsplit(':', 1)
        conn = socket.create_connection((host, int(port)), timeout=120)
        conn.setsockopt(socket.IPPROTO_TCP, socket.TCP_NODELAY, 1)
        pool[addr] = conn
    return conn


def _drop_channel(addr, conn):
    _channels.pool.pop(addr, None)
    try:
        conn.close()
    except OSError:
        pass


def call(addr, method, **kw):
    """One request to the endpoint at 'host:port' (accepts 'grpc://host:port',
    the form in ``cluster.targets``, scheduler.py:284)."""
    if addr.startswith('grpc://'):
        addr = addr[len('grpc://'):]
    while True:
        pool = getattr(_channels, 'pool', None) or {}
        reused = addr in pool
        conn = _channel(addr)
        try:
            send(conn, (method, kw))
        except OSError:
            # the request never left: a stale kept-alive connection (endpoint
            # restarted).  Reconnect ONCE, and only in this case -- a request that was
            # delivered is never re-sent (an `apply` would run twice and add a global
            # step; a recv timeout behind a sync-mode straggler is not a lost request)
            _drop_channel(addr, conn)
            if not reused:
                raise
            continue
        try:
            status, payload = recv(conn)
        except (OSError, AssertionError, EOFError):
            _drop_channel(addr, conn)
            raise
        break
    if status != 'ok':
        raise RuntimeError('endpoint %s: %s failed: %s' % (addr, method, payload))
    return payload


class Endpoint(object):

    def __init__(self, job_name, task_index, cluster_def, gpus=0):
        self.job_name, self.task_index = job_name, int(task_index)
        self.cluster_def = cluster_def
        self.gpus = gpus
        self.lock = threading.RLock()
        self.shards = {}          # key -> [psx.Shard, applies enqueued, global steps enqueued]
        self.values = {}          # tiny value store for plumbing graphs (plus.py)
        self.stop_event = threading.Event()
        self._device = None
        self._stream = None

    # ---- GPU, created lazily so plumbing-only clusters never touch CUDA ------
    def device(self):
        if self._device is None:
            from . import psx
            n = psx.device_count()
            if os.environ.get('CUDA_VISIBLE_DEVICES') and self.gpus:
                self._device = 0
            else:
                self._device = self.task_index % n
            psx.init(self._device)
        return self._device

    def stream(self):
        if self._stream is None:
            import torch
            self._stream = torch.cuda.Stream(device=self.device())
        return self._stream

    # ---- requests -------------------------------------------------------------
    def do_hello(self):
        return {'job_name': self.job_name, 'task_index': self.task_index,
                'pid': os.getpid()}

    def do_device(self):
        return self.device()

    def do_create_shard(self, key, nelem, opt, hyper, n_slots, wire=0):
        """Idempotent: the chief creates, everyone else gets the same handle."""
        from . import psx
        with self.lock:
            if key not in self.shards:
                lr, b1, b2, eps = hyper
                shard = psx.Shard(self.device(), nelem, opt, lr, b1, b2, eps, n_slots, wire)
                self.shards[key] = [shard, 0, 0, None]
            return self.shards[key][0].export()

    def do_shard_handle(self, key):
        with self.lock:
            return self.shards[key][0].export() if key in self.shards else None

    def do_register_client(self, key, slot, handle):
        # under the lock and with the stream drained: an apply enqueued from another
        # connection's thread may still hold the mapping this replaces (revived worker)
        with self.lock:
            shard = self.shards[key][0]
            self.stream().synchronize()
            shard.register_client(slot, handle)

    def do_unregister_client(self, key, slot):
        """A worker is closing its session: stop publishing into its HBM BEFORE it
        frees the block (the PS keeps serving the other workers)."""
        with self.lock:
            if key in self.shards:
                self.stream().synchronize()
                self.shards[key][0].unregister_client(slot)

    def do_round_bind(self, key, slot, grad_handle, param_handle, elem_off):
        with self.lock:
            shard = self.shards[key][0]
            self.stream().synchronize()
            shard.round_bind(slot, grad_handle, param_handle, elem_off)

    def do_serve(self, key, mode, replicas_to_aggregate=1, idle_sleep_us=20):
        """Start the shard's request-free serving loop (psx_serve_start): from now on
        pushes are consumed as they arrive, with no request per step -- the
        reference's default discipline (examples/mnist/mnist_replica.py:198-205) or,
        mode SYNC_MEAN, SyncReplicasOptimizer on the device (:148-162).  Idempotent:
        every worker asks at its first step, the first one starts it."""
        with self.lock:
            entry = self.shards[key]
            want = (int(mode), int(replicas_to_aggregate))
            if entry[3] is None:
                entry[0].serve_start(want[0], want[1], idle_sleep_us)
                entry[3] = want + (int(idle_sleep_us),)
            elif entry[3][:2] != want:
                raise RuntimeError('shard %r is served with mode/aggregate %r, asked for %r'
                                   % (key, entry[3][:2], want))
            return True

    def _paused(self):
        """Context: every served shard's loop stopped (one consistent cut across
        var / m / v / state), restarted afterwards."""
        ep = self

        class _P(object):
            def __enter__(self_):
                self_.served = [e for e in ep.shards.values() if e[3] is not None]
                for e in self_.served:
                    e[0].serve_stop()

            def __exit__(self_, *exc):
                for e in self_.served:
                    e[0].serve_start(*e[3])
                return False
        return _P()

    def do_serve_stats(self, key):
        return self.shards[key][0].serve_stats()

    def do_apply(self, key, mode, first_slot, count, wait_seq, fused=False):
        """Enqueue wait(flags) + the fused reduce/apply kernel on this task's
        stream; returns the apply_seq the caller's pull must wait for and the
        global_step that apply produces (every apply of a shard goes through this
        method, so the host-side count is exact and needs no device sync)."""
        from . import psx
        with self.lock:
            entry = self.shards[key]
            fn = entry[0].round if fused else entry[0].apply
            fn(mode, first_slot, count, wait_seq, self.stream())
            entry[1] += 1
            entry[2] += count if mode == psx.MODE_ASYNC_ORDERED else 1
            return {'applied': entry[1], 'global_step': entry[2]}

    def do_set_values(self, key, which, off, data):
        import numpy as np
        self.shards[key][0].set_values(which, np.frombuffer(data, np.float32), off)

    def do_get_values(self, key, which, off, n):
        self.stream().synchronize()
        return self.shards[key][0].get_values(which, off, n).tobytes()

    def do_state(self, key):
        self.stream().synchronize()
        return self.shards[key][0].state()

    def do_set_hyper(self, key, hyper):
        self.shards[key][0].set_hyper(*hyper)

    # checkpoints of this task's shards (tf.train.Supervisor(logdir=...) stand-in,
    # examples/mnist/mnist_replica.py:165-170); one file per PS task
    class _Hosted(object):
        def __init__(self, shard):
            self.shard = shard
            self.spec = type('Spec', (), {'nelem': shard.nelem, 'off': 0})()

    def _as_cluster(self):
        view = type('View', (), {})()
        view.servers = {key: self._Hosted(entry[0]) for key, entry in self.shards.items()}
        return view

    def do_save(self, path):
        """One consistent cut: the lock keeps other workers' applies out between the
        var / m / v / state reads (async training), the stream is drained first."""
        from . import checkpoint
        with self.lock, self._paused():
            self.stream().synchronize()
            return checkpoint.save(self._as_cluster(), path, self.task_index,
                                   len(self.cluster_def.get('ps', [])) or 1)

    def do_restore(self, path):
        from . import checkpoint
        with self.lock, self._paused():
            self.stream().synchronize()
            fn = checkpoint.restore(self._as_cluster(), path, self.task_index,
                                    len(self.cluster_def.get('ps', [])) or 1)
            for entry in self.shards.values():
                st = entry[0].state()
                entry[2] = st['global_step']
        return fn

    def do_put(self, name, value):
        with self.lock:
            self.values[name] = value

    def do_get(self, name, default=None):
        with self.lock:
            return self.values.get(name, default)

    # plumbing graph of examples/plus.py: constants on ps tasks, add on a worker
    def do_eval(self, node):
        where = node[1]
        if where is not None and where != (self.job_name, self.task_index):
            return call(self.cluster_def[where[0]][where[1]], 'eval', node=node)
        if node[0] == 'const':
            return node[2]
        if node[0] == 'add':
            return self.do_eval(node[2]) + self.do_eval(node[3])
        raise ValueError('unknown node %r' % (node[0],))

    def do_call(self, fn, kwargs):
        """Run ``module:function(endpoint, **kwargs)`` in this task (the in-graph
        examples drive their workers this way)."""
        import importlib
        mod, name = fn.split(':')
        allowed = os.environ.get('TFMESOS_CALL_MODULES', 'examples,tfmesos_b200,tests').split(',')
        if not any(mod == a or mod.startswith(a + '.') for a in allowed if a):
            raise PermissionError('module %r is not in TFMESOS_CALL_MODULES (%s)'
                                  % (mod, ','.join(allowed)))
        return getattr(importlib.import_module(mod), name)(self, **kwargs)

    def do_shutdown(self):
        self.stop_event.set()

    # ---- server loop ------------------------------------------------------------
    def handle(self, conn):
        """Serve requests on one (kept-alive) connection until the peer closes."""
        try:
            conn.setsockopt(socket.IPPROTO_TCP, socket.TCP_NODELAY, 1)
            while not self.stop_event.is_set():
                try:
                    method, kw = recv(conn)
                except (AssertionError, OSError, EOFError):
                    break                      # peer closed
                try:
                    result = getattr(self, 'do_' + method)(**kw)
                    send(conn, ('ok', result))
                except Exception:
                    send(conn, ('err', traceback.format_exc()))
        except Exception:
            logger.exception('bad request')
        finally:
            conn.close()

    def serve(self, listener):
        listener.listen(128)
        listener.settimeout(0.2)
        while not self.stop_event.is_set():
            try:
                conn, _ = listener.accept()
            except socket.timeout:
                continue
            t = threading.Thread(target=self.handle, args=(conn,))
            t.daemon = True
            t.start()
        for entry in self.shards.values():
            try:
                entry[0].destroy()
            except Exception:
                pass
        listener.close()


def serve(config, reserved_socket):
    """Entry from tfmesos_b200.server when cmd is None (fine-grained mode)."""
    ep = Endpoint(config['job_name'], config['task_index'], config['cluster_def'],
                  gpus=config.get('gpus', 0))
    ep.serve(reserved_socket)
