# coding: utf-8
"""Per-task bootstrap: ``python -m tfmesos_b200.server <task_id> <sched host:port>``.

Same contract as tfmesos/server.py:14-113:
  * reserve a port, connect back, send ``(task_id, "host:port")``, receive the
    config dict, answer ``'ok'``                                   (:18-49)
  * cmd is None  -> become the task's endpoint and serve until killed (:51-66;
    the reference builds a tf.train.Server here -- this is the doorway to the
    hot path, now the B200 PS/worker endpoint of tfmesos_b200.endpoint)
  * cmd given    -> run ``initializer``, expand ``{ps_hosts} {worker_hosts}
    {job_name} {task_index}``, export ``TFMESOS_*``, run the command in the
    scheduler's cwd, tee its stdout (optionally to tfrun's collector with a
    ``[job:idx] `` prefix), run ``finalizer``                         (:67-113)
"""
import logging
import os
import socket
import subprocess
import sys

from .utils import bind_advertised, local_hostname, recv, send

logger = logging.getLogger(__name__)


def reserve_port():
    """Bind (never listen): the port is reserved for whatever the task serves
    next, as the reference does for TF's gRPC server (server.py:18-21)."""
    fd = socket.socket()
    fd.setsockopt(socket.SOL_SOCKET, socket.SO_REUSEADDR, 1)
    bind_advertised(fd)
    return fd, '%s:%s' % (local_hostname(), fd.getsockname()[1])


def expand_command(cmd, cluster_def, job_name, task_index):
    """The four substitutions of replica mode and the TFMESOS_* environment
    (server.py:72-92); job names 'ps' and 'worker' are hard-wired there too."""
    ps_hosts = ','.join(cluster_def['ps'])
    worker_hosts = ','.join(cluster_def['worker'])
    env = {
        'PYTHONUNBUFFERED': '1',
        'TFMESOS_PS_HOSTS': ps_hosts,
        'TFMESOS_WORKER_HOSTS': worker_hosts,
        'TFMESOS_JOB_NAME': job_name,
        'TFMESOS_TASK_INDEX': str(task_index),
        'TFMESOS_DISTRIBUTED': '1',
    }
    line = cmd.format(ps_hosts=ps_hosts, worker_hosts=worker_hosts,
                      job_name=job_name, task_index=task_index)
    return line, env


def run_command(config, forward_fd):
    extra = config['extra_config'] or {}
    if extra.get('initializer') is not None:
        subprocess.check_call(extra['initializer'], shell=True)
    line, extra_env = expand_command(config['cmd'], config['cluster_def'],
                                     config['job_name'], config['task_index'])
    env = os.environ.copy()
    env.update(extra_env)
    prefix = ('[%s:%s] ' % (config['job_name'], config['task_index'])).encode('ascii')
    out = getattr(sys.stdout, 'buffer', sys.stdout)
    try:
        child = subprocess.Popen(line, shell=True, cwd=config['cwd'],
                                 stdout=subprocess.PIPE, env=env)
        for chunk in iter(child.stdout.readline, b''):
            out.write(chunk)
            out.flush()
            if forward_fd:
                forward_fd.sendall(prefix + chunk)
        return child.wait()
    finally:
        if extra.get('finalizer') is not None:
            logger.info('Running clean up command %s', extra['finalizer'])
            subprocess.check_call(extra['finalizer'], shell=True)
        if forward_fd:
            forward_fd.close()


def main(argv):
    task_id, sched_addr = argv[1:3]
    host, port = sched_addr.rsplit(':', 1)
    reserved, my_addr = reserve_port()
    conn = socket.create_connection((host, int(port)))
    send(conn, (task_id, my_addr))
    config = recv(conn)

    forward_fd = None
    target = '/job:%s/task:%s' % (config['job_name'], config['task_index'])
    forwards = config['forward_addresses']
    if forwards and target in forwards:
        forward_fd = socket.create_connection(tuple(forwards[target]))

    if config['cmd'] is None:
        reserved.listen(128)      # be reachable before the scheduler hands out targets
    send(conn, 'ok')
    conn.close()

    if config['cmd'] is None:
        from . import endpoint
        try:
            endpoint.serve(config, reserved)
        except KeyboardInterrupt:
            pass
        return 0
    reserved.close()
    return run_command(config, forward_fd)


if __name__ == '__main__':
    sys.exit(main(sys.argv))
