"""Replica mode end to end WITHOUT a GPU: tfrun expands {ps_hosts} {worker_hosts}
{job_name} {task_index} (tfmesos/server.py:72-98), the ps task re-binds the port
its bootstrap reserved (server.py:18-21) through train.Server and serves its
endpoint, the worker reaches it -- with the control sockets bound to the advertised
address (default) and to every interface (TFMESOS_BIND=all, the reference's way)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

PROBE = r'''
import sys, time
from tfmesos_b200 import endpoint
from tfmesos_b200 import train as tf
ps_hosts, worker_hosts, job, idx = sys.argv[1].split(","), sys.argv[2].split(","), sys.argv[3], int(sys.argv[4])
cluster = tf.ClusterSpec({"ps": ps_hosts, "worker": worker_hosts})
if job == "ps":
    tf.Server(cluster, job_name="ps", task_index=idx).join()
else:
    for k, a in enumerate(ps_hosts):
        for _ in range(200):
            try:
                r = endpoint.call(a, "hello")
                print("reached ps %d: %s/%d" % (k, r["job_name"], r["task_index"]), flush=True)
                break
            except OSError:
                time.sleep(0.05)
'''


@pytest.mark.parametrize("bind", ["advertised", "all"])
def test_ps_tasks_rebind_their_reserved_port_and_the_worker_reaches_them(tmp_path, bind):
    probe = tmp_path / "probe.py"
    probe.write_text(PROBE)
    env = dict(os.environ, TFMESOS_BIND=bind,
               PYTHONPATH=ROOT + os.pathsep + os.environ.get("PYTHONPATH", ""))
    r = subprocess.run([sys.executable, os.path.join(ROOT, "script", "tfrun"), "-w", "1", "-s", "2",
                        "--", sys.executable, str(probe), "{ps_hosts}", "{worker_hosts}",
                        "{job_name}", "{task_index}"],
                       env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=180)
    out = r.stdout.decode()
    assert r.returncode == 0, (out[-800:], r.stderr.decode()[-1500:])
    assert "reached ps 0: ps/0" in out and "reached ps 1: ps/1" in out
