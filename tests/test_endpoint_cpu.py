"""The per-task endpoint's control requests that need no GPU: the plumbing graph
of plus.py, the value store used for the chief/wait handshake, error reporting,
kept-alive connections."""
import socket
import threading

import pytest

from tfmesos_b200 import endpoint
from tfmesos_b200 import train as tf


def _start(n_ps=2, n_worker=1):
    socks = []
    for _ in range(n_ps + n_worker):
        s = socket.socket()
        s.setsockopt(socket.SOL_SOCKET, socket.SO_REUSEADDR, 1)
        s.bind(("127.0.0.1", 0))
        s.listen(128)              # reachable before the serving thread is scheduled
        socks.append(s)
    addrs = ["127.0.0.1:%d" % s.getsockname()[1] for s in socks]
    cluster_def = {"ps": addrs[:n_ps], "worker": addrs[n_ps:]}
    eps = []
    for i, s in enumerate(socks):
        job, idx = ("ps", i) if i < n_ps else ("worker", i - n_ps)
        ep = endpoint.Endpoint(job, idx, cluster_def)
        t = threading.Thread(target=ep.serve, args=(s,))
        t.daemon = True
        t.start()
        eps.append(ep)
    return cluster_def, eps


def _stop(eps):
    for ep in eps:
        ep.stop_event.set()


def test_plus_graph_is_evaluated_across_tasks():
    cluster_def, eps = _start(2, 2)
    try:
        with tf.device('/job:ps/task:0'):
            a = tf.constant(10)
        with tf.device('/job:ps/task:1'):
            b = tf.constant(32)
        with tf.device('/job:worker/task:1'):
            op = a + b
        assert op[1] == ("worker", 1) and a[1] == ("ps", 0)
        with tf.Session("grpc://" + cluster_def["worker"][0]) as sess:
            assert sess.run(op) == 42
    finally:
        _stop(eps)


def test_value_store_and_hello_over_one_kept_alive_connection():
    cluster_def, eps = _start(1, 0)
    try:
        addr = cluster_def["ps"][0]
        assert endpoint.call(addr, "get", name="initialized", default=False) is False
        endpoint.call(addr, "put", name="initialized", value=True)
        assert endpoint.call(addr, "get", name="initialized") is True
        hello = endpoint.call(addr, "hello")
        assert hello["job_name"] == "ps" and hello["task_index"] == 0
        pool = endpoint._channels.pool
        assert list(pool) .count(addr) == 1            # one socket served all four requests
    finally:
        _stop(eps)


def test_errors_come_back_as_runtime_error_with_the_remote_traceback():
    cluster_def, eps = _start(1, 0)
    try:
        addr = cluster_def["ps"][0]
        with pytest.raises(RuntimeError, match="no_such_method"):
            endpoint.call(addr, "no_such_method")
        with pytest.raises(RuntimeError, match="KeyError"):
            endpoint.call(addr, "register_client", key=(9, 9), slot=0, handle=b"")
        assert endpoint.call(addr, "hello")["pid"] > 0     # the connection survived
    finally:
        _stop(eps)


def test_client_reconnects_after_the_endpoint_restarts_on_the_same_port():
    s = socket.socket()
    s.setsockopt(socket.SOL_SOCKET, socket.SO_REUSEADDR, 1)
    s.bind(("127.0.0.1", 0))
    s.listen(128)
    port = s.getsockname()[1]
    addr = "127.0.0.1:%d" % port
    ep = endpoint.Endpoint("ps", 0, {"ps": [addr], "worker": []})
    t = threading.Thread(target=ep.serve, args=(s,))
    t.daemon = True
    t.start()
    assert endpoint.call(addr, "hello")["task_index"] == 0
    ep.stop_event.set()
    t.join(5)
    endpoint._channels.pool[addr].close()                # what a dead peer looks like
    s2 = socket.socket()
    s2.setsockopt(socket.SOL_SOCKET, socket.SO_REUSEADDR, 1)
    s2.bind(("127.0.0.1", port))
    s2.listen(128)
    ep2 = endpoint.Endpoint("ps", 0, {"ps": [addr], "worker": []})
    t2 = threading.Thread(target=ep2.serve, args=(s2,))
    t2.daemon = True
    t2.start()
    try:
        assert endpoint.call(addr, "hello")["job_name"] == "ps"
    finally:
        ep2.stop_event.set()


def test_a_delivered_request_is_never_sent_twice():
    """call() used to re-send after ANY socket error, so an `apply` whose
    reply was lost ran twice (an extra global step).  Now only a send() failure on a
    reused connection reconnects; a request that left is never repeated."""
    from tfmesos_b200.utils import recv
    seen = []
    srv = socket.socket()
    srv.bind(("127.0.0.1", 0))
    srv.listen(8)
    addr = "127.0.0.1:%d" % srv.getsockname()[1]

    def serve():
        for _ in range(2):
            try:
                srv.settimeout(3)
                conn, _ = srv.accept()
            except OSError:
                return
            try:
                seen.append(recv(conn))        # take the request ...
            except Exception:
                pass
            conn.close()                       # ... and drop the connection without replying

    t = threading.Thread(target=serve)
    t.daemon = True
    t.start()
    with pytest.raises((OSError, AssertionError, EOFError)):
        endpoint.call(addr, "apply", key=(0, 0))
    t.join(5)
    srv.close()
    assert len(seen) == 1 and seen[0][0] == "apply"


def test_call_only_runs_allow_listed_modules():
    cluster_def, eps = _start(1, 0)
    try:
        with pytest.raises(RuntimeError, match="TFMESOS_CALL_MODULES"):
            endpoint.call(cluster_def["ps"][0], "call", fn="os:getcwd", kwargs={})
    finally:
        _stop(eps)
