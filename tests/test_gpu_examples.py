"""The reference's training examples (BASELINE configs #2, #3) run end to end
through tfrun / cluster() on the CUDA path, and their parameters compared with
an all-CPU oracle run (numpy gradients + oracle optimizer) on the same seeded
synthetic batches.  Tolerance: the gradients come from torch on the GPU vs numpy
on the CPU (different summation order), so parameters are compared with
rtol 2e-3 / atol 2e-4 after a few steps; the PS arithmetic itself is bit-exact
(tests/test_gpu_parity.py)."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from oracle import ps_oracle as o

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
F = np.float32
NAMES = ["hid_w", "hid_b", "sm_w", "sm_b"]


def tfrun(args, timeout=600):
    env = dict(os.environ, PYTHONPATH=ROOT + os.pathsep + os.environ.get("PYTHONPATH", ""))
    r = subprocess.run([sys.executable, os.path.join(ROOT, "script", "tfrun")] + args,
                       env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=timeout)
    assert r.returncode == 0, (r.stdout.decode()[-1500:], r.stderr.decode()[-3000:])
    return r.stdout.decode()


def replica_cmd(extra):
    return ["--", sys.executable, os.path.join(ROOT, "examples", "mnist", "mnist_replica.py"),
            "--ps_hosts", "{ps_hosts}", "--worker_hosts", "{worker_hosts}",
            "--job_name", "{job_name}", "--worker_index", "{task_index}"] + extra


def oracle_mnist_replica(n_workers, rounds, mode, aggregate=None):
    """The same run on the CPU: variables placed on 1 ps task, seeded batches
    rng(1234 + worker), numpy gradients, oracle Adam(0.01)."""
    from tfmesos_b200 import engine
    variables = [("global_step", ()), ("hid_w", (784, 100)), ("hid_b", (100,)),
                 ("sm_w", (100, 10)), ("sm_b", (10,))]
    lay = engine.VariableLayout(variables, 1)
    shard = o.Shard(lay.bucket_nelem[0], o.ADAM, lr=0.01)
    init = {"hid_w": o.truncated_normal(np.random.default_rng(1), (784, 100), 1.0 / 28),
            "sm_w": o.truncated_normal(np.random.default_rng(2), (100, 10), 0.1)}
    for k, v in init.items():
        _, off, _, numel = lay.entries[k]
        shard.var[off:off + numel] = v.ravel()
    rngs = [np.random.default_rng(1234 + w) for w in range(n_workers)]
    for _ in range(rounds):
        slots = np.zeros((n_workers, shard.n), F)
        cur = {}
        for k in NAMES:
            _, off, shape, numel = lay.entries[k]
            cur[k] = shard.var[off:off + numel].reshape(shape).copy()
        for w in range(n_workers):
            x = rngs[w].random((100, 784)).astype(F)
            y = np.eye(10, dtype=F)[rngs[w].integers(0, 10, 100)]
            _, *gs = o.mlp_grads(cur["hid_w"], cur["hid_b"], cur["sm_w"], cur["sm_b"], x, y)
            for k, g in zip(NAMES, gs):
                _, off, _, numel = lay.entries[k]
                slots[w, off:off + numel] = g.ravel()
        shard.round(slots if aggregate is None else slots[:aggregate], mode)
    out = {}
    for k in NAMES:
        _, off, shape, numel = lay.entries[k]
        out[k] = shard.var[off:off + numel].reshape(shape)
    return out, shard.step


def test_mnist_replica_one_worker_matches_oracle(tmp_path):
    """README.rst:92 runs it `-w 1 -s 1`."""
    dump = str(tmp_path / "final.npz")
    out = tfrun(["-w", "1", "-s", "1", "-Gw", "1"] + replica_cmd(
        ["--train_steps", "12", "--dump", dump]))
    assert "training step 12 done (global step: 12)" in out
    assert "Training elapsed time" in out
    got = np.load(dump)
    want, step = oracle_mnist_replica(1, 12, o.ASYNC_ORDERED)
    assert int(got["global_step"]) == step == 12
    for k in NAMES:
        np.testing.assert_allclose(got[k], want[k], rtol=2e-3, atol=2e-4, err_msg=k)


def test_mnist_replica_sync_replicas_two_workers_matches_oracle(tmp_path):
    """SyncReplicasOptimizer path (mnist_replica.py:148-154): mean of the two
    workers' gradients, one apply and one global step per round."""
    dump = str(tmp_path / "final.npz")
    out = tfrun(["-w", "2", "-s", "1"] + replica_cmd(
        ["--train_steps", "8", "--sync_replicas", "--dump", dump]))
    assert "global step: 8" in out
    got = np.load(dump)
    want, step = oracle_mnist_replica(2, 8, o.SYNC_MEAN)
    assert int(got["global_step"]) == step == 8
    for k in NAMES:
        np.testing.assert_allclose(got[k], want[k], rtol=2e-3, atol=2e-4, err_msg=k)


def test_mnist_replica_sync_replicas_to_aggregate_two_of_three(tmp_path):
    """replicas_to_aggregate < num_workers (mnist_replica.py:64-67,109-113): the PS
    averages the first 2 gradients to ARRIVE each round and drops the third as
    stale (device-side SyncReplicas, psx_serve_start SYNC_MEAN).  Worker 2 is made
    a deliberate straggler (it sleeps before every push), so it is always workers
    0 and 1 that are aggregated -- the oracle's schedule; one global step per round."""
    dump = str(tmp_path / "final.npz")
    out = tfrun(["-w", "3", "-s", "1"] + replica_cmd(
        ["--train_steps", "6", "--sync_replicas", "--replicas_to_aggregate", "2",
         "--straggler", "2", "--straggle_ms", "1500", "--dump", dump]))
    assert "global step: 6" in out
    got = np.load(dump)
    want, step = oracle_mnist_replica(3, 6, o.SYNC_MEAN, aggregate=2)
    assert int(got["global_step"]) == step == 6
    for k in NAMES:
        np.testing.assert_allclose(got[k], want[k], rtol=2e-3, atol=2e-4, err_msg=k)


def test_mnist_replica_async_two_workers_two_ps_runs():
    """Default async mode with variables spread over 2 ps tasks (global_step->0,
    hid_w->1, hid_b->0, sm_w->1, sm_b->0): every worker push is its own global
    step; the loop stops once global_step >= train_steps (overshoot < n_workers)."""
    out = tfrun(["-w", "2", "-s", "2", "--worker-logs", "*"] + replica_cmd(["--train_steps", "30"]))
    steps = [int(m) for m in re.findall(r"\(global step: (\d+)\)", out)]
    assert steps and 30 <= max(steps) <= 31
    val = float(re.search(r"validation cross entropy = ([-+0-9.eE]+|nan|inf)", out).group(1))
    assert np.isfinite(val)


def test_matrix_factorization_two_iterations_match_oracle():
    """examples/matrix_factorization.py through cluster(): W on ps:0, H on ps:1,
    SGD(0.1), session on worker:1."""
    sys.path.insert(0, ROOT)
    import tfmesos_b200
    from tfmesos_b200 import train as tf
    rows, cols, rank = 64, 48, 8
    matrix = np.random.default_rng(0).random((rows, cols))
    with tfmesos_b200.cluster([{"name": "ps", "num": 2}, {"name": "worker", "num": 2}],
                              quiet=True) as c:
        with tf.Session(c.targets['/job:worker/task:1']) as session:
            session.call("examples.matrix_factorization:nmf_setup", matrix=matrix, rank=rank,
                         learning_rate=0.01, seed=1)
            for _ in range(2):
                mat_w, mat_h, loss = session.call("examples.matrix_factorization:nmf_run")
    scale = 2 * np.sqrt(matrix.mean() / rank)
    rng = np.random.default_rng(1)
    W = (rng.random((rows, rank)) * scale).astype(F)
    H = (rng.random((rank, cols)) * scale).astype(F)
    R = matrix.astype(F)
    for _ in range(2):
        _, dW, dH = o.nmf_grads(W, H, R)
        o.sgd_apply(W, dW, 0.01)
        o.sgd_apply(H, dH, 0.01)
    np.testing.assert_allclose(mat_w, W, rtol=2e-3, atol=1e-4)
    np.testing.assert_allclose(mat_h, H, rtol=2e-3, atol=1e-4)
    # the reference's 1e13 penalty makes the loss overflow float32 as soon as an
    # entry turns negative (matrix_factorization.py:10,34-36): compare with the
    # oracle's loss rather than insisting on a finite value
    with np.errstate(all="ignore"):
        want_loss, _, _ = o.nmf_grads(W, H, R)
    if np.isfinite(want_loss):
        np.testing.assert_allclose(loss, want_loss, rtol=1e-2)
    else:
        assert not np.isfinite(loss)


def test_in_graph_mnist_runs_and_learns():
    """examples/mnist/mnist.py: client threads drive 2 worker tasks, async SGD."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "examples", "mnist", "mnist.py"),
                        "-w", "2", "-s", "1", "--steps", "60"],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=600)
    assert r.returncode == 0, r.stderr.decode()[-3000:]
    acc = float(r.stdout.decode().strip().splitlines()[-1])
    assert 0.5 < acc <= 1.0


@pytest.mark.multigpu
def test_mnist_replica_pinned_workers_reach_ps_tasks_on_gpus_they_cannot_see():
    """`tfrun -Gw 1`: every worker is pinned to ONE GPU through
    CUDA_VISIBLE_DEVICES, the two ps tasks sit on GPU 0 and GPU 1 -- so worker:0
    maps a shard that lives on a GPU outside its visible set (and whose ordinal
    in the ps task's numbering does not exist in the worker's).  CUDA IPC maps it
    regardless; ordinals from another process must never be interpreted locally."""
    out = tfrun(["-w", "2", "-s", "2", "-Gw", "1", "--worker-logs", "*"] +
                replica_cmd(["--train_steps", "24"]))
    steps = [int(m) for m in re.findall(r"\(global step: (\d+)\)", out)]
    assert steps and 24 <= max(steps) <= 25
    val = float(re.search(r"validation cross entropy = ([-+0-9.eE]+|nan|inf)", out).group(1))
    assert np.isfinite(val)
