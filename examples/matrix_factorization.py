# coding: utf-8
"""Non-negative matrix factorisation by gradient descent, the B200 edition of the
reference's examples/matrix_factorization.py: 2 ps + 2 workers, W pinned to
ps:0 and H to ps:1 (matrix_factorization.py:21-28), loss = |R - WH|_F^2 +
1e13 * (sum(|W|-W) + sum(|H|-H)), GradientDescentOptimizer(0.1), a single
session on worker:1 drives every iteration and reads W, H and the loss back.

    python examples/matrix_factorization.py [rows cols rank iters]
"""
from __future__ import print_function

import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np  # noqa: E402

from tfmesos_b200 import cluster  # noqa: E402
from tfmesos_b200 import train as tf  # noqa: E402

INFINITY = 10e+12


def nmf_setup(ep, matrix, rank, learning_rate, seed):
    """Runs inside worker:1 (Session.call): build the 'graph' -- W on ps:0, H on
    ps:1 -- and keep the session in the task."""
    import torch
    torch.backends.cuda.matmul.allow_tf32 = False
    device = ep.device()
    torch.cuda.set_device(device)
    rows, cols = matrix.shape
    scale = 2 * np.sqrt(matrix.mean() / rank)
    rng = np.random.default_rng(seed)
    init = {"W": (rng.random((rows, rank)) * scale).astype(np.float32),
            "H": (rng.random((rank, cols)) * scale).astype(np.float32)}
    spec = tf.ClusterSpec({"ps": ep.cluster_def["ps"], "worker": ["local"]})
    sess = tf.ParameterClient(spec, [("W", (rows, rank)), ("H", (rank, cols))],
                              tf.GradientDescentOptimizer(learning_rate), 0, device=device,
                              placement={"W": 0, "H": 1}, init=init)
    ep.values["nmf"] = (sess, torch.from_numpy(matrix.astype(np.float32)).cuda())
    return True


def nmf_loss(W, H, R):
    import torch
    f_norm = torch.sum((R - W @ H) ** 2)
    nn_w = torch.sum(torch.abs(W) - W)
    nn_h = torch.sum(torch.abs(H) - H)
    constraint = INFINITY * (nn_w + nn_h)
    return f_norm + constraint, constraint


def nmf_run(ep):
    """One session.run([loss, constraint, optimizer]) + W.eval(), H.eval(),
    loss.eval() (matrix_factorization.py:43-49)."""
    import torch
    sess, R = ep.values["nmf"]
    W = sess.params["W"].detach().requires_grad_(True)
    H = sess.params["H"].detach().requires_grad_(True)
    loss, _ = nmf_loss(W, H, R)
    gW, gH = torch.autograd.grad(loss, [W, H])
    sess.grads["W"].copy_(gW)
    sess.grads["H"].copy_(gH)
    sess.minimize()
    with torch.no_grad():
        new_loss, _ = nmf_loss(sess.params["W"], sess.params["H"], R)
    return sess.read("W"), sess.read("H"), float(new_loss)


def main(argv):
    rows = int(argv[1]) if len(argv) > 1 else 1000
    cols = int(argv[2]) if len(argv) > 2 else 1000
    rank = int(argv[3]) if len(argv) > 3 else 200
    max_iter = int(argv[4]) if len(argv) > 4 else 100
    matrix = np.random.default_rng(0).random((rows, cols))
    jobs_def = [
        {"name": "ps", "num": 2},
        {"name": "worker", "num": 2},
    ]
    with cluster(jobs_def, quiet=True) as c:
        with tf.Session(c.targets['/job:worker/task:1']) as session:
            session.call("examples.matrix_factorization:nmf_setup", matrix=matrix, rank=rank,
                         learning_rate=0.1 / (rows * cols / 1e6) if rows * cols > 1e6 else 0.1,
                         seed=1)
            for i in range(max_iter):
                mat_w, mat_h, loss = session.call("examples.matrix_factorization:nmf_run")
                print("loss#%d: %s" % (i, loss))

    err = np.power(matrix - np.matmul(mat_w, mat_h), 2)
    print("err mean: %s" % err.mean())
    print("loss: %s" % loss)


if __name__ == '__main__':
    main(sys.argv)
