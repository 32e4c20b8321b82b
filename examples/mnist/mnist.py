"""In-graph replicated softmax regression, the B200 edition of the reference's
examples/mnist/mnist.py: one client, N worker tasks, one thread + one session
per worker; W, b, global_step placed by replica_device_setter(ps_tasks=nserver);
GradientDescentOptimizer(0.005); async updates; batches are fed from the client
(synthetic -- no network to download MNIST).

    python examples/mnist/mnist.py -w 2 -s 1 [--steps 10000]
"""
import argparse
import os
import sys
from threading import RLock, Thread

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))

import numpy as np  # noqa: E402

from tfmesos_b200 import cluster  # noqa: E402
from tfmesos_b200 import train as tf  # noqa: E402

VARIABLES = [("W", (784, 10)), ("b", (10,)), ("global_step", ())]


def open_session(ep, worker_index, nworker, learning_rate):
    """Runs inside worker i: attach to the PS tasks (chief = worker 0 runs the
    init_op: zeros, mnist.py:44-45)."""
    import torch
    torch.backends.cuda.matmul.allow_tf32 = False
    device = ep.device()
    torch.cuda.set_device(device)
    spec = tf.ClusterSpec({"ps": ep.cluster_def["ps"], "worker": ep.cluster_def["worker"]})
    ep.values["mnist"] = tf.ParameterClient(spec, VARIABLES,
                                            tf.GradientDescentOptimizer(learning_rate),
                                            worker_index, device=device)
    return True


def train_step(ep, batch_xs, batch_ys):
    """sess.run([steps[i], global_step], feed_dict=...) (mnist.py:71)."""
    import torch
    sess = ep.values["mnist"]
    x = torch.from_numpy(batch_xs).cuda()
    y_ = torch.from_numpy(batch_ys).cuda()
    W = sess.params["W"].detach().requires_grad_(True)
    b = sess.params["b"].detach().requires_grad_(True)
    y = torch.softmax(x @ W + b, 1)
    cross_entropy = -(y_ * torch.log(y)).sum()
    gW, gb = torch.autograd.grad(cross_entropy, [W, b])
    sess.grads["W"].copy_(gW)
    sess.grads["b"].copy_(gb)
    return sess.minimize()


def accuracy(ep, xs, ys):
    import torch
    sess = ep.values["mnist"]
    sess.pull()
    x = torch.from_numpy(xs).cuda()
    y = torch.softmax(x @ sess.params["W"] + sess.params["b"], 1)
    return float((y.argmax(1).cpu().numpy() == ys.argmax(1)).mean())


def main():
    parser = argparse.ArgumentParser()
    parser.add_argument('-w', '--nworker', type=int, default=1)
    parser.add_argument('-s', '--nserver', type=int, default=1)
    parser.add_argument('-Gw', '--worker-gpus', type=int, default=0)
    parser.add_argument('--steps', type=int, default=10000)
    args, _ = parser.parse_known_args()
    nworker, nserver = args.nworker, args.nserver
    jobs_def = [
        {"name": "ps", "num": nserver},
        {"name": "worker", "num": nworker, "gpus": args.worker_gpus},
    ]
    lock = RLock()
    rng = np.random.default_rng(0)
    # synthetic "digits" with MNIST-like statistics: sparse strokes, values in [0, 1]
    centers = ((rng.random((10, 784)) < 0.2) * 0.8).astype(np.float32)

    def next_batch(n):
        labels = rng.integers(0, 10, n)
        xs = np.clip(centers[labels] + 0.1 * rng.standard_normal((n, 784)), 0.0, 1.0)
        return xs.astype(np.float32), np.eye(10, dtype=np.float32)[labels]

    with cluster(jobs_def, quiet=True) as c:
        sessions = [tf.Session(c.targets['/job:worker/task:%d' % i]) for i in range(nworker)]
        for i, s in enumerate(sessions):
            s.call("examples.mnist.mnist:open_session", worker_index=i, nworker=nworker,
                   learning_rate=0.005)
        stop = []

        def train(i):
            step = 0
            while not stop and step < args.steps:
                with lock:
                    batch_xs, batch_ys = next_batch(100)
                step = sessions[i].call("examples.mnist.mnist:train_step",
                                        batch_xs=batch_xs, batch_ys=batch_ys)
            stop.append(i)

        threads = [Thread(target=train, args=(i,)) for i in range(nworker)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        xs, ys = next_batch(2000)
        print(sessions[0].call("examples.mnist.mnist:accuracy", xs=xs, ys=ys))


if __name__ == '__main__':
    main()
