"""Between-graph replicated MNIST training, the B200 edition of the reference's
examples/mnist/mnist_replica.py (same flags, same model, same optimizer):

    tfrun -w 2 -s 1 -Gw 1 -- python examples/mnist/mnist_replica.py \
        --ps_hosts {ps_hosts} --worker_hosts {worker_hosts} \
        --job_name {job_name} --worker_index {task_index}

784 -> 100 ReLU -> 10 softmax, Adam(0.01), batch 100, 200 global steps, async by
default, --sync_replicas for SyncReplicasOptimizer-style mean aggregation.
Variables are placed on the PS tasks by replica_device_setter; each step the
worker PULLs them, computes gradients on its GPU, PUSHes them and the PS applies
Adam -- all of that through libpsx.so.  Batches are synthetic (no network).
"""
from __future__ import print_function

import argparse
import math
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))

import numpy as np  # noqa: E402

from tfmesos_b200 import psx  # noqa: E402
from tfmesos_b200 import train as tf  # noqa: E402

IMAGE_PIXELS = 28


def parse(argv):
    p = argparse.ArgumentParser()
    p.add_argument("--worker_index", type=int, default=0)
    p.add_argument("--ps_hosts", type=str, default="")
    p.add_argument("--worker_hosts", type=str, default="")
    p.add_argument("--job_name", type=str, default="")
    p.add_argument("--replicas_to_aggregate", type=int, default=None)
    p.add_argument("--hidden_units", type=int, default=100)
    p.add_argument("--train_steps", type=int, default=200)
    p.add_argument("--batch_size", type=int, default=100)
    p.add_argument("--learning_rate", type=float, default=0.01)
    p.add_argument("--sync_replicas", action="store_true")
    # B200 build only (not in the reference):
    p.add_argument("--lag_step", action="store_true",
                   help="never block the host on the step's global_step: report the newest "
                        "value already copied back (one step behind)")
    p.add_argument("--straggler", type=int, default=-1,
                   help="worker index that sleeps --straggle_ms before every push (makes "
                        "'whose gradient is stale' deterministic in tests)")
    p.add_argument("--straggle_ms", type=float, default=400.0)
    p.add_argument("--device_batches", action="store_true",
                   help="draw the synthetic batches on the GPU (torch) instead of numpy + H2D")
    p.add_argument("--quiet_steps", action="store_true",
                   help="print the per-step line every 100 steps only (timing runs)")
    p.add_argument("--dump", type=str, default=None,
                   help="chief writes the final variables here (.npz), for parity tests")
    return p.parse_args(argv)


def truncated_normal(rng, shape, stddev):
    out = rng.standard_normal(shape)
    bad = np.abs(out) > 2.0
    while bad.any():
        out[bad] = rng.standard_normal(int(bad.sum()))
        bad = np.abs(out) > 2.0
    return (out * stddev).astype(np.float32)


def main(argv):
    FLAGS = parse(argv)
    ps_hosts = FLAGS.ps_hosts.split(",")
    worker_hosts = FLAGS.worker_hosts.split(",")
    cluster = tf.ClusterSpec({"ps": ps_hosts, "worker": worker_hosts})

    if FLAGS.job_name == "ps":
        server = tf.Server(cluster, job_name="ps", task_index=FLAGS.worker_index)
        server.join()
        sys.exit(0)

    import torch
    torch.backends.cuda.matmul.allow_tf32 = False
    num_workers = len(worker_hosts)
    is_chief = FLAGS.worker_index == 0
    print("Worker index = %d" % FLAGS.worker_index)
    print("Number of workers = %d" % num_workers)
    if os.environ.get("CUDA_VISIBLE_DEVICES"):
        device = 0                       # the scheduler pinned our slice
    else:
        device = FLAGS.worker_index % torch.cuda.device_count()
    torch.cuda.set_device(device)

    # creation order = placement order (replica_device_setter)
    variables = [("global_step", ()),
                 ("hid_w", (IMAGE_PIXELS * IMAGE_PIXELS, FLAGS.hidden_units)),
                 ("hid_b", (FLAGS.hidden_units,)),
                 ("sm_w", (FLAGS.hidden_units, 10)),
                 ("sm_b", (10,))]
    init = {
        "hid_w": truncated_normal(np.random.default_rng(1),
                                  (IMAGE_PIXELS * IMAGE_PIXELS, FLAGS.hidden_units),
                                  1.0 / IMAGE_PIXELS),
        "sm_w": truncated_normal(np.random.default_rng(2), (FLAGS.hidden_units, 10),
                                 1.0 / math.sqrt(FLAGS.hidden_units)),
    }
    opt = tf.AdamOptimizer(FLAGS.learning_rate)
    if is_chief:
        print("Worker %d: Initializing session..." % FLAGS.worker_index)
    else:
        print("Worker %d: Waiting for session to be initialized..." % FLAGS.worker_index)
    sess = tf.ParameterClient(cluster, variables, opt, FLAGS.worker_index, device=device,
                              init=init)
    print("Worker %d: Session initialization complete." % FLAGS.worker_index)
    mode = psx.MODE_SYNC_MEAN if FLAGS.sync_replicas else psx.MODE_ASYNC_ORDERED

    rng = np.random.default_rng(1234 + FLAGS.worker_index)
    names = ["hid_w", "hid_b", "sm_w", "sm_b"]

    def cross_entropy(ps, x, y_):
        hid = torch.relu(x @ ps[0] + ps[1])
        y = torch.softmax(hid @ ps[2] + ps[3], 1)
        return -(y_ * torch.log(torch.clamp(y, 1e-10, 1.0))).sum()

    time_begin = time.time()
    print("Training begins @ %f" % time_begin)
    local_step, step = 0, 0
    while step < FLAGS.train_steps:
        if FLAGS.device_batches:
            x = torch.rand(FLAGS.batch_size, IMAGE_PIXELS * IMAGE_PIXELS, device="cuda")
            y_ = torch.nn.functional.one_hot(
                torch.randint(0, 10, (FLAGS.batch_size,), device="cuda"), 10).float()
        else:
            batch_xs = rng.random((FLAGS.batch_size, IMAGE_PIXELS * IMAGE_PIXELS)).astype(np.float32)
            batch_ys = np.eye(10, dtype=np.float32)[rng.integers(0, 10, FLAGS.batch_size)]
            x = torch.from_numpy(batch_xs).cuda()
            y_ = torch.from_numpy(batch_ys).cuda()
        ps = [sess.params[k].detach().requires_grad_(True) for k in names]
        loss = cross_entropy(ps, x, y_)
        grads = torch.autograd.grad(loss, ps)
        for k, g in zip(names, grads):
            sess.grads[k].copy_(g)
        if FLAGS.straggler == FLAGS.worker_index:
            torch.cuda.synchronize()
            time.sleep(FLAGS.straggle_ms / 1e3)
        step = sess.minimize(mode, FLAGS.replicas_to_aggregate if FLAGS.sync_replicas else None,
                             fetch_step=not FLAGS.lag_step)
        local_step += 1
        if is_chief and (not FLAGS.quiet_steps or local_step % 100 == 0):
            print("%f: Worker %d: training step %d done (global step: %d)"
                  % (time.time(), FLAGS.worker_index, local_step, step))

    step = sess.global_step() if FLAGS.lag_step else step
    if is_chief:
        torch.cuda.synchronize()
        time_end = time.time()
        print("Training ends @ %f" % time_end)
        print("Training elapsed time: %f s" % (time_end - time_begin))
        vrng = np.random.default_rng(99)
        vx = torch.from_numpy(vrng.random((500, 784)).astype(np.float32)).cuda()
        vy = torch.from_numpy(np.eye(10, dtype=np.float32)[vrng.integers(0, 10, 500)]).cuda()
        sess.pull()
        val = cross_entropy([sess.params[k] for k in names], vx, vy).item()
        print("After %d training step(s), validation cross entropy = %g"
              % (FLAGS.train_steps, val))
        if FLAGS.dump:
            np.savez(FLAGS.dump, global_step=step, **{k: sess.read(k) for k in names})
    sess.close()


if __name__ == "__main__":
    main(sys.argv[1:])
