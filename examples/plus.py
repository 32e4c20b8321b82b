# coding: utf-8
"""Plumbing check, the B200 edition of the reference's examples/plus.py:
2 ps + 2 workers, a=10 lives on ps:0, b=32 on ps:1, a+b is computed on worker:1
and fetched through a session on worker:0.  Prints 42 (README.rst:65)."""
from __future__ import print_function

import logging
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from tfmesos_b200 import cluster
from tfmesos_b200 import train as tf


def main(argv):
    jobs_def = [
        {"name": "ps", "num": 2},
        {"name": "worker", "num": 2},
    ]
    master = argv[1] if len(argv) > 1 else None      # accepted, unused: no Mesos here
    with cluster(jobs_def, master=master, quiet=False) as c:
        with tf.device('/job:ps/task:0'):
            a = tf.constant(10)

        with tf.device('/job:ps/task:1'):
            b = tf.constant(32)

        with tf.device("/job:worker/task:1"):
            op = a + b

        with tf.Session(c.targets['/job:worker/task:0']) as sess:
            print(sess.run(op))


if __name__ == '__main__':
    logging.basicConfig()
    main(sys.argv)
