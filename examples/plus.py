# coding: utf-8
"""Plumbing check (the B200 edition of the reference's examples/plus.py): with
2 ps + 2 workers, the constant 10 lives on ps:0, 32 on ps:1, their sum is
computed on worker:1 and fetched through a session on worker:0.
Prints 42 (README.rst:65 -- the only expected value the reference states)."""
from __future__ import print_function

import logging
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from tfmesos_b200 import Job, cluster  # noqa: E402
from tfmesos_b200 import train as tf  # noqa: E402

PLACEMENT = {'/job:ps/task:0': 10, '/job:ps/task:1': 32}
ADD_ON = '/job:worker/task:1'
FETCH_FROM = '/job:worker/task:0'


def build_graph():
    terms = []
    for where, value in sorted(PLACEMENT.items()):
        with tf.device(where):
            terms.append(tf.constant(value))
    with tf.device(ADD_ON):
        return terms[0] + terms[1]


def main(argv):
    master = argv[1] if len(argv) > 1 else None      # accepted, unused: no Mesos here
    with cluster([Job('ps', 2), Job('worker', 2)], master=master, quiet=False) as c:
        op = build_graph()
        with tf.Session(c.targets[FETCH_FROM]) as sess:
            print(sess.run(op))


if __name__ == '__main__':
    logging.basicConfig()
    main(sys.argv)
