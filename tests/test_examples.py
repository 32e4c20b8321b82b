"""The reference's examples, run through this repo's cluster()/tfrun surface."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_plus_example_prints_42():
    """README.rst:65: "should result in an output of 42" -- the only expected
    value the reference states.  2 ps + 2 workers, no GPU involved."""
    out = subprocess.run([sys.executable, os.path.join(ROOT, "examples", "plus.py")],
                         stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=120)
    assert out.returncode == 0, out.stderr.decode()
    assert out.stdout.decode().strip().splitlines()[-1] == "42"
