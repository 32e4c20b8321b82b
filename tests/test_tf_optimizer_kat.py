"""TensorFlow's own published optimizer known-answers (adam_test.py /
gradient_descent_test.py ``testBasic``, TF r0.12) held against both oracle
restatements (CPU) and the CUDA kernels (GPU) at TF's own float32 tolerance.
The vectors and their provenance: tests/golden/make_tf_optimizer_kat.py."""
import json
import os

import numpy as np
import pytest

from oracle import ps_oracle as o

F = np.float32
KAT = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden",
                                  "tf_optimizer_kat.json")))
RTOL, ATOL = KAT["tolerance"]["rtol"], KAT["tolerance"]["atol"]
INP = KAT["inputs"]
VAR = np.array(INP["var0"] + INP["var1"], F)       # one 4-element bucket: var0 | var1
GRAD = np.array(INP["grads0"] + INP["grads1"], F)


def _expect_adam(step):
    return np.array(step["var0"] + step["var1"], np.float64)


@pytest.mark.parametrize("cls", [o.Shard, o.CShard])
def test_oracle_reproduces_tf_sgd_testbasic(cls):
    sh = cls(4, o.SGD, lr=KAT["sgd"]["learning_rate"])
    sh.var[:] = VAR
    sh.round(GRAD[None, :], o.SUM)
    want = np.array(KAT["sgd"]["var0"] + KAT["sgd"]["var1"])
    np.testing.assert_allclose(sh.var, want, rtol=RTOL, atol=ATOL)
    assert sh.step == 1


@pytest.mark.parametrize("cls", [o.Shard, o.CShard])
def test_oracle_reproduces_tf_adam_testbasic(cls):
    a = KAT["adam"]
    sh = cls(4, o.ADAM, lr=a["learning_rate"], b1=a["beta1"], b2=a["beta2"], eps=a["epsilon"])
    sh.var[:] = VAR
    for st in a["steps"]:
        # TF asserts the stored powers BEFORE running update t
        np.testing.assert_allclose([sh.b1p, sh.b2p], st["beta_powers_before"], rtol=RTOL)
        sh.round(GRAD[None, :], o.SUM)
        np.testing.assert_allclose(sh.var, _expect_adam(st), rtol=RTOL, atol=ATOL)
        np.testing.assert_allclose(sh.m[:2], st["m0"], rtol=1e-5, atol=ATOL)
        np.testing.assert_allclose(sh.v[:2], st["v0"], rtol=1e-5, atol=1e-9)
    assert sh.step == 3


@pytest.mark.gpu
@pytest.mark.parametrize("fused", [False, True])
def test_cuda_kernels_reproduce_tf_testbasic(fused):
    """The same vectors through libpsx.so: push -> fused reduce+apply -> pull, and
    the one-kernel psx_round."""
    import torch
    from tfmesos_b200 import engine, psx
    psx.init(0)
    a = KAT["adam"]
    for opt, steps in ((engine.GradientDescentOptimizer(KAT["sgd"]["learning_rate"]), None),
                       (engine.AdamOptimizer(a["learning_rate"], a["beta1"], a["beta2"],
                                             a["epsilon"]), a["steps"])):
        cl = engine.LocalCluster([("v", (4,))], 1, 1, opt, fused=fused)
        cl.set_variable("v", VAR)
        cl.workers[0].grads["v"].copy_(torch.from_numpy(GRAD))
        if steps is None:
            cl.round(psx.MODE_SUM)
            torch.cuda.synchronize()
            want = np.array(KAT["sgd"]["var0"] + KAT["sgd"]["var1"])
            np.testing.assert_allclose(cl.get_variable("v"), want, rtol=RTOL, atol=ATOL)
            np.testing.assert_allclose(cl.workers[0].params["v"].cpu().numpy(), want,
                                       rtol=RTOL, atol=ATOL)
        else:
            for st in steps:
                s = cl.servers[(0, 0)].shard.state()
                np.testing.assert_allclose([s["beta1_power"], s["beta2_power"]],
                                           st["beta_powers_before"], rtol=RTOL)
                cl.round(psx.MODE_SUM)
                torch.cuda.synchronize()
                np.testing.assert_allclose(cl.get_variable("v"), _expect_adam(st),
                                           rtol=RTOL, atol=ATOL)
            assert cl.global_step() == 3
        cl.close()
