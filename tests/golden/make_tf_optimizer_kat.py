#!/usr/bin/env python
"""Writes tests/golden/tf_optimizer_kat.json: the known-answer vectors of
TensorFlow's OWN optimizer unit tests, the only externally authored anchor the
PS data path has (the reference, douban/tfmesos, ships no test for it and
TensorFlow 0.12 -- requirements.txt:10 -- cannot be installed here).

Sources (TensorFlow r0.12, tensorflow/python/training/):

* ``gradient_descent_test.py`` ``GradientDescentOptimizerTest.testBasic``:
  var0 = [1.0, 2.0], var1 = [3.0, 4.0], grads0 = [0.1, 0.1], grads1 = [0.01, 0.01],
  GradientDescentOptimizer(3.0), one step; the test asserts
  var0 == [1.0 - 3.0 * 0.1, 2.0 - 3.0 * 0.1] and var1 == [3.0 - 3.0 * 0.01, 4.0 - 3.0 * 0.01].

* ``adam_test.py`` ``AdamOptimizerTest.testBasic``: the same four vectors,
  AdamOptimizer() with its defaults (lr 0.001, beta1 0.9, beta2 0.999, eps 1e-8),
  three steps; before step t the test asserts beta1_power == 0.9**t and
  beta2_power == 0.999**t (t = 1..3) and after each step compares the variables
  with the test's own reference implementation

      def adam_update_numpy(param, g_t, t, m, v, alpha=0.001, beta1=0.9,
                            beta2=0.999, epsilon=1e-8):
        alpha_t = alpha * np.sqrt(1 - beta2**t) / (1 - beta1**t)
        m_t = beta1 * m + (1 - beta1) * g_t
        v_t = beta2 * v + (1 - beta2) * g_t * g_t
        param_t = param - alpha_t * m_t / (np.sqrt(v_t) + epsilon)
        return param_t, m_t, v_t

  through assertAllCloseAccordingToType (float32: rtol = atol = 1e-6).

This script evaluates exactly those published formulas (``adam_update_numpy``
verbatim, float32 inputs as in the test's float32 leg) -- it does NOT import the
oracle -- and stores the expected values; tests/test_tf_optimizer_kat.py then
holds both oracle restatements (CPU) and the CUDA kernels (GPU) to TF's own
tolerance.  This is the ceiling of what can be pinned without TensorFlow.
"""
import json
import os

import numpy as np


def adam_update_numpy(param, g_t, t, m, v, alpha=0.001, beta1=0.9, beta2=0.999, epsilon=1e-8):
    alpha_t = alpha * np.sqrt(1 - beta2 ** t) / (1 - beta1 ** t)
    m_t = beta1 * m + (1 - beta1) * g_t
    v_t = beta2 * v + (1 - beta2) * g_t * g_t
    param_t = param - alpha_t * m_t / (np.sqrt(v_t) + epsilon)
    return param_t, m_t, v_t


def main():
    var0 = np.array([1.0, 2.0], np.float32)
    var1 = np.array([3.0, 4.0], np.float32)
    g0 = np.array([0.1, 0.1], np.float32)
    g1 = np.array([0.01, 0.01], np.float32)
    out = {"source": "tensorflow r0.12 python/training/{gradient_descent,adam}_test.py testBasic",
           "tolerance": {"rtol": 1e-6, "atol": 1e-6,
                         "why": "assertAllCloseAccordingToType, float32 leg"},
           "inputs": {"var0": var0.tolist(), "var1": var1.tolist(),
                      "grads0": g0.tolist(), "grads1": g1.tolist()}}
    out["sgd"] = {"learning_rate": 3.0, "steps": 1,
                  "var0": [1.0 - 3.0 * 0.1, 2.0 - 3.0 * 0.1],
                  "var1": [3.0 - 3.0 * 0.01, 4.0 - 3.0 * 0.01]}
    m0 = v0 = m1 = v1 = 0.0
    p0, p1 = var0, var1
    steps = []
    for t in range(1, 4):
        powers = [0.9 ** t, 0.999 ** t]          # asserted BEFORE the update of step t
        p0, m0, v0 = adam_update_numpy(p0, g0, t, m0, v0)
        p1, m1, v1 = adam_update_numpy(p1, g1, t, m1, v1)
        steps.append({"t": t, "beta_powers_before": powers,
                      "var0": np.asarray(p0, np.float64).tolist(),
                      "var1": np.asarray(p1, np.float64).tolist(),
                      "m0": np.asarray(m0, np.float64).tolist(),
                      "v0": np.asarray(v0, np.float64).tolist()})
    out["adam"] = {"learning_rate": 0.001, "beta1": 0.9, "beta2": 0.999, "epsilon": 1e-8,
                   "steps": steps}
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "tf_optimizer_kat.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
    print(path)


if __name__ == "__main__":
    main()
