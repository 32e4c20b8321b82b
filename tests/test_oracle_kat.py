"""Pins the oracle itself.  The reference has no golden vectors for the PS
update path ("parity unpinned", oracle/ps_oracle.c header), so the oracle is
pinned by (1) hand-computed known answers, (2) agreement between its two
independent restatements (numpy / C), (3) torch autograd for the model maths,
(4) the placement tables of SURVEY.md 8(a3)."""
import numpy as np
import pytest

from oracle import ps_oracle as o

F = np.float32


def test_sgd_known_answer_exact_binary_fractions():
    # mnist.py:55 semantics, values chosen to be exact in binary32
    var = np.array([1.0, -2.0, 0.5, 4.0], F)
    g = np.array([0.5, 0.25, -1.0, 8.0], F)
    o.sgd_apply(var, g, 0.5)
    assert var.tolist() == [0.75, -2.125, 1.0, 0.0]


def test_sgd_three_steps_known_answer():
    s = o.Shard(4, o.SGD, lr=0.25)
    s.var[:] = [1, 2, 3, 4]
    for k in range(3):
        s.round(np.array([[1, -1, 2, 0]], F), o.ASYNC_ORDERED)
    assert s.var.tolist() == [0.25, 2.75, 1.5, 4.0]
    assert s.step == 3


def test_adam_first_step_known_answer():
    """Step 1 with m=v=0 and TF's formula: alpha = lr*sqrt(1-b2)/(1-b1);
    m = (1-b1) g; v = (1-b2) g^2; var -= m*alpha/(sqrt(v)+eps).  For |g| >> eps
    this is var - lr*sign(g) to ~1e-7."""
    s = o.Shard(4, o.ADAM, lr=0.01)
    s.var[:] = [1, 1, 1, 1]
    g = np.array([[0.5, -0.5, 2.0, -8.0]], F)
    s.round(g, o.ASYNC_ORDERED)
    # float64 evaluation of the same algebra
    b1, b2, eps, lr = 0.9, 0.999, 1e-8, 0.01
    alpha = lr * np.sqrt(1 - b2) / (1 - b1)
    m = (1 - b1) * g[0].astype(np.float64)
    v = (1 - b2) * g[0].astype(np.float64) ** 2
    want = 1 - m * alpha / (np.sqrt(v) + eps)
    np.testing.assert_allclose(s.var, want, rtol=0, atol=2e-7)
    np.testing.assert_allclose(s.var, 1 - 0.01 * np.sign(g[0]), atol=1e-6)
    assert s.b1p == F(0.9) * F(0.9) and s.b2p == F(0.999) * F(0.999)
    assert s.step == 1


def test_adam_epsilon_is_outside_bias_correction():
    """TF's 'epsilon hat' form differs measurably from torch.optim.Adam's
    sqrt(v/(1-b2^t))+eps when |g| ~ eps."""
    s = o.Shard(1, o.ADAM, lr=0.01)
    g = np.array([[1e-8]], F)
    s.round(g, o.ASYNC_ORDERED)
    b1, b2, eps, lr = 0.9, 0.999, 1e-8, 0.01
    tf_form = -(lr * np.sqrt(1 - b2) / (1 - b1)) * ((1 - b1) * 1e-8) / (np.sqrt((1 - b2) * 1e-16) + eps)
    torch_form = -lr * 1e-8 / (1e-8 + eps)
    assert abs(s.var[0] - tf_form) < 1e-5 * abs(tf_form)
    assert abs(tf_form - torch_form) > 1e-3      # the two conventions disagree here


def test_async_ordered_sgd_equals_sum_only_to_tolerance():
    rng = np.random.default_rng(0)
    slots = rng.standard_normal((4, 1000)).astype(F)
    a, b = o.Shard(1000, o.SGD, lr=0.1), o.Shard(1000, o.SGD, lr=0.1)
    a.round(slots, o.ASYNC_ORDERED)
    b.round(slots, o.SUM)
    np.testing.assert_allclose(a.var, b.var, rtol=1e-5, atol=1e-6)
    assert a.step == 4 and b.step == 1


def test_sync_mean_is_sum_over_w_divided():
    slots = np.array([[1, 2], [3, 6], [5, 1]], F)
    s = o.Shard(2, o.SGD, lr=1.0)
    s.round(slots, o.SYNC_MEAN)
    assert s.var.tolist() == [-3.0, -3.0]


@pytest.mark.parametrize("opt", [o.SGD, o.ADAM])
@pytest.mark.parametrize("mode", [o.ASYNC_ORDERED, o.SUM, o.SYNC_MEAN])
@pytest.mark.parametrize("W", [1, 2, 5])
def test_numpy_and_c_restatements_agree_bit_for_bit(opt, mode, W):
    rng = np.random.default_rng(10 * W + mode)
    n = 4099
    a, b = o.Shard(n, opt, lr=0.01), o.CShard(n, opt, lr=0.01)
    init = rng.standard_normal(n).astype(F)
    a.var[:] = init
    b.var[:] = init
    for r in range(6):
        scale = F(10.0 ** rng.integers(-6, 3))
        slots = (rng.standard_normal((W, n)) * scale).astype(F)
        a.round(slots, mode)
        b.round(slots, mode)
    assert np.array_equal(a.var, b.var)
    assert np.array_equal(a.m, b.m) and np.array_equal(a.v, b.v)
    assert a.step == b.step and a.b1p == b.b1p and a.b2p == b.b2p


def test_threaded_cpu_ps_round_equals_scalar_oracle():
    n, W = 300000, 3
    base = o.CpuPsBaseline(n, W, o.ADAM, lr=0.01, threads=5)
    ref = o.CShard(n, o.ADAM, lr=0.01)
    ref.var[:] = base.var
    for r in range(3):
        base.round(o.SUM)
        ref.round(np.stack(base.grads), o.SUM)
    assert np.array_equal(base.var, ref.var)
    assert all(np.array_equal(p, base.var) for p in base.params)
    assert base.state[0] == ref.b1p


def test_bf16_rounding_known_answers():
    # ties-to-even at the 8th mantissa bit; 1+2^-8 is a tie -> 1.0 ; 1+3*2^-8 -> 1+2^-6
    x = np.array([1.0, 1.0 + 2 ** -8, 1.0 + 3 * 2 ** -8, -2.5, 3.0e38, 1e-40], F)
    h = o.f32_to_bf16(x)
    back = o.bf16_to_f32(h)
    assert back[0] == 1.0 and back[1] == 1.0 and back[2] == F(1.0 + 2 ** -6)
    assert back[3] == -2.5
    lib = o.c_lib()
    for xi, hi in zip(x, h):
        assert lib.psx_oracle_f32_to_bf16(float(xi)) == int(hi)
    import torch
    t = torch.from_numpy(x).to(torch.bfloat16).view(torch.int16).numpy().view(np.uint16)
    assert np.array_equal(t, h)


def test_placement_tables():
    # mnist_replica.py:121-134 with 2 ps / 1 ps; mnist.py:44-46
    p = o.replica_device_setter_placement
    assert p(o.MNIST_MLP_VARS, 2) == {"global_step": 0, "hid_w": 1, "hid_b": 0, "sm_w": 1, "sm_b": 0}
    assert set(p(o.MNIST_MLP_VARS, 1).values()) == {0}
    assert p(o.MNIST_SOFTMAX_VARS, 2) == {"W": 0, "b": 1, "global_step": 0}
    assert p(o.MNIST_SOFTMAX_VARS, 3) == {"W": 0, "b": 1, "global_step": 2}


def test_model_gradients_against_torch_autograd():
    import torch
    rng = np.random.default_rng(5)
    x = rng.random((100, 784)).astype(F)
    y = np.eye(10, dtype=F)[rng.integers(0, 10, 100)]
    # softmax regression (mnist.py:44-50)
    W = (rng.standard_normal((784, 10)) * 0.01).astype(F)
    b = np.zeros(10, F)
    loss, dW, db = o.softmax_regression_grads(W, b, x, y)
    tW, tb = torch.tensor(W, requires_grad=True), torch.tensor(b, requires_grad=True)
    tl = -(torch.tensor(y) * torch.log(torch.softmax(torch.tensor(x) @ tW + tb, 1))).sum()
    tl.backward()
    np.testing.assert_allclose(loss, tl.item(), rtol=1e-5)
    np.testing.assert_allclose(dW, tW.grad.numpy(), rtol=1e-3, atol=1e-4)
    np.testing.assert_allclose(db, tb.grad.numpy(), rtol=1e-3, atol=1e-4)
    # MLP (mnist_replica.py:124-145)
    hw = o.truncated_normal(np.random.default_rng(1), (784, 100), 1.0 / 28)
    hb = np.zeros(100, F)
    sw = o.truncated_normal(np.random.default_rng(2), (100, 10), 0.1)
    sb = np.zeros(10, F)
    assert np.abs(hw).max() <= 2.0 / 28 + 1e-7
    loss, dhw, dhb, dsw, dsb = o.mlp_grads(hw, hb, sw, sb, x, y)
    ts = [torch.tensor(a, requires_grad=True) for a in (hw, hb, sw, sb)]
    h = torch.relu(torch.tensor(x) @ ts[0] + ts[1])
    p = torch.softmax(h @ ts[2] + ts[3], 1)
    tl = -(torch.tensor(y) * torch.log(torch.clamp(p, 1e-10, 1.0))).sum()
    tl.backward()
    np.testing.assert_allclose(loss, tl.item(), rtol=1e-5)
    for mine, t in zip((dhw, dhb, dsw, dsb), ts):
        np.testing.assert_allclose(mine, t.grad.numpy(), rtol=2e-3, atol=2e-4)
    # NMF (matrix_factorization.py:30-36)
    R = rng.random((50, 40)).astype(F)
    Wm = (rng.random((50, 8)) * 0.5).astype(F)
    Hm = (rng.random((8, 40)) * 0.5).astype(F)
    Wm[0, 0] = -0.1
    loss, dWm, dHm = o.nmf_grads(Wm, Hm, R)
    tWm, tHm = torch.tensor(Wm, requires_grad=True), torch.tensor(Hm, requires_grad=True)
    tl = ((torch.tensor(R) - tWm @ tHm) ** 2).sum() + 10e12 * (
        (tWm.abs() - tWm).sum() + (tHm.abs() - tHm).sum())
    tl.backward()
    np.testing.assert_allclose(dWm, tWm.grad.numpy(), rtol=1e-3, atol=1e-3)
    np.testing.assert_allclose(dHm, tHm.grad.numpy(), rtol=1e-3, atol=1e-3)


def test_sgd_matches_torch_optim_sgd_and_adam_differs_from_torch_adam():
    """plain SGD is the same algebra as torch.optim.SGD (to 1 ulp:
    torch's CPU kernel evaluates p + (-lr)*g with a fused multiply-add, Eigen and
    the oracle round the product first); TF-form Adam is NOT torch.optim.Adam
    (epsilon placement), which is why torch's optimizer is never used as oracle."""
    import torch
    rng = np.random.default_rng(3)
    p0 = rng.standard_normal(1000).astype(F)
    gs = [rng.standard_normal(1000).astype(F) for _ in range(5)]
    s = o.Shard(1000, o.SGD, lr=0.005)
    s.var[:] = p0
    tp = torch.tensor(p0.copy(), requires_grad=True)
    opt = torch.optim.SGD([tp], lr=0.005)
    for g in gs:
        s.round(g[None, :], o.ASYNC_ORDERED)
        tp.grad = torch.tensor(g)
        opt.step()
    np.testing.assert_allclose(s.var, tp.detach().numpy(), rtol=1e-6, atol=1e-8)
    a = o.Shard(1000, o.ADAM, lr=0.01)
    a.var[:] = p0
    tq = torch.tensor(p0.copy(), requires_grad=True)
    topt = torch.optim.Adam([tq], lr=0.01, betas=(0.9, 0.999), eps=1e-8)
    for g in gs:
        small = (g * F(1e-8)).astype(F)           # |g| ~ eps: the two conventions split
        a.round(small[None, :], o.ASYNC_ORDERED)
        tq.grad = torch.tensor(small)
        topt.step()
    assert not np.allclose(a.var, tq.detach().numpy(), rtol=1e-3, atol=0)


def test_cpu_quota_is_read_from_cgroup_v2_and_v1(tmp_path):
    """The CPU arm sizes its thread pool to the container's CPU quota (the GPU boxes
    show 128 CPUs under a 16-core CFS quota: profiles/r30)."""
    (tmp_path / "cpu.max").write_text("1600000 100000\n")
    assert o.cpu_quota_cores(str(tmp_path)) == 16.0
    (tmp_path / "cpu.max").write_text("max 100000\n")
    assert o.cpu_quota_cores(str(tmp_path)) is None
    v1 = tmp_path / "v1"
    (v1 / "cpu").mkdir(parents=True)
    (v1 / "cpu" / "cpu.cfs_quota_us").write_text("-1\n")
    (v1 / "cpu" / "cpu.cfs_period_us").write_text("100000\n")
    assert o.cpu_quota_cores(str(v1)) is None
    (v1 / "cpu" / "cpu.cfs_quota_us").write_text("250000\n")
    assert o.cpu_quota_cores(str(v1)) == 2.5
    assert 1 <= o.usable_threads() <= len(__import__("os").sched_getaffinity(0))


@pytest.mark.parametrize("opt", [o.SGD, o.ADAM])
@pytest.mark.parametrize("mode", [o.SUM, o.SYNC_MEAN])
def test_sparse_rows_round_numpy_and_c_restatements_agree_and_match_a_hand_case(opt, mode):
    """Index-list (IndexedSlices) round, SURVEY 8f-3: the two restatements agree bit
    for bit on overlapping random lists; rows nobody pushed are untouched; a hand
    case: two workers pushing the same row are summed in worker order."""
    n_rows, d, W = 300, 12, 3
    rng = np.random.default_rng(4)
    a = o.Shard(n_rows * d, opt, lr=0.05)
    b = o.CShard(n_rows * d, opt, lr=0.05)
    init = rng.standard_normal(n_rows * d).astype(F)
    a.var[:] = init
    b.var[:] = init
    touched = set()
    for _ in range(3):
        idx, rows = [], []
        for w in range(W):
            i = np.sort(rng.choice(n_rows, size=int(rng.integers(1, 80)), replace=False))
            idx.append(i.astype(np.int64))
            rows.append((rng.standard_normal((i.size, d)) * 0.1).astype(F))
            touched |= set(i.tolist())
        o.rows_round(a, d, idx, rows, mode)
        o.c_rows_round(b, d, idx, rows, mode)
    assert np.array_equal(a.var.view(np.uint32), b.var.view(np.uint32))
    assert np.array_equal(a.m.view(np.uint32), b.m.view(np.uint32))
    assert np.array_equal(a.v.view(np.uint32), b.v.view(np.uint32))
    assert a.step == b.step == 3 and a.b1p == b.b1p and a.b2p == b.b2p
    untouched = sorted(set(range(n_rows)) - touched)
    assert untouched, "the test wants some rows nobody pushed"
    got = a.var.reshape(n_rows, d)[untouched]
    assert np.array_equal(got, init.reshape(n_rows, d)[untouched])
    # hand case (SGD, SUM): row 1 pushed by workers 0 and 2 -> var -= lr * (g0 + g2)
    if opt == o.SGD and mode == o.SUM:
        s = o.Shard(3 * 2, o.SGD, lr=0.5)
        s.var[:] = 1.0
        o.rows_round(s, 2, [np.array([1]), np.array([0]), np.array([1, 2])],
                     [np.array([[0.2, 0.4]], F), np.array([[1.0, 1.0]], F),
                      np.array([[0.6, 0.0], [2.0, 2.0]], F)], o.SUM)
        want = np.array([[0.5, 0.5], [1 - 0.5 * F(0.2 + F(0.6)), 1 - 0.5 * 0.4], [0.0, 0.0]], F)
        np.testing.assert_allclose(s.var.reshape(3, 2), want, rtol=1e-6)
