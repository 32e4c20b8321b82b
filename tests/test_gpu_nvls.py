"""NVSwitch multicast primitives (experimental, 2 GPUs): multimem.st stores into
every GPU's copy; multimem.ld_reduce returns the sum over the copies (with two
members the sum has one possible order, so it is bit-exact here; with more it is
only tolerance-checked by design -- DESIGN.md section 6)."""
import numpy as np
import pytest

from tfmesos_b200 import psx

pytestmark = [pytest.mark.gpu, pytest.mark.multigpu]
F = np.float32


def _mc(nbytes):
    import torch
    psx.init(0)
    if not (psx.nvls_supported(0) and psx.nvls_supported(1)):
        pytest.skip("this box does not expose NVSwitch multicast")
    try:
        return psx.MulticastBuffer([0, 1], nbytes)
    except RuntimeError as exc:
        pytest.skip("multicast object could not be created here: %s" % str(exc)[:200])


def test_multicast_broadcast_reaches_every_copy_and_reduce_sums_them():
    import torch
    n = 1 << 20
    mc = _mc(n * 4)
    try:
        t = [mc.tensor(0), mc.tensor(1)]
        rng = np.random.default_rng(41)
        a = rng.standard_normal(n).astype(F)
        b = rng.standard_normal(n).astype(F)
        t[0].copy_(torch.from_numpy(a))
        t[1].copy_(torch.from_numpy(b))
        torch.cuda.synchronize(0)
        torch.cuda.synchronize(1)
        for member in (0, 1):
            out = torch.zeros(n, device="cuda:%d" % member)
            with torch.cuda.device(member):
                mc.reduce(member, out.data_ptr(), n * 4)
                torch.cuda.synchronize(member)
            assert np.array_equal(out.cpu().numpy().view(np.uint32), (a + b).view(np.uint32))
        src = torch.from_numpy(rng.standard_normal(n).astype(F)).to("cuda:1")
        with torch.cuda.device(1):
            mc.broadcast(1, src.data_ptr(), n * 4)
            torch.cuda.synchronize(1)
        torch.cuda.synchronize(0)
        for member in (0, 1):
            assert torch.equal(t[member].cpu(), src.cpu()), member
        # a sub-range leaves the rest untouched
        part = torch.full((4096,), 7.0, device="cuda:0")
        with torch.cuda.device(0):
            mc.broadcast(0, part.data_ptr(), 4096 * 4, off=8192 * 4)
            torch.cuda.synchronize(0)
        torch.cuda.synchronize(1)
        want = src.cpu().clone()
        want[8192:8192 + 4096] = 7.0
        for member in (0, 1):
            assert torch.equal(t[member].cpu(), want), member
    finally:
        mc.destroy()


def test_multicast_argument_checks():
    mc = _mc(1 << 16)
    try:
        with pytest.raises(RuntimeError, match="16-byte"):
            mc.broadcast(0, 16, 100, off=0)
        with pytest.raises(RuntimeError, match="member"):
            mc.reduce(5, 16, 16)
    finally:
        mc.destroy()
