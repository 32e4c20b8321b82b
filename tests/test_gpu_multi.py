"""Cross-process (CUDA IPC) and cross-GPU (NVLink P2P) paths of the data plane,
still through the C ABI, still bit-exact against the oracle.

* one GPU, two processes: PS process and worker process share GPU 0 through
  IPC handles shipped over a pipe (what the rendez-vous socket carries)
* two GPUs (skipped on a 1-GPU box): PS on GPU 0, workers on GPU 0 and GPU 1,
  each on its own stream -- the flags + stream memops do all the ordering
"""
import multiprocessing as mp
import traceback

import numpy as np
import pytest

from oracle import ps_oracle as o
from tfmesos_b200 import engine, psx

pytestmark = pytest.mark.gpu
F = np.float32
N = 79510
ROUNDS = 4


def _grad(w, r, n=N):
    return (np.random.default_rng(1000 * w + r).standard_normal(n) * 0.1).astype(F)


def _worker_proc(conn, device, slot, n_workers, mode):
    """Child process = one worker task (tfmesos/server.py would have spawned it)."""
    try:
        import torch
        psx.init(device)
        handle = conn.recv()
        client = psx.Client(handle, device, slot)
        conn.send(client.export())
        assert conn.recv() == "registered"
        grad = torch.zeros(N, device="cuda:%d" % device)
        param = torch.zeros(N, device="cuda:%d" % device)
        stream = torch.cuda.Stream(device=device)
        with torch.cuda.device(device), torch.cuda.stream(stream):
            for r in range(1, ROUNDS + 1):
                grad.copy_(torch.from_numpy(_grad(slot, r)).to(grad.device))
                client.push(grad.data_ptr(), N, seq=r, stream=stream)
                client.pull(param.data_ptr(), N, wait_seq=r, stream=stream)
            stream.synchronize()
        conn.send(param.cpu().numpy())
        assert conn.recv() == "bye"
        client.close()
    except Exception:
        conn.send("ERROR " + traceback.format_exc())


def _run_ps_with_remote_workers(worker_devices, mode):
    import torch
    ctx = mp.get_context("spawn")
    W = len(worker_devices)
    psx.init(0)
    shard = psx.Shard(0, N, psx.OPT_ADAM, lr=0.01, n_slots=W)
    init = np.random.default_rng(5).standard_normal(N).astype(F)
    shard.set_values(psx.VAR, init)
    procs, conns = [], []
    try:
        for w, dev in enumerate(worker_devices):
            parent, child = ctx.Pipe()
            p = ctx.Process(target=_worker_proc, args=(child, dev, w, W, mode))
            p.start()
            procs.append(p)
            conns.append(parent)
            parent.send(shard.export())
        for w, c in enumerate(conns):
            assert c.poll(120), "worker %d never answered" % w
            h = c.recv()
            assert not isinstance(h, str), h
            shard.register_client(w, h)
        for c in conns:
            c.send("registered")
        ps_stream = torch.cuda.Stream(device=0)
        for r in range(1, ROUNDS + 1):
            shard.apply(mode, 0, W, wait_seq=r, stream=ps_stream)
        ref = o.CShard(N, o.ADAM, lr=0.01)
        ref.var[:] = init
        for r in range(1, ROUNDS + 1):
            ref.round(np.stack([_grad(w, r) for w in range(W)]), mode)
        finals = []
        for w, c in enumerate(conns):
            assert c.poll(180), "worker %d hung" % w
            got = c.recv()
            assert not isinstance(got, str), got
            finals.append(got)
            c.send("bye")
        ps_stream.synchronize()
        got_var = shard.get_values(psx.VAR)
        assert np.array_equal(got_var.view(np.uint32), ref.var.view(np.uint32))
        for w, got in enumerate(finals):
            assert np.array_equal(got.view(np.uint32), ref.var.view(np.uint32)), "worker %d" % w
        assert shard.state()["apply_seq"] == ROUNDS
    finally:
        for p in procs:
            p.join(30)
            if p.is_alive():
                p.kill()
        shard.destroy()


@pytest.mark.parametrize("mode", [psx.MODE_SUM, psx.MODE_ASYNC_ORDERED])
def test_ipc_ps_and_two_worker_processes_on_one_gpu(mode):
    _run_ps_with_remote_workers([0, 0], mode)


@pytest.mark.multigpu
@pytest.mark.parametrize("mode", [psx.MODE_SUM, psx.MODE_SYNC_MEAN])
def test_ipc_workers_on_other_gpus_over_nvlink(mode):
    import torch
    devs = [i % torch.cuda.device_count() for i in range(1, 4)]
    _run_ps_with_remote_workers(devs, mode)


@pytest.mark.multigpu
@pytest.mark.parametrize("fused", [False, True])
@pytest.mark.parametrize("stripes", [1, 2])
def test_two_gpus_one_process_bit_exact(fused, stripes):
    """PS bucket on GPU 0 (or striped over GPU 0+1), one worker per GPU."""
    import torch
    psx.init(0)
    variables = [("hid_w", (784, 100)), ("hid_b", (100,)), ("sm_w", (100, 10)), ("sm_b", (10,))]
    ps_devices = [0] if stripes == 1 else [[0, 1]]
    cl = engine.LocalCluster(variables, 1, 2, engine.AdamOptimizer(0.01),
                             ps_devices=ps_devices, worker_devices=[0, 1], fused=fused)
    nb = cl.layout.bucket_nelem[0]
    ref = o.CShard(nb, o.ADAM, lr=0.01)
    rng = np.random.default_rng(21)
    try:
        init = rng.standard_normal(nb).astype(F)
        for name, (task, off, shape, numel) in cl.layout.entries.items():
            cl.set_variable(name, init[off:off + numel].reshape(shape))
            ref.var[off:off + numel] = init[off:off + numel]
        streams = {d: torch.cuda.Stream(device=d) for d in (0, 1)}
        ps_streams = {d: torch.cuda.Stream(device=d) for d in (0, 1)}
        for r in range(1, 4):
            slots = np.zeros((2, nb), F)
            for w in range(2):
                for name, (task, off, shape, numel) in cl.layout.entries.items():
                    g = (rng.standard_normal(numel) * 0.1).astype(F)
                    slots[w, off:off + numel] = g
                    cl.workers[w].grads[name].copy_(torch.from_numpy(g).view(shape))
            torch.cuda.synchronize(0)
            torch.cuda.synchronize(1)
            for w in range(2):
                if fused:
                    cl.workers[w].signal(r, streams[w])
                else:
                    cl.workers[w].push(r, streams[w])
            for key, ps in cl.servers.items():
                st = ps_streams[ps.spec.device]
                (ps.round if fused else ps.apply)(psx.MODE_SUM, r, st)
            for w in range(2):
                if fused:
                    cl.workers[w].wait_applied(r, streams[w])
                else:
                    cl.workers[w].pull(r, streams[w])
            ref.round(slots, o.SUM)
        for d in (0, 1):
            streams[d].synchronize()
            ps_streams[d].synchronize()
        for name, (task, off, shape, numel) in cl.layout.entries.items():
            want = ref.var[off:off + numel].reshape(shape)
            assert np.array_equal(cl.get_variable(name).view(np.uint32), want.view(np.uint32)), name
            for w in range(2):
                got = cl.workers[w].params[name].cpu().numpy()
                assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), (w, name)
    finally:
        cl.close()
