"""NVLink byte counters (ncu nvlrx__/nvltx__) for the kernels that cross GPUs, from
ONE process driving two GPUs -- ncu replays kernels, which a multi-process run with
cross-process flags cannot survive, so the multi-rank bench lines carry algorithmic
NVLink bytes and THIS capture shows what the algorithmic model is worth:

  A. the unicast one-kernel round (k_apply<..,SCATTER,PeerSrc>): a 5e7-element f32
     Adam bucket striped over GPU 0 and GPU 1, one worker per GPU.  Per launch the
     model says: S = 1e8 B of remote gradients in + S of parameters out.
  B. the NVLS primitives (k_mc_reduce = multimem.ld_reduce, k_mc_broadcast =
     multimem.st) on a 2-GPU multicast buffer: the model behind the B(1 + 1/N)
     accounting of the NVLS round says the requester's OWN copy also travels to the
     switch (egress S) and the multicast store comes back to the sender (ingress S).

    ncu --metrics gpu__time_duration.sum,nvlrx__bytes.sum,nvltx__bytes.sum,\\
nvlrx__bytes_data_user.sum,nvltx__bytes_data_user.sum,dram__bytes_read.sum,dram__bytes_write.sum \\
        --clock-control none -k regex:'k_apply|k_mc_' --csv --log-file out.csv \\
        python tools/prof_nvlink.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

from tfmesos_b200 import engine, psx  # noqa: E402

N = 50_000_000


def main():
    assert torch.cuda.device_count() >= 2
    psx.init(0)
    cl = engine.LocalCluster([("W", (N,))], 1, 2, engine.AdamOptimizer(0.01),
                             ps_devices=[[0, 1]], worker_devices=[0, 1], fused=True)
    for w in cl.workers:
        w.grad_flat[0].fill_(0.01 * (w.index + 1))
    streams = {d: torch.cuda.Stream(device=d) for d in (0, 1)}
    ps_streams = {d: torch.cuda.Stream(device=d) for d in (0, 1)}
    for r in range(1, 3):
        torch.cuda.synchronize(0)
        torch.cuda.synchronize(1)
        for w in range(2):
            cl.workers[w].signal(r, streams[w])
        for key, ps in cl.servers.items():
            ps.round(psx.MODE_SUM, r, ps_streams[ps.spec.device])
        for w in range(2):
            cl.workers[w].wait_applied(r, streams[w])
        for d in (0, 1):
            streams[d].synchronize()
            ps_streams[d].synchronize()
    print("unicast round: 2 rounds, stripe = %d elements per GPU" % cl.topo.shards[0].nelem)
    cl.close()

    if all(psx.nvls_supported(d) for d in (0, 1)):
        nbytes = 256 << 20
        mc = psx.MulticastBuffer([0, 1], nbytes)
        for d in (0, 1):
            mc.tensor(d).fill_(1.0)
        dst = torch.zeros(nbytes // 4, device="cuda:0")
        src = torch.ones(nbytes // 4, device="cuda:0")
        s0 = torch.cuda.Stream(device=0)
        for d in (0, 1):
            torch.cuda.synchronize(d)
        mc.reduce(0, dst.data_ptr(), nbytes, 0, s0)       # GPU 0 reduces the whole buffer
        s0.synchronize()
        mc.broadcast(0, src.data_ptr(), nbytes, 0, s0)    # GPU 0 multicasts the whole buffer
        s0.synchronize()
        for d in (0, 1):
            torch.cuda.synchronize(d)
        print("nvls primitives: %d bytes, reduce sum check %.1f" % (nbytes, float(dst[0])))
        mc.destroy()


if __name__ == "__main__":
    main()
