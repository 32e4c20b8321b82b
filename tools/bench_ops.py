"""Cost of the synchronisation primitives themselves (1 GPU): host time per ABI
call and device time per op, for k_signal, satisfied cuStreamWaitValue32 waits,
and a tiny psx_round."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

from tfmesos_b200 import engine, psx  # noqa: E402


def main():
    psx.init(0)
    cl = engine.LocalCluster([("a", (4096,))], 1, 1, engine.AdamOptimizer(0.01), fused=True)
    c = cl.workers[0].clients[(0, 0)]
    sh = cl.servers[(0, 0)].shard
    st = torch.cuda.Stream()
    out = {}

    def measure(name, fn, n=200):
        for _ in range(20):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record(st)
        for _ in range(n):
            fn()
        e1.record(st)
        host = (time.perf_counter() - t0) / n * 1e6
        torch.cuda.synchronize()
        out[name] = {"host_us_per_call": host, "device_us_per_op": e0.elapsed_time(e1) * 1e3 / n}

    seq = [0]

    def sig():
        seq[0] += 1
        c.signal(seq[0], st)

    measure("k_signal", sig)
    measure("wait_applied_already_satisfied", lambda: c.wait_applied(0, st))
    measure("wait_slots_already_satisfied", lambda: sh.wait_slots(0, 1, 1, st))

    def rnd():
        seq[0] += 1
        c.signal(seq[0], st)
        sh.round(psx.MODE_SUM, 0, 1, seq[0], st)
        c.wait_applied(0, st)

    measure("signal+round(4096 elems)+wait", rnd)
    batch = psx.Batch([dict(op=psx.OP_SIGNAL, id=c.id, stream=st),
                       dict(op=psx.OP_ROUND, id=sh.id, a=psx.MODE_SUM, b=0, c=1, stream=st),
                       dict(op=psx.OP_WAIT_APPLIED, id=c.id, stream=st, uses_seq=False)])

    def rnd_batch():
        seq[0] += 1
        batch.run(seq[0])

    measure("same round as one psx_batch", rnd_batch)
    g = torch.zeros(4096, device="cuda")
    measure("torch tiny kernel (g.add_(1)) for scale", lambda: g.add_(1))
    print(json.dumps(out))
    cl.close()


if __name__ == "__main__":
    main()
