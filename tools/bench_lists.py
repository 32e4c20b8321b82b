"""Micro-benchmark of the tensor-list path on a ResNet-50-shaped parameter list
(161 tensors, 25.5 M f32): one TMA launch vs one ld/st launch vs 161 per-variable
psx_push calls vs the flat bucket copy.  CUDA events, L2 flushed between reps."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

from tfmesos_b200 import engine, psx  # noqa: E402


def resnet50_shapes():
    try:
        import torchvision
        m = torchvision.models.resnet50()
        return [(n.replace(".", "_"), tuple(p.shape)) for n, p in m.named_parameters()]
    except Exception:
        shapes, c_in = [("conv1", (64, 3, 7, 7)), ("bn1_w", (64,)), ("bn1_b", (64,))], 64
        for li, (c, blocks) in enumerate([(64, 3), (128, 4), (256, 6), (512, 3)]):
            for b in range(blocks):
                pre = "l%d_%d_" % (li, b)
                shapes += [(pre + "c1", (c, c_in, 1, 1)), (pre + "n1w", (c,)), (pre + "n1b", (c,)),
                           (pre + "c2", (c, c, 3, 3)), (pre + "n2w", (c,)), (pre + "n2b", (c,)),
                           (pre + "c3", (4 * c, c, 1, 1)), (pre + "n3w", (4 * c,)), (pre + "n3b", (4 * c,))]
                if b == 0:
                    shapes += [(pre + "ds", (4 * c, c_in, 1, 1)), (pre + "dsw", (4 * c,)), (pre + "dsb", (4 * c,))]
                c_in = 4 * c
        shapes += [("fc_w", (1000, 2048)), ("fc_b", (1000,))]
        return shapes


def main():
    psx.init(0)
    shapes = resnet50_shapes()
    cl = engine.LocalCluster(shapes, 1, 1, engine.GradientDescentOptimizer(0.1))
    wk = cl.workers[0]
    tensors = {n: torch.randn(s, device="cuda") for n, s in shapes}
    bind = engine.TensorListBinding(wk, tensors)
    nbytes = sum(t.numel() for t in tensors.values()) * 4
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    client = wk.clients[(0, 0)]

    def per_variable():
        for name, (task, off, shape, numel) in cl.layout.entries.items():
            client.push(tensors[name].data_ptr(), numel, off)

    cases = {
        "list_tma_push": lambda: bind.push(0, True),
        "list_ldst_push": lambda: bind.push(0, False),
        "list_tma_pull": lambda: bind.pull(0, True),
        "list_ldst_pull": lambda: bind.pull(0, False),
        "per_variable_push_161_launches": per_variable,
        "flat_bucket_push": lambda: wk.push(0),
    }
    out = {"tensors": len(shapes), "bytes": nbytes}
    for name, fn in cases.items():
        for _ in range(3):
            fn()
        times = []
        for _ in range(10):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1))
        times.sort()
        ms = times[len(times) // 2]
        out[name] = {"ms": ms, "GBps_read_plus_write": 2 * nbytes / ms / 1e6}
    print(json.dumps(out))
    bind.close()
    cl.close()


if __name__ == "__main__":
    main()
