"""CPU parameter server over loopback gRPC -- the transport the reference selects
(``protocol='grpc'``, tfmesos/scheduler.py:186; ``tf.train.Server``,
tfmesos/server.py:52-61).  TEST / BASELINE INFRASTRUCTURE ONLY: imported by
tests/ and by bench.py's CPU legs, never by tfmesos_b200/.

What TensorFlow 0.12 does per worker step and what is mirrored:
one RPC per variable per direction -- ``Pull(name) -> bytes`` (RecvTensor of the
variable), ``Push(name, bytes)`` (RecvTensor of its gradient) -- raw tensor
bytes as payload (no protobuf TensorProto encode: a favour to the baseline),
and the apply on the PS host cores with the oracle's C kernels (the Eigen
expressions' algebra), one apply per pushed gradient (async, use_locking=False).
The PS is its own OS process, like a ps task.

This is "a CPU restatement of the TF-0.12 PS path, not TensorFlow itself"
(BASELINE.md section 3); the memcpy-transport variant in ps_oracle.c is the best
case of the same path.
"""
import ctypes
import multiprocessing as mp
import struct
import time
from concurrent import futures

import numpy as np

F = np.float32
_OPTS = [("grpc.max_receive_message_length", -1), ("grpc.max_send_message_length", -1)]


def _serve(conn, variables, opt_adam, lr, threads):
    """PS task: {name: numel}; state in numpy, arithmetic in ps_oracle.c."""
    import grpc

    from oracle import ps_oracle as o
    lib = o.c_lib()
    state = {}
    for name, n in variables.items():
        state[name] = {"var": np.zeros(n, F), "m": np.zeros(n, F), "v": np.zeros(n, F),
                       "pow": np.array([0.9, 0.999], F)}
    fp = ctypes.POINTER(ctypes.c_float)

    def ptr(a):
        return a.ctypes.data_as(fp)

    def split(req):
        (k,) = struct.unpack_from(">I", req, 0)
        return req[4:4 + k].decode(), memoryview(req)[4 + k:]

    def push(req, ctx):
        name, payload = split(req)
        st = state[name]
        g = np.frombuffer(payload, F)
        if opt_adam:
            lib.psx_oracle_adam(ptr(st["var"]), ptr(st["m"]), ptr(st["v"]), ptr(g), g.size,
                                lr, 0.9, 0.999, 1e-8, float(st["pow"][0]), float(st["pow"][1]))
            st["pow"][0] *= F(0.9)
            st["pow"][1] *= F(0.999)
        else:
            lib.psx_oracle_sgd(ptr(st["var"]), ptr(g), g.size, lr)
        return b"ok"

    def pull(req, ctx):
        name, _ = split(req)
        return state[name]["var"].tobytes()

    def assign(req, ctx):
        name, payload = split(req)
        state[name]["var"][:] = np.frombuffer(payload, F)
        return b"ok"

    class Handler(grpc.GenericRpcHandler):
        def service(self, details):
            fn = {"/ps/Push": push, "/ps/Pull": pull, "/ps/Assign": assign}.get(details.method)
            return grpc.unary_unary_rpc_method_handler(fn) if fn else None

    server = grpc.server(futures.ThreadPoolExecutor(max_workers=max(2, threads)), options=_OPTS)
    server.add_generic_rpc_handlers((Handler(),))
    port = server.add_insecure_port("127.0.0.1:0")
    server.start()
    conn.send(port)
    conn.recv()                 # "stop"
    server.stop(0)


class GrpcCpuPs(object):
    """One PS task process + an in-process worker stub."""

    def __init__(self, variables, opt_adam=True, lr=0.01, threads=4):
        import grpc
        ctx = mp.get_context("spawn")
        self.parent, child = ctx.Pipe()
        self.proc = ctx.Process(target=_serve, args=(child, dict(variables), opt_adam, lr, threads))
        self.proc.start()
        assert self.parent.poll(120), "grpc PS did not start"
        port = self.parent.recv()
        self.channel = grpc.insecure_channel("127.0.0.1:%d" % port, options=_OPTS)
        self._push = self.channel.unary_unary("/ps/Push")
        self._pull = self.channel.unary_unary("/ps/Pull")
        self._assign = self.channel.unary_unary("/ps/Assign")
        self.variables = dict(variables)

    @staticmethod
    def _frame(name, payload=b""):
        nb = name.encode()
        return struct.pack(">I", len(nb)) + nb + payload

    def assign(self, name, value):
        self._assign(self._frame(name, np.ascontiguousarray(value, F).tobytes()))

    def push(self, name, grad):
        """PUSH one gradient: serialise, RPC, apply on the PS."""
        self._push(self._frame(name, np.ascontiguousarray(grad, F).tobytes()))

    def pull(self, name):
        """PULL one variable into a fresh worker-side array."""
        return np.frombuffer(self._pull(self._frame(name)), F)

    def step(self, grads):
        """One worker step: push every gradient, pull every variable (one RPC per
        variable per direction)."""
        for name, g in grads.items():
            self.push(name, g)
        return {name: self.pull(name) for name in grads}

    def close(self):
        try:
            self.channel.close()
            self.parent.send("stop")
        finally:
            self.proc.join(10)
            if self.proc.is_alive():
                self.proc.kill()


def time_round(nelem, steps=3, warmup=1, opt_adam=True):
    """push+pull GB/s of the gRPC CPU-PS path for one variable of ``nelem`` f32."""
    ps = GrpcCpuPs({"v": nelem}, opt_adam=opt_adam)
    try:
        g = (np.random.default_rng(7).standard_normal(nelem) * 1e-2).astype(F)
        for _ in range(warmup):
            ps.step({"v": g})
        t0 = time.perf_counter()
        for _ in range(steps):
            ps.step({"v": g})
        dt = (time.perf_counter() - t0) / steps
    finally:
        ps.close()
    return {"value": nelem * 8 / dt / 1e9, "unit": "GB/s", "ms_per_step": dt * 1e3,
            "sample": "%d parameters, 1 worker, loopback gRPC raw-bytes payloads, "
                      "one RPC per variable per direction" % nelem}
