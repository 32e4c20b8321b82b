"""N>1 host logic on CPU with the gloo backend, world_size 2: every rank derives
the same topology, every shard has exactly one hosting rank, the handle exchange
delivers every rank's blobs to every rank; and the reference arm of bench.py
obeys the torchrun contract (rank 0 prints one JSON line, the others exit 0)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import json, os, sys
sys.path.insert(0, %r)
import torch.distributed as dist
from tfmesos_b200 import engine
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
layout = engine.VariableLayout([("W", (100000, 20)), ("H", (20, 1000))], 2, {"W": 0, "H": 1})
topo = engine.torchrun_topology(layout, world, stripes=int(os.environ["STRIPES"]))
mine = {s.key: ("handle-of-%%d-%%d-from-rank-%%d" %% (s.task, s.stripe, rank)).encode()
        for s in topo.shards_on(rank)}
merged = engine.merge_across_ranks(mine)
clients = engine.merge_across_ranks({(k, rank): b"c%%d" %% rank for k in merged})
out = {"rank": rank, "shards": [repr(s) for s in topo.shards],
       "hosted": sorted(map(list, mine)), "merged": sorted(map(list, merged)),
       "owners": {"%%d/%%d" %% k: v.decode().rsplit("-", 1)[1] for k, v in merged.items()},
       "n_clients": len(clients)}
with open(os.path.join(os.environ["OUT_DIR"], "rank%%d.json" %% rank), "w") as f:
    json.dump(out, f)
dist.barrier()
dist.destroy_process_group()
''' % ROOT


@pytest.mark.parametrize("stripes", [1, 2, 5])
def test_topology_and_handle_exchange_world_size_2(tmp_path, stripes):
    script = tmp_path / "w.py"
    script.write_text(WORKER)
    env = dict(os.environ, STRIPES=str(stripes), OUT_DIR=str(tmp_path))
    port = 29600 + stripes
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
                        "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", str(port), str(script)],
                       env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=240)
    assert r.returncode == 0, r.stderr.decode()[-2000:]
    a = json.load(open(tmp_path / "rank0.json"))
    b = json.load(open(tmp_path / "rank1.json"))
    assert a["shards"] == b["shards"]                       # deterministic topology
    assert a["merged"] == b["merged"] == sorted(a["hosted"] + b["hosted"])
    assert not [k for k in a["hosted"] if k in b["hosted"]]  # exactly one host per shard
    assert len(a["merged"]) == len(a["shards"])
    assert a["owners"] == b["owners"]
    for key, owner in a["owners"].items():
        t, j = map(int, key.split("/"))
        assert int(owner) == (t + j) % 2                     # stripe j of task t -> GPU (t+j) mod N
    assert a["n_clients"] == 2 * len(a["shards"])


def test_reference_arm_under_torchrun_prints_one_line_from_rank0():
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
                        "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29650", os.path.join(ROOT, "bench.py"),
                        "--impl", "reference", "--gpus", "2", "--steps", "2", "--warmup", "1",
                        "--cpu-sample-elems", "2000000"],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=300)
    assert r.returncode == 0, r.stderr.decode()[-2000:]
    lines = [l for l in r.stdout.decode().splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 2 and d["metric"] == "ps_push_pull_GBps"
    assert d["cpu_baseline"]["kind"] == "port" and d["e2e"]["h2d_bytes_per_step"] == 0
    assert "2 worker" in d["cpu_baseline"]["sample"]


def test_gpu_arm_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("box has a GPU")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1"],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=120)
    assert r.returncode != 0
    assert b"no CUDA device" in r.stderr + r.stdout


ARENA_WORKER = r'''
import json, os, sys
sys.path.insert(0, %r)
import torch.distributed as dist
from tfmesos_b200 import engine, psx
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
FAIL = os.environ.get("FAIL_RANK", "")


class FakeMember(object):
    """psx.McMember without CUDA: the 'multicast object' is a pipe whose write end the
    creator keeps; importing = receiving a descriptor that refers to the same pipe."""
    def __init__(self, fd, created):
        self.fd, self.created, self.added = fd, created, False
    @classmethod
    def create(cls, device, n, nbytes):
        if FAIL == "create":
            raise RuntimeError("cuMulticastCreate failed (simulated)")
        r, w = os.pipe()
        os.write(w, b"mc-object-of-%%d-members" %% n)
        cls.keep = w
        return cls(r, True)
    @classmethod
    def import_fd(cls, device, n, nbytes, fd):
        if FAIL == str(rank):
            raise RuntimeError("cuMemImportFromShareableHandle failed (simulated)")
        m = cls(os.dup(fd), False)
        return m
    def add_device(self):
        self.added = True
    def bind(self):
        return 0x1000, 0x2000, 1 << 20
    def destroy(self):
        os.close(self.fd)


psx.McMember = FakeMember


def bcast(obj):
    box = [obj]
    dist.broadcast_object_list(box, src=0)
    return box[0]


out = {"rank": rank}
try:
    arena = engine.McArena(0, 1 << 20, rank, world, bcast)
    out["err"] = None
except (RuntimeError, OSError) as exc:
    arena, out["err"] = None, str(exc)
errs = [None] * world
dist.all_gather_object(errs, out["err"])
out["all_errs"] = errs
if arena is not None and not any(errs):
    # every member's descriptor refers to the SAME object (fstat identity of the pipe)
    st = os.fstat(arena.mcx.fd)
    ids = [None] * world
    dist.all_gather_object(ids, (st.st_dev, st.st_ino))
    out["same_object"] = len(set(ids)) == 1
    out["added"] = arena.mcx.added
    arena.bind()
    out["size"] = arena.size
    arena.destroy()
elif arena is not None:
    arena.destroy()
with open(os.path.join(os.environ["OUT_DIR"], "arena%%d.json" %% rank), "w") as f:
    json.dump(out, f)
dist.barrier()
dist.destroy_process_group()
''' % ROOT


def _run_arena(tmp_path, world, fail, port):
    script = tmp_path / "a.py"
    script.write_text(ARENA_WORKER)
    env = dict(os.environ, OUT_DIR=str(tmp_path), FAIL_RANK=fail)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
                        "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
                        "--master-port", str(port), str(script)],
                       env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=240)
    assert r.returncode == 0, r.stderr.decode()[-2500:]
    return [json.load(open(tmp_path / ("arena%d.json" % k))) for k in range(world)]


def test_multicast_descriptor_reaches_every_rank_over_scm_rights(tmp_path):
    """engine.McArena's plumbing without CUDA: rank 0 'creates' the multicast object,
    its descriptor travels to the other PROCESSES over an abstract AF_UNIX socket
    (SCM_RIGHTS), every rank ends up holding the same kernel object and adds itself."""
    outs = _run_arena(tmp_path, 3, "", 29671)
    assert all(o["err"] is None for o in outs)
    assert all(o["same_object"] and o["added"] and o["size"] == 1 << 20 for o in outs)


@pytest.mark.parametrize("fail", ["create", "2"])
def test_a_failing_rank_does_not_leave_the_others_waiting(tmp_path, fail):
    """NVLS is chosen once, at set-up: if building the team fails on ANY rank (here:
    the creator, or an importer), every rank gets out of the exchange and learns about
    it (TorchrunCluster._agree turns that into NvlsUnavailable on all of them)."""
    outs = _run_arena(tmp_path, 3, fail, 29672 if fail == "create" else 29673)
    assert any(o["err"] for o in outs)
    assert all(any(o["all_errs"]) for o in outs)          # everybody knows
