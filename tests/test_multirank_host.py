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
