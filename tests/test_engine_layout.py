"""Host-side sharding logic (no GPU): variable placement is the reference's
round-robin rule, buckets and stripes tile the parameter space exactly."""
import pytest

from tfmesos_b200 import engine
from oracle import ps_oracle as o

MLP = [("global_step", ()), ("hid_w", (784, 100)), ("hid_b", (100,)),
       ("sm_w", (100, 10)), ("sm_b", (10,))]


@pytest.mark.parametrize("ps_tasks", [1, 2, 3, 5])
def test_placement_equals_oracle_round_robin(ps_tasks):
    lay = engine.VariableLayout(MLP, ps_tasks)
    names = [n for n, _ in MLP]
    assert dict(lay.placement()) == o.replica_device_setter_placement(names, ps_tasks)


def test_replica_device_setter_accepts_cluster_like_the_reference():
    place = engine.replica_device_setter(cluster={"ps": ["a:1", "b:2"], "worker": ["c:3"]})
    assert [place(n) for n in "abcd"] == [0, 1, 0, 1]
    assert engine.replica_device_setter(ps_tasks=0)("x") is None


def test_explicit_device_pinning_like_matrix_factorization():
    lay = engine.VariableLayout([("W", (1000, 200)), ("H", (200, 1000))], 2,
                                placement=o.NMF_PLACEMENT)
    assert dict(lay.placement()) == {"W": 0, "H": 1}
    assert lay.bucket_nelem == [200000, 200000]


def test_buckets_do_not_overlap_and_are_aligned():
    lay = engine.VariableLayout(MLP, 2)
    for task in range(2):
        spans = sorted((e[1], e[1] + e[3]) for e in lay.entries.values() if e[0] == task)
        for (a0, a1), (b0, b1) in zip(spans, spans[1:]):
            assert a1 <= b0
        assert all(s[0] % engine.ALIGN == 0 for s in spans)
        assert lay.bucket_nelem[task] == spans[-1][1]


@pytest.mark.parametrize("n,stripes", [(1, 1), (7850, 2), (79510, 4), (200_000_000, 8),
                                       (1024, 8), (25_557_032, 6)])
def test_stripes_tile_the_bucket(n, stripes):
    rs = engine.stripe_ranges(n, stripes)
    assert 1 <= len(rs) <= stripes
    assert rs[0][0] == 0 and sum(c for _, c in rs) == n
    for (a, ca), (b, cb) in zip(rs, rs[1:]):
        assert a + ca == b and b % engine.STRIPE_ALIGN == 0


def test_topology_is_deterministic_and_pins_shards():
    lay = engine.VariableLayout([("W", (1000, 200)), ("H", (200, 1000))], 2,
                                placement={"W": 0, "H": 1})
    t1 = engine.Topology(lay, [[0, 1], 1], [0, 1, 2, 3])
    t2 = engine.Topology(lay, [[0, 1], 1], [0, 1, 2, 3])
    assert [repr(s) for s in t1.shards] == [repr(s) for s in t2.shards]
    assert [(s.task, s.stripe, s.device) for s in t1.shards] == [(0, 0, 0), (0, 1, 1), (1, 0, 1)]
    assert t1.n_workers == 4 and len(t1.shards_on(1)) == 2


def test_torchrun_topology_default_and_dedicated_ps_ranks():
    """Default: every bucket striped over all ranks, worker r on rank r.  Explicit
    roles: BASELINE config #3 as written -- 2 ps + 4 workers in tfrun's first-fit
    order (ps tasks first: /root/reference/script/tfrun:58-75, scheduler.py:252-275)."""
    lay = engine.VariableLayout([("W", (1000, 200)), ("H", (200, 1000))], 2,
                                placement={"W": 0, "H": 1})
    t = engine.torchrun_topology(lay, 4)
    assert [s.device for s in t.shards_of(0)] == [0, 1, 2, 3]
    assert [s.device for s in t.shards_of(1)] == [1, 2, 3, 0]
    assert t.worker_devices == [0, 1, 2, 3]
    d = engine.torchrun_topology(lay, 8, ps_ranks=[0, 1], worker_ranks=[2, 3, 4, 5])
    assert [(s.task, s.device) for s in d.shards] == [(0, 0), (1, 1)]
    assert d.worker_devices == [2, 3, 4, 5] and d.n_workers == 4
    assert d.shards_on(2) == [] and len(d.shards_on(0)) == 1
    # whole-variable placement AND striping over the two PS GPUs
    s2 = engine.torchrun_topology(lay, 8, ps_ranks=[[0, 1], [1, 0]], worker_ranks=[2, 3, 4, 5])
    assert [(s.task, s.stripe, s.device) for s in s2.shards] == \
        [(0, 0, 0), (0, 1, 1), (1, 0, 1), (1, 1, 0)]
    # more stripes than listed ranks cycle over them (pipelining granularity)
    s4 = engine.torchrun_topology(lay, 8, stripes=4, ps_ranks=[[0, 1], [1]],
                                  worker_ranks=[2, 3])
    assert [s.device for s in s4.shards_of(0)] == [0, 1, 0, 1]
