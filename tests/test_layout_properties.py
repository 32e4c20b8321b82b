"""Property tests (hypothesis) of the host-side sharding arithmetic: whatever the
variable list, PS-task count and stripe count, buckets and stripes tile the
parameter space exactly once, on the promised alignments, and the per-rank shard
order is a rotation that never sends two workers to the same GPU at once."""
from hypothesis import given, settings
from hypothesis import strategies as st

from tfmesos_b200 import engine

shapes = st.lists(st.tuples(st.integers(1, 40), st.integers(1, 300)), min_size=1, max_size=12)


@settings(max_examples=60, deadline=None)
@given(shapes=shapes, ps_tasks=st.integers(1, 5))
def test_buckets_hold_every_variable_exactly_once(shapes, ps_tasks):
    variables = [("v%d" % i, s) for i, s in enumerate(shapes)]
    lay = engine.VariableLayout(variables, ps_tasks)
    assert list(lay.placement().values()) == [i % ps_tasks for i in range(len(variables))]
    for task in range(ps_tasks):
        spans = sorted((e[1], e[1] + e[3]) for e in lay.entries.values() if e[0] == task)
        for (a0, a1), (b0, b1) in zip(spans, spans[1:]):
            assert a1 <= b0                                  # no overlap
        assert all(a % engine.ALIGN == 0 for a, _ in spans)  # 128-byte aligned
        assert lay.bucket_nelem[task] == (spans[-1][1] if spans else 0)
    assert sum(e[3] for e in lay.entries.values()) == sum(a * b for a, b in shapes)


@settings(max_examples=100, deadline=None)
@given(n=st.integers(1, 3_000_000), stripes=st.integers(1, 40))
def test_stripes_tile_any_bucket(n, stripes):
    rs = engine.stripe_ranges(n, stripes)
    assert 1 <= len(rs) <= stripes
    assert rs[0][0] == 0 and sum(c for _, c in rs) == n
    pos = 0
    for off, cnt in rs:
        assert off == pos and cnt > 0 and off % engine.STRIPE_ALIGN == 0
        pos += cnt


@settings(max_examples=40, deadline=None)
@given(world=st.integers(1, 8), ps_tasks=st.integers(1, 3), extra=st.integers(0, 2))
def test_rotated_shard_order_is_incast_free(world, ps_tasks, extra):
    """Worker r visits the GPUs in the order r, r+1, ...: at step k of the walk the
    workers target `world` different GPUs (a permutation)."""
    stripes = world * (1 + extra)
    variables = [("p%d" % t, (4096 * stripes, 3)) for t in range(ps_tasks)]
    lay = engine.VariableLayout(variables, ps_tasks)
    topo = engine.torchrun_topology(lay, world, stripes)
    per_gpu = {}
    for s in topo.shards:
        per_gpu.setdefault(s.device, []).append(s.key)
    orders = []
    for r in range(world):
        order = sorted(topo.shards, key=lambda s, r=r: ((s.device - r) % world, s.task, s.stripe))
        assert sorted(s.key for s in order) == sorted(s.key for s in topo.shards)
        orders.append([s.device for s in order])
    counts = {len(v) for v in per_gpu.values()}
    if len(counts) == 1:                       # same number of shards on every GPU
        k = counts.pop()
        for step in range(0, len(topo.shards), k):
            assert sorted(o[step] for o in orders) == list(range(world))
