"""CPU: host-side logic of the render layer -- plan construction and the multi-rank shard /
gather helpers over a world_size-2 gloo group (the N>1 path of bench.py uses the same code)."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_mixing_console_plan_matches_reference_recording():
    import json
    import numpy as np
    from grafx_b200.render import mixing_console_plan, plan_from_dict

    z = np.load(os.path.join(ROOT, "tests", "golden", "render_mix3_3d.npz"))
    ref = plan_from_dict(json.loads(bytes(z["meta"]).decode())["plan"])
    mine = mixing_console_plan(3, ["eq", "compressor", "reverb"])
    assert mine.num_nodes == ref.num_nodes and mine.max_order == ref.max_order
    for a, b in zip(mine.iter_list[1:], ref.iter_list[1:]):
        assert a.node_type == b.node_type
        assert [(r.method, tuple(r.idx)) for r in a.source_reads] == [(r.method, tuple(r.idx)) for r in b.source_reads]
        assert [g.method for g in a.aggregations] == [g.method for g in b.aggregations]
        assert (a.dest_write.method, tuple(a.dest_write.idx)) == (b.dest_write.method, tuple(b.dest_write.idx))
        assert (a.parameter_read.method, tuple(a.parameter_read.idx)) == (b.parameter_read.method, tuple(b.parameter_read.idx))


def test_shard_bounds_cover_and_balance():
    from grafx_b200.render import shard_bounds

    for n in (0, 1, 7, 128, 129):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from grafx_b200.render.parallel import gather_batch, max_over_ranks, shard_batch

    full = torch.arange(7 * 3, dtype=torch.float32).view(7, 3)
    local = shard_batch(full)                       # ragged: 4 + 3 rows
    back = gather_batch(local * 2, total=7)
    ok = torch.equal(back, full * 2) and local.shape[0] == (4 if rank == 0 else 3)
    t = max_over_ranks(10.0 + rank)
    q.put((rank, ok, t))
    dist.destroy_process_group()


def test_gloo_world2_shard_gather_and_max():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 1000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(60)
    assert all(ok for _, ok, _ in res)
    assert all(t == 11.0 for _, _, t in res)


def test_first_order_fold_decision():
    """Host logic of the source fold (render/graph.py:_first_order_folds): offered only when the first render order is one
    processor call on exactly the source slice, without aggregation, and the processor says its first kernel is the
    biquad cascade."""
    import grafx_b200.processors as P
    from grafx_b200.render import mixing_console_plan
    from grafx_b200.render.graph import _first_order_folds
    from grafx_b200.render.plan import RenderData, _AggregationData as Agg, _SingleRenderData as It, _TensorAccessData as Acc

    eq = P.ParametricEqualizer(num_filters=3, processor_channel="stereo", backend="lfilter")
    rd = mixing_console_plan(4, ["eq", "compressor"])
    assert _first_order_folds({"eq": eq, "compressor": P.Compressor()}, rd, 4)
    assert not _first_order_folds({"eq": P.ParametricEqualizer(processor_channel="midside", backend="lfilter")}, rd, 4)
    assert not _first_order_folds({"eq": P.ParametricEqualizer(backend="fsm")}, rd, 4)
    assert not _first_order_folds({"eq": P.Compressor()}, rd, 4)                      # no cascade in front
    assert _first_order_folds({"eq": P.LowPassFilter(backend="lfilter")}, rd, 4)
    assert _first_order_folds({"eq": P.GraphicEqualizer(backend="lfilter")}, rd, 4)
    assert not _first_order_folds({"eq": eq}, rd, 5)                                  # reads only part of the sources
    eq._gfx_no_source_fold = True
    assert not _first_order_folds({"eq": eq}, rd, 4)
    del eq._gfx_no_source_fold
    first = rd.iter_list[1]
    summed = RenderData("beam", rd.num_nodes, rd.max_order, True, [rd.iter_list[0],
        It("eq", first.source_reads, [Agg("sum")], first.parameter_read, first.dest_write)] + list(rd.iter_list[2:]))
    assert not _first_order_folds({"eq": eq}, summed, 4)
    indexed = RenderData("beam", rd.num_nodes, rd.max_order, True, [rd.iter_list[0],
        It("eq", [Acc("index", torch.arange(4))], [Agg("none")], first.parameter_read, first.dest_write)] + list(rd.iter_list[2:]))
    assert not _first_order_folds({"eq": eq}, indexed, 4)
    assert not _first_order_folds({"mix": eq}, rd, 4)                                 # first order is not a processor
