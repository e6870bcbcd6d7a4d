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
