"""GOP-wise sharding and the statistics gather (the one collective of the path), world_size 2 over gloo on CPU."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from lssvc_b200 import gop


def test_units_cover_every_frame_once():
    units = gop.work_units(n_seq=8, n_frames=96, gop_size=32)          # BASELINE.json configs[2]
    assert len(units) == 24
    seen = set()
    for u in units:
        for f, intra in gop.frames_of(u):
            assert (u.seq, f) not in seen
            seen.add((u.seq, f))
            assert intra == (f % 32 == 0)
    assert len(seen) == 8 * 96
    for world in (1, 2, 4, 8):
        shards = [gop.shard(units, world, r) for r in range(world)]
        assert sorted(sum(shards, [])) == sorted(units)
        assert {len(s) for s in shards} == {24 // world}


def test_ragged_last_gop():
    units = gop.work_units(1, 30, 12)
    assert [u.n_frames for u in units] == [12, 12, 6]
    assert gop.work_units(0, 10, 4) == [] and gop.work_units(2, 0, 4) == []


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    units = gop.shard(gop.work_units(3, 10, 4), world, rank)           # 9 units: rank 0 gets 5, rank 1 gets 4 (ragged)
    rows = []
    for u in units:
        for f, intra in gop.frames_of(u):
            rows.append((u.seq, f, float(intra), 100.0 * u.seq + f, 1000.0 + f, 0.5, 0.25))
    table = gop.gather_stats(rows, dist)
    q.put((rank, table))
    dist.barrier()
    dist.destroy_process_group()


def _run_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    try:
        got = dict(q.get(timeout=240) for _ in range(2))
        for p in procs:
            p.join(timeout=60)
            assert p.exitcode == 0
        return got
    finally:
        for p in procs:            # only the exact processes started here
            if p.is_alive():
                p.kill()


def test_gather_stats_world2_gloo():
    try:
        got = _run_world2()
    except Exception:              # a rendezvous port taken between _free_port() and bind, or a cold first import: one retry
        got = _run_world2()
    assert torch.equal(got[0], got[1])
    t = got[0]
    assert t.shape == (30, 7)
    assert [(int(a), int(b)) for a, b in t[:, :2].tolist()] == [(s, f) for s in range(3) for f in range(10)]
    assert torch.equal(t[:, 3], t[:, 0] * 100 + t[:, 1])
    assert int(t[:, 2].sum()) == 9
    s = gop.summarize(t, pixels_el=64, pixels_bl=16)
    assert s["frames"] == 30 and s["bpp_el"] > 0


def test_gather_single_rank_no_dist():
    t = gop.gather_stats([(0, 1, 0, 5.0, 6.0, 0.1, 0.2), (0, 0, 1, 7.0, 8.0, 0.1, 0.2)])
    assert t[:, 1].tolist() == [0.0, 1.0]
    assert gop.gather_stats([]).shape == (0, 7)
