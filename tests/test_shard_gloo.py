"""Multi-process host logic on CPU (gloo, world_size 2): stream / station sharding with no data-path collective,
result gathering to rank 0 and the max-over-ranks timing rule.  The per-rank 'work' is the oracle's CPU chain on the
rank's own streams (the CUDA product cannot run here), which also shows that a sharded run reproduces the
single-process result stream by stream."""
import os
import socket
import sys

import numpy as np
import pytest

from conftest import ROOT, RATES, load_package


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_streams, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import importlib
    import torch.distributed as dist
    load_package()
    shard = importlib.import_module("radiofm_b200.shard")
    synth = importlib.import_module("radiofm_b200.synth")
    from oracle import port as oport
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard.shard_range(n_streams, rank, world)
    fs, ds, blk = RATES["1.0M"]
    sums = {}
    for s in range(lo, hi):
        iq, _ = synth.make_station_u8(fs, blk, stream_id=s, rds=False)
        d = oport.OracleFmDecoder(fs, -0.15 * fs, downsample=ds)
        a = d.process_u8(iq)
        sums[s] = (int(a.size), float(np.abs(a).sum()))
    dist.barrier()
    t = shard.max_over_ranks(dist, 1.0 + rank)          # slowest rank defines the step time
    gathered = shard.gather_to_root(dist, {"range": (lo, hi), "sums": sums, "tmax": t})
    if rank == 0:
        q.put(gathered)
    dist.barrier()
    dist.destroy_process_group()


def test_shard_ranges_cover_exactly_once():
    shard = __import__("importlib").import_module("radiofm_b200.shard") if load_package() else None
    for n in (0, 1, 2, 7, 100, 4096, 4097):
        for world in (1, 2, 3, 4, 8):
            seen = []
            for r in range(world):
                lo, hi = shard.shard_range(n, r, world)
                assert 0 <= lo <= hi <= n
                seen += list(range(lo, hi))
            assert seen == list(range(n))
            sizes = [shard.shard_range(n, r, world)[1] - shard.shard_range(n, r, world)[0] for r in range(world)]
            assert max(sizes) - min(sizes) <= 1
            for u in (0, n // 2, n - 1):
                if 0 <= u < n:
                    lo, hi = shard.shard_range(n, shard.owner_of(u, n, world), world)
                    assert lo <= u < hi
    with pytest.raises(ValueError):
        shard.shard_range(4, 2, 2)


def test_two_rank_sharded_run_matches_single_process(port, synth):
    import torch.multiprocessing as mp
    n_streams, world = 5, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    p = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, p, n_streams, q)) for r in range(world)]
    for pr in procs:
        pr.start()
    gathered = q.get(timeout=180)
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    assert [g["range"] for g in gathered] == [(0, 2), (2, 5)]
    assert all(g["tmax"] == 2.0 for g in gathered)
    merged = {}
    for g in gathered:
        assert not (set(g["sums"]) & set(merged))
        merged.update(g["sums"])
    assert sorted(merged) == list(range(n_streams))
    fs, ds, blk = RATES["1.0M"]
    for s in range(n_streams):
        iq, _ = synth.make_station_u8(fs, blk, stream_id=s, rds=False)
        a = port.OracleFmDecoder(fs, -0.15 * fs, downsample=ds).process_u8(iq)
        assert merged[s] == (int(a.size), float(np.abs(a).sum()))
