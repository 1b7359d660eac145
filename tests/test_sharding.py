"""CPU: sharding of (window, pair) units and the record all_gather, world_size 2 over gloo."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from radio_mapper_b200 import sharding


def test_split_even_covers_everything():
    for n in (0, 1, 5, 8, 120, 2016):
        for parts in (1, 2, 3, 8):
            spans = [sharding.split_even(n, parts, i) for i in range(parts)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_shard_units_policy():
    assert sharding.shard_units(5, 120, 1, 0) == (list(range(5)), None)
    # enough windows: shard windows, all pairs
    w0, p0 = sharding.shard_units(8, 2016, 8, 3)
    assert w0 == [3] and p0 is None
    w, p = sharding.shard_units(5, 120, 2, 1)
    assert w == [3, 4] and p is None
    # fewer windows than ranks: every window, a slice of the pairs (252 each for 2016 pairs on 8 ranks)
    w, p = sharding.shard_units(1, 2016, 8, 7)
    assert w == [0] and (p.start, p.stop) == (1764, 2016)
    covered = []
    for r in range(8):
        _, sl = sharding.shard_units(1, 2016, 8, r)
        covered += list(range(sl.start, sl.stop))
    assert covered == list(range(2016))


def _worker(rank, world, port, n_windows, n_pairs, n_buoys, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        full = torch.arange(n_windows * n_pairs * 4, dtype=torch.int32).reshape(n_windows, n_pairs, 4)
        full_en = torch.arange(n_windows * n_buoys, dtype=torch.int64).reshape(n_windows, n_buoys)
        windows, pair_slice = sharding.shard_units(n_windows, n_pairs, world, rank)
        local = full[windows]
        if pair_slice is not None:
            local = local[:, pair_slice].contiguous()
        rec, en = sharding.gather_records(local.contiguous(), full_en[windows].contiguous(), n_windows, n_pairs, world, rank)
        q.put((rank, bool(torch.equal(rec, full)), bool(torch.equal(en, full_en)), sharding.world_and_rank()))
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _run(n_windows, n_pairs, world=2):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_windows, n_pairs, 4, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return sorted(out)


def test_gather_records_by_window_gloo():
    out = _run(n_windows=5, n_pairs=6)
    assert [(r, a, b) for r, a, b, _ in out] == [(0, True, True), (1, True, True)]
    assert out[1][3] == (2, 1)


def test_gather_records_by_pair_gloo():
    out = _run(n_windows=1, n_pairs=7)
    assert [(r, a, b) for r, a, b, _ in out] == [(0, True, True), (1, True, True)]
