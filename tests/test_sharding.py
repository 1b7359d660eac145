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


def test_tile_pairs_partition_and_buoy_sets():
    """Blocks of the pair matrix: every pair on exactly one rank, local indices consistent with the rank's buoy
    list, and at 64 buoys / 8 ranks each rank transforms half of the buoys for ~1/8 of the pairs."""
    for n_buoys, world in [(64, 8), (64, 4), (64, 2), (64, 1), (16, 8), (16, 3), (5, 8), (3, 2), (2, 2)]:
        tiles = sharding.tile_pairs(n_buoys, world)
        assert len(tiles) == world
        full = [(i, j) for i in range(n_buoys) for j in range(i + 1, n_buoys)]
        seen = np.concatenate([t["global_index"] for t in tiles]) if full else np.empty(0, np.int64)
        assert sorted(seen.tolist()) == list(range(len(full)))
        for t in tiles:
            assert t["local_pairs"].shape == (len(t["global_index"]), 2)
            for (li, lj), g in zip(t["local_pairs"], t["global_index"]):
                assert (int(t["buoys"][li]), int(t["buoys"][lj])) == full[int(g)]
    tiles = sharding.tile_pairs(64, 8)
    assert max(len(t["buoys"]) for t in tiles) == 32
    assert max(len(t["global_index"]) for t in tiles) <= 256 and min(len(t["global_index"]) for t in tiles) >= 240


def _tile_worker(rank, world, port, n_buoys, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n_pairs = n_buoys * (n_buoys - 1) // 2
        full = torch.arange(2 * n_pairs * 4, dtype=torch.int32).reshape(2, n_pairs, 4)
        ok = True
        # several tilings in one process, some built and dropped on the fly: the permutation cache is keyed by
        # CONTENT, so a recycled id() or an equal padded width can never return a stale table
        for nb in (n_buoys, n_buoys + 1, n_buoys, n_buoys + 2):
            n_pairs = nb * (nb - 1) // 2
            full = torch.arange(2 * n_pairs * 4, dtype=torch.int32).reshape(2, n_pairs, 4)
            for tiles in (sharding.tile_pairs(nb, world), sharding.tiles_for(nb, world)):
                mine = full[:, torch.from_numpy(tiles[rank]["global_index"])].contiguous()
                got = sharding.gather_tiled_records(mine, tiles, n_pairs, world)
                ok = ok and bool(torch.equal(got, full))
                del tiles
        assert sharding.tiles_for(n_buoys, world) is sharding.tiles_for(n_buoys, world)
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


def test_gather_tiled_records_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_tile_worker, args=(r, 2, port, 9, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert out == [(0, True), (1, True)]
