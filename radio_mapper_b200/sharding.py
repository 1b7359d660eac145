"""Work sharding over the ranks of one box (one process per GPU, torch.distributed).

The (window, pair) units of the correlate stage are independent given the spectra, so there
is no data-path collective: ranks take disjoint windows when there are at least as many
windows as ranks (zero redundant forward FFTs), otherwise disjoint contiguous slices of the
i<j pair list (each rank recomputes the forward FFTs of its window — cheaper than receiving
spectra over NVLink).  Only the 16-byte peak records are exchanged, with one
all_gather (NCCL on GPUs; gloo in the CPU tests).
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import numpy as np
import torch
import torch.distributed as dist


def world_and_rank() -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(), dist.get_rank()
    return 1, 0


def split_even(n: int, parts: int, index: int) -> Tuple[int, int]:
    """[start, stop) of slice `index` when n items are dealt into `parts` contiguous slices whose
    sizes differ by at most one (larger slices first)."""
    base, extra = divmod(n, parts)
    start = index * base + min(index, extra)
    return start, start + base + (1 if index < extra else 0)


def shard_units(n_windows: int, n_pairs: int, world: int, rank: int) -> Tuple[List[int], Optional[slice]]:
    """-> (windows this rank processes, pair slice or None for all pairs)."""
    if world <= 1:
        return list(range(n_windows)), None
    if n_windows >= world:
        a, b = split_even(n_windows, world, rank)
        return list(range(a, b)), None
    # fewer windows than ranks: every rank takes every window, and a slice of the pairs
    a, b = split_even(n_pairs, world, rank)
    return list(range(n_windows)), slice(a, b)


def _all_gather_flat(flat: torch.Tensor, world: int) -> torch.Tensor:
    """ONE collective: every rank contributes `flat` (1-D, same length everywhere); returns [world, len]."""
    buf = flat.new_empty((world, flat.numel()))
    try:
        dist.all_gather_into_tensor(buf, flat)
    except (RuntimeError, AttributeError, NotImplementedError):          # backend without the tensor form
        dist.all_gather(list(buf.unbind(0)), flat)
    return buf


def gather_records(rec: torch.Tensor, energy: torch.Tensor, n_windows: int, n_pairs: int, world: int, rank: int):
    """Assemble the full [W, P, 4] record tensor (and [W, B] energies) on every rank with ONE all-gather.

    rec: this rank's [w_local, p_local, 4] int32; energy: [w_local, B] int64.  Records and energies travel in one
    padded int32 buffer (an int64 energy is two int32 words), so a step costs a single latency-bound collective."""
    by_window = n_windows >= world
    if by_window:
        sizes = [split_even(n_windows, world, r) for r in range(world)]
        width = max(b - a for a, b in sizes)
        n_rec = width * rec.shape[1] * rec.shape[2]
        n_en = width * energy.shape[1] * 2
        flat = rec.new_zeros(n_rec + n_en)
        flat[: rec.numel()] = rec.reshape(-1)
        flat[n_rec: n_rec + 2 * energy.numel()] = energy.contiguous().view(torch.int32).reshape(-1)
        buf = _all_gather_flat(flat, world)
        recs = buf[:, :n_rec].reshape(world, width, rec.shape[1], rec.shape[2])
        ens = buf[:, n_rec:].contiguous().view(torch.int64).reshape(world, width, energy.shape[1])
        full_rec = torch.cat([recs[r, : b - a] for r, (a, b) in enumerate(sizes)], dim=0)
        full_en = torch.cat([ens[r, : b - a] for r, (a, b) in enumerate(sizes)], dim=0)
        return full_rec, full_en
    sizes = [split_even(n_pairs, world, r) for r in range(world)]
    width = max(b - a for a, b in sizes)
    pad_rec = rec.new_zeros((rec.shape[0], width, rec.shape[2]))
    pad_rec[:, : rec.shape[1]] = rec
    buf = _all_gather_flat(pad_rec.reshape(-1), world).reshape(world, rec.shape[0], width, rec.shape[2])
    full_rec = torch.cat([buf[r, :, : b - a] for r, (a, b) in enumerate(sizes)], dim=1)
    return full_rec, energy          # every rank computed all energies of its windows


# ---------------------------------------------------------------------------------------------
# tiled pair sharding: blocks of the upper-triangular pair matrix per rank, so a rank
# transforms only the buoys its blocks touch instead of all of them
# ---------------------------------------------------------------------------------------------
FWD_COST_IN_PAIRS = 1.4     # forward FFT of one buoy ~ 1.4 pair correlations (B200, N = 2^20: 12.5 us vs 9.1 us)


def _pair_index(i: int, j: int, n: int) -> int:
    """Position of (i, j), i < j, in the i-major list `engine.pair_table(n)`."""
    return i * n - i * (i + 1) // 2 + (j - i - 1)


def tile_pairs(n_buoys: int, world: int):
    """Deal the i<j pairs to `world` ranks as blocks of a k x k grouping of the buoys.

    Returns a list (one entry per rank) of dicts:
        buoys          sorted int64 array of the buoys the rank has to transform
        local_pairs    int32[P_r, 2]: (i, j) as indices into `buoys`
        global_index   int64[P_r]: position of each pair in the full i-major pair list
    Every pair appears on exactly one rank.  The number of groups k and the block-to-rank assignment minimise
    the most loaded rank's cost = pairs + FWD_COST_IN_PAIRS * buoys (longest-processing-time greedy); the result
    depends only on (n_buoys, world), so every rank computes the same table."""
    if world <= 1:
        idx = np.arange(n_buoys * (n_buoys - 1) // 2, dtype=np.int64)
        pairs = np.array([(i, j) for i in range(n_buoys) for j in range(i + 1, n_buoys)], dtype=np.int32).reshape(-1, 2)
        return [dict(buoys=np.arange(n_buoys, dtype=np.int64), local_pairs=pairs, global_index=idx)]
    best = None
    for k in range(1, min(n_buoys, 4 * world) + 1):
        if k * (k + 1) // 2 < world and k < n_buoys:
            continue
        bounds = [split_even(n_buoys, k, g) for g in range(k)]
        blocks = []
        for a in range(k):
            for b in range(a, k):
                na, nb = bounds[a][1] - bounds[a][0], bounds[b][1] - bounds[b][0]
                cnt = na * (na - 1) // 2 if a == b else na * nb
                if cnt:
                    blocks.append((cnt, a, b))
        blocks.sort(key=lambda t: (-t[0], t[1], t[2]))
        groups = [set() for _ in range(world)]
        npairs = [0] * world
        owner = []
        for cnt, a, b in blocks:
            def cost(r):
                g = groups[r] | {a, b}
                return npairs[r] + cnt + FWD_COST_IN_PAIRS * sum(bounds[x][1] - bounds[x][0] for x in g)
            r = min(range(world), key=lambda q: (cost(q), q))
            groups[r] |= {a, b}
            npairs[r] += cnt
            owner.append(r)
        worst = max(npairs[r] + FWD_COST_IN_PAIRS * sum(bounds[x][1] - bounds[x][0] for x in groups[r]) for r in range(world))
        if best is None or worst < best[0]:
            best = (worst, bounds, blocks, owner)
    _, bounds, blocks, owner = best
    out = []
    for r in range(world):
        mine = [(a, b) for (cnt, a, b), o in zip(blocks, owner) if o == r]
        buoys = sorted({x for a, b in mine for g in (a, b) for x in range(bounds[g][0], bounds[g][1])})
        pos = {x: t for t, x in enumerate(buoys)}
        lp, gi = [], []
        for a, b in sorted(mine):
            for i in range(bounds[a][0], bounds[a][1]):
                for j in range(bounds[b][0], bounds[b][1]):
                    if i < j:
                        lp.append((pos[i], pos[j]))
                        gi.append(_pair_index(i, j, n_buoys))
        out.append(dict(buoys=np.array(buoys, dtype=np.int64), local_pairs=np.array(lp, dtype=np.int32).reshape(-1, 2),
                        global_index=np.array(gi, dtype=np.int64)))
    return out


_tile_cache: dict = {}
_gather_perm_cache: dict = {}


def tiles_for(n_buoys: int, world: int):
    """tile_pairs(n_buoys, world), computed once per (n_buoys, world): the table depends on nothing else."""
    key = (int(n_buoys), int(world))
    tiles = _tile_cache.get(key)
    if tiles is None:
        if len(_tile_cache) >= 16:
            _tile_cache.clear()
        tiles = _tile_cache[key] = tile_pairs(*key)
        for r, t in enumerate(tiles):
            t["key"] = key + (r,)              # content key: (n_buoys, world, rank) identifies the tile
    return tiles


def _tiles_key(tiles, n_pairs: int, world: int):
    """Content key of a tiling (never id(): ids are reused after garbage collection)."""
    keys = tuple(t.get("key") for t in tiles)
    if all(k is not None for k in keys):
        return keys
    import hashlib
    h = hashlib.sha1()
    for t in tiles:
        h.update(np.ascontiguousarray(t["global_index"]).tobytes())
        h.update(b"|")
    return (n_pairs, world, h.hexdigest())


def gather_tiled_records(rec: torch.Tensor, tiles, n_pairs: int, world: int):
    """rec: this rank's [W, P_r, 4] int32 records for its tile.  Returns [W, n_pairs, 4] in the i-major pair
    order on every rank: one padded all_gather into a single buffer, then one index_select with a cached
    permutation (the index lists are known to every rank)."""
    width = max(len(t["global_index"]) for t in tiles)
    key = (_tiles_key(tiles, n_pairs, world), str(rec.device), width)
    perm = _gather_perm_cache.get(key)
    if perm is None:
        src = np.empty(n_pairs, dtype=np.int64)                 # position of pair g in the concatenated padded buffers
        for r, t in enumerate(tiles):
            src[t["global_index"]] = r * width + np.arange(len(t["global_index"]))
        if len(_gather_perm_cache) >= 32:
            _gather_perm_cache.clear()
        perm = _gather_perm_cache[key] = torch.from_numpy(src).to(rec.device)
    pad = rec.new_zeros((rec.shape[0], width, rec.shape[2]))
    pad[:, : rec.shape[1]] = rec
    buf = _all_gather_flat(pad.reshape(-1), world).reshape(world, rec.shape[0], width, rec.shape[2])
    cat = buf.permute(1, 0, 2, 3).reshape(rec.shape[0], world * width, rec.shape[2])
    return cat.index_select(1, perm)
