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


def gather_records(rec: torch.Tensor, energy: torch.Tensor, n_windows: int, n_pairs: int, world: int, rank: int):
    """Assemble the full [W, P, 4] record tensor (and [W, B] energies) on every rank.

    rec: this rank's [w_local, p_local, 4] int32; energy: [w_local, B] int64."""
    by_window = n_windows >= world
    if by_window:
        sizes = [split_even(n_windows, world, r) for r in range(world)]
        width = max(b - a for a, b in sizes)
        pad_rec = rec.new_zeros((width,) + tuple(rec.shape[1:]))
        pad_rec[: rec.shape[0]] = rec
        pad_en = energy.new_zeros((width,) + tuple(energy.shape[1:]))
        pad_en[: energy.shape[0]] = energy
        recs = [torch.empty_like(pad_rec) for _ in range(world)]
        ens = [torch.empty_like(pad_en) for _ in range(world)]
        dist.all_gather(recs, pad_rec)
        dist.all_gather(ens, pad_en)
        full_rec = torch.cat([recs[r][: b - a] for r, (a, b) in enumerate(sizes)], dim=0)
        full_en = torch.cat([ens[r][: b - a] for r, (a, b) in enumerate(sizes)], dim=0)
        return full_rec, full_en
    sizes = [split_even(n_pairs, world, r) for r in range(world)]
    width = max(b - a for a, b in sizes)
    pad_rec = rec.new_zeros((rec.shape[0], width, rec.shape[2]))
    pad_rec[:, : rec.shape[1]] = rec
    recs = [torch.empty_like(pad_rec) for _ in range(world)]
    dist.all_gather(recs, pad_rec)
    full_rec = torch.cat([recs[r][:, : b - a] for r, (a, b) in enumerate(sizes)], dim=1)
    return full_rec, energy          # every rank computed all energies of its windows
