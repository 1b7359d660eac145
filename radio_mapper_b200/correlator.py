"""Batched multi-buoy correlator: the device-side half of `TDoAProcessor.correlate_iq`.

One `Correlator` serves a fixed (n_buoys, n_samples) shape on one GPU.  Per window it runs
    rmx_fft_forward_cu8   (fused cu8 unpack + zero-pad + forward FFT of all buoys)
    rmx_signal_energy     (exact per-buoy energy, for the coherence / confidence value)
    rmx_xcorr_pairs_peak  (conj-multiply + inverse FFT + arg-max + parabolic, all pairs)
and brings back only the 16-byte peak records.  Multi-GPU: windows (or, with fewer windows
than ranks, pairs) are sharded over the ranks of the default torch.distributed group; each
rank recomputes the spectra it needs and only peak records cross NVLink (one all_gather).
"""
from __future__ import annotations

import ctypes
from typing import Optional

import numpy as np
import torch

from . import _native, engine, sharding

_lib = _native.load()

RECORD_DTYPE = np.dtype([("lag", "<i4"), ("peak", "<f4"), ("frac", "<f4"), ("coherence", "<f4")])


def as_u8_tensor(iq_u8) -> torch.Tensor:
    if isinstance(iq_u8, np.ndarray):
        if iq_u8.dtype != np.uint8:
            raise TypeError("cu8 IQ must be uint8")
        return torch.from_numpy(np.ascontiguousarray(iq_u8))
    if not isinstance(iq_u8, torch.Tensor) or iq_u8.dtype != torch.uint8:
        raise TypeError("cu8 IQ must be a uint8 numpy array or torch tensor")
    return iq_u8


WORKSPACE_FRACTION = 0.5      # default cap of the correlation workspace: this share of the free device memory
STAGING_WINDOWS = 4           # host -> device staging ring (windows); copies run this far ahead of the kernels


class Correlator:
    def __init__(self, n_buoys: int, n_samples: int, device=None, workspace_pairs: Optional[int] = None):
        """workspace_pairs: pairs correlated per chunk (8*fft_len bytes of workspace each).  Default: all pairs if
        they fit WORKSPACE_FRACTION of the free device memory, else as many as do (64 buoys at N = 2^22 would
        otherwise ask for 135 GB); librmx walks the pair list in chunks of that size."""
        if n_buoys < 2:
            raise ValueError("need at least two buoys to correlate")
        self.n_buoys, self.n_samples = int(n_buoys), int(n_samples)
        self.plan = engine.Plan(self.n_buoys, self.n_samples, device=device)
        self.device = self.plan.device
        self.pairs_host = engine.pair_table(self.n_buoys)
        self.pairs = torch.from_numpy(self.pairs_host).to(self.device)
        self.n_pairs = len(self.pairs_host)
        self.spectra = torch.empty((self.n_buoys, self.plan.fft_len), dtype=torch.complex64, device=self.device)
        if workspace_pairs is None:
            per_pair = max(1, self.plan.workspace_bytes(1))
            with torch.cuda.device(self.device):
                free, _total = torch.cuda.mem_get_info()
            fit = int(free * WORKSPACE_FRACTION) // per_pair
            if fit < self.n_pairs:
                workspace_pairs = max(1, fit)
        self.workspace_pairs = workspace_pairs
        self._staging: Optional[torch.Tensor] = None
        self._rec_host = None        # page-locked result slots of run_iter
        self._en_host = None
        self._copy_stream = None
        self._split = None           # plans / pair subsets of the split first window
        self._tile_plans = {}        # forward plans per tile size (run_device_tile)
        self._tile_dev = {}          # device copies of a tile's buoy list and local pair table, keyed by CONTENT
        self.launches = 0            # kernels launched by the last run()

    # -- device-resident core ---------------------------------------------------------------
    def run_device(self, iq_dev: torch.Tensor, windows, pair_slice: Optional[slice] = None,
                   records: Optional[torch.Tensor] = None, energy: Optional[torch.Tensor] = None, pairs=None,
                   on_window=None):
        """iq_dev: CUDA uint8[B, W, 2N].  Processes the listed windows; returns (records int32
        [len(windows), P', 4], energy uint64[len(windows), B]) on the device."""
        if pairs is None:
            pairs = self.pairs if pair_slice is None else self.pairs[pair_slice].contiguous()
        nw = len(windows)
        if records is None:
            records = torch.empty((nw, pairs.shape[0], 4), dtype=torch.int32, device=self.device)
        if energy is None:
            energy = torch.empty((nw, self.n_buoys), dtype=torch.int64, device=self.device)
        stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        n_passes = len(self.plan.pass_lengths)
        for k, w in enumerate(windows):
            view = iq_dev[:, w, :]
            self.plan.forward(view, out=self.spectra)
            _native.check(_lib.rmx_signal_energy(ctypes.c_void_p(view.data_ptr()), view.stride(0) if self.n_buoys > 1 else 0,
                                                 self.n_buoys, self.n_samples, ctypes.c_void_p(energy[k].data_ptr()), stream),
                          "rmx_signal_energy")
            self.plan.xcorr_pairs_peak(self.spectra, pairs, out=records[k], max_pairs_in_flight=self.workspace_pairs)
            self.launches += n_passes + 1 + n_passes + 1
            if on_window is not None:
                on_window(k, records[k], energy[k])
        return records, energy

    def run_device_tile(self, iq_dev: torch.Tensor, windows, tile):
        """Pair-tiled form of run_device (sharding.tile_pairs): only the buoys of `tile` are transformed and only
        its pairs correlated.  Returns (records int32[len(windows), P_tile, 4], energy int64[len(windows), B] of
        ALL buoys -- the energy pass reads 2 bytes per sample and is not worth sharding)."""
        # content key (ids are reused after garbage collection): the tile's own key when it came from
        # sharding.tiles_for, else its bytes
        key = tile.get("key") or (tile["buoys"].tobytes(), np.ascontiguousarray(tile["local_pairs"]).tobytes())
        cached = self._tile_dev.get(key)
        if cached is None:
            if len(self._tile_dev) >= 64:
                self._tile_dev.clear()
            cached = self._tile_dev[key] = (torch.from_numpy(tile["buoys"]).to(self.device),
                                            torch.from_numpy(np.ascontiguousarray(tile["local_pairs"])).to(self.device))
        buoys, pairs_local = cached
        nb = int(buoys.numel())
        nw = len(windows)
        n_local = int(tile["local_pairs"].shape[0])
        records = torch.empty((nw, n_local, 4), dtype=torch.int32, device=self.device)
        energy = torch.empty((nw, self.n_buoys), dtype=torch.int64, device=self.device)
        stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        if nb:
            key = ("tile", nb)
            sub_plan = self._tile_plans.get(key)
            if sub_plan is None:
                sub_plan = self._tile_plans[key] = engine.Plan(nb, self.n_samples, device=self.device)
        n_passes = len(self.plan.pass_lengths)
        for k, w in enumerate(windows):
            view = iq_dev[:, w, :]
            _native.check(_lib.rmx_signal_energy(ctypes.c_void_p(view.data_ptr()), view.stride(0) if self.n_buoys > 1 else 0,
                                                 self.n_buoys, self.n_samples, ctypes.c_void_p(energy[k].data_ptr()), stream),
                          "rmx_signal_energy")
            self.launches += 1
            if nb == 0 or n_local == 0:
                continue
            sub = view.index_select(0, buoys)                       # [nb, 2N] contiguous copy of this tile's rows
            sub_plan.forward(sub, out=self.spectra[:nb])
            self.plan.xcorr_pairs_peak(self.spectra[:nb], pairs_local, out=records[k],
                                       max_pairs_in_flight=self.workspace_pairs)
            self.launches += n_passes + n_passes + 1
        return records, energy

    # -- host-facing call ---------------------------------------------------------------------
    def run(self, iq_u8: torch.Tensor, max_lag: Optional[int] = None, distributed: bool = False) -> np.ndarray:
        """iq_u8: uint8[B, W, 2N] on the host (pinned for async copies) or on the device.
        Returns host records [W, P] (RECORD_DTYPE)."""
        rows = list(self.run_iter(iq_u8, max_lag=max_lag, distributed=distributed))
        if not rows:
            return np.empty((0, self.n_pairs), dtype=RECORD_DTYPE)
        return np.stack(rows, axis=0)

    def run_iter(self, iq_u8: torch.Tensor, max_lag: Optional[int] = None, distributed: bool = False):
        """Generator form of `run`: yields the [P] record array of each window as soon as that window's kernels
        and its 16-byte-per-pair read-back have finished, while the GPU is already working on the windows behind
        it (every launch of the call is enqueued before the first yield).  The caller's per-window host work --
        e.g. building TDoAMeasurement objects -- therefore overlaps the GPU."""
        if iq_u8.shape[0] != self.n_buoys or iq_u8.shape[2] != 2 * self.n_samples:
            raise ValueError("expected uint8[%d, W, %d], got %s" % (self.n_buoys, 2 * self.n_samples, tuple(iq_u8.shape)))
        n_windows = iq_u8.shape[1]
        if max_lag != self.plan.max_lag:
            self.plan.set_max_lag(max_lag)
        self.launches = 0
        world, rank = sharding.world_and_rank() if distributed else (1, 0)
        with torch.cuda.device(self.device):
            if world > 1:
                rec_dev, en_dev = self._run_distributed(iq_u8, n_windows, world, rank)
                rec = rec_dev.cpu().numpy()
                en = en_dev.cpu().numpy()
                out = self._finish(rec, en)
                for w in range(out.shape[0]):
                    yield out[w]
                return
            # single GPU: results leave through page-locked slots right behind each window's kernels
            if self._rec_host is None or self._rec_host.shape[0] < n_windows:
                self._rec_host = torch.empty((n_windows, self.n_pairs, 4), dtype=torch.int32).pin_memory()
                self._en_host = torch.empty((n_windows, self.n_buoys), dtype=torch.int64).pin_memory()
            compute = torch.cuda.current_stream()
            events = [None] * n_windows

            def on_window(k, rec_k, en_k):
                self._rec_host[k].copy_(rec_k, non_blocking=True)
                self._en_host[k].copy_(en_k, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(compute)
                events[k] = ev

            windows = list(range(n_windows))
            if iq_u8.is_cuda:
                self.run_device(iq_u8, windows, None, on_window=on_window)
            else:
                self._run_from_host(iq_u8, windows, None, on_window=on_window)
            for k in range(n_windows):
                events[k].synchronize()
                yield self._finish(self._rec_host[k].numpy()[None], self._en_host[k].numpy()[None])[0]

    def _run_distributed(self, iq_u8, n_windows, world, rank):
        if n_windows < world:
            # fewer windows than ranks: blocks of the pair matrix per rank -- a rank transforms
            # only the buoys its blocks touch and one all-gather assembles the records
            tiles = sharding.tiles_for(self.n_buoys, world)
            windows = list(range(n_windows))
            if iq_u8.is_cuda:
                rec_dev, en_dev = self.run_device_tile(iq_u8, windows, tiles[rank])
            else:
                rec_dev, en_dev = self._run_from_host(iq_u8, windows, None, tile=tiles[rank])
            return sharding.gather_tiled_records(rec_dev, tiles, self.n_pairs, world), en_dev
        windows, pair_slice = sharding.shard_units(n_windows, self.n_pairs, world, rank)
        if iq_u8.is_cuda:
            rec_dev, en_dev = self.run_device(iq_u8, windows, pair_slice)
        else:
            rec_dev, en_dev = self._run_from_host(iq_u8, windows, pair_slice)
        return sharding.gather_records(rec_dev, en_dev, n_windows, self.n_pairs, world, rank)

    def _run_from_host(self, iq_u8: torch.Tensor, windows, pair_slice, tile=None, on_window=None):
        """Host cu8 -> device, one window at a time on a copy stream, so the H2D transfer of window
        w+1 overlaps the FFT / correlate kernels of window w (pinned host memory makes the copies
        asynchronous; pageable memory still works, just without overlap).  Only THIS rank's windows are staged,
        through a ring of STAGING_WINDOWS device slots (slot reuse waits for the kernels that read it)."""
        depth = max(1, min(len(windows), STAGING_WINDOWS))
        shape = (self.n_buoys, depth, 2 * self.n_samples)
        if self._staging is None or tuple(self._staging.shape) != shape:
            self._staging = None
            self._staging = torch.empty(shape, dtype=torch.uint8, device=self.device)
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=self.device)
        compute = torch.cuda.current_stream()
        pairs = self.pairs if pair_slice is None else self.pairs[pair_slice].contiguous()
        n_rec = pairs.shape[0] if tile is None else int(tile["local_pairs"].shape[0])
        records = torch.empty((len(windows), n_rec, 4), dtype=torch.int32, device=self.device)
        energy = torch.empty((len(windows), self.n_buoys), dtype=torch.int64, device=self.device)
        self._copy_stream.wait_stream(compute)              # staging may still be read by earlier kernels
        # The first window has nothing to hide its copy behind, so it is cut into growing groups of buoys
        # (2, 2, 4, 8, ...): each group is transformed, and correlated with everything already on the device,
        # while the next group is still on the bus.
        split = tile is None and pair_slice is None and self.n_buoys >= 4 and len(windows) > 0 and iq_u8.is_pinned()
        cuts = self._split_cuts() if split else []

        def copy_window(k):
            """enqueue the copy of windows[k] into slot k % depth; returns (event, cut events)"""
            cut_events = []
            with torch.cuda.stream(self._copy_stream):
                for b in range(self.n_buoys):               # contiguous rows: plain async memcpys
                    self._staging[b, k % depth].copy_(iq_u8[b, windows[k]], non_blocking=True)
                    if split and k == 0 and (b + 1) in cuts:
                        ev = torch.cuda.Event()
                        ev.record(self._copy_stream)
                        cut_events.append(ev)
                ev = torch.cuda.Event()
                ev.record(self._copy_stream)
            return ev, cut_events

        pending = {}
        for k in range(min(depth, len(windows))):
            pending[k] = copy_window(k)
        for k in range(len(windows)):
            ev, cut_events = pending.pop(k)
            slot = k % depth
            if split and k == 0:
                self._run_split_window(slot, cuts, cut_events, records[0], energy[0])
            else:
                compute.wait_event(ev)
                if tile is not None:
                    rec_k, en_k = self.run_device_tile(self._staging, [slot], tile)
                    records[k].copy_(rec_k[0])
                    energy[k].copy_(en_k[0])
                else:
                    self.run_device(self._staging, [slot], pair_slice, records=records[k:k + 1], energy=energy[k:k + 1], pairs=pairs)
            if on_window is not None:
                on_window(k, records[k], energy[k])
            if k + depth < len(windows):
                self._copy_stream.wait_stream(compute)      # the slot is free once this window's kernels are done
                pending[k + depth] = copy_window(k + depth)
        return records, energy

    def _split_cuts(self):
        """Group boundaries of the split first window.  Up to 8 buoys: 2, 3, 4, ... (one buoy per group after the
        first two -- with few, long signals every buoy's arrival unlocks milliseconds of pair work: cfg5's 8 x 134 MB
        window reaches 90 % of the device-resident rate instead of 80 % with doubling groups); more buoys: doubling
        groups 2, 4, 8, ... so the number of small launches stays logarithmic.  The last group takes the remainder."""
        cuts, c = [], 2
        while c < self.n_buoys:
            cuts.append(c)
            c = c + 1 if self.n_buoys <= 8 else c * 2
        cuts.append(self.n_buoys)
        return cuts

    def _run_split_window(self, w: int, cuts, cut_events, records_w: torch.Tensor, energy_w: torch.Tensor):
        """One window group by group: buoys [first, cut) as soon as their bytes are on the device (forward FFT,
        energy), then every pair (i < j) with first <= j < cut.  Same kernels on the same spectra as
        run_device, so the records are identical."""
        compute = torch.cuda.current_stream()
        stream = ctypes.c_void_p(compute.cuda_stream)
        if self._split is None:
            dev = self.device
            groups, first = [], 0
            plans = {}
            for cut in cuts:
                count = cut - first
                if count not in plans:
                    plans[count] = engine.Plan(count, self.n_samples, device=dev)
                sel = np.nonzero((self.pairs_host[:, 1] >= first) & (self.pairs_host[:, 1] < cut))[0]
                idx = torch.from_numpy(sel).to(dev)
                groups.append((first, count, plans[count], idx, self.pairs[idx].contiguous()))
                first = cut
            self._split = groups
        view = self._staging[:, w, :]
        n_passes = len(self.plan.pass_lengths)
        for (first, count, plan, idx, prs), ev in zip(self._split, cut_events):
            compute.wait_event(ev)
            part = view[first:first + count]
            plan.forward(part, out=self.spectra[first:first + count])
            _native.check(_lib.rmx_signal_energy(ctypes.c_void_p(part.data_ptr()), part.stride(0) if count > 1 else 0, count,
                                                 self.n_samples, ctypes.c_void_p(energy_w[first:first + count].data_ptr()), stream),
                          "rmx_signal_energy")
            rec = self.plan.xcorr_pairs_peak(self.spectra, prs, max_pairs_in_flight=self.workspace_pairs)
            records_w.index_copy_(0, idx, rec)
            self.launches += n_passes + 1 + n_passes + 1 + 1

    def _finish(self, rec: np.ndarray, energy_x4: np.ndarray) -> np.ndarray:
        out = np.empty(rec.shape[:2], dtype=RECORD_DTYPE)
        out["lag"] = rec[..., 0]
        out["peak"] = rec[..., 1].view(np.float32)
        out["frac"] = rec[..., 2].view(np.float32)
        e = energy_x4.astype(np.float64) / 4.0                                  # sum |x|^2 per (window, buoy)
        pi, pj = self.pairs_host[:, 0], self.pairs_host[:, 1]
        denom = np.sqrt(e[:, pi] * e[:, pj])
        with np.errstate(divide="ignore", invalid="ignore"):
            coh = np.where(denom > 0, out["peak"].astype(np.float64) / denom, 0.0)
        out["coherence"] = np.clip(coh, 0.0, 1.0).astype(np.float32)
        return out
