"""B200-native drop-in for the reference's `tdoa_processor` module.

Public surface mirrors /root/reference/tdoa_processor.py (same class, method and field names,
same argument meaning and return types — citations below are to that file):

    BuoyPosition :25   SignalDetection :34   TDoAMeasurement :47   TriangulationResult :57
    GeodeticCalculator :71      TDoACalculator :138      HyperbolicPositioning :212
    TDoAProcessor :330  (+ alias TDOAProcessor, the spelling BASELINE.json uses)

What is new: `TDoAProcessor.correlate_iq` turns raw multi-buoy cu8 IQ windows into
`TDoAMeasurement`s on the GPU (batched FFT -> pairwise conj-multiply + inverse FFT -> arg-max
lag with parabolic refinement), replacing the timestamp subtraction of :166 with a measured
sample lag.  The tiny geodesy / least-squares multilateration stays on the host.
"""
from __future__ import annotations

import itertools
import logging
import math
from dataclasses import dataclass
from datetime import datetime, timezone
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import scipy.optimize

logger = logging.getLogger(__name__)


# ----------------------------------------------------------------------------------------
# records (field order and defaults follow the reference dataclasses)
# ----------------------------------------------------------------------------------------
@dataclass
class BuoyPosition:
    buoy_id: str
    lat: float
    lng: float
    altitude: float = 0.0
    timing_accuracy_ns: int = 100000


@dataclass
class SignalDetection:
    buoy_id: str
    frequency_mhz: float
    signal_strength_dbm: float
    timestamp_utc: str
    gps_timestamp_ns: int
    lat: float
    lng: float
    confidence: float
    signal_type: str = "unknown"


@dataclass
class TDoAMeasurement:
    buoy1_id: str
    buoy2_id: str
    time_difference_ns: int          # buoy2 - buoy1; positive: buoy2 received later (:51)
    distance_difference_m: float
    confidence: float
    frequency_mhz: float


@dataclass
class TriangulationResult:
    estimated_lat: float
    estimated_lng: float
    estimated_altitude: float
    accuracy_meters: float
    confidence: float
    frequency_mhz: float
    signal_type: str
    timestamp_utc: str
    contributing_buoys: List[str]
    tdoa_measurements: List[TDoAMeasurement]
    method: str

    @property
    def accuracy_estimate_meters(self) -> float:
        """Name `central_processor.py:434` reads (the reference dataclass lacks it, so its live
        triangulation path raises AttributeError)."""
        return self.accuracy_meters


# ----------------------------------------------------------------------------------------
# geodesy (spherical earth of radius 6378137 m, as :74-136)
# ----------------------------------------------------------------------------------------
class GeodeticCalculator:
    EARTH_RADIUS_M = 6378137.0

    @staticmethod
    def lat_lng_to_xyz(lat: float, lng: float, alt: float = 0.0) -> Tuple[float, float, float]:
        phi, lam = math.radians(lat), math.radians(lng)
        rho = GeodeticCalculator.EARTH_RADIUS_M + alt
        return (rho * math.cos(phi) * math.cos(lam),
                rho * math.cos(phi) * math.sin(lam),
                rho * math.sin(phi))

    @staticmethod
    def xyz_to_lat_lng(x: float, y: float, z: float) -> Tuple[float, float, float]:
        horizontal = math.sqrt(x * x + y * y)
        lat = math.degrees(math.atan2(z, horizontal))
        lng = math.degrees(math.atan2(y, x))
        alt = math.sqrt(x * x + y * y + z * z) - GeodeticCalculator.EARTH_RADIUS_M
        return lat, lng, alt

    @staticmethod
    def distance_3d(lat1: float, lng1: float, alt1: float,
                    lat2: float, lng2: float, alt2: float) -> float:
        a = GeodeticCalculator.lat_lng_to_xyz(lat1, lng1, alt1)
        b = GeodeticCalculator.lat_lng_to_xyz(lat2, lng2, alt2)
        return math.sqrt((b[0] - a[0]) ** 2 + (b[1] - a[1]) ** 2 + (b[2] - a[2]) ** 2)

    @staticmethod
    def bearing_distance(lat1: float, lng1: float, lat2: float, lng2: float) -> Tuple[float, float]:
        p1, p2 = math.radians(lat1), math.radians(lat2)
        dl = math.radians(lng2 - lng1)
        hav = math.sin((p2 - p1) / 2) ** 2 + math.cos(p1) * math.cos(p2) * math.sin(dl / 2) ** 2
        distance = GeodeticCalculator.EARTH_RADIUS_M * 2 * math.atan2(math.sqrt(hav), math.sqrt(1 - hav))
        east = math.sin(dl) * math.cos(p2)
        north = math.cos(p1) * math.sin(p2) - math.sin(p1) * math.cos(p2) * math.cos(dl)
        bearing = (math.degrees(math.atan2(east, north)) + 360) % 360
        return bearing, distance


# ----------------------------------------------------------------------------------------
# pairwise time differences
# ----------------------------------------------------------------------------------------
class TDoACalculator:
    SPEED_OF_LIGHT = 299792458.0

    def __init__(self):
        self.logger = logging.getLogger(__name__ + ".TDoACalculator")

    def calculate_tdoa_measurements(self, detections: List[SignalDetection],
                                    buoy_positions: Dict[str, BuoyPosition]) -> List[TDoAMeasurement]:
        """All i<j pairs of `detections` (list order) whose frequencies agree within 0.01 MHz
        and whose buoys are registered; dt = t_j - t_i in ns, dd = dt * c  (:156-193)."""
        if len(detections) < 2:
            self.logger.warning("Need at least 2 detections for TDoA calculation")
            return []
        out: List[TDoAMeasurement] = []
        for first, second in itertools.combinations(detections, 2):
            if abs(first.frequency_mhz - second.frequency_mhz) > 0.01:
                continue
            dt_ns = second.gps_timestamp_ns - first.gps_timestamp_ns
            dd_m = (dt_ns / 1e9) * self.SPEED_OF_LIGHT
            pos1, pos2 = buoy_positions.get(first.buoy_id), buoy_positions.get(second.buoy_id)
            if not pos1 or not pos2:
                continue
            conf = min(first.confidence, second.confidence) * self._calculate_timing_confidence(pos1, pos2)
            out.append(TDoAMeasurement(first.buoy_id, second.buoy_id, dt_ns, dd_m, conf, first.frequency_mhz))
            self.logger.debug("TDoA %s-%s dT=%.1fus dD=%.1fm", first.buoy_id, second.buoy_id, dt_ns / 1000, dd_m)
        return out

    def _calculate_timing_confidence(self, buoy1: BuoyPosition, buoy2: BuoyPosition) -> float:
        """exp(-sqrt(s1^2 + s2^2) / 100 us), capped at 1  (:200-210)."""
        rss_ns = math.sqrt(buoy1.timing_accuracy_ns ** 2 + buoy2.timing_accuracy_ns ** 2)
        return min(math.exp(-rss_ns / 100000), 1.0)

    # ---- new: measured lags -> measurements -------------------------------------------------
    def measurements_from_lags(self, buoy_ids: Sequence[str], pairs: np.ndarray, lag: np.ndarray,
                               frac: np.ndarray, strength_conf: np.ndarray, sample_rate: float,
                               frequency_mhz: float,
                               buoy_positions: Optional[Dict[str, BuoyPosition]] = None,
                               start_ns: Optional[Sequence[int]] = None) -> List[TDoAMeasurement]:
        """Seam between the GPU lag search and :166-170:  dt_ns = round((lag+frac)/fs * 1e9),
        dd = dt/1e9 * c.  Timing confidence uses the registered buoys when available.
        start_ns: GPS time of each buoy's first sample; when the captures did not start together the arrival-time
        difference is the lag plus the difference of the capture starts (t_j0 - t_i0), in integer nanoseconds."""
        total = (lag.astype(np.float64) + frac.astype(np.float64)) / float(sample_rate) * 1e9
        dt = np.rint(total).astype(np.int64)
        pr = np.asarray(pairs)
        if start_ns is not None:
            t0 = np.asarray(start_ns, dtype=np.int64)
            dt = dt + (t0[pr[:, 1]] - t0[pr[:, 0]])
        # same float64 expression per element as :169-170 (dt/1e9 * c), vectorised; the per-pair timing confidence
        # depends only on the two buoys, so it is evaluated once per pair with the scalar routine of :200-210
        dd = (dt.astype(np.float64) / 1e9) * self.SPEED_OF_LIGHT
        conf = strength_conf.astype(np.float64)
        if buoy_positions:
            acc = [buoy_positions[b].timing_accuracy_ns if b in buoy_positions else None for b in buoy_ids]
            cache: Dict[tuple, float] = {}
            scale = np.ones(len(pr), dtype=np.float64)
            for k, (i, j) in enumerate(pr.tolist()):
                a1, a2 = acc[i], acc[j]
                if a1 is None or a2 is None:
                    continue
                v = cache.get((a1, a2))
                if v is None:
                    v = cache[(a1, a2)] = min(math.exp(-math.sqrt(a1 ** 2 + a2 ** 2) / 100000), 1.0)
                scale[k] = v
            conf = conf * scale
        names = list(buoy_ids)
        return [TDoAMeasurement(names[i], names[j], t, d, c, frequency_mhz)
                for (i, j), t, d, c in zip(pr.tolist(), dt.tolist(), dd.tolist(), conf.tolist())]


# ----------------------------------------------------------------------------------------
# multilateration (host; tiny)
# ----------------------------------------------------------------------------------------
class HyperbolicPositioning:
    def __init__(self):
        self.logger = logging.getLogger(__name__ + ".HyperbolicPositioning")

    def triangulate_position(self, measurements: List[TDoAMeasurement],
                             buoy_positions: Dict[str, BuoyPosition]) -> Optional[TriangulationResult]:
        """BFGS on sum_m (|x-b2| - |x-b1| - dd_m)^2 / (conf_m + 0.1) in ECEF, started at the
        centroid of the buoys  (:218-328)."""
        if len(measurements) < 2:
            self.logger.warning("Need at least 2 TDoA measurements for triangulation")
            return None
        involved = set()
        for m in measurements:
            involved.update((m.buoy1_id, m.buoy2_id))
        if len(involved) < 3:
            self.logger.warning("Need at least 3 buoys for 2D triangulation")
            return None
        ecef = {}
        for bid in involved:
            pos = buoy_positions.get(bid)
            if pos is None:
                self.logger.error("Missing position for buoy %s", bid)
                return None
            ecef[bid] = GeodeticCalculator.lat_lng_to_xyz(pos.lat, pos.lng, pos.altitude)

        first_xyz = np.array([ecef[m.buoy1_id] for m in measurements], dtype=np.float64)
        second_xyz = np.array([ecef[m.buoy2_id] for m in measurements], dtype=np.float64)
        measured = np.array([m.distance_difference_m for m in measurements], dtype=np.float64)
        inv_weight = 1.0 / (np.array([m.confidence for m in measurements], dtype=np.float64) + 0.1)

        def cost(x):
            # accumulate in measurement order with Python floats, like the reference's sum()
            total = 0.0
            for k in range(len(measured)):
                d1 = math.sqrt((x[0] - first_xyz[k, 0]) ** 2 + (x[1] - first_xyz[k, 1]) ** 2 + (x[2] - first_xyz[k, 2]) ** 2)
                d2 = math.sqrt((x[0] - second_xyz[k, 0]) ** 2 + (x[1] - second_xyz[k, 1]) ** 2 + (x[2] - second_xyz[k, 2]) ** 2)
                total += ((d2 - d1) - measured[k]) ** 2 * inv_weight[k]
            return total

        pts = list(ecef.values())
        start = [sum(p[a] for p in pts) / len(pts) for a in range(3)]
        try:
            sol = scipy.optimize.minimize(cost, start, method="BFGS", options={"maxiter": 1000})
            if not sol.success:
                self.logger.warning("Optimization failed: %s", sol.message)
                return None
            lat, lng, alt = GeodeticCalculator.xyz_to_lat_lng(*sol.x)
            result = TriangulationResult(
                estimated_lat=lat, estimated_lng=lng, estimated_altitude=alt,
                accuracy_meters=math.sqrt(sol.fun / len(measurements)),
                confidence=sum(m.confidence for m in measurements) / len(measurements),
                frequency_mhz=measurements[0].frequency_mhz, signal_type="unknown",
                timestamp_utc=datetime.now(timezone.utc).isoformat(),
                contributing_buoys=list(involved), tdoa_measurements=measurements, method="hyperbolic")
            self.logger.info("Triangulation successful: (%.6f, %.6f) +-%.1fm, confidence %.2f",
                             lat, lng, result.accuracy_meters, result.confidence)
            return result
        except Exception as exc:  # same contract as the reference: log and return None
            self.logger.error("Triangulation failed: %s", exc)
            return None


    def triangulate_position_robust(self, measurements: List[TDoAMeasurement],
                                    buoy_positions: Dict[str, BuoyPosition]) -> Optional[TriangulationResult]:
        """Same model as `triangulate_position`, solved as a scaled least-squares problem in a
        local east/north/up frame centred on the buoys.  The reference's BFGS works
        on raw ECEF coordinates (~6.4e6 m) and stops on precision loss in most geometries
        (Documents/TDOA_README.md:49-52); this variant is offered beside it, not instead of it."""
        if len(measurements) < 2:
            return None
        involved = sorted({b for m in measurements for b in (m.buoy1_id, m.buoy2_id)})
        if len(involved) < 3 or any(b not in buoy_positions for b in involved):
            return None
        ecef = {b: np.array(GeodeticCalculator.lat_lng_to_xyz(buoy_positions[b].lat, buoy_positions[b].lng,
                                                              buoy_positions[b].altitude)) for b in involved}
        origin = np.mean(list(ecef.values()), axis=0)
        up = origin / np.linalg.norm(origin)
        east = np.cross([0.0, 0.0, 1.0], up)
        east /= np.linalg.norm(east)
        north = np.cross(up, east)
        frame = np.stack([east, north, up])                       # rows: local axes
        local = {b: frame @ (p - origin) for b, p in ecef.items()}
        p1 = np.array([local[m.buoy1_id] for m in measurements])
        p2 = np.array([local[m.buoy2_id] for m in measurements])
        dd = np.array([m.distance_difference_m for m in measurements])
        wgt = 1.0 / np.sqrt(np.array([m.confidence for m in measurements]) + 0.1)
        height = float(np.mean([q[2] for q in local.values()]))

        def residuals(xy):
            x = np.array([xy[0], xy[1], height])
            return (np.linalg.norm(x - p2, axis=1) - np.linalg.norm(x - p1, axis=1) - dd) * wgt

        sol = scipy.optimize.least_squares(residuals, [0.0, 0.0], x_scale=1000.0, method="lm" if len(dd) >= 2 else "trf")
        if not sol.success:
            return None
        xyz = origin + frame.T @ np.array([sol.x[0], sol.x[1], height])
        lat, lng, alt = GeodeticCalculator.xyz_to_lat_lng(*xyz)
        cost = float(np.sum(sol.fun ** 2))
        return TriangulationResult(
            estimated_lat=lat, estimated_lng=lng, estimated_altitude=alt,
            accuracy_meters=math.sqrt(cost / len(measurements)),
            confidence=sum(m.confidence for m in measurements) / len(measurements),
            frequency_mhz=measurements[0].frequency_mhz, signal_type="unknown",
            timestamp_utc=datetime.now(timezone.utc).isoformat(), contributing_buoys=list(involved),
            tdoa_measurements=measurements, method="least_squares_enu")


# ----------------------------------------------------------------------------------------
# processor
# ----------------------------------------------------------------------------------------
class TDoAProcessor:
    """Coordinates grouping, pairwise TDoA and multilateration; owns the GPU correlator."""

    def __init__(self):
        self.logger = logging.getLogger(__name__ + ".TDoAProcessor")
        self.tdoa_calculator = TDoACalculator()
        self.hyperbolic_positioner = HyperbolicPositioning()
        self.buoy_positions: Dict[str, BuoyPosition] = {}
        self.correlation_window_s = 10.0
        self.min_buoys_for_triangulation = 3
        self._correlators = {}          # (B, N, device) -> radio_mapper_b200.correlator.Correlator

    def register_buoy(self, buoy_position: BuoyPosition):
        self.buoy_positions[buoy_position.buoy_id] = buoy_position
        self.logger.info("Registered buoy %s at (%.6f, %.6f)", buoy_position.buoy_id,
                         buoy_position.lat, buoy_position.lng)

    def process_signal_detections(self, detections: List[SignalDetection]) -> List[TriangulationResult]:
        """Group by frequency, keep the last `correlation_window_s`, pairwise TDoA, multilaterate
        (:351-403)."""
        if not detections:
            return []
        self.logger.info("Processing %d signal detections", len(detections))
        results: List[TriangulationResult] = []
        for frequency, group in self._group_by_frequency(detections).items():
            recent = self._filter_by_time_window(group)
            if len(recent) < self.min_buoys_for_triangulation:
                self.logger.debug("Insufficient detections for %s MHz (%d < %d)", frequency, len(recent),
                                  self.min_buoys_for_triangulation)
                continue
            measurements = self.tdoa_calculator.calculate_tdoa_measurements(recent, self.buoy_positions)
            if len(measurements) < 2:
                self.logger.debug("Insufficient TDoA measurements for %s MHz", frequency)
                continue
            fix = self.hyperbolic_positioner.triangulate_position(measurements, self.buoy_positions)
            if fix is None:
                continue
            types = [d.signal_type for d in recent]
            fix.signal_type = max(set(types), key=types.count)
            results.append(fix)
            if fix.signal_type == "emergency":
                self.logger.warning("EMERGENCY SIGNAL TRIANGULATED: %s MHz at (%.6f, %.6f) +-%.1fm", frequency,
                                    fix.estimated_lat, fix.estimated_lng, fix.accuracy_meters)
        return results

    def triangulate_signal(self, detections: List[SignalDetection]) -> Optional[TriangulationResult]:
        """Entry point `central_processor.py:418` calls (missing in the reference): the first
        triangulation result of `process_signal_detections`, or None."""
        results = self.process_signal_detections(detections)
        return results[0] if results else None

    def _group_by_frequency(self, detections: List[SignalDetection],
                            frequency_tolerance_mhz: float = 0.01) -> Dict[float, List[SignalDetection]]:
        """First-fit grouping: a detection joins the first existing group whose key is within
        the tolerance, otherwise it opens a group keyed by its own frequency (:405-425)."""
        groups: Dict[float, List[SignalDetection]] = {}
        for det in detections:
            key = next((f for f in groups if abs(det.frequency_mhz - f) <= frequency_tolerance_mhz), None)
            if key is None:
                groups[det.frequency_mhz] = [det]
            else:
                groups[key].append(det)
        return groups

    def _filter_by_time_window(self, detections: List[SignalDetection]) -> List[SignalDetection]:
        """Detections (sorted by timestamp) no older than correlation_window_s before the newest
        (:427-445)."""
        if not detections:
            return []
        ordered = sorted(detections, key=lambda d: d.gps_timestamp_ns)
        cutoff = ordered[-1].gps_timestamp_ns - int(self.correlation_window_s * 1e9)
        return [d for d in ordered if d.gps_timestamp_ns >= cutoff]

    def get_buoy_network_status(self) -> Dict:
        return {
            "registered_buoys": len(self.buoy_positions),
            "buoy_list": [dict(buoy_id=p.buoy_id, lat=p.lat, lng=p.lng, timing_accuracy_ns=p.timing_accuracy_ns)
                          for p in self.buoy_positions.values()],
            "min_buoys_required": self.min_buoys_for_triangulation,
            "correlation_window_s": self.correlation_window_s,
            "triangulation_ready": len(self.buoy_positions) >= self.min_buoys_for_triangulation,
        }

    # ------------------------------------------------------------------------------------
    # new batched GPU entry points
    # ------------------------------------------------------------------------------------
    def _correlator(self, n_buoys: int, n_samples: int, device):
        from .correlator import Correlator        # imports the CUDA engine; fails loudly without it
        key = (int(n_buoys), int(n_samples), str(device))
        cor = self._correlators.get(key)
        if cor is None:
            cor = self._correlators[key] = Correlator(n_buoys, n_samples, device=device)
        return cor

    def correlate_iq_records(self, iq_u8, max_lag: Optional[int] = None, device=None, distributed: bool = False):
        """iq_u8: uint8[B, W, 2N] (torch tensor, host-pinned or CUDA; or numpy) — B buoys, W
        windows of N complex samples in rtl_sdr cu8 format.  Returns a host structured array
        [W, P] with fields lag, peak, frac, coherence for the pairs i<j in list order."""
        from .correlator import as_u8_tensor
        t = as_u8_tensor(iq_u8)
        if t.ndim == 2:
            t = t[:, None, :]
        if t.ndim != 3 or t.shape[2] % 2:
            raise ValueError("iq_u8 must be uint8[B, W, 2N]")
        cor = self._correlator(t.shape[0], t.shape[2] // 2, device)
        return cor.run(t, max_lag=max_lag, distributed=distributed)

    def correlate_iq(self, iq_u8, buoy_ids: Sequence[str], sample_rate: float = 2048000,
                     frequency_mhz: float = 0.0, max_lag: Optional[int] = None, device=None,
                     distributed: bool = False) -> List[TDoAMeasurement]:
        """Cross-correlate every buoy pair of every window on the GPU and return one
        `TDoAMeasurement` per (window, pair), window-major, pairs in the i<j order of
        `calculate_tdoa_measurements`.  time_difference_ns = round((lag + frac) / fs * 1e9)."""
        from .correlator import as_u8_tensor
        from .engine import pair_table
        t = as_u8_tensor(iq_u8)
        if t.ndim == 2:
            t = t[:, None, :]
        if t.ndim != 3 or t.shape[2] % 2:
            raise ValueError("iq_u8 must be uint8[B, W, 2N]")
        n_buoys = len(buoy_ids)
        if t.shape[0] != n_buoys:
            raise ValueError("buoy_ids has %d entries but the IQ block has a different buoy count" % n_buoys)
        pairs = pair_table(n_buoys)
        cor = self._correlator(t.shape[0], t.shape[2] // 2, device)
        out: List[TDoAMeasurement] = []
        # windows arrive one by one while the GPU works on the ones behind them: the per-window object building
        # below overlaps the kernels
        for rec in cor.run_iter(t, max_lag=max_lag, distributed=distributed):
            out.extend(self.tdoa_calculator.measurements_from_lags(
                buoy_ids, pairs, rec["lag"], rec["frac"], rec["coherence"], sample_rate, frequency_mhz,
                self.buoy_positions))
        return out

    def correlate_stream(self, source, buoy_ids: Sequence[str], sample_rate: float = 2048000,
                         frequency_mhz: float = 0.0, max_lag: Optional[int] = None, device=None, depth: int = 3,
                         max_windows: Optional[int] = None):
        """Streaming form of `correlate_iq`: `source` is an `ingest.Cu8FileSource` (one raw rtl_sdr capture
        per buoy, `sdr_capture.py:26`), `ingest.Cu8PipeSource` (live `rtl_sdr ... -` pipes,
        `iq_stream_client.py:101-116`) or `ingest.ArraySource`.  Windows go through a pinned host ring and
        are copied to the GPU while the previous window is being correlated.  Yields, per window, the
        list of `TDoAMeasurement`s `correlate_iq` would return for it."""
        from . import ingest
        from .engine import pair_table
        if len(buoy_ids) != source.n_buoys:
            raise ValueError("buoy_ids has %d entries but the source delivers %d buoys" % (len(buoy_ids), source.n_buoys))
        cor = self._correlator(source.n_buoys, source.samples_per_window, device)
        key = ("stream", id(cor), int(depth))
        sc = self._correlators.get(key)
        if sc is None:
            sc = self._correlators[key] = ingest.StreamingCorrelator(cor, depth=depth)
        pairs = pair_table(source.n_buoys)
        for rec in sc.run(source, max_lag=max_lag, max_windows=max_windows):
            yield self.tdoa_calculator.measurements_from_lags(
                buoy_ids, pairs, rec["lag"], rec["frac"], rec["coherence"], sample_rate, frequency_mhz,
                self.buoy_positions)

    def correlate_window_frames(self, frames, buoy_ids: Sequence[str], samples_per_window: int,
                                max_lag: Optional[int] = None, device=None, depth: int = 4):
        """Central-side consumer of the binary `cu8_window` frames (wire.py; SURVEY §8f-1: raw IQ windows next to the
        reference's JSON `signal_detection` frames, central_processor.py:305-335).  `frames` is an iterable of binary
        frames (or wire.Cu8Window objects) from the buoys in any interleaving; every time all `buoy_ids` have
        delivered a window it is correlated on the GPU and (window_index, List[TDoAMeasurement]) is yielded, with
        time_difference_ns = lag/fs plus the difference of the buoys' capture-start GPS timestamps."""
        from . import wire
        from .engine import pair_table
        asm = wire.WindowAssembler(buoy_ids, samples_per_window, depth=depth)
        pairs = pair_table(len(buoy_ids))
        for w, block, stamps, fs, fc in asm.feed(frames):
            rec = self.correlate_iq_records(block, max_lag=max_lag, device=device)
            yield w, self.tdoa_calculator.measurements_from_lags(
                list(buoy_ids), pairs, rec["lag"][0], rec["frac"][0], rec["coherence"][0], fs, fc / 1e6,
                self.buoy_positions, start_ns=stamps)

    def triangulate_iq(self, iq_u8, buoy_ids: Sequence[str], sample_rate: float = 2048000,
                       frequency_mhz: float = 0.0, signal_type: str = "unknown",
                       max_lag: Optional[int] = None, robust: bool = False, device=None,
                       distributed: bool = False) -> List[Optional[TriangulationResult]]:
        """correlate_iq + host multilateration, one result (or None) per window.  robust=True uses
        `triangulate_position_robust` instead of the reference's BFGS; device / distributed as in correlate_iq."""
        meas = self.correlate_iq(iq_u8, buoy_ids, sample_rate, frequency_mhz, max_lag=max_lag, device=device,
                                 distributed=distributed)
        solve = self.hyperbolic_positioner.triangulate_position_robust if robust else \
            self.hyperbolic_positioner.triangulate_position
        per_window = len(meas) // max(1, (len(buoy_ids) * (len(buoy_ids) - 1)) // 2)
        n_pairs = len(meas) // max(1, per_window)
        fixes = []
        for w in range(per_window):
            fix = solve(meas[w * n_pairs:(w + 1) * n_pairs], self.buoy_positions)
            if fix is not None:
                fix.signal_type = signal_type
            fixes.append(fix)
        return fixes


# spelling used by BASELINE.json's north_star
TDOAProcessor = TDoAProcessor
