"""Top-level shim so `import buoy_node` resolves to the B200-native drop-in of the reference module's hot path
(`SignalDetector._detect_real_signals`; the reference keeps this module at its repository root)."""
from radio_mapper_b200.buoy_node import (CaptureError, GPSTimeSource, SignalDetection, SignalDetector,  # noqa: F401
                                         rtl_sdr_capture)
