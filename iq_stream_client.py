"""Top-level shim so `import iq_stream_client` resolves to the B200-native drop-in of the reference module's hot
path (`RealTimeSDRCapture.read_iq_samples`, `SignalDetector.detect_signals`; the reference keeps this module at its
repository root)."""
from radio_mapper_b200.iq_stream_client import RealTimeSDRCapture, SignalDetection, SignalDetector  # noqa: F401
