"""Top-level shim so `import signal_analyzer` resolves to the B200-native drop-in
(the reference keeps this module at its repository root)."""
from radio_mapper_b200.signal_analyzer import *  # noqa: F401,F403
from radio_mapper_b200.signal_analyzer import (SignalAnalyzer, analyze_iq_file, analyze_spectrum,  # noqa: F401
                                               calculate_signal_stats, load_iq_data, plot_spectrum)
