"""radio_mapper_b200 — B200-native hot path of physiii/radio-mapper.

Drop-in host modules (same names and signatures as the reference):
    radio_mapper_b200.tdoa_processor   (TDoAProcessor / TDOAProcessor, dataclasses, geodesy)
    radio_mapper_b200.signal_analyzer  (load_iq_data, analyze_spectrum, ..., SignalAnalyzer)
    radio_mapper_b200.detectors        (block detectors of buoy_node / iq_stream_client)
Device engine: radio_mapper_b200.engine (ctypes over librmx.so, include/rmx.h).

Importing the package does not touch CUDA; the GPU modules load librmx.so on first use and
raise if it has not been built (there is no CPU fallback for the hot path).
"""
__version__ = "0.1.0"
