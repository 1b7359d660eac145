"""Top-level shim so `import tdoa_processor` resolves to the B200-native drop-in
(the reference keeps this module at its repository root)."""
from radio_mapper_b200.tdoa_processor import *  # noqa: F401,F403
from radio_mapper_b200.tdoa_processor import TDOAProcessor, TDoAProcessor  # noqa: F401
