"""smarties_b200 — B200-native V-RACER learner hot path of cselab/smarties behind a C-ABI.

Only what the path needs: `csrc/` (CUDA kernels + C-ABI, built into libsmarties_b200.so),
`learner.py` (ctypes binding mirroring the reference's learner interface), `settings.py`
(the settings/*.json surface) and `synth.py` (synthetic replay buffers for tests and bench).
"""
from .settings import HyperParameters  # noqa: F401
from .learner import Learner, SmartiesB200Error, load_library, LIB_PATH, EXPORTS  # noqa: F401
