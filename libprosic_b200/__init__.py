"""libprosic_b200 -- B200-native (sm_100a) implementation of Varlociraptor's per-site likelihood
hot path behind a C ABI (include/vlr_gpu.h).  This package is the thin host-side mirror used by
the parity tests and the benchmark; the product is csrc/ + the shared library.
"""
from ._ffi import (HMM_DEFAULT, HMM_LITERAL, HMM_PATH_CLOSE, HMM_SUM3_APPROX, HMM_SUM3_APPROX3, REALIGN_FAST,
                   VlrError)  # noqa: F401
from .context import Context, GapParams  # noqa: F401
