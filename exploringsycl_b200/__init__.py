"""exploringsycl_b200 -- B200-native (sm_100a) backend for TeaLeaf's CG / Chebyshev / PPCG / Jacobi
heat-conduction solvers behind the reference's kernel_interface.h plugin API.

Only what the hot path needs lives here: csrc/ (CUDA kernels + the C-ABI of include/tealeaf_b200.h)
and tealeaf.py (the host-side mirror of the reference interface). No CPU fallback exists.
"""
from ._lib import TeaLeafError, lib, LIB_PATH  # noqa: F401
from .tealeaf import (Chunk, Comms, Settings, State, TeaLeaf, read_config, read_config_clean,  # noqa: F401
                      settings_overload, write_to_visit, decompose_field, get_checking_value)
