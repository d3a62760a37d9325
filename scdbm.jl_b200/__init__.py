"""scdbm.jl_b200 -- B200-native Gibbs / CD / PCD / mean-field / AIS hot path of
scDBM.jl behind the reference's BoltzmannMachines-style API.

The compute lives in `libscdbm_b200.so` (hand-written sm_100a CUDA; C ABI in
include/scdbm_b200.h).  This package is the host-side mirror of the Julia API
over ctypes.  There is no CPU fallback: importing works without a GPU (so that
symbols can be inspected), every compute call raises without one.
"""
from ._lib import (ScdbmError, LIB_PATH, PROTOTYPES, lib,  # noqa: F401
                   VIS_BERNOULLI, VIS_NEGBIN, MAT_BINARY, MAT_COUNT, MAT_REAL,
                   OUT_INPUT, OUT_POTENTIAL, OUT_SAMPLE, ENGINE_AUTO, ENGINE_SIMT, ENGINE_TC)
from .api import *  # noqa: F401,F403
from .api import Context, DeviceModel, Mat, DeviceParticles, RBMTrainer  # noqa: F401
from . import hostrng  # noqa: F401
