"""inria_wbc_b200 -- B200-native batched whole-body-control cycle behind the inria_wbc controller API.

The product is `libiwbc_b200.so` (hand-written sm_100a CUDA + C++ host code behind the C ABI of
include/iwbc_batch.h).  This package is the thin ctypes binding used by the tests and bench.py (the C++ mirror of the reference's controller classes is
include/inria_wbc_b200/batch_pos_tracker.hpp).  There is no CPU fallback:
the compute calls raise if the CUDA library or a GPU is missing.
"""
from ._lib import (Problem, Batch, Results, Behavior, CollisionModel, Stabilizer, MultiBatch, shard_range, qp_solve_batch, lib, lib_path, IwbcError,  # noqa: F401
                   ETC_PATH, ROBOTS_PATH, STATUS_OPTIMAL, ACTIVE_WORDS, measure_fp64_tflops)
