"""mcintegration.jl_b200 — B200-native (sm_100a) sampling-and-accumulation hot path of
numericalEFT/MCIntegration.jl behind the reference's `integrate` / `Configuration` surface.

The compute lives in libmcib200.so (hand-written CUDA, C ABI in include/mcib200.h).  This package is
the thin host-side mirror that the reference's Julia driver plays; there is no CPU fallback.
Import name: `mcintegration_jl_b200` (see mcintegration_jl_b200.py at the repo root).
"""
from . import _abi, integrands
from .configuration import Configuration, ProblemSpec, problem_from_config, maps_from_config, maps_to_config
from .engine import CudaEngine, IterationStats, MCIError, fp64_peak_tflops
from .main import integrate, drive, block_statistics
from .statistics import Result, average, mean_std, report
from .variables import CompositeVar, Continuous, Discrete, FermiK

__all__ = [
    "integrate", "report", "Result", "Configuration", "Continuous", "Discrete", "CompositeVar", "FermiK", "integrands",
    "CudaEngine", "IterationStats", "MCIError", "ProblemSpec", "average", "mean_std", "drive",
]
