"""suews_b200 -- B200-native SUEWS per-grid-cell timestep physics (CUDA, sm_100a).

Host side: plain Python + ctypes over the C ABI in include/suews_b200.h (no PyTorch).
    SuewsBatch      many cells resident in HBM, advanced many timesteps per launch
    SuewsMultiBatch the same over several GPUs of one box (contiguous blocks of cells, final gather)
    forcing         met-forcing reader / hourly->timestep disaggregation
    supy_driver     drop-in for supy_driver.suews_driver (suews_cal_multitsteps, ...)
"""
from .batch import SuewsBatch, SuewsMultiBatch, OUT_AGG, OUT_ALL_STAGES, OUT_FLUX_ONLY, OUT_FULL, OUT_SUM7, OUT_SUM82  # noqa: F401
from ._lib import SuewsError  # noqa: F401

__version__ = "0.1.0"
