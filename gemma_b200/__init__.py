"""gemma_b200 -- B200-native univariate LMM association engine behind GEMMA's -gk/-eigen/-lmm seams.

The compute lives in csrc/ (hand-written sm_100a CUDA behind the C ABI of include/gemma_b200.h);
this package only holds the host-side mirror of that ABI (api.py), the synthetic genotype
generator shared by tests and bench (synth.py) and the SNP sharding helper (shard.py).
"""
from .api import Context, GB200Error, SUMSTAT_DTYPE, load_library, LIB_PATH  # noqa: F401

__all__ = ["Context", "GB200Error", "SUMSTAT_DTYPE", "load_library", "LIB_PATH"]
