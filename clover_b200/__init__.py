"""clover_b200 -- B200-native (sm_100a) implementation of Clover's video-language hot path.

Importing the package is cheap and GPU-free (integer tables, registry, synthetic data).  Every
compute entry point goes through the C-ABI library ``libclover_b200.so`` (clover_b200/_lib.py) and
raises if it is missing or no CUDA device is present: there is no CPU fallback.
"""
__version__ = "0.1.0"
