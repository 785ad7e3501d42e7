"""rmdetect_b200 — B200-native scan hot path for RMDetect module models.

Host side (Python, mirrors the reference's lib/ interface for this path) over a
C-ABI CUDA library (`include/rmdetect_b200.h`, built from `csrc/`).  Importing
the package never touches the GPU; the native library is loaded on first use
and its absence is an error, never a CPU fallback.
"""
__version__ = "0.1.0"
