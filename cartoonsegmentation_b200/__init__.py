"""cartoonsegmentation_b200 -- B200 (sm_100a) implementation of CartoonSegmentation's per-image hot path.

Python/PyTorch host code mirroring the reference call surface, over hand-written CUDA kernels behind a C ABI
(`include/csb200.h`, `libcsb200.so`).  No CPU fallback: every op raises if the CUDA library is missing.
"""
__version__ = "0.1.0"
