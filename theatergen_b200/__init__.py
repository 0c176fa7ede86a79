"""theatergen_b200 -- B200-native (sm_100a) implementation of TheaterGen's denoising hot path.

Python host code mirrors the reference's surface (ip_adapter/, models/pipelines.py, utils/) and calls
hand-written CUDA kernels through the C-ABI library libtheatergen.so (include/theatergen.h).
There is no CPU fallback: importing the kernels without the built library raises.
"""
__version__ = "0.1.0"
