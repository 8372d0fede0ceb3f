"""B200-native engine for the ray-batch hot path of nuggy875/NeRF_pytorch_paeng.

The package mirrors the reference's Python call surface (rays, nerf_process, model, config, train,
test, utils) and routes every body to hand-written sm_100a CUDA kernels in libnerf_b200.so through
the C ABI declared in include/nerf_b200.h.  There is no CPU or eager-PyTorch fallback.
"""
from ._lib import NB_BF16, NB_FP32, NBError  # noqa: F401

__version__ = '0.1.0'
