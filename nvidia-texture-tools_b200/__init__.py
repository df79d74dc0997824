"""nvidia-texture-tools_b200 — B200-native (sm_100a) BCn block compression + mip-chain generation.

The product is `lib/libnvtt_b200.so` (C ABI declared in include/nvtt_b200.h, CUDA kernels in csrc/).  This Python
package is only the thin ctypes view used by tests/ and bench.py; the C++ mirror of the nvtt:: API lives in host/.
The directory name is not a valid Python identifier: load it with `nvtt_b200_loader.load()` at the repo root.
"""
from .capi import *  # noqa: F401,F403
from . import synth  # noqa: E402,F401
from . import sharding  # noqa: E402,F401
