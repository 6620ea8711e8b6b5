"""vidtome_b200 — B200-native (sm_100a) implementation of VidToMe's cross-frame token-merge hot path.

Drop-in for the reference package `vidtome` on that path: same exports (vidtome/__init__.py:1-4).
The compute runs in hand-written CUDA (tcgen05 / TMA) behind a C-ABI shared library; there is no
CPU or eager-PyTorch fallback — importing is cheap, but the first use requires libvidtome_b200.so
(`python -m vidtome_b200._build`) and a CUDA device.
"""
from . import merge, patch
from .patch import apply_patch, remove_patch, update_patch, collect_from_patch

__version__ = "0.1.0"
__all__ = ["merge", "patch", "apply_patch", "remove_patch", "update_patch", "collect_from_patch"]
