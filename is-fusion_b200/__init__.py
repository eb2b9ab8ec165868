"""isfusion_b200 — B200-native (sm_100a) implementation of the IS-Fusion fusion-neck hot path.

Host side mirrors the reference's mmdet3d module / operator API (same names, argument meaning and
error behaviour); all compute goes through the C-ABI CUDA library ``libisfusion_b200.so``
(``include/isfusion_b200.h``).  There is no CPU fallback.
"""
__version__ = '0.1.0'
