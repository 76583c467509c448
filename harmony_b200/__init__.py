"""harmony_b200 -- B200-native (CUDA sm_100a) BLS12-381 aggregate/verify backend for Harmony's consensus hot path.

Only the hot path lives here: csrc/ (kernels + C ABI), host/ (C++ mirror of crypto/bls), bls.py (ctypes binding).
There is no CPU fallback: importing `harmony_b200.bls` and calling init() without a CUDA device raises.
"""
__all__ = ["bls"]
