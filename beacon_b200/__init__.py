"""beacon_b200 — B200-native batched environment dynamics for the beacon flow-control benchmark.

Hand-written sm_100a CUDA behind a C-ABI (include/beacon_b200.h, beacon_b200/lib/libbeacon_b200.so),
driven from Python with torch CUDA tensors.  No CPU fallback.
"""
from ._capi import BeaconError, LIB_PATH  # noqa: F401
from .batched import BatchedEnv  # noqa: F401

__all__ = ["BatchedEnv", "BeaconError", "LIB_PATH"]   # also: .vector (VectorEnv, SeparableShkadov), .dist, .peer (LearnerBuffer), .envs
