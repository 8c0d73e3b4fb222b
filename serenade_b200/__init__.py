"""serenade_b200 — B200-native VMIS-kNN `predict_next` (the hot path of bolcom/serenade).

The compute lives in ``libvmis_b200.so`` (hand-written sm_100a CUDA behind the C ABI of
``include/vmis.h``).  This package is the thin Python mirror of the reference's
``VMISIndex`` / ``predict`` surface used by the tests and the benchmark.
"""
from .vmis import (Batcher, Server, md5, VMISIndex, VmisError, load_library, predict, predict_batch, read_sessions_csv, synth_queries,  # noqa: F401
                   synth_sessions, DEVICE_NONE)

__all__ = ["Batcher", "Server", "md5", "VMISIndex", "VmisError", "load_library", "predict", "predict_batch", "read_sessions_csv", "synth_sessions", "synth_queries",
           "DEVICE_NONE"]
