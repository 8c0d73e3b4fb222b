"""Synthetic click-log generator through its own small host library, ``libvmis_synth.so``.

The same generator is compiled into ``libvmis_b200.so`` (``vmis_synth_sessions`` / ``vmis_synth_queries`` of
include/vmis.h); this module exists so that code which must not map the product library — the CPU arm of
``bench.py --impl reference`` — can still build the identical workload.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libvmis_synth.so")
_lib = None
_u64p, _u32p = C.POINTER(C.c_uint64), C.POINTER(C.c_uint32)


def _load():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            raise ImportError(f"{_LIB_PATH} is missing: build it with `make -C serenade_b200/csrc`")
        L = C.CDLL(_LIB_PATH)
        L.vmis_synth_sessions.restype = C.c_int
        L.vmis_synth_sessions.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, _u64p, _u64p, _u32p, _u64p]
        L.vmis_synth_queries.restype = C.c_int
        L.vmis_synth_queries.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32, _u64p, _u32p]
        _lib = L
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def synth_sessions(seed, n_items, n_sessions):
    """Synthetic training sessions → (items u64, sess_off u64, sess_ts u32)."""
    L = _load()
    total = C.c_uint64()
    if L.vmis_synth_sessions(seed, n_items, n_sessions, None, None, None, C.byref(total)) != 0:
        raise ValueError("vmis_synth_sessions: bad arguments")
    items = np.empty(total.value, dtype=np.uint64)
    off = np.empty(n_sessions + 1, dtype=np.uint64)
    ts = np.empty(n_sessions, dtype=np.uint32)
    if L.vmis_synth_sessions(seed, n_items, n_sessions, _p(items, C.c_uint64), _p(off, C.c_uint64), _p(ts, C.c_uint32),
                             C.byref(total)) != 0:
        raise ValueError("vmis_synth_sessions: bad arguments")
    return items, off, ts


def synth_queries(seed, n_items, n_q, max_items_in_session=4):
    """Synthetic evolving sessions (evaluator.rs:46-57 shape) → CSR (q_items u64, q_off u32)."""
    L = _load()
    q_items = np.empty(n_q * max_items_in_session, dtype=np.uint64)
    q_off = np.empty(n_q + 1, dtype=np.uint32)
    if L.vmis_synth_queries(seed, n_items, n_q, max_items_in_session, _p(q_items, C.c_uint64), _p(q_off, C.c_uint32)) != 0:
        raise ValueError("vmis_synth_queries: bad arguments")
    return q_items[:q_off[-1]].copy(), q_off
