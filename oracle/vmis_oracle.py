"""ctypes loader for the CPU oracle (oracle/vmis_oracle.cpp).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  Nothing in serenade_b200/
imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "libvmis_oracle.so")

FAITHFUL, CANONICAL = 0, 1


def build(force=False):
    src = os.path.join(_HERE, "vmis_oracle.cpp")
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB)
        u64p, u32p, f64p, f32p = (C.POINTER(C.c_uint64), C.POINTER(C.c_uint32), C.POINTER(C.c_double),
                                  C.POINTER(C.c_float))
        L.vo_index_from_sessions.restype = C.c_void_p
        L.vo_index_from_sessions.argtypes = [u64p, u64p, u32p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_double]
        L.vo_index_from_csv.restype = C.c_void_p
        L.vo_index_from_csv.argtypes = [C.c_char_p, C.c_size_t, C.c_double, C.c_size_t]
        L.vo_index_from_parts.restype = C.c_void_p
        L.vo_index_from_parts.argtypes = [u64p, u64p, u32p, f64p, C.POINTER(C.c_uint8), C.c_size_t, u64p, u64p, u32p,
                                          C.c_size_t]
        L.vo_index_free.argtypes = [C.c_void_p]
        for f in ("vo_num_sessions", "vo_num_items", "vo_kept_pairs", "vo_max_len"):
            getattr(L, f).restype = C.c_size_t
            getattr(L, f).argtypes = [C.c_void_p]
        L.vo_session_ts.restype = C.c_uint32
        L.vo_session_ts.argtypes = [C.c_void_p, C.c_uint32]
        L.vo_items_for_session.restype = C.c_size_t
        L.vo_items_for_session.argtypes = [C.c_void_p, C.c_uint32, u64p, C.c_size_t]
        L.vo_idf.restype = C.c_int
        L.vo_idf.argtypes = [C.c_void_p, C.c_uint64, f64p]
        L.vo_postings.restype = C.c_size_t
        L.vo_postings.argtypes = [C.c_void_p, C.c_uint64, u32p, C.c_size_t]
        L.vo_find_attributes.restype = C.c_int
        L.vo_find_attributes.argtypes = [C.c_void_p, C.c_uint64]
        L.vo_set_attributes.argtypes = [C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.c_int]
        L.vo_find_neighbors.restype = C.c_int
        L.vo_find_neighbors.argtypes = [C.c_void_p, u64p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_int, u32p, f64p]
        L.vo_predict.restype = C.c_int
        L.vo_predict.argtypes = [C.c_void_p, u64p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_size_t, C.c_int, C.c_int,
                                 u64p, f64p]
        L.vo_predict_batch.restype = C.c_double
        L.vo_predict_batch.argtypes = [C.c_void_p, u64p, u32p, C.c_uint32, C.c_size_t, C.c_size_t, C.c_size_t,
                                       C.c_int, C.c_int, C.c_int, u64p, f64p, u32p, f32p]
        L.vo_heap_kat.restype = C.c_int
        L.vo_heap_kat.argtypes = [C.c_int, u64p]
        _lib = L
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


class OracleIndex:
    """Mirror of VMISIndex (vmis_index.rs:28-35) over the CPU restatement."""

    def __init__(self, handle):
        if not handle:
            raise RuntimeError("oracle index construction failed")
        self._h = C.c_void_p(handle)

    @classmethod
    def new_from_csv(cls, path, m, idf_weighting, max_len=0):
        return cls(lib().vo_index_from_csv(os.fsencode(path), m, float(idf_weighting), max_len))

    @classmethod
    def from_sessions(cls, items, off, ts, m, max_len, idf_weighting):
        items = np.ascontiguousarray(items, dtype=np.uint64)
        off = np.ascontiguousarray(off, dtype=np.uint64)
        ts = np.ascontiguousarray(ts, dtype=np.uint32)
        return cls(lib().vo_index_from_sessions(_p(items, C.c_uint64), _p(off, C.c_uint64), _p(ts, C.c_uint32),
                                                len(ts), m, max_len, float(idf_weighting)))

    @classmethod
    def from_parts(cls, item_ids, post_off, post_sessions, idf, attr, items, off, ts):
        """VMISIndex::new (vmis_index.rs:85-313): pre-computed posting lists / idf / attributes, dense sessions."""
        item_ids = np.ascontiguousarray(item_ids, dtype=np.uint64)
        post_off = np.ascontiguousarray(post_off, dtype=np.uint64)
        post_sessions = np.ascontiguousarray(post_sessions, dtype=np.uint32)
        idf = np.ascontiguousarray(idf, dtype=np.float64)
        attr = None if attr is None else np.ascontiguousarray(attr, dtype=np.uint8)
        items = np.ascontiguousarray(items, dtype=np.uint64)
        off = np.ascontiguousarray(off, dtype=np.uint64)
        ts = np.ascontiguousarray(ts, dtype=np.uint32)
        return cls(lib().vo_index_from_parts(_p(item_ids, C.c_uint64), _p(post_off, C.c_uint64),
                                             _p(post_sessions, C.c_uint32), _p(idf, C.c_double),
                                             None if attr is None else _p(attr, C.c_uint8), len(item_ids),
                                             _p(items, C.c_uint64), _p(off, C.c_uint64), _p(ts, C.c_uint32), len(ts)))

    def __del__(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.vo_index_free(self._h)
            self._h = None

    num_sessions = property(lambda s: lib().vo_num_sessions(s._h))
    num_items = property(lambda s: lib().vo_num_items(s._h))
    kept_pairs = property(lambda s: lib().vo_kept_pairs(s._h))
    max_len = property(lambda s: lib().vo_max_len(s._h))

    def session_ts(self, s):
        return lib().vo_session_ts(self._h, s)

    def items_for_session(self, s):
        buf = np.zeros(4096, dtype=np.uint64)
        n = lib().vo_items_for_session(self._h, s, _p(buf, C.c_uint64), len(buf))
        return buf[:n].copy()

    def idf(self, item):
        out = C.c_double()
        if lib().vo_idf(self._h, item, C.byref(out)) != 0:
            raise KeyError(item)
        return out.value

    def postings(self, item, cap=1 << 16):
        buf = np.zeros(cap, dtype=np.uint32)
        n = lib().vo_postings(self._h, item, _p(buf, C.c_uint32), cap)
        return buf[:n].copy()

    def find_attributes(self, item):
        a = lib().vo_find_attributes(self._h, item)
        return None if not (a & 1) else {"is_for_sale": bool(a & 2), "is_adult": bool(a & 4)}

    def set_attributes(self, item, exists=True, for_sale=True, adult=False):
        lib().vo_set_attributes(self._h, item, int(exists), int(for_sale), int(adult))

    def find_neighbors(self, ev, k, m, mode=CANONICAL):
        ev = np.ascontiguousarray(ev, dtype=np.uint64)
        sess = np.zeros(max(k, 1), dtype=np.uint32)
        sim = np.zeros(max(k, 1), dtype=np.float64)
        n = lib().vo_find_neighbors(self._h, _p(ev, C.c_uint64), len(ev), k, m, mode, _p(sess, C.c_uint32),
                                    _p(sim, C.c_double))
        return sess[:n].copy(), sim[:n].copy()

    def predict(self, ev, k, m, how_many, enable_business_logic=False, mode=CANONICAL):
        ev = np.ascontiguousarray(ev, dtype=np.uint64)
        ids = np.zeros(max(how_many, 1), dtype=np.uint64)
        sc = np.zeros(max(how_many, 1), dtype=np.float64)
        n = lib().vo_predict(self._h, _p(ev, C.c_uint64), len(ev), k, m, how_many, int(enable_business_logic), mode,
                             _p(ids, C.c_uint64), _p(sc, C.c_double))
        return ids[:n].copy(), sc[:n].copy()

    def predict_batch(self, q_items, q_off, k, m, how_many, enable_business_logic=False, mode=CANONICAL, threads=1,
                      want_latency=False, want_outputs=True):
        q_items = np.ascontiguousarray(q_items, dtype=np.uint64)
        q_off = np.ascontiguousarray(q_off, dtype=np.uint32)
        n_q = len(q_off) - 1
        ids = np.zeros((n_q, how_many), dtype=np.uint64) if want_outputs else None
        sc = np.zeros((n_q, how_many), dtype=np.float64) if want_outputs else None
        cnt = np.zeros(n_q, dtype=np.uint32)
        lat = np.zeros(n_q, dtype=np.float32) if want_latency else None
        secs = lib().vo_predict_batch(self._h, _p(q_items, C.c_uint64), _p(q_off, C.c_uint32), n_q, k, m, how_many,
                                      int(enable_business_logic), mode, threads,
                                      _p(ids, C.c_uint64) if want_outputs else None,
                                      _p(sc, C.c_double) if want_outputs else None,
                                      _p(cnt, C.c_uint32), _p(lat, C.c_float) if want_latency else None)
        return ids, sc, cnt, secs, lat


def heap_kat(which):
    out = np.zeros(8, dtype=np.uint64)
    n = lib().vo_heap_kat(which, _p(out, C.c_uint64))
    return [int(x) for x in out[:n]]
