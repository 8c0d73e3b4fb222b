"""ctypes binding of include/vmis.h, shaped like the reference's Rust surface.

Reference interface mirrored here (file:line under the reference tree):
  * ``VMISIndex::new_from_csv``            src/vmisknn/vmis_index.rs:38
  * trait ``SimilarityComputationNew``     src/vmisknn/similarity_indexed.rs:8-24
      ``items_for_session``, ``idf``, ``find_neighbors``, ``find_attributes``
  * ``vmisknn::predict``                   src/vmisknn/mod.rs:118-125

There is no CPU implementation behind these calls: if ``libvmis_b200.so`` is missing the
import of the library fails loudly, and every query needs a B200 (sm_100) device.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.environ.get("VMIS_LIB", os.path.join(_HERE, "libvmis_b200.so"))   # VMIS_LIB: tuning builds only

DEVICE_NONE = -1
ATTR_EXISTS, ATTR_FOR_SALE, ATTR_ADULT = 1, 2, 4


class VmisError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"vmis error {code}: {msg}")
        self.code = code


class _Stats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("n_sessions", "n_sessions_kept", "n_items", "n_pairs_kept", "n_postings",
                                          "max_len", "m_build", "device_bytes")] + [("idf_weighting", C.c_double)]


class _PrebuiltInfo(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("item_files", "session_files", "item_records", "session_records",
                                          "lists_reordered", "duplicate_postings", "pruned_postings")] + [("m_carry", C.c_uint32),
                                                                                       ("prebuilt", C.c_uint32)]


_lib = None
_u64p, _u32p, _f64p, _u8p = C.POINTER(C.c_uint64), C.POINTER(C.c_uint32), C.POINTER(C.c_double), C.POINTER(C.c_uint8)


def load_library():
    """Load libvmis_b200.so (built by ``__graft_entry__.build()`` / ``make -C serenade_b200/csrc``)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        raise ImportError(f"{_LIB_PATH} is missing: build it with `make -C serenade_b200/csrc` "
                          "(there is no Python/CPU fallback for the VMIS-kNN kernels)")
    L = C.CDLL(_LIB_PATH)
    vp, sz, i32, u32, u64, f64 = C.c_void_p, C.c_size_t, C.c_int, C.c_uint32, C.c_uint64, C.c_double
    sig = {
        "vmis_index_from_csv": (vp, [C.c_char_p, sz, f64, i32]),
        "vmis_index_from_csv_ex": (vp, [C.c_char_p, sz, f64, sz, i32]),
        "vmis_sessions_from_csv": (vp, [C.c_char_p]),
        "vmis_sessions_view": (i32, [vp, C.POINTER(_u64p), C.POINTER(_u64p), C.POINTER(_u32p), C.POINTER(sz)]),
        "vmis_sessions_free": (None, [vp]),
        "vmis_index_from_sessions": (vp, [_u64p, _u64p, _u32p, sz, sz, sz, f64, i32]),
        "vmis_index_from_sessions_attrs": (vp, [_u64p, _u64p, _u32p, sz, sz, sz, f64, _u64p, _u8p, sz, i32]),
        "vmis_index_from_sessions_sharded": (vp, [_u64p, _u64p, _u32p, sz, sz, sz, f64, i32, u32, u32]),
        "vmis_index_from_device_sessions": (vp, [vp, vp, vp, sz, sz, sz, f64, i32, u32, u32]),
        "vmis_index_synth": (vp, [u64, u64, u64, sz, sz, f64, i32, u32, u32]),
        "vmis_index_from_avro": (vp, [C.c_char_p, i32]),
        "vmis_index_from_avro_sharded": (vp, [C.c_char_p, i32, u32, u32]),
        "vmis_index_from_avro_ex": (vp, [C.c_char_p, i32, u32, u32, sz]),
        "vmis_index_from_parts_ex": (vp, [_u64p, _u64p, _u32p, _f64p, _u8p, sz, _u64p, _u64p, _u32p, sz, i32, u32, u32, sz]),
        "vmis_index_from_parts": (vp, [_u64p, _u64p, _u32p, _f64p, _u8p, sz, _u64p, _u64p, _u32p, sz, i32, u32, u32]),
        "vmis_index_prebuilt_info": (i32, [vp, C.POINTER(_PrebuiltInfo)]),
        "vmis_index_to_avro": (i32, [vp, C.c_char_p, C.c_char_p, u32]),
        "vmis_index_save": (i32, [vp, C.c_char_p]),
        "vmis_index_load": (vp, [C.c_char_p, i32]),
        "vmis_index_export_shard": (i32, [vp, vp]),
        "vmis_index_attach_shard": (i32, [vp, u32, vp]),
        "vmis_index_attach_shard_ptr": (i32, [vp, u32, vp]),
        "vmis_index_shard_ptr": (vp, [vp]),
        "vmis_index_set_attributes": (i32, [vp, _u64p, _u8p, sz]),
        "vmis_index_free": (None, [vp]),
        "vmis_index_stats": (i32, [vp, C.POINTER(_Stats)]),
        "vmis_predict_batch": (i32, [vp, _u64p, _u32p, u32, u32, u32, u32, i32, _u64p, _f64p, _u32p, vp]),
        "vmis_predict_batch_device": (i32, [vp, vp, vp, u32, u32, u32, u32, i32, vp, vp, vp, vp, vp]),
        "vmis_predict": (i32, [vp, _u64p, sz, sz, sz, sz, i32, _u64p, _f64p]),
        "vmis_find_neighbors_batch": (i32, [vp, _u64p, _u32p, u32, u32, u32, _u32p, _f64p, _u32p, vp]),
        "vmis_items_for_session": (_u64p, [vp, u32, C.POINTER(sz)]),
        "vmis_idf": (i32, [vp, u64, _f64p]),
        "vmis_find_attributes": (i32, [vp, u64]),
        "vmis_postings": (sz, [vp, u64, _u32p, sz]),
        "vmis_session_timestamp": (i32, [vp, u32, _u32p]),
        "vmis_synth_sessions": (i32, [u64, u64, u64, _u64p, _u64p, _u32p, _u64p]),
        "vmis_synth_queries": (i32, [u64, u64, u32, u32, _u64p, _u32p]),
        "vmis_batcher_create": (vp, [vp, u32, u32, u32, i32, u32, u32]),
        "vmis_batcher_predict": (i32, [vp, _u64p, sz, _u64p, _f64p]),
        "vmis_batcher_stats": (i32, [vp, _u64p, _u64p]),
        "vmis_batcher_destroy": (None, [vp]),
        "vmis_batcher_load_test": (C.c_longlong, [vp, _u64p, _u32p, u32, u32, f64, u32, C.POINTER(C.c_float), sz, _f64p]),
        "vmis_predict_latency_test": (C.c_longlong, [vp, _u64p, _u32p, u32, u32, u32, u32, i32, u32, C.POINTER(C.c_float)]),
        "vmis_server_create": (vp, [vp, u32, u32, u32, u32, i32, u32, u32, u64, u64]),
        "vmis_server_recommend": (i32, [vp, C.c_char_p, u64, i32, _u64p, _f64p]),
        "vmis_server_session_window": (i32, [vp, C.c_char_p, u64, i32, _u64p, sz]),
        "vmis_server_stored_items": (i32, [vp, C.c_char_p, _u64p, sz]),
        "vmis_server_set_clock": (i32, [vp, u64]),
        "vmis_server_stats": (i32, [vp, _u64p, _u64p, _u64p]),
        "vmis_server_destroy": (None, [vp]),
        "vmis_md5": (None, [C.c_char_p, sz, _u8p]),
        "vmis_last_error": (C.c_char_p, []),
        "vmis_last_error_code": (i32, []),
        "vmis_version": (C.c_char_p, []),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype, fn.argtypes = res, args
    _lib = L
    return L


EXPORTED_SYMBOLS = ("vmis_index_from_csv", "vmis_index_from_csv_ex", "vmis_sessions_from_csv", "vmis_sessions_view",
                    "vmis_sessions_free", "vmis_index_from_sessions", "vmis_index_from_sessions_attrs",
                    "vmis_index_from_sessions_sharded", "vmis_index_export_shard", "vmis_index_attach_shard",
                    "vmis_index_attach_shard_ptr", "vmis_index_shard_ptr", "vmis_index_from_device_sessions",
                    "vmis_index_synth", "vmis_index_from_avro", "vmis_index_from_avro_sharded", "vmis_index_from_avro_ex", "vmis_index_from_parts", "vmis_index_from_parts_ex",
                    "vmis_index_prebuilt_info", "vmis_index_to_avro", "vmis_index_save", "vmis_index_load",
                    "vmis_index_set_attributes", "vmis_index_free", "vmis_index_stats", "vmis_predict_batch",
                    "vmis_predict_batch_device", "vmis_predict", "vmis_find_neighbors_batch", "vmis_items_for_session",
                    "vmis_idf", "vmis_find_attributes", "vmis_postings", "vmis_session_timestamp",
                    "vmis_synth_sessions", "vmis_synth_queries", "vmis_batcher_create", "vmis_batcher_predict",
                    "vmis_batcher_stats", "vmis_batcher_destroy", "vmis_batcher_load_test", "vmis_predict_latency_test", "vmis_server_create", "vmis_server_recommend",
                    "vmis_server_session_window", "vmis_server_stored_items", "vmis_server_set_clock", "vmis_server_stats",
                    "vmis_server_destroy", "vmis_md5", "vmis_last_error", "vmis_last_error_code",
                    "vmis_version")


def _check(rc):
    if rc < 0:
        raise VmisError(rc, load_library().vmis_last_error().decode(errors="replace"))
    return rc


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def _csr(sessions):
    """list of item lists -> (items u64, off u32)"""
    lens = np.fromiter((len(s) for s in sessions), dtype=np.int64, count=len(sessions))
    off = np.zeros(len(sessions) + 1, dtype=np.uint32)
    np.cumsum(lens, out=off[1:])
    items = np.fromiter((int(i) for s in sessions for i in s), dtype=np.uint64, count=int(off[-1]))
    return items, off


class VMISIndex:
    """``VMISIndex`` (vmis_index.rs:28-35) resident in HBM; implements ``SimilarityComputationNew``."""

    def __init__(self, handle):
        if not handle:
            raise VmisError(load_library().vmis_last_error_code(), load_library().vmis_last_error().decode(errors="replace"))
        self._h = C.c_void_p(handle)

    # -- constructors -------------------------------------------------------------------------------
    @classmethod
    def new_from_csv(cls, path_to_training, m_most_recent_sessions, idf_weighting, max_len=0, device=0):
        """vmis_index.rs:38 (max_len=0 → p99.5 of the session lengths, :67)."""
        L = load_library()
        return cls(L.vmis_index_from_csv_ex(os.fsencode(path_to_training), m_most_recent_sessions,
                                            float(idf_weighting), max_len, device))

    @classmethod
    def from_sessions(cls, items, sess_off, sess_ts, m_most_recent_sessions, max_len, idf_weighting, device=0,
                      attributes=None):
        """prepare_hashmap (vmis_index.rs:422) + struct assembly (:75-82).  attributes: optional (item ids, flags)."""
        L = load_library()
        items = np.ascontiguousarray(items, dtype=np.uint64)
        sess_off = np.ascontiguousarray(sess_off, dtype=np.uint64)
        sess_ts = np.ascontiguousarray(sess_ts, dtype=np.uint32)
        if attributes is not None:
            a_items = np.ascontiguousarray(attributes[0], dtype=np.uint64)
            a_flags = np.ascontiguousarray(attributes[1], dtype=np.uint8)
            return cls(L.vmis_index_from_sessions_attrs(_p(items, C.c_uint64), _p(sess_off, C.c_uint64),
                                                        _p(sess_ts, C.c_uint32), len(sess_ts), m_most_recent_sessions,
                                                        max_len, float(idf_weighting), _p(a_items, C.c_uint64),
                                                        _p(a_flags, C.c_uint8), len(a_items), device))
        return cls(L.vmis_index_from_sessions(_p(items, C.c_uint64), _p(sess_off, C.c_uint64),
                                              _p(sess_ts, C.c_uint32), len(sess_ts), m_most_recent_sessions, max_len,
                                              float(idf_weighting), device))

    @classmethod
    def from_sessions_sharded(cls, items, sess_off, sess_ts, m_most_recent_sessions, max_len, idf_weighting, device,
                              shard, n_shards):
        """Item-sharded postings (BASELINE config 5): this handle owns shard `shard` of `n_shards`; attach the
        peers with connect_shards() (one process per GPU) or attach_shard_ptr() (same process)."""
        L = load_library()
        items = np.ascontiguousarray(items, dtype=np.uint64)
        sess_off = np.ascontiguousarray(sess_off, dtype=np.uint64)
        sess_ts = np.ascontiguousarray(sess_ts, dtype=np.uint32)
        return cls(L.vmis_index_from_sessions_sharded(_p(items, C.c_uint64), _p(sess_off, C.c_uint64),
                                                      _p(sess_ts, C.c_uint32), len(sess_ts), m_most_recent_sessions,
                                                      max_len, float(idf_weighting), device, shard, n_shards))

    @classmethod
    def from_device_sessions(cls, d_items_ptr, d_off_ptr, d_ts_ptr, n_sessions, m_most_recent_sessions, max_len,
                             idf_weighting, device=0, shard=0, n_shards=1):
        """on-device prepare_hashmap over session arrays resident in HBM (raw device pointers)"""
        L = load_library()
        return cls(L.vmis_index_from_device_sessions(C.c_void_p(d_items_ptr), C.c_void_p(d_off_ptr), C.c_void_p(d_ts_ptr),
                                                     n_sessions, m_most_recent_sessions, max_len, float(idf_weighting),
                                                     device, shard, n_shards))

    @classmethod
    def synth(cls, seed, n_items, n_sessions, m_most_recent_sessions, max_len, idf_weighting, device=0, shard=0,
              n_shards=1):
        """synthetic workload generated and indexed entirely on the device (BASELINE configs 4-5)"""
        L = load_library()
        return cls(L.vmis_index_synth(seed, n_items, n_sessions, m_most_recent_sessions, max_len, float(idf_weighting),
                                      device, shard, n_shards))

    @classmethod
    def new(cls, base_path, device=0, shard=0, n_shards=1, max_session_len=0):
        """``VMISIndex::new(base_path)`` (vmis_index.rs:85): the production index from ``<base_path>/itemindex/*.avro``
        and ``<base_path>/sessionindex/*.avro``.  max_session_len > 0 drops longer sessions from the posting lists."""
        return cls(load_library().vmis_index_from_avro_ex(os.fsencode(base_path), device, shard, n_shards, max_session_len))

    @classmethod
    def from_parts(cls, item_ids, post_off, post_sessions, idf, attr, items, sess_off, sess_ts, device=0, shard=0,
                   n_shards=1, max_session_len=0):
        """The pre-computed index of ``VMISIndex::new`` from arrays in memory (posting lists, idf, attributes as given)."""
        L = load_library()
        item_ids = np.ascontiguousarray(item_ids, dtype=np.uint64)
        post_off = np.ascontiguousarray(post_off, dtype=np.uint64)
        post_sessions = np.ascontiguousarray(post_sessions, dtype=np.uint32)
        idf = np.ascontiguousarray(idf, dtype=np.float64)
        attr = None if attr is None else np.ascontiguousarray(attr, dtype=np.uint8)
        items = np.ascontiguousarray(items, dtype=np.uint64)
        sess_off = np.ascontiguousarray(sess_off, dtype=np.uint64)
        sess_ts = np.ascontiguousarray(sess_ts, dtype=np.uint32)
        return cls(L.vmis_index_from_parts_ex(_p(item_ids, C.c_uint64), _p(post_off, C.c_uint64),
                                              _p(post_sessions, C.c_uint32), _p(idf, C.c_double),
                                              None if attr is None else _p(attr, C.c_uint8), len(item_ids),
                                              _p(items, C.c_uint64), _p(sess_off, C.c_uint64), _p(sess_ts, C.c_uint32),
                                              len(sess_ts), device, shard, n_shards, max_session_len))

    def prebuilt_info(self):
        pi = _PrebuiltInfo()
        _check(load_library().vmis_index_prebuilt_info(self._h, C.byref(pi)))
        return {n: getattr(pi, n) for n, _ in _PrebuiltInfo._fields_}

    def to_avro(self, base_path, codec="deflate", n_files=1):
        """write the index in the production on-disk format that ``VMISIndex::new`` reads (itemindex/ + sessionindex/)"""
        _check(load_library().vmis_index_to_avro(self._h, os.fsencode(base_path), codec.encode(), n_files))

    def save(self, path):
        """serialise the HBM arrays of this handle (fast restart)"""
        _check(load_library().vmis_index_save(self._h, os.fsencode(path)))

    @classmethod
    def load(cls, path, device=0):
        return cls(load_library().vmis_index_load(os.fsencode(path), device))

    def export_shard(self):
        buf = C.create_string_buffer(64)
        _check(load_library().vmis_index_export_shard(self._h, buf))
        return buf.raw

    def attach_shard(self, shard, handle_bytes):
        _check(load_library().vmis_index_attach_shard(self._h, shard, C.create_string_buffer(handle_bytes, 64)))

    def attach_shard_ptr(self, shard, device_ptr):
        _check(load_library().vmis_index_attach_shard_ptr(self._h, shard, C.c_void_p(device_ptr)))

    def shard_ptr(self):
        return load_library().vmis_index_shard_ptr(self._h)

    def connect_shards(self, rank, world, group=None):
        """exchange the CUDA IPC handles of all ranks' shards through torch.distributed and attach the peers"""
        import torch.distributed as dist
        handles = [None] * world
        dist.all_gather_object(handles, self.export_shard(), group=group)
        for s, h in enumerate(handles):
            if s != rank:
                self.attach_shard(s, h)

    def close(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.vmis_index_free(self._h)
        self._h = None

    __del__ = close

    @property
    def handle(self):
        return self._h

    def stats(self):
        st = _Stats()
        _check(load_library().vmis_index_stats(self._h, C.byref(st)))
        return {n: getattr(st, n) for n, _ in _Stats._fields_}

    def set_attributes(self, items, flags):
        items = np.ascontiguousarray(items, dtype=np.uint64)
        flags = np.ascontiguousarray(flags, dtype=np.uint8)
        _check(load_library().vmis_index_set_attributes(self._h, _p(items, C.c_uint64), _p(flags, C.c_uint8),
                                                        len(items)))

    # -- trait SimilarityComputationNew ---------------------------------------------------------------
    def items_for_session(self, session_idx):
        n = C.c_size_t()
        ptr = load_library().vmis_items_for_session(self._h, session_idx, C.byref(n))
        if not ptr:
            raise IndexError(session_idx)
        return np.ctypeslib.as_array(ptr, shape=(n.value,)).copy() if n.value else np.zeros(0, np.uint64)

    def idf(self, item_id):
        out = C.c_double()
        if load_library().vmis_idf(self._h, int(item_id), C.byref(out)) != 0:
            raise KeyError(item_id)   # the reference panics (vmis_index.rs:322)
        return out.value

    def find_attributes(self, item_id):
        a = load_library().vmis_find_attributes(self._h, int(item_id))
        return None if not (a & ATTR_EXISTS) else {"is_for_sale": bool(a & ATTR_FOR_SALE),
                                                    "is_adult": bool(a & ATTR_ADULT)}

    def postings(self, item_id, cap=1 << 16):
        buf = np.zeros(cap, dtype=np.uint32)
        n = load_library().vmis_postings(self._h, int(item_id), _p(buf, C.c_uint32), cap)
        return buf[:min(n, cap)].copy()

    def session_timestamp(self, session_idx):
        out = C.c_uint32()
        _check(load_library().vmis_session_timestamp(self._h, session_idx, C.byref(out)))
        return out.value

    def find_neighbors(self, evolving_session, k, m):
        """-> (session ids, similarities), best first (vmis_index.rs:325-415)."""
        sess, sim, cnt = self.find_neighbors_batch([evolving_session], k, m)
        return sess[0, :cnt[0]].copy(), sim[0, :cnt[0]].copy()

    def find_neighbors_batch(self, sessions, k, m):
        q_items, q_off = sessions if isinstance(sessions, tuple) else _csr(sessions)
        n_q = len(q_off) - 1
        sess = np.zeros((n_q, max(k, 1)), dtype=np.uint32)
        sim = np.zeros((n_q, max(k, 1)), dtype=np.float64)
        cnt = np.zeros(n_q, dtype=np.uint32)
        _check(load_library().vmis_find_neighbors_batch(self._h, _p(q_items, C.c_uint64), _p(q_off, C.c_uint32), n_q,
                                                        k, m, _p(sess, C.c_uint32), _p(sim, C.c_double),
                                                        _p(cnt, C.c_uint32), None))
        return sess, sim, cnt


def predict_batch(index, sessions, k, m, how_many, enable_business_logic=False, out=None, stream=None):
    """Batched ``predict``: sessions is a list of item-id lists or a CSR pair (q_items u64, q_off u32).
    Returns (ids [n_q, how_many] u64, scores f64, counts u32)."""
    q_items, q_off = sessions if isinstance(sessions, tuple) else _csr(sessions)
    q_items = np.ascontiguousarray(q_items, dtype=np.uint64)
    q_off = np.ascontiguousarray(q_off, dtype=np.uint32)
    n_q = len(q_off) - 1
    if out is None:
        ids = np.zeros((n_q, max(how_many, 1)), dtype=np.uint64)
        sc = np.zeros((n_q, max(how_many, 1)), dtype=np.float64)
        cnt = np.zeros(n_q, dtype=np.uint32)
    else:
        ids, sc, cnt = out
    _check(load_library().vmis_predict_batch(index.handle, _p(q_items, C.c_uint64), _p(q_off, C.c_uint32), n_q, k, m,
                                             how_many, int(enable_business_logic), _p(ids, C.c_uint64),
                                             _p(sc, C.c_double), _p(cnt, C.c_uint32), stream))
    return ids, sc, cnt


def predict(index, evolving_session, k, m, how_many, enable_business_logic=False):
    """``vmisknn::predict`` (mod.rs:118-125) → list of (item_id, score), score-descending — the order of
    ``recommendations.into_sorted_vec()`` at the reference call sites (recommend_resource.rs:58-62)."""
    ev = np.ascontiguousarray(evolving_session, dtype=np.uint64)
    ids = np.zeros(max(how_many, 1), dtype=np.uint64)
    sc = np.zeros(max(how_many, 1), dtype=np.float64)
    n = _check(load_library().vmis_predict(index.handle, _p(ev, C.c_uint64), len(ev), k, m, how_many,
                                           int(enable_business_logic), _p(ids, C.c_uint64), _p(sc, C.c_double)))
    return [(int(ids[i]), float(sc[i])) for i in range(n)]


def read_sessions_csv(path):
    """``read_from_file`` (vmis_index.rs:591-752) → (items u64, sess_off u64, sess_ts u32), parsed once for many index
    builds (``VMISIndex.from_sessions(items, off, ts, m, 0, idf_weighting)`` == ``VMISIndex.new_from_csv``)."""
    L = load_library()
    h = L.vmis_sessions_from_csv(os.fsencode(path))
    if not h:
        raise VmisError(L.vmis_last_error_code(), L.vmis_last_error().decode(errors="replace"))
    h = C.c_void_p(h)
    try:
        pi, po, pt, n = _u64p(), _u64p(), _u32p(), C.c_size_t()
        _check(L.vmis_sessions_view(h, C.byref(pi), C.byref(po), C.byref(pt), C.byref(n)))
        off = np.ctypeslib.as_array(po, shape=(n.value + 1,)).copy()
        ts = np.ctypeslib.as_array(pt, shape=(n.value,)).copy() if n.value else np.zeros(0, np.uint32)
        items = np.ctypeslib.as_array(pi, shape=(int(off[-1]),)).copy() if off[-1] else np.zeros(0, np.uint64)
        return items, off, ts
    finally:
        L.vmis_sessions_free(h)


def synth_sessions(seed, n_items, n_sessions):
    """Synthetic training sessions → (items u64, sess_off u64, sess_ts u32)."""
    L = load_library()
    total = C.c_uint64()
    _check(L.vmis_synth_sessions(seed, n_items, n_sessions, None, None, None, C.byref(total)))
    items = np.empty(total.value, dtype=np.uint64)
    off = np.empty(n_sessions + 1, dtype=np.uint64)
    ts = np.empty(n_sessions, dtype=np.uint32)
    _check(L.vmis_synth_sessions(seed, n_items, n_sessions, _p(items, C.c_uint64), _p(off, C.c_uint64),
                                 _p(ts, C.c_uint32), C.byref(total)))
    return items, off, ts


def synth_queries(seed, n_items, n_q, max_items_in_session=4):
    """Synthetic evolving sessions (evaluator.rs:46-57 shape) → CSR (q_items u64, q_off u32)."""
    L = load_library()
    q_items = np.empty(n_q * max_items_in_session, dtype=np.uint64)
    q_off = np.empty(n_q + 1, dtype=np.uint32)
    _check(L.vmis_synth_queries(seed, n_items, n_q, max_items_in_session, _p(q_items, C.c_uint64),
                                _p(q_off, C.c_uint32)))
    return q_items[:q_off[-1]].copy(), q_off


class Batcher:
    """Micro-batching front for the online call shape (one evolving session per caller thread, as actix workers
    call predict at recommend_resource.rs:56)."""

    def __init__(self, index, k, m, how_many, enable_business_logic=False, max_batch=4096, max_wait_us=200):
        self._index = index                       # keep the index alive
        self._how_many = how_many
        self._b = C.c_void_p(load_library().vmis_batcher_create(index.handle, k, m, how_many, int(enable_business_logic),
                                                               max_batch, max_wait_us))
        if not self._b:
            raise VmisError(-1, "vmis_batcher_create failed")

    def predict(self, evolving_session):
        ev = np.ascontiguousarray(evolving_session, dtype=np.uint64)
        ids = np.zeros(max(self._how_many, 1), dtype=np.uint64)
        sc = np.zeros(max(self._how_many, 1), dtype=np.float64)
        n = _check(load_library().vmis_batcher_predict(self._b, _p(ev, C.c_uint64), len(ev), _p(ids, C.c_uint64),
                                                       _p(sc, C.c_double)))
        return [(int(ids[i]), float(sc[i])) for i in range(n)]

    def stats(self):
        a, b = C.c_uint64(), C.c_uint64()
        _check(load_library().vmis_batcher_stats(self._b, C.byref(a), C.byref(b)))
        return {"batches": a.value, "requests": b.value}

    def load_test(self, sessions, target_rps, duration_ms=1000, n_threads=64):
        """open-loop replay of the evolving sessions (CSR pair) at target_rps → (latencies in us, achieved rps)"""
        q_items, q_off = sessions
        q_items = np.ascontiguousarray(q_items, dtype=np.uint64)
        q_off = np.ascontiguousarray(q_off, dtype=np.uint32)
        cap = int(target_rps * duration_ms / 1e3) + 1
        lat = np.zeros(cap, dtype=np.float32)
        rps = C.c_double()
        n = load_library().vmis_batcher_load_test(self._b, _p(q_items, C.c_uint64), _p(q_off, C.c_uint32), len(q_off) - 1,
                                                  n_threads, float(target_rps), int(duration_ms),
                                                  _p(lat, C.c_float), cap, C.byref(rps))
        _check(int(n) if n < 0 else 0)
        return lat[:min(int(n), cap)], rps.value

    def close(self):
        if getattr(self, "_b", None) and _lib is not None:
            _lib.vmis_batcher_destroy(self._b)
        self._b = None

    __del__ = close


class Server:
    """``GET /v1/recommend`` without the HTTP layer (recommend_resource.rs:20-65): evolving-session window over an
    in-process store with RocksDBSessionStore semantics (sessions/mod.rs), then ``predict`` through a micro-batcher."""

    def __init__(self, index, k, m, how_many, max_items_in_session, enable_business_logic=False, max_batch=4096,
                 max_wait_us=200, session_ttl_secs=30 * 60, max_session_idle_secs=20 * 60):
        self._index = index
        self._how_many = how_many
        self._cap = max(max_items_in_session, 1)
        self._s = C.c_void_p(load_library().vmis_server_create(index.handle, k, m, how_many, max_items_in_session,
                                                              int(enable_business_logic), max_batch, max_wait_us,
                                                              session_ttl_secs, max_session_idle_secs))
        if not self._s:
            raise VmisError(-1, "vmis_server_create failed")

    def recommend(self, session_id, item_id, user_consent=True):
        """v1_recommend → list of item ids, best first"""
        ids = np.zeros(max(self._how_many, 1), dtype=np.uint64)
        n = _check(load_library().vmis_server_recommend(self._s, session_id.encode(), int(item_id), int(user_consent),
                                                        _p(ids, C.c_uint64), None))
        return [int(x) for x in ids[:n]]

    def session_window(self, session_id, item_id, user_consent=True):
        buf = np.zeros(self._cap + 1, dtype=np.uint64)
        n = _check(load_library().vmis_server_session_window(self._s, session_id.encode(), int(item_id), int(user_consent),
                                                             _p(buf, C.c_uint64), len(buf)))
        return [int(x) for x in buf[:n]]

    def stored_items(self, session_id):
        buf = np.zeros(self._cap + 1, dtype=np.uint64)
        n = _check(load_library().vmis_server_stored_items(self._s, session_id.encode(), _p(buf, C.c_uint64), len(buf)))
        return [int(x) for x in buf[:n]]

    def set_clock(self, epoch_secs):
        _check(load_library().vmis_server_set_clock(self._s, int(epoch_secs)))

    def stats(self):
        a, b, c = C.c_uint64(), C.c_uint64(), C.c_uint64()
        _check(load_library().vmis_server_stats(self._s, C.byref(a), C.byref(b), C.byref(c)))
        return {"sessions": a.value, "batches": b.value, "requests": c.value}

    def close(self):
        if getattr(self, "_s", None) and _lib is not None:
            _lib.vmis_server_destroy(self._s)
        self._s = None

    __del__ = close


def md5(data):
    out = np.zeros(16, dtype=np.uint8)
    load_library().vmis_md5(data, len(data), _p(out, C.c_uint8))
    return bytes(out)
