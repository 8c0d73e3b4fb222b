"""Minimal Avro object-container WRITER for the tests (the image has no avro / fastavro package).

Writes the two record layouts VMISIndex::new reads (vmis_index.rs:184-192, :249-255) with the variations real
producers show: null / deflate / snappy block codecs, several blocks per file, several files per directory,
nullable-union column types (Spark), shuffled field order and extra columns the loader has to skip.
"""
import json
import os
import struct
import zlib

import numpy as np

SYNC = bytes(range(16))


def zz(n):
    """zigzag varint of a signed integer"""
    n = int(n)
    u = (n << 1) ^ (n >> 63)
    u &= (1 << 64) - 1
    out = bytearray()
    while True:
        b = u & 0x7F
        u >>= 7
        if u:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def s64(x):
    """u64 → the i64 the Avro long carries (`ItemId as u64` on the way back, vmis_index.rs:222)"""
    x = int(x)
    return x - (1 << 64) if x >= 1 << 63 else x


def s32(x):
    x = int(x)
    return x - (1 << 32) if x >= 1 << 31 else x


def _bytes(b):
    return zz(len(b)) + b


def _array(items, enc, block=0):
    """Avro array; block > 0 cuts it into several blocks, every second one with the negative-count + size form"""
    if not len(items):
        return zz(0)
    out = bytearray()
    step = block if block else len(items)
    for bi, i in enumerate(range(0, len(items), step)):
        chunk = b"".join(enc(x) for x in items[i:i + step])
        n = len(items[i:i + step])
        if block and bi % 2 == 1:
            out += zz(-n) + zz(len(chunk)) + chunk
        else:
            out += zz(n) + chunk
    return bytes(out + zz(0))


def snappy_block(raw):
    import pyarrow as pa
    comp = pa.Codec("snappy").compress(raw, asbytes=True)
    return comp + struct.pack(">I", zlib.crc32(raw) & 0xFFFFFFFF)


def write_container(path, schema, records, codec="null", records_per_block=100, sync=SYNC):
    """records: list of already encoded record bodies"""
    meta = {"avro.schema": json.dumps(schema).encode(), "avro.codec": codec.encode()}
    out = bytearray(b"Obj\x01")
    out += zz(len(meta))
    for k, v in meta.items():
        out += _bytes(k.encode()) + _bytes(v)
    out += zz(0) + sync
    for i in range(0, len(records), records_per_block):
        chunk = records[i:i + records_per_block]
        raw = b"".join(chunk)
        if codec == "deflate":
            c = zlib.compressobj(6, zlib.DEFLATED, -15)
            raw = c.compress(raw) + c.flush()
        elif codec == "snappy":
            raw = snappy_block(raw)
        elif codec != "null":
            pass                                           # unknown codec names are written verbatim (error tests)
        out += zz(len(chunk)) + zz(len(raw)) + raw + sync
    with open(path, "wb") as f:
        f.write(out)


def item_schema(style):
    """style 'plain': the five columns in struct order; 'spark': nullable unions, shuffled order, extra columns"""
    if style == "plain":
        return {"type": "record", "name": "ItemIndex", "fields": [
            {"name": "ItemId", "type": "long"},
            {"name": "session_indices_time_ordered", "type": {"type": "array", "items": "int"}},
            {"name": "idf", "type": "double"},
            {"name": "ForSale", "type": "boolean"},
            {"name": "IsAdult", "type": "boolean"}]}
    return {"type": "record", "name": "topLevelRecord", "namespace": "com.example", "fields": [
        {"name": "Title", "type": ["null", "string"], "default": None},
        {"name": "idf", "type": ["double", "null"]},
        {"name": "IsAdult", "type": ["null", "boolean"]},
        {"name": "Stats", "type": {"type": "record", "name": "Stats", "fields": [
            {"name": "views", "type": "long"}, {"name": "tags", "type": {"type": "map", "values": "string"}},
            {"name": "kind", "type": {"type": "enum", "name": "Kind", "symbols": ["A", "B"]}},
            {"name": "digest", "type": {"type": "fixed", "name": "MD5", "size": 16}},
            {"name": "ratio", "type": "float"}, {"name": "blob", "type": "bytes"}]}},
        {"name": "ItemId", "type": ["null", "long"]},
        {"name": "ForSale", "type": "boolean"},
        {"name": "session_indices_time_ordered", "type": ["null", {"type": "array", "items": ["int", "null"]}]},
        {"name": "again", "type": ["null", "Stats"]}]}


def encode_item(style, item_id, sessions, idf, for_sale, adult, n=0):
    if style == "plain":
        return (zz(s64(item_id)) + _array(list(sessions), lambda s: zz(s32(s))) +
                struct.pack("<d", idf) + bytes([1 if for_sale else 0]) + bytes([1 if adult else 0]))
    stats = (zz(n) + _array([("k%d" % n, "v" * (n % 5))], lambda kv: _bytes(kv[0].encode()) + _bytes(kv[1].encode())) +
             zz(n % 2) + bytes(range(16)) + struct.pack("<f", 0.5) + _bytes(b"\x00\x01\x02" * (n % 4)))
    title = zz(0) if n % 3 == 0 else zz(1) + _bytes(("item %d " % item_id * 8).encode())
    return (title + zz(0) + struct.pack("<d", idf) + zz(1) + bytes([1 if adult else 0]) + stats +
            zz(1) + zz(s64(item_id)) + bytes([1 if for_sale else 0]) +
            zz(1) + _array(list(sessions), lambda s: zz(0) + zz(s32(s)), block=3) +
            (zz(0) if n % 2 else zz(1) + stats))


def session_schema(style):
    if style == "plain":
        return {"type": "record", "name": "SessionIndex", "fields": [
            {"name": "SessionIndex", "type": "int"},
            {"name": "item_ids_asc", "type": {"type": "array", "items": "long"}},
            {"name": "Time", "type": "int"}]}
    return {"type": "record", "name": "topLevelRecord", "fields": [
        {"name": "Time", "type": ["int", "null"]},
        {"name": "Visitor", "type": ["null", "string"]},
        {"name": "item_ids_asc", "type": ["null", {"type": "array", "items": ["null", "long"]}]},
        {"name": "SessionIndex", "type": ["null", {"type": "int", "logicalType": "date"}]}]}


def encode_session(style, index, items, time, n=0):
    if style == "plain":
        return zz(index) + _array(list(items), lambda i: zz(s64(i))) + zz(s32(time))
    return (zz(0) + zz(s32(time)) + (zz(0) if n % 2 else zz(1) + _bytes(b"visitor-%d" % n)) +
            zz(1) + _array(list(items), lambda i: zz(1) + zz(s64(i)), block=2) + zz(1) + zz(index))


def write_index_dir(base, parts, style="plain", codec="null", files=2, records_per_block=64):
    """parts: dict(item_ids, post_off, post_sessions, idf, attr, items, off, ts) → base/itemindex, base/sessionindex"""
    os.makedirs(os.path.join(base, "itemindex"), exist_ok=True)
    os.makedirs(os.path.join(base, "sessionindex"), exist_ok=True)
    ids, po, ps = parts["item_ids"], parts["post_off"], parts["post_sessions"]
    recs = [encode_item(style, ids[i], ps[po[i]:po[i + 1]], float(parts["idf"][i]), bool(parts["attr"][i] & 2),
                        bool(parts["attr"][i] & 4), n=i) for i in range(len(ids))]
    for f in range(files):
        write_container(os.path.join(base, "itemindex", "part-%05d.avro" % f), item_schema(style), recs[f::files], codec,
                        records_per_block)
    off, it, ts = parts["off"], parts["items"], parts["ts"]
    srecs = [encode_session(style, s, it[off[s]:off[s + 1]], ts[s], n=s) for s in range(len(ts)) if off[s + 1] > off[s]]
    for f in range(files):
        write_container(os.path.join(base, "sessionindex", "part-%05d.avro" % f), session_schema(style), srecs[f::files],
                        codec, records_per_block)
    # files without the .avro suffix are ignored by the loader (vmis_index.rs:209, :271)
    with open(os.path.join(base, "itemindex", "_SUCCESS"), "wb") as f:
        f.write(b"")
    with open(os.path.join(base, "sessionindex", ".part-00000.avro.crc"), "wb") as f:
        f.write(b"crc")


def parts_from_oracle(oix, items, off, ts):
    """The arrays VMISIndex::new would load for the index `oix` built from the sessions (items, off, ts): posting
    lists, idf and attributes per indexed item, all sessions dense by index."""
    ids = [int(i) for i in np.unique(items) if len(oix.postings(int(i)))]
    post_off, post = [0], []
    idf, attr = [], []
    for i in ids:
        p = oix.postings(i)
        post.extend(int(x) for x in p)
        post_off.append(len(post))
        idf.append(oix.idf(i))
        a = oix.find_attributes(i)
        attr.append(0 if a is None else 1 | (2 if a["is_for_sale"] else 0) | (4 if a["is_adult"] else 0))
    return dict(item_ids=np.array(ids, dtype=np.uint64), post_off=np.array(post_off, dtype=np.uint64),
                post_sessions=np.array(post, dtype=np.uint32), idf=np.array(idf, dtype=np.float64),
                attr=np.array(attr, dtype=np.uint8), items=np.asarray(items, dtype=np.uint64),
                off=np.asarray(off, dtype=np.uint64), ts=np.asarray(ts, dtype=np.uint32))


# ---- independent reader for the two plain layouts (checks the C++ WRITER without going through the C++ reader) ----
def _rd_long(b, p):
    u, shift = 0, 0
    while True:
        c = b[p]
        p += 1
        u |= (c & 0x7F) << shift
        shift += 7
        if not c & 0x80:
            break
    return (u >> 1) ^ -(u & 1), p


def read_container_plain(path, kind):
    """→ list of records of an `item` / `session` container written with the plain schema (null or deflate codec)"""
    b = open(path, "rb").read()
    assert b[:4] == b"Obj\x01"
    p, meta = 4, {}
    while True:
        n, p = _rd_long(b, p)
        if n == 0:
            break
        for _ in range(abs(n)):
            kl, p = _rd_long(b, p); k = b[p:p + kl]; p += kl
            vl, p = _rd_long(b, p); meta[k.decode()] = b[p:p + vl]; p += vl
    sync = b[p:p + 16]
    p += 16
    schema = json.loads(meta["avro.schema"])
    codec = meta.get("avro.codec", b"null").decode()
    out = []
    while p < len(b):
        count, p = _rd_long(b, p)
        size, p = _rd_long(b, p)
        raw = b[p:p + size]
        p += size
        assert b[p:p + 16] == sync
        p += 16
        if codec == "deflate":
            raw = zlib.decompress(raw, -15)
        q = 0
        for _ in range(count):
            def arr(q):
                vals = []
                while True:
                    n, q = _rd_long(raw, q)
                    if n == 0:
                        return vals, q
                    if n < 0:
                        n = -n
                        _, q = _rd_long(raw, q)
                    for _ in range(n):
                        v, q = _rd_long(raw, q)
                        vals.append(v)
            if kind == "item":
                item, q = _rd_long(raw, q)
                sess, q = arr(q)
                idf = struct.unpack("<d", raw[q:q + 8])[0]
                q += 8
                out.append((item & ((1 << 64) - 1), [s & 0xFFFFFFFF for s in sess], idf, bool(raw[q]), bool(raw[q + 1])))
                q += 2
            else:
                idx, q = _rd_long(raw, q)
                items, q = arr(q)
                t, q = _rd_long(raw, q)
                out.append((idx, [i & ((1 << 64) - 1) for i in items], t & 0xFFFFFFFF))
        assert q == len(raw)
    return schema, out
