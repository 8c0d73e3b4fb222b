"""CPU tests of the Avro index loader (VMISIndex::new, vmis_index.rs:85-313): container framing, codecs, schema
variations, "last record wins", and the load-time checks.  Host-only handles (VMIS_DEVICE_NONE): no GPU needed."""
import os

import numpy as np
import pytest

import avro_util as au
from util import random_index_data


def _parts(oracle, rng, n_sessions=300, n_items=60, m=12, max_len=6):
    items, off, ts = random_index_data(rng, n_sessions, n_items, max_len=8, id_scale=1_000_003)
    oix = oracle.OracleIndex.from_sessions(items, off, ts, m, max_len, 2.0)
    parts = au.parts_from_oracle(oix, items, off, ts)
    # attributes as the production index carries them (ForSale / IsAdult, :190-191)
    parts["attr"] = np.array([1 | (2 if i % 5 else 0) | (4 if i % 7 == 0 else 0) for i in range(len(parts["item_ids"]))],
                             dtype=np.uint8)
    return oix, parts


def _check_loaded(gix, oix, parts):
    ids, po, ps = parts["item_ids"], parts["post_off"], parts["post_sessions"]
    for i, item in enumerate(ids):
        item = int(item)
        assert gix.idf(item) == parts["idf"][i]                       # bit-identical: the stored double, not a recomputation
        np.testing.assert_array_equal(gix.postings(item), ps[po[i]:po[i + 1]])
        a = gix.find_attributes(item)
        assert a == {"is_for_sale": bool(parts["attr"][i] & 2), "is_adult": bool(parts["attr"][i] & 4)}
    off = parts["off"]
    for s in range(0, len(parts["ts"]), 7):
        np.testing.assert_array_equal(gix.items_for_session(s), parts["items"][off[s]:off[s + 1]])
        assert gix.session_timestamp(s) == parts["ts"][s]
    assert gix.find_attributes(12345678901234) is None
    with pytest.raises(KeyError):
        gix.idf(12345678901234)


@pytest.mark.parametrize("style,codec", [("plain", "null"), ("plain", "deflate"), ("spark", "snappy"), ("spark", "deflate"),
                                         ("plain", "snappy")])
def test_avro_round_trip(sb, oracle, tmp_path, style, codec):
    rng = np.random.default_rng(5)
    oix, parts = _parts(oracle, rng)
    au.write_index_dir(str(tmp_path), parts, style=style, codec=codec, files=3, records_per_block=17)
    gix = sb.VMISIndex.new(str(tmp_path), device=sb.DEVICE_NONE)
    _check_loaded(gix, oix, parts)
    info = gix.prebuilt_info()
    assert info["prebuilt"] == 1 and info["item_files"] == 3 and info["session_files"] == 3
    assert info["item_records"] == len(parts["item_ids"])
    assert info["lists_reordered"] == 0 and info["duplicate_postings"] == 0
    st = gix.stats()
    assert st["n_items"] == len(parts["item_ids"]) and st["n_postings"] == len(parts["post_sessions"])
    assert st["m_build"] == 12
    # lists are the 12 most recent sessions of their item: the first-match position can be carried for m <= 12
    assert info["m_carry"] == 12


def test_from_parts_matches_avro(sb, oracle, tmp_path):
    rng = np.random.default_rng(6)
    oix, parts = _parts(oracle, rng)
    gix = sb.VMISIndex.from_parts(parts["item_ids"], parts["post_off"], parts["post_sessions"], parts["idf"], parts["attr"],
                                  parts["items"], parts["off"], parts["ts"], device=sb.DEVICE_NONE)
    _check_loaded(gix, oix, parts)
    assert gix.prebuilt_info()["m_carry"] == 12


def test_avro_last_record_wins(sb, oracle, tmp_path):
    rng = np.random.default_rng(7)
    oix, parts = _parts(oracle, rng)
    au.write_index_dir(str(tmp_path), parts, files=1)
    # a later file repeats item 0 with another idf and flags, and session 3 with another timestamp (HashMap::insert
    # at :222-226 and the Vec store at :289-290 keep the last one)
    ids, po, ps = parts["item_ids"], parts["post_off"], parts["post_sessions"]
    rec = au.encode_item("plain", ids[0], ps[po[0]:po[1]], 9.5, False, True)
    au.write_container(os.path.join(str(tmp_path), "itemindex", "zz-late.avro"), au.item_schema("plain"), [rec])
    assert po[1] - po[0] >= 2
    s = int(ps[int(po[1]) - 1])                                       # the OLDEST session of item 0's list
    off = parts["off"]
    srec = au.encode_session("plain", s, parts["items"][off[s]:off[s + 1]], 4_000_000_000)   # `Time as u32` of a negative i32
    au.write_container(os.path.join(str(tmp_path), "sessionindex", "zz-late.avro"), au.session_schema("plain"), [srec])
    gix = sb.VMISIndex.new(str(tmp_path), device=sb.DEVICE_NONE)
    assert gix.idf(int(ids[0])) == 9.5
    assert gix.find_attributes(int(ids[0])) == {"is_for_sale": False, "is_adult": True}
    assert gix.session_timestamp(s) == 4_000_000_000
    assert gix.prebuilt_info()["item_records"] == len(ids) + 1
    # session s is now the most recent one: it moved to the front of every list that holds it → lists re-ordered
    assert gix.prebuilt_info()["lists_reordered"] >= 1
    assert int(gix.postings(int(ids[0]))[0]) == s


def test_avro_normalises_list_order_and_reports_m_carry(sb, oracle):
    rng = np.random.default_rng(8)
    oix, parts = _parts(oracle, rng)
    po, ps = parts["post_off"], parts["post_sessions"].copy()
    # list of item 0 reversed (time ascending) and with a repeated entry
    a, b = int(po[0]), int(po[1])
    assert b - a >= 3
    rev = ps[a:b][::-1].copy()
    ps[a:b] = rev
    gix = sb.VMISIndex.from_parts(parts["item_ids"], po, ps, parts["idf"], parts["attr"], parts["items"], parts["off"],
                                  parts["ts"], device=sb.DEVICE_NONE)
    np.testing.assert_array_equal(gix.postings(int(parts["item_ids"][0])), parts["post_sessions"][a:b])
    assert gix.prebuilt_info()["lists_reordered"] == 1
    # drop the MOST RECENT session from a truncated list: it is no longer a most-recent prefix → m_carry 0
    full = [i for i in range(len(parts["item_ids"])) if po[i + 1] - po[i] == 12]
    i = full[0]
    keep = np.ones(len(ps), dtype=bool)
    keep[int(po[i])] = False
    po2 = po.copy()
    po2[i + 1:] -= 1
    gix2 = sb.VMISIndex.from_parts(parts["item_ids"], po2, parts["post_sessions"][keep], parts["idf"], parts["attr"],
                                   parts["items"], parts["off"], parts["ts"], device=sb.DEVICE_NONE)
    assert gix2.prebuilt_info()["m_carry"] == 0


def _expect_error(sb, path, needle):
    with pytest.raises(sb.VmisError) as e:
        sb.VMISIndex.new(path, device=sb.DEVICE_NONE)
    assert needle in str(e.value), str(e.value)


def test_avro_errors(sb, oracle, tmp_path):
    rng = np.random.default_rng(9)
    oix, parts = _parts(oracle, rng, n_sessions=80, n_items=30)
    _expect_error(sb, str(tmp_path / "missing"), "cannot read directory")
    base = tmp_path / "a"
    au.write_index_dir(str(base), parts, files=1)
    item_file = str(base / "itemindex" / "part-00000.avro")
    good = open(item_file, "rb").read()
    # not a container
    open(item_file, "wb").write(b"PAR1" + good[4:])
    _expect_error(sb, str(base), "not an Avro object container")
    # truncated file
    open(item_file, "wb").write(good[:len(good) // 2])
    _expect_error(sb, str(base), "part-00000.avro")
    # damaged sync marker of the last block
    open(item_file, "wb").write(good[:-1] + bytes([good[-1] ^ 0xFF]))
    _expect_error(sb, str(base), "sync marker")
    # unsupported codec
    ids, po, ps = parts["item_ids"], parts["post_off"], parts["post_sessions"]
    recs = [au.encode_item("plain", ids[i], ps[po[i]:po[i + 1]], 1.0, True, False) for i in range(len(ids))]
    au.write_container(item_file, au.item_schema("plain"), recs, codec="zstandard")
    _expect_error(sb, str(base), "unsupported avro.codec")
    # a column is missing
    schema = au.item_schema("plain")
    schema["fields"] = [f for f in schema["fields"] if f["name"] != "IsAdult"]
    au.write_container(item_file, schema, [r[:-1] for r in recs])
    _expect_error(sb, str(base), "lacks one of")
    # null in a required column
    schema = au.item_schema("plain")
    schema["fields"][2]["type"] = ["null", "double"]
    au.write_container(item_file, schema, [au.zz(s) + au._array([], None) + au.zz(0) + b"\x01\x00" for s in range(3)])
    _expect_error(sb, str(base), "null where a double is required")
    # corrupt snappy payload (CRC)
    au.write_container(item_file, au.item_schema("plain"), recs, codec="snappy")
    blob = bytearray(open(item_file, "rb").read())
    blob[-21] ^= 0x01                                                 # last CRC byte of the last block
    open(item_file, "wb").write(bytes(blob))
    _expect_error(sb, str(base), "CRC")


def test_avro_semantic_checks(sb, oracle):
    rng = np.random.default_rng(10)
    oix, parts = _parts(oracle, rng, n_sessions=80, n_items=30)
    args = lambda p: (p["item_ids"], p["post_off"], p["post_sessions"], p["idf"], p["attr"], p["items"], p["off"], p["ts"])
    # a posting names a session index beyond the session table (the reference panics at mod.rs:138)
    bad = dict(parts)
    bad["post_sessions"] = parts["post_sessions"].copy()
    bad["post_sessions"][0] = len(parts["ts"]) + 5
    with pytest.raises(sb.VmisError, match="no sessionindex record"):
        sb.VMISIndex.from_parts(*args(bad), device=sb.DEVICE_NONE)
    # a posting names a session that does not hold the item
    bad = dict(parts)
    ps = parts["post_sessions"].copy()
    item0 = int(parts["item_ids"][0])
    off = parts["off"]
    other = next(s for s in map(int, ps) if item0 not in parts["items"][off[s]:off[s + 1]])
    ps[0] = other
    bad["post_sessions"] = ps
    with pytest.raises(sb.VmisError, match="does not contain the item"):
        sb.VMISIndex.from_parts(*args(bad), device=sb.DEVICE_NONE)
    # a referenced session holds an item without an itemindex record (the reference panics at vmis_index.rs:322)
    bad = dict(parts)
    victim = len(parts["item_ids"]) - 1
    sel = np.arange(len(parts["item_ids"])) != victim
    po = parts["post_off"]
    keep = np.ones(len(parts["post_sessions"]), dtype=bool)
    keep[int(po[victim]):int(po[victim + 1])] = False
    bad["item_ids"], bad["idf"], bad["attr"] = parts["item_ids"][sel], parts["idf"][sel], parts["attr"][sel]
    bad["post_sessions"] = parts["post_sessions"][keep]
    lens = np.diff(po.astype(np.int64))[sel]
    bad["post_off"] = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
    with pytest.raises(sb.VmisError, match="no itemindex record"):
        sb.VMISIndex.from_parts(*args(bad), device=sb.DEVICE_NONE)


def test_snappy_decoder_copy_elements(sb, tmp_path):
    """hand-made snappy stream with every element type: literals and 1/2/4-byte-offset copies"""
    import json
    import struct
    import zlib
    sess = [[1000, 2000 + s % 2] for s in range(40)]
    recs = [au.encode_session("plain", s, sess[s], 100 + s) for s in range(40)]
    raw = b"".join(recs)
    n = len(raw)
    out = bytearray()
    v = n
    while True:                                                     # varint of the uncompressed length
        b = v & 0x7F
        v >>= 7
        out.append(b | 0x80 if v else b)
        if not v:
            break
    i, n_copies, kinds = 0, 0, set()
    while i < n:
        best = None
        for ln in (11, 8, 4):
            if i + ln <= n:
                j = raw.rfind(raw[i:i + ln], 0, i + ln - 1)
                if 0 <= j < i:
                    best = (ln, i - j)
                    break
        if best:
            ln, offs = best
            kind = n_copies % 3
            n_copies += 1
            if kind == 0 and offs < 2048:
                out += bytes([1 | ((ln - 4) << 2) | ((offs >> 8) << 5), offs & 0xFF]); kinds.add("c1")
            elif kind == 1:
                out += bytes([2 | ((ln - 1) << 2)]) + struct.pack("<H", offs); kinds.add("c2")
            else:
                out += bytes([3 | ((ln - 1) << 2)]) + struct.pack("<I", offs); kinds.add("c4")
            i += ln
        else:
            ln = min(3, n - i)                                       # short literal: tag = (len - 1) << 2
            out += bytes([(ln - 1) << 2]) + raw[i:i + ln]
            i += ln
    assert {"c1", "c2", "c4"} <= kinds
    body = bytes(out) + struct.pack(">I", zlib.crc32(raw) & 0xFFFFFFFF)
    meta = (au.zz(2) + au._bytes(b"avro.schema") + au._bytes(json.dumps(au.session_schema("plain")).encode()) +
            au._bytes(b"avro.codec") + au._bytes(b"snappy") + au.zz(0))
    blob = b"Obj\x01" + meta + au.SYNC + au.zz(len(recs)) + au.zz(len(body)) + body + au.SYNC
    os.makedirs(tmp_path / "sessionindex")
    os.makedirs(tmp_path / "itemindex")
    open(tmp_path / "sessionindex" / "s.avro", "wb").write(blob)
    ids = [1000, 2000, 2001]
    irecs = [au.encode_item("plain", i, [s for s in range(40) if i in sess[s]][::-1], 1.5, True, False) for i in ids]
    au.write_container(str(tmp_path / "itemindex" / "i.avro"), au.item_schema("plain"), irecs, codec="snappy")
    gix = sb.VMISIndex.new(str(tmp_path), device=sb.DEVICE_NONE)
    for s in range(40):
        np.testing.assert_array_equal(gix.items_for_session(s), sess[s])
        assert gix.session_timestamp(s) == 100 + s
    np.testing.assert_array_equal(gix.postings(2001), list(range(39, 0, -2)))
    assert gix.prebuilt_info()["lists_reordered"] == 0


def test_avro_loader_survives_corruption(sb, oracle, tmp_path):
    """byte flips, truncations and insertions anywhere in a container: the loader either loads or fails with a
    VmisError — it never crashes and never hangs"""
    import random
    rng = np.random.default_rng(12)
    oix, parts = _parts(oracle, rng, n_sessions=120, n_items=30)
    random.seed(3)
    outcomes = {"ok": 0, "err": 0}
    for codec, style in [("null", "spark"), ("deflate", "plain"), ("snappy", "spark")]:
        base = tmp_path / codec
        au.write_index_dir(str(base), parts, style=style, codec=codec, files=1, records_per_block=40)
        for sub in ("itemindex", "sessionindex"):
            f = base / sub / "part-00000.avro"
            good = f.read_bytes()
            for it in range(40):
                b = bytearray(good)
                if it % 3 == 0:
                    for _ in range(random.randint(1, 4)):
                        b[random.randrange(len(b))] = random.randrange(256)
                elif it % 3 == 1:
                    b = b[:random.randrange(1, len(b))]
                else:
                    pos = random.randrange(len(b))
                    b[pos:pos] = bytes(random.randrange(256) for _ in range(random.randint(1, 9)))
                f.write_bytes(bytes(b))
                try:
                    sb.VMISIndex.new(str(base), device=sb.DEVICE_NONE).close()
                    outcomes["ok"] += 1
                except sb.VmisError:
                    outcomes["err"] += 1
            f.write_bytes(good)
    assert outcomes["err"] > 150 and outcomes["ok"] + outcomes["err"] == 240


@pytest.mark.parametrize("codec", ["null", "deflate"])
def test_avro_export_is_the_reference_layout(sb, oracle, tmp_path, codec):
    """vmis_index_to_avro: the files hold exactly the oracle's index (posting lists as session indices, idf, flags,
    sessions) in the record layouts of vmis_index.rs:184-192 / :249-255 — checked with an independent Python decoder —
    and load back into an identical index."""
    import glob
    rng = np.random.default_rng(21)
    items, off, ts = random_index_data(rng, 300, 50, max_len=8, id_scale=(1 << 40) + 7)
    oix = oracle.OracleIndex.from_sessions(items, off, ts, 12, 6, 2.0)
    hix = sb.VMISIndex.from_sessions(items, off, ts, 12, 6, 2.0, device=sb.DEVICE_NONE)
    ids = [int(i) for i in np.unique(items) if len(oix.postings(int(i)))]
    hix.set_attributes(ids[:5], [sb.vmis.ATTR_EXISTS | sb.vmis.ATTR_ADULT] * 5)       # not for sale, adult
    hix.to_avro(str(tmp_path), codec, 3)
    item_recs, sess_recs = {}, {}
    for f in sorted(glob.glob(str(tmp_path / "itemindex" / "*.avro"))):
        schema, recs = au.read_container_plain(f, "item")
        assert [x["name"] for x in schema["fields"]] == ["ItemId", "session_indices_time_ordered", "idf", "ForSale", "IsAdult"]
        item_recs.update({r[0]: r for r in recs})
    for f in sorted(glob.glob(str(tmp_path / "sessionindex" / "*.avro"))):
        schema, recs = au.read_container_plain(f, "session")
        assert [x["name"] for x in schema["fields"]] == ["SessionIndex", "item_ids_asc", "Time"]
        sess_recs.update({r[0]: r for r in recs})
    assert sorted(item_recs) == ids
    for n, i in enumerate(ids):
        _, sess, idf, for_sale, adult = item_recs[i]
        assert sess == [int(x) for x in oix.postings(i)] and idf == oix.idf(i)
        assert (for_sale, adult) == ((False, True) if n < 5 else (True, False))
    assert len(sess_recs) == len(ts)
    for s in range(len(ts)):
        assert sess_recs[s][1] == [int(x) for x in items[off[s]:off[s + 1]]] and sess_recs[s][2] == int(ts[s])
    back = sb.VMISIndex.new(str(tmp_path), device=sb.DEVICE_NONE)
    for i in ids:
        np.testing.assert_array_equal(back.postings(i), hix.postings(i))
        assert back.idf(i) == hix.idf(i) and back.find_attributes(i) == hix.find_attributes(i)
    assert back.prebuilt_info()["m_carry"] == 12 and back.prebuilt_info()["lists_reordered"] == 0
