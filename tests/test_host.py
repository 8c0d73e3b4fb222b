"""CPU tests of the product's host side: the C-ABI library loads and exports every symbol include/vmis.h
declares, the index builders agree with the oracle, and query entry points fail loudly without a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from util import random_index_data

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_are_exported(sb):
    hdr = open(os.path.join(ROOT, "include", "vmis.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(vmis_[a-z_0-9]+)\s*\(", hdr))
    declared -= {"vmis_index", "vmis_stats", "vmis_query_stats"}
    assert len(declared) >= 19
    lib = C.CDLL(os.path.join(ROOT, "serenade_b200", "libvmis_b200.so"))
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/vmis.h but not exported"
    from serenade_b200.vmis import EXPORTED_SYMBOLS
    assert declared == set(EXPORTED_SYMBOLS)
    assert b"sm_100a" in sb.load_library().vmis_version()


def test_rust_binding_matches_header():
    """bindings/rust/src/lib.rs cannot be compiled here (no Rust toolchain): every function of its extern "C" block
    must be declared in include/vmis.h with the same number of parameters and be exported by the library."""
    hdr = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "vmis.h")).read(), flags=re.S)
    c_decl = {m.group(1): len([a for a in m.group(2).split(",") if a.strip() and a.strip() != "void"])
              for m in re.finditer(r"\b(vmis_[a-z_0-9]+)\s*\(([^)]*)\)\s*;", hdr)}
    rs = open(os.path.join(ROOT, "bindings", "rust", "src", "lib.rs")).read()
    block = rs[rs.index('extern "C" {'):]
    block = block[:block.index("\n}\n")]
    rust_decl = {m.group(1): len([a for a in m.group(2).split(",") if ":" in a])
                 for m in re.finditer(r"pub fn (vmis_[a-z_0-9]+)\(([^)]*)\)", block, flags=re.S)}
    assert len(rust_decl) >= 11
    lib = C.CDLL(os.path.join(ROOT, "serenade_b200", "libvmis_b200.so"))
    for name, n_args in rust_decl.items():
        assert name in c_decl, f"{name} is not declared in include/vmis.h"
        assert c_decl[name] == n_args, f"{name}: {n_args} parameters in Rust, {c_decl[name]} in C"
        assert hasattr(lib, name)
    # the trait method must bridge to the handle's attributes, not to a Rust-side map (similarity_indexed.rs:23)
    assert "vmis_find_attributes(self.h" in rs


def test_index_from_sessions_with_attributes(sb, oracle):
    """vmis_index_from_sessions_attrs: attributes as (external id, flags) pairs; unlisted items keep the CSV default"""
    items, off, ts = random_index_data(np.random.default_rng(3), 200, 40)
    ids = np.unique(items)[:5]
    flags = np.array([sb.vmis.ATTR_FOR_SALE | sb.vmis.ATTR_ADULT, 0, sb.vmis.ATTR_FOR_SALE, sb.vmis.ATTR_ADULT, 0], dtype=np.uint8)
    hix = sb.VMISIndex.from_sessions(items, off, ts, 10, 8, 1.0, device=sb.DEVICE_NONE, attributes=(ids, flags))
    assert hix.find_attributes(int(ids[0])) == {"is_for_sale": True, "is_adult": True}
    assert hix.find_attributes(int(ids[1])) is None
    assert hix.find_attributes(int(ids[2])) == {"is_for_sale": True, "is_adult": False}
    assert hix.find_attributes(int(ids[3])) == {"is_for_sale": False, "is_adult": True}
    other = int(np.unique(items)[7])
    assert hix.find_attributes(other) == {"is_for_sale": True, "is_adult": False}      # vmis_index.rs:514-518


def test_error_code_is_cleared_by_the_next_successful_call(sb):
    """vmis_last_error_code() used to stick after a failure"""
    items, off, ts = random_index_data(np.random.default_rng(1), 50, 10)
    hix = sb.VMISIndex.from_sessions(items, off, ts, 10, 8, 1.0, device=sb.DEVICE_NONE)
    lib = sb.load_library()
    with pytest.raises(KeyError):
        hix.idf(0xDEADBEEFDEADBEEF)
    assert lib.vmis_last_error_code() != 0
    assert hix.idf(int(items[0])) > 0
    assert lib.vmis_last_error_code() == 0


def test_library_has_no_oracle_dependency():
    """the product must not link or load anything under oracle/"""
    import subprocess
    so = os.path.join(ROOT, "serenade_b200", "libvmis_b200.so")
    needed = subprocess.run(["readelf", "-d", so], capture_output=True, text=True).stdout
    assert "oracle" not in needed
    for dirpath, _, files in os.walk(os.path.join(ROOT, "serenade_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r'#include\s*"[^"]*oracle|^\s*(from|import)\s+oracle|libvmis_oracle|dlopen', txt,
                                     flags=re.M), f


def test_queries_fail_loudly_without_gpu(sb, oracle):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    items, off, ts = random_index_data(np.random.default_rng(0), 50, 10)
    with pytest.raises(sb.VmisError) as e:           # device build without a device
        sb.VMISIndex.from_sessions(items, off, ts, 10, 8, 1.0, device=0)
    assert "no CPU fallback" in str(e.value) or "CUDA" in str(e.value)
    hix = sb.VMISIndex.from_sessions(items, off, ts, 10, 8, 1.0, device=sb.DEVICE_NONE)
    with pytest.raises(sb.VmisError) as e:
        sb.predict(hix, [int(items[0])], 5, 10, 5)
    assert e.value.code == -3
    with pytest.raises(sb.VmisError):
        hix.find_neighbors([int(items[0])], 5, 10)


def test_toy_host_index_matches_oracle(sb, oracle, toy_dir):
    train = os.path.join(toy_dir, "train.txt")
    hix = sb.VMISIndex.new_from_csv(train, 1502, 1.0, max_len=0, device=sb.DEVICE_NONE)
    st = hix.stats()
    assert st["max_len"] == 15                       # restated t-digest p99.5 == exact p99.5 on the toy data
    oix = oracle.OracleIndex.new_from_csv(train, 1502, 1.0, 15)
    assert st["n_sessions"] == oix.num_sessions == 23753
    assert st["n_items"] == oix.num_items == 17919
    assert st["n_pairs_kept"] == oix.kept_pairs == 77651
    rng = np.random.default_rng(1)
    for s in rng.integers(0, st["n_sessions"], size=300):
        assert np.array_equal(hix.items_for_session(int(s)), oix.items_for_session(int(s)))
        assert hix.session_timestamp(int(s)) == oix.session_ts(int(s))
    all_items = np.unique(np.concatenate([oix.items_for_session(int(s)) for s in range(0, 23753, 7)]))
    checked = 0
    for it in all_items[::5]:
        try:
            want = oix.idf(int(it))
        except KeyError:
            with pytest.raises(KeyError):
                hix.idf(int(it))
            continue
        assert hix.idf(int(it)) == want              # bit-exact f64
        assert np.array_equal(hix.postings(int(it)), oix.postings(int(it)))
        assert hix.find_attributes(int(it)) == oix.find_attributes(int(it))
        checked += 1
    assert checked > 500


@pytest.mark.parametrize("seed", range(4))
def test_random_host_index_matches_oracle(sb, oracle, seed):
    """truncation to m, timestamp ties (higher session idx first), max_len pruning"""
    rng = np.random.default_rng(seed)
    items, off, ts = random_index_data(rng, 400, 25, max_len=9, unique_ts=bool(seed % 2), id_scale=1 << 33)
    m, max_len = int(rng.integers(1, 30)), int(rng.integers(2, 9))
    hix = sb.VMISIndex.from_sessions(items, off, ts, m, max_len, 1.7, device=sb.DEVICE_NONE)
    oix = oracle.OracleIndex.from_sessions(items, off, ts, m, max_len, 1.7)
    st = hix.stats()
    assert st["n_pairs_kept"] == oix.kept_pairs and st["n_items"] == oix.num_items
    for it in np.unique(items):
        try:
            want = oix.idf(int(it))
        except KeyError:                             # item only in pruned sessions
            with pytest.raises(KeyError):
                hix.idf(int(it))
            assert len(hix.postings(int(it))) == 0
            continue
        assert hix.idf(int(it)) == want
        assert np.array_equal(hix.postings(int(it)), oix.postings(int(it)))


def test_csv_reader_quirks(sb, oracle, tmp_path):
    """read_from_file (vmis_index.rs:591-752): header skipped, stable grouping by session id, first occurrence of a
    duplicate item, clock only moved by non-duplicate rows, items sorted, LAST ROW DROPPED (:666-667,675-686)."""
    p = tmp_path / "t.txt"
    p.write_text("SessionId\tItemId\tTime\n"
                 "7\t30\t100.0\n"
                 "3\t10\t50.4\n"
                 "7\t20\t90.0\n"
                 "7\t30\t500.0\n"      # duplicate item: ignored, clock not moved
                 "3\t11\t60.6\n"
                 "9\t1\t70.0\n"
                 "9\t2\t80.0\n")       # last sorted row: lost
    hix = sb.VMISIndex.new_from_csv(str(p), 10, 1.0, max_len=10, device=sb.DEVICE_NONE)
    oix = oracle.OracleIndex.new_from_csv(str(p), 10, 1.0, 10)
    assert hix.stats()["n_sessions"] == oix.num_sessions == 3
    want = [([10, 11], 61), ([20, 30], 100), ([1], 70)]
    for s, (its, t) in enumerate(want):
        assert list(hix.items_for_session(s)) == its == list(oix.items_for_session(s))
        assert hix.session_timestamp(s) == t == oix.session_ts(s)
    with pytest.raises(sb.VmisError) as e:
        sb.VMISIndex.new_from_csv(str(tmp_path / "missing.txt"), 10, 1.0, device=sb.DEVICE_NONE)
    assert e.value.code == -2


def test_synthetic_generator(sb):
    items, off, ts = sb.synth_sessions(42, 5000, 20000)
    items2, off2, ts2 = sb.synth_sessions(42, 5000, 20000)
    assert np.array_equal(items, items2) and np.array_equal(off, off2) and np.array_equal(ts, ts2)
    lens = np.diff(off.astype(np.int64))
    assert lens.min() >= 1 and lens.max() <= 34 and 4.5 < lens.mean() < 6.0
    assert len(np.unique(ts)) == len(ts)                                  # unique timestamp per session
    for s in range(0, 20000, 97):                                          # distinct, ascending inside a session
        seg = items[off[s]:off[s + 1]]
        assert (np.diff(seg.astype(np.int64)) > 0).all()
    cnt = np.sort(np.unique(items, return_counts=True)[1])[::-1]
    assert cnt[0] > 20 * np.median(cnt)                                    # Zipf head
    q_items, q_off = sb.synth_queries(43, 5000, 1000, 4)
    L = np.diff(q_off.astype(np.int64))
    assert L.min() >= 1 and L.max() <= 4 and q_off[-1] == len(q_items)
    assert np.isin(q_items, np.unique(sb.synth_sessions(42, 5000, 200000)[0])).mean() > 0.9


def test_argument_errors(sb):
    items, off, ts = random_index_data(np.random.default_rng(2), 30, 8)
    with pytest.raises(sb.VmisError):
        sb.VMISIndex.from_sessions(items, off, ts, 0, 8, 1.0, device=sb.DEVICE_NONE)   # m = 0
    dup = np.array([5, 5], dtype=np.uint64)
    with pytest.raises(sb.VmisError):
        sb.VMISIndex.from_sessions(dup, np.array([0, 2], dtype=np.uint64), np.array([1], dtype=np.uint32), 5, 8, 1.0,
                                   device=sb.DEVICE_NONE)
    hix = sb.VMISIndex.from_sessions(items, off, ts, 5, 8, 1.0, device=sb.DEVICE_NONE)
    with pytest.raises(IndexError):
        hix.items_for_session(10 ** 6)
    with pytest.raises(KeyError):
        hix.idf(123456789)
    assert hix.find_attributes(123456789) is None
    hix.set_attributes(np.array([int(items[0])], dtype=np.uint64), np.array([6], dtype=np.uint8))
    assert hix.find_attributes(int(items[0])) == {"is_for_sale": True, "is_adult": True}
    hix.set_attributes(np.array([int(items[0])], dtype=np.uint64), np.array([0], dtype=np.uint8))
    assert hix.find_attributes(int(items[0])) is None


def test_sharded_posting_layout_matches_unsharded(sb):
    """item-sharded postings (config 5): same lists whichever shard an item lands in"""
    items, off, ts = random_index_data(np.random.default_rng(8), 500, 40, max_len=7)
    full = sb.VMISIndex.from_sessions(items, off, ts, 25, 7, 1.0, device=sb.DEVICE_NONE)
    for n_shards in (2, 3, 8):
        sh = sb.VMISIndex.from_sessions_sharded(items, off, ts, 25, 7, 1.0, sb.DEVICE_NONE, 0, n_shards)
        assert sh.stats()["n_postings"] == full.stats()["n_postings"]
        for it in np.unique(items):
            assert np.array_equal(sh.postings(int(it)), full.postings(int(it)))
    with pytest.raises(sb.VmisError):
        sb.VMISIndex.from_sessions_sharded(items, off, ts, 25, 7, 1.0, sb.DEVICE_NONE, 2, 2)
    with pytest.raises(sb.VmisError):
        sb.VMISIndex.from_sessions_sharded(items, off, ts, 25, 7, 1.0, sb.DEVICE_NONE, 0, 9)


def test_cpp_host_mirror(tmp_path):
    """include/vmis.hpp (the C++ mirror of VMISIndex / predict) against a host-only handle"""
    import subprocess
    exe = tmp_path / "host_mirror_test"
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-o", str(exe), os.path.join(ROOT, "tests", "cpp", "host_mirror_test.cpp"),
                           "-L" + os.path.join(ROOT, "serenade_b200"), "-lvmis_b200",
                           "-Wl,-rpath," + os.path.join(ROOT, "serenade_b200")])
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0 and "host mirror ok" in out.stdout, out.stdout + out.stderr


def test_sessions_parsed_once_give_the_same_index(sb, toy_dir):
    """vmis_sessions_from_csv + from_sessions(max_len=0) == new_from_csv (the HPO loop parses the training file once)"""
    train = os.path.join(toy_dir, "train.txt")
    items, off, ts = sb.read_sessions_csv(train)
    a = sb.VMISIndex.new_from_csv(train, 500, 2.0, device=sb.DEVICE_NONE)
    b = sb.VMISIndex.from_sessions(items, off, ts, 500, 0, 2.0, device=sb.DEVICE_NONE)
    assert a.stats() == b.stats() and a.stats()["n_sessions"] == len(ts) == 23753
    for item in (13598, 2835, 10, 12068):
        np.testing.assert_array_equal(a.postings(item), b.postings(item))
        assert a.idf(item) == b.idf(item)
    np.testing.assert_array_equal(a.items_for_session(5189), items[off[5189]:off[5189 + 1]])
    with pytest.raises(sb.VmisError):
        sb.read_sessions_csv("/nonexistent/train.txt")


def test_parallel_tsv_reader_matches_the_serial_restatement(sb, oracle, tmp_path):
    """a few MB of rows (several parser chunks): sessions interleaved in the file, duplicate rows, float timestamps,
    blank / malformed lines — session for session equal to the oracle's line-by-line read_from_file"""
    rng = np.random.default_rng(4)
    n = 260_000
    sid = rng.integers(0, 40_000, size=n)
    item = rng.integers(1, 3_000, size=n) * 1_000_003
    tm = rng.integers(1_500_000_000, 1_600_000_000, size=n)
    path = str(tmp_path / "train.tsv")
    with open(path, "w") as f:
        f.write("SessionId\tItemId\tTime\n")
        for i in range(n):
            if i % 50_000 == 7:
                f.write("\n")                                    # blank line
            if i % 70_000 == 11:
                f.write("not\ta\trow\n")                         # unparsable: skipped (vmis_index.rs:610-616 prints and goes on)
            t = f"{tm[i]}.0" if i % 3 == 0 else (f"{tm[i]}.4" if i % 3 == 1 else str(tm[i]))
            f.write(f"{sid[i]}\t{item[i]}\t{t}\n")
    assert os.path.getsize(path) > 4 << 20
    items, off, ts = sb.read_sessions_csv(path)
    oix = oracle.OracleIndex.new_from_csv(path, 100, 1.0, 50)
    assert len(ts) == oix.num_sessions
    for s in range(0, len(ts), 1):
        if s % 97 and s < len(ts) - 3:
            continue                                             # every 97th session and the last three (last-row quirk)
        np.testing.assert_array_equal(items[off[s]:off[s + 1]], oix.items_for_session(s))
        assert ts[s] == oix.session_ts(s)
    assert int(off[-1]) == sum(len(oix.items_for_session(s)) for s in range(len(ts)))


def test_tsv_parser_is_as_strict_as_csv_serde(sb, oracle, tmp_path):
    """rows the reference's csv + serde (usize, usize, f64) parse rejects are skipped here too — signs,
    blanks, a 4th column, overflow — and the float → usize cast saturates like Rust's `as` (vmis_index.rs:604-609)."""
    good = ["1\t10\t100.0", "1\t11\t101.4", "2\t10\t200", "2\t12\t201", "3\t13\t300", "3\t14\t3e2", "4\t15\t400", "4\t16\t401"]
    bad = ["-5\t10\t100", " 6\t10\t100", "7\t-1\t100", "8\t10\t100\t9", "9\t10\tabc", "99999999999999999999999\t1\t1",
           "10\t10\t 5", "11\t10", "12\t10\t0x10"]
    p = tmp_path / "t.txt"
    rows = []
    for i, g in enumerate(good):
        rows.append(g)
        if i < len(bad):
            rows.append(bad[i])
    rows += bad[len(good):]
    p.write_text("SessionId\tItemId\tTime\n" + "\n".join(rows) + "\n")
    items, off, ts = sb.read_sessions_csv(str(p))
    q = tmp_path / "clean.txt"
    q.write_text("SessionId\tItemId\tTime\n" + "\n".join(good) + "\n")
    items2, off2, ts2 = sb.read_sessions_csv(str(q))
    assert np.array_equal(items, items2) and np.array_equal(off, off2) and np.array_equal(ts, ts2)
    assert len(ts) == 4 and list(ts[:3]) == [101, 201, 300]          # "+"-less plain rows; 101.4 rounds to 101
    # saturating casts: negative and NaN times become 0
    r = tmp_path / "neg.txt"
    r.write_text("SessionId\tItemId\tTime\n1\t10\t-7.5\n1\t11\tNaN\n2\t12\t5\n2\t13\t6\n")
    _, _, ts3 = sb.read_sessions_csv(str(r))
    assert list(ts3) == [0, 5]           # session 2 keeps item 12 only: the last sorted row is dropped (:666-686)
    # a file whose only data row is swallowed by the last-row quirk is an error, not an empty index
    one = tmp_path / "one.txt"
    one.write_text("SessionId\tItemId\tTime\n1\t10\t5\n")
    with pytest.raises(sb.VmisError) as e:
        sb.read_sessions_csv(str(one))
    assert e.value.code == -2


def test_server_refuses_windows_longer_than_the_kernel_limit(sb):
    items, off, ts = random_index_data(np.random.default_rng(5), 50, 10)
    hix = sb.VMISIndex.from_sessions(items, off, ts, 10, 8, 1.0, device=sb.DEVICE_NONE)
    with pytest.raises(sb.VmisError):
        sb.Server(hix, 5, 10, 5, max_items_in_session=129)
