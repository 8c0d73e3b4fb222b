// serenade_b200/csrc/index_host.cpp — host-side construction of the VMIS index:
// TSV reader (read_from_file, vmis_index.rs:591-752), session-length p99.5
// (vmis_index.rs:693-716) and the CSR builder that replaces prepare_hashmap
// (vmis_index.rs:422-528) with the flat, time-ranked layout of vmis_device.h.
#include "vmis_host.h"
#include "avro_reader.h"

#include <algorithm>
#include <cerrno>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <numeric>
#include <thread>

namespace vmis {
namespace {

inline uint64_t mix64(uint64_t x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
  return x;
}

// growable u64 -> u32 map used while building the item dictionary
struct IdMap {
  std::vector<uint64_t> keys; std::vector<uint32_t> vals; size_t n = 0, mask = 0;
  explicit IdMap(size_t cap) { size_t c = 1024; while (c < cap * 2) c <<= 1; keys.assign(c, 0); vals.assign(c, kEmpty); mask = c - 1; }
  void grow() {
    std::vector<uint64_t> ok; std::vector<uint32_t> ov; ok.swap(keys); ov.swap(vals);
    size_t c = (mask + 1) * 2; keys.assign(c, 0); vals.assign(c, kEmpty); mask = c - 1;
    for (size_t i = 0; i < ov.size(); ++i) if (ov[i] != kEmpty) {
      size_t h = mix64(ok[i]) & mask; while (vals[h] != kEmpty) h = (h + 1) & mask; keys[h] = ok[i]; vals[h] = ov[i];
    }
  }
  uint32_t get_or_add(uint64_t k) {
    if ((n + 1) * 2 > mask + 1) grow();
    size_t h = mix64(k) & mask;
    while (vals[h] != kEmpty) { if (keys[h] == k) return vals[h]; h = (h + 1) & mask; }
    keys[h] = k; vals[h] = (uint32_t)n; return (uint32_t)n++;
  }
};

// ---- t-digest (merging digest, max_size centroids), the published algorithm of the `tdigest`
// crate the reference calls at vmis_index.rs:693-716 (TDigest::new_with_size(100).merge_unsorted(v)
// then estimate_quantile(q)).  The crate is not vendored in the reference tree.
struct Centroid { double mean, weight; };
inline double k_to_q(double k, double d) {
  const double r = k / d;
  if (r >= 0.5) { const double b = 1.0 - r; return 1.0 - 2.0 * b * b; }
  return 2.0 * r * r;
}
struct TDigest {
  std::vector<Centroid> c; double count = 0, mn = 0, mx = 0;
  void merge_sorted(const std::vector<double>& v, size_t max_size) {
    if (v.empty()) return;
    count = (double)v.size(); mn = v.front(); mx = v.back();
    double k_limit = 1.0;
    double q_limit_times_count = k_to_q(k_limit, (double)max_size) * count; k_limit += 1.0;
    Centroid curr{v[0], 1.0};
    double weight_so_far = 1.0, sums = 0.0, weights = 0.0;
    auto flush = [&](Centroid& x) { const double ns = sums + x.weight * x.mean; const double nw = x.weight + weights;
                                    x.weight = nw; x.mean = ns / nw; sums = 0.0; weights = 0.0; };
    for (size_t i = 1; i < v.size(); ++i) {
      weight_so_far += 1.0;
      if (weight_so_far <= q_limit_times_count) { sums += v[i]; weights += 1.0; }
      else {
        flush(curr); c.push_back(curr);
        q_limit_times_count = k_to_q(k_limit, (double)max_size) * count; k_limit += 1.0;
        curr = Centroid{v[i], 1.0};
      }
    }
    flush(curr); c.push_back(curr);
    std::stable_sort(c.begin(), c.end(), [](const Centroid& a, const Centroid& b) { return a.mean < b.mean; });
  }
  double estimate_quantile(double q) const {
    if (c.empty()) return 0.0;
    const double rank = q * count; size_t pos; double t;
    if (q > 0.5) {
      if (q >= 1.0) return mx;
      pos = 0; t = count;
      for (size_t k = c.size(); k-- > 0;) { t -= c[k].weight; if (rank >= t) { pos = k; break; } }
    } else {
      if (q <= 0.0) return mn;
      pos = c.size() - 1; t = 0.0;
      for (size_t k = 0; k < c.size(); ++k) { if (rank < t + c[k].weight) { pos = k; break; } t += c[k].weight; }
    }
    double delta = 0.0, lo = mn, hi = mx;
    if (c.size() > 1) {
      if (pos == 0) { delta = c[1].mean - c[0].mean; hi = c[1].mean; }
      else if (pos == c.size() - 1) { delta = c[pos].mean - c[pos - 1].mean; lo = c[pos - 1].mean; }
      else { delta = (c[pos + 1].mean - c[pos - 1].mean) / 2.0; lo = c[pos - 1].mean; hi = c[pos + 1].mean; }
    }
    const double value = c[pos].mean + ((rank - t) / c[pos].weight - 0.5) * delta;
    return std::min(std::max(value, lo), hi);
  }
};

}  // namespace

size_t session_length_p99_5(const Sessions& s) {
  std::vector<double> lens(s.size());
  for (size_t i = 0; i < s.size(); ++i) lens[i] = (double)(s.off[i + 1] - s.off[i]);
  std::sort(lens.begin(), lens.end());
  TDigest d; d.merge_sorted(lens, 100);
  return (size_t)std::llround(d.estimate_quantile(0.995));
}

bool read_sessions_from_csv(const std::string& path, Sessions* out, std::string* err) {
  FILE* f = std::fopen(path.c_str(), "rb");
  if (!f) { *err = "cannot open " + path; return false; }
  // The whole file in memory, every line turned into a C string; the lines are parsed on all host cores
  // (chunks in file order, so the order of the rows is the file's) with the same strtoull / strtod calls a
  // line-by-line reader would make.
  std::vector<char> buf;
  {
    std::fseek(f, 0, SEEK_END);
    const long sz = std::ftell(f);
    std::fseek(f, 0, SEEK_SET);
    if (sz < 0) { std::fclose(f); *err = "cannot read " + path; return false; }
    buf.resize((size_t)sz + 1);
    const size_t got = std::fread(buf.data(), 1, (size_t)sz, f);
    std::fclose(f);
    buf.resize(got + 1);
    buf[got] = 0;
  }
  const size_t size = buf.size() - 1;
  struct Row { uint64_t session, item, time; };
  const size_t nt = std::max<size_t>(1, std::min<size_t>(std::thread::hardware_concurrency(), size / (1 << 20) + 1));
  std::vector<std::vector<Row>> parts(nt);
  std::vector<size_t> bad_rows(nt, 0);
  {
    std::vector<std::thread> th;
    auto terminate_lines = [&](size_t t) {
      for (size_t i = size * t / nt, e = size * (t + 1) / nt; i < e; ++i) if (buf[i] == '\n') buf[i] = 0;
    };
    for (size_t t = 1; t < nt; ++t) th.emplace_back(terminate_lines, t);
    terminate_lines(0);
    for (auto& x : th) x.join();
    th.clear();
    auto parse = [&](size_t t) {
      size_t pos = size * t / nt;
      const size_t end = size * (t + 1) / nt;
      if (t == 0) { while (pos < size && buf[pos] != 0) ++pos; ++pos; }           // header row (:597)
      else if (pos > 0 && buf[pos - 1] != 0) { while (pos < size && buf[pos] != 0) ++pos; ++pos; }   // inside a line: it belongs to the previous chunk
      std::vector<Row>& rows = parts[t];
      rows.reserve((end - std::min(pos, end)) / 24 + 16);
      while (pos < end) {                                                          // lines STARTING in [begin, end)
        const char* p = buf.data() + pos;
        const size_t len = std::strlen(p);
        pos += len + 1;
        if (len == 0 || *p == '\r') continue;
        // (usize, usize, f64) through csv + serde (:604-609): exactly three fields; the integers are digits with an
        // optional '+' (no sign, no blanks, no overflow); the float is what Rust's f64::from_str takes.  Anything else
        // is the reference's "Unable to parse input!" → the row is skipped.
        auto uint_token = [](const char* q) { return (*q >= '0' && *q <= '9') || (*q == '+' && q[1] >= '0' && q[1] <= '9'); };
        char* e = nullptr;
        if (!uint_token(p)) { ++bad_rows[t]; continue; }
        errno = 0;
        const unsigned long long sid = std::strtoull(p, &e, 10);
        if (errno == ERANGE || *e != '\t') { ++bad_rows[t]; continue; }
        p = e + 1;
        if (!uint_token(p)) { ++bad_rows[t]; continue; }
        errno = 0;
        const unsigned long long iid = std::strtoull(p, &e, 10);
        if (errno == ERANGE || *e != '\t') { ++bad_rows[t]; continue; }
        p = e + 1;
        if (*p == ' ' || *p == '\t' || *p == '\0' || *p == '\r' || (p[0] == '0' && (p[1] == 'x' || p[1] == 'X'))) { ++bad_rows[t]; continue; }
        const double tm = std::strtod(p, &e);
        if (e == p || !(*e == '\0' || (*e == '\r' && e[1] == '\0'))) { ++bad_rows[t]; continue; }   // a 4th field or trailing junk
        // `raw.2.round() as usize` (:607-609): Rust's float → integer cast saturates (NaN and negatives give 0)
        const double r = std::round(tm);
        const uint64_t tsec = !(r > 0.0) ? 0ull : r >= 18446744073709551615.0 ? ~0ull : (uint64_t)r;
        rows.push_back(Row{sid, iid, tsec});
      }
    };
    for (size_t t = 1; t < nt; ++t) th.emplace_back(parse, t);
    parse(0);
    for (auto& x : th) x.join();
  }
  size_t bad = 0, total = 0;
  for (size_t t = 0; t < nt; ++t) { bad += bad_rows[t]; total += parts[t].size(); }
  if (bad) std::fprintf(stderr, "vmis: %zu unparsable rows skipped in %s\n", bad, path.c_str());
  std::vector<Row> rows(total);
  std::vector<size_t> begin(nt + 1, 0);
  for (size_t t = 0; t < nt; ++t) begin[t + 1] = begin[t] + parts[t].size();
  {
    // rows grouped by session id, file order kept inside a session (stable, :620-633): every chunk is copied into
    // place and stable-sorted on its own thread, then neighbouring runs are merged (std::inplace_merge is stable)
    std::vector<std::thread> th;
    auto by_session = [](const Row& a, const Row& b) { return a.session < b.session; };
    auto sort_part = [&](size_t t) {
      std::copy(parts[t].begin(), parts[t].end(), rows.begin() + begin[t]);
      std::vector<Row>().swap(parts[t]);
      std::stable_sort(rows.begin() + begin[t], rows.begin() + begin[t + 1], by_session);
    };
    for (size_t t = 1; t < nt; ++t) th.emplace_back(sort_part, t);
    sort_part(0);
    for (auto& x : th) x.join();
    for (size_t width = 1; width < nt; width *= 2) {
      th.clear();
      for (size_t lo = 0; lo + width < nt; lo += 2 * width) {
        const size_t mid = lo + width, hi = std::min(nt, lo + 2 * width);
        th.emplace_back([&, lo, mid, hi]() {
          std::inplace_merge(rows.begin() + begin[lo], rows.begin() + begin[mid], rows.begin() + begin[hi], by_session);
        });
      }
      for (auto& x : th) x.join();
    }
  }
  if (rows.empty()) { *err = "no rows in " + path; return false; }
  const size_t n = rows.size();
  out->items.clear(); out->off.assign(1, 0); out->ts.clear();
  std::vector<uint64_t> cur; cur.push_back(rows[0].item);
  uint64_t max_ts = rows[0].time;
  auto close_session = [&]() {
    std::sort(cur.begin(), cur.end());                               // items ascending (:676-678)
    out->items.insert(out->items.end(), cur.begin(), cur.end());
    out->off.push_back(out->items.size());
    out->ts.push_back((uint32_t)max_ts);
  };
  for (size_t i = 1; i < n; ++i) {
    // The reference closes the running session at a session change AND at the very last row, whose
    // own item is then dropped (:666-667, :675-686).  Reproduced on purpose.
    const bool same = rows[i].session == rows[i - 1].session && i != n - 1;
    if (same) {
      if (std::find(cur.begin(), cur.end(), rows[i].item) == cur.end()) {   // first occurrence only (:668)
        cur.push_back(rows[i].item);
        if (rows[i].time > max_ts) max_ts = rows[i].time;             // only non-duplicate rows move the clock (:671)
      }
    } else {
      close_session();
      cur.clear(); cur.push_back(rows[i].item); max_ts = rows[i].time;
    }
  }
  if (out->ts.empty()) {
    // e.g. a single data row: the last-row handling drops it and nothing is left.  The reference goes on to panic
    // (empty percentile digest, vmis_index.rs:689-716); an empty index is never what the caller meant.
    *err = "no training session survives read_from_file in " + path + " (the last sorted row is always dropped, vmis_index.rs:666-686)";
    return false;
  }
  return true;
}

uint32_t host_lookup_item(const FlatIndex& f, uint64_t item) {
  if (f.item_hash.empty()) {
    // device-built handle: the hash table lives in HBM only; the dictionary is sorted, and the accessors that come
    // here (idf, find_attributes, postings) are not on any hot path
    const auto it = std::lower_bound(f.item_key.begin(), f.item_key.end(), item);
    return it != f.item_key.end() && *it == item ? (uint32_t)(it - f.item_key.begin()) : kEmpty;
  }
  const uint32_t mask = (uint32_t)f.item_hash.size() - 1;
  uint32_t h = (uint32_t)mix64(item) & mask;
  for (;;) {
    const ItemHashEntry& e = f.item_hash[h];
    if (e.val == kEmpty) return kEmpty;
    if (e.key == item) return e.val;
    h = (h + 1) & mask;
  }
}

bool build_flat_index(const Sessions& s, size_t m, size_t max_len, double idf_weighting, uint32_t n_shards,
                      FlatIndex* out, std::string* err) {
  const size_t S = s.size();
  if (S >= 0xFFFFFFFFull) { *err = "too many sessions"; return false; }
  if (m == 0) { *err = "m must be >= 1"; return false; }
  if (n_shards == 0 || n_shards > (uint32_t)kMaxShards) { *err = "n_shards must be in [1, 8]"; return false; }
  FlatIndex& F = *out;
  F = FlatIndex();
  F.m_build = (uint32_t)std::min<size_t>(m, 0xFFFFFFFFu);
  F.m_carry = F.m_build;
  F.max_len = (uint32_t)std::min<size_t>(max_len, 0xFFFFFFFFu);
  F.idf_weighting = idf_weighting;
  F.n_shards = n_shards;

  // 1. kept sessions (len <= max_len, :452) ranked by (timestamp, session idx) ascending.  The rank
  //    replaces the timestamp: "more recent" == "larger rank", ties broken like the stable sort +
  //    reverse of :497-503 (higher session idx first).
  std::vector<uint64_t> order; order.reserve(S);
  uint64_t P = 0;
  for (size_t i = 0; i < S; ++i) {
    const uint64_t len = s.off[i + 1] - s.off[i];
    if (len <= max_len && len > 0) { order.push_back(((uint64_t)s.ts[i] << 32) | (uint64_t)i); P += len; }
  }
  std::sort(order.begin(), order.end());
  const size_t Sk = order.size();
  F.rank_to_orig.resize(Sk);
  for (size_t r = 0; r < Sk; ++r) F.rank_to_orig[r] = (uint32_t)(order[r] & 0xFFFFFFFFull);
  F.n_pairs_kept = P;

  // 2. item dictionary: temp ids in first-seen order, document frequencies
  IdMap ids(1 << 16);
  std::vector<uint32_t> entry_tid(P);
  std::vector<uint32_t> df;
  {
    size_t e = 0;
    for (size_t r = 0; r < Sk; ++r) {
      const uint32_t o = F.rank_to_orig[r];
      for (uint64_t p = s.off[o]; p < s.off[o + 1]; ++p) {
        const uint32_t t = ids.get_or_add(s.items[p]);
        if (t == df.size()) df.push_back(0);
        ++df[t]; entry_tid[e++] = t;
      }
    }
  }
  const size_t I = df.size();
  std::vector<uint64_t> tkey(I);
  for (size_t h = 0; h <= ids.mask; ++h) if (ids.vals[h] != kEmpty) tkey[ids.vals[h]] = ids.keys[h];
  std::vector<uint32_t> by_key(I); std::iota(by_key.begin(), by_key.end(), 0u);
  std::sort(by_key.begin(), by_key.end(), [&](uint32_t a, uint32_t b) { return tkey[a] < tkey[b]; });
  std::vector<uint32_t> dense_of_temp(I);
  F.item_key.resize(I);
  for (size_t d = 0; d < I; ++d) { dense_of_temp[by_key[d]] = (uint32_t)d; F.item_key[d] = tkey[by_key[d]]; }

  // 3. session -> items (dense, ascending), 16-byte aligned starts
  F.sess_ref.resize(Sk);
  F.sess_items.clear(); F.sess_items.reserve(P + P / 3 + 16);
  {
    size_t e = 0; std::vector<uint32_t> tmp;
    for (size_t r = 0; r < Sk; ++r) {
      const uint32_t o = F.rank_to_orig[r];
      const uint32_t len = (uint32_t)(s.off[o + 1] - s.off[o]);
      tmp.resize(len);
      for (uint32_t t = 0; t < len; ++t) tmp[t] = dense_of_temp[entry_tid[e + t]];
      std::sort(tmp.begin(), tmp.end());
      for (uint32_t t = 1; t < len; ++t) if (tmp[t] == tmp[t - 1]) { *err = "duplicate item inside a training session"; return false; }
      const size_t start = F.sess_items.size();
      if ((start >> 2) > 0xFFFFFFFFull) { *err = "session item array too large"; return false; }
      F.sess_ref[r] = make_uint2((uint32_t)(start >> 2), len);
      F.sess_items.insert(F.sess_items.end(), tmp.begin(), tmp.end());
      while (F.sess_items.size() & 3) F.sess_items.push_back(kEmpty);
      e += len;
    }
  }
  std::vector<uint32_t>().swap(entry_tid);

  // 4. postings: per item the min(df, m) most recent kept sessions, rank descending (:497-504)
  F.post_ref.resize(I);
  {
    // item d belongs to shard d % n_shards; offsets are relative to the shard's own array
    std::vector<uint64_t> shard_size(n_shards, 0);
    for (size_t d = 0; d < I; ++d) {
      const uint32_t len = (uint32_t)std::min<uint64_t>(df[by_key[d]], m);
      uint64_t& sz = shard_size[d % n_shards];
      if ((sz >> 2) > 0xFFFFFFFFull) { *err = "posting array too large"; return false; }
      F.post_ref[d] = make_uint2((uint32_t)(sz >> 2), len);
      F.n_postings += len;
      sz += (len + 3u) & ~3u;
    }
    F.shard_begin.assign(n_shards + 1, 0);
    for (uint32_t sh = 0; sh < n_shards; ++sh) F.shard_begin[sh + 1] = F.shard_begin[sh] + shard_size[sh];
    const uint64_t pos = F.shard_begin[n_shards];
    F.postings.assign(pos, kEmpty);
    std::vector<uint32_t> fill(I, 0);
    for (size_t r = Sk; r-- > 0;) {
      const uint2 ref = F.sess_ref[r];
      const uint32_t* it = &F.sess_items[(size_t)ref.x * 4];
      for (uint32_t t = 0; t < ref.y; ++t) {
        const uint32_t d = it[t];
        if (fill[d] < F.post_ref[d].y)
          F.postings[F.shard_begin[d % n_shards] + (size_t)F.post_ref[d].x * 4 + fill[d]++] = (uint32_t)r;
      }
    }
  }

  // 5. idf = ln(P_kept / df) * idf_weighting (:509-513); default attributes (:514-518)
  F.idf.resize(I); F.attr.assign(I, (uint8_t)(VMIS_ATTR_EXISTS | VMIS_ATTR_FOR_SALE));
  for (size_t d = 0; d < I; ++d) F.idf[d] = std::log((double)P / (double)df[by_key[d]]) * idf_weighting;

  // 6. device item hash
  size_t cap = 16; while (cap < I * 2) cap <<= 1;
  F.item_hash.assign(cap, ItemHashEntry{0, kEmpty, 0});
  for (size_t d = 0; d < I; ++d) {
    uint32_t h = (uint32_t)mix64(F.item_key[d]) & (uint32_t)(cap - 1);
    while (F.item_hash[h].val != kEmpty) h = (h + 1) & (uint32_t)(cap - 1);
    F.item_hash[h] = ItemHashEntry{F.item_key[d], (uint32_t)d, 0};
  }
  return true;
}

// VMISIndex::new (vmis_index.rs:85-313) hands the query path posting lists, idf and attributes that were
// computed offline; this turns them into the flat, time-ranked layout of vmis_device.h and checks the
// properties the kernel's closed forms rely on (see avro_reader.h::PrebuiltInfo).
bool build_flat_index_prebuilt(const PrebuiltIndex& p, uint32_t n_shards, FlatIndex* out, PrebuiltInfo* info,
                               std::string* err) {
  const Sessions& s = p.sessions;
  const size_t S = s.size(), I = p.item_ids.size();
  if (S >= 0xFFFFFFFFull || I >= 0xFFFFFFFFull) { *err = "too many sessions / items"; return false; }
  if (I == 0) { *err = "no item records"; return false; }
  if (n_shards == 0 || n_shards > (uint32_t)kMaxShards) { *err = "n_shards must be in [1, 8]"; return false; }
  if (p.post_off.size() != I + 1 || p.idf.size() != I || p.attr.size() != I) { *err = "inconsistent item arrays"; return false; }
  FlatIndex& F = *out;
  F = FlatIndex();
  F.n_shards = n_shards;
  F.idf_weighting = 0.0;                                  // unknown: idf values come from the files
  PrebuiltInfo pi;

  // 1. item dictionary: dense index = rank of the external id
  std::vector<uint32_t> by_key(I); std::iota(by_key.begin(), by_key.end(), 0u);
  std::sort(by_key.begin(), by_key.end(), [&](uint32_t a, uint32_t b) { return p.item_ids[a] < p.item_ids[b]; });
  F.item_key.resize(I); F.idf.resize(I); F.attr.resize(I);
  for (size_t d = 0; d < I; ++d) {
    F.item_key[d] = p.item_ids[by_key[d]]; F.idf[d] = p.idf[by_key[d]]; F.attr[d] = p.attr[by_key[d]];
    if (d > 0 && F.item_key[d] == F.item_key[d - 1]) { *err = "duplicate item id " + std::to_string(F.item_key[d]); return false; }
  }
  size_t cap = 16; while (cap < I * 2) cap <<= 1;
  F.item_hash.assign(cap, ItemHashEntry{0, kEmpty, 0});
  for (size_t d = 0; d < I; ++d) {
    uint32_t h = (uint32_t)mix64(F.item_key[d]) & (uint32_t)(cap - 1);
    while (F.item_hash[h].val != kEmpty) h = (h + 1) & (uint32_t)(cap - 1);
    F.item_hash[h] = ItemHashEntry{F.item_key[d], (uint32_t)d, 0};
  }

  // 2. sessions named by at least one posting list are kept and ranked by (timestamp, session idx) ascending
  std::vector<uint8_t> used(S, 0);
  for (uint32_t sid : p.post_sessions) {
    if (sid >= S || s.off[sid + 1] == s.off[sid]) {
      *err = "a posting list names session " + std::to_string(sid) + ", which has no sessionindex record (the reference panics at mod.rs:138)";
      return false;
    }
    used[sid] = 1;
  }
  std::vector<uint64_t> order;
  for (size_t i = 0; i < S; ++i) if (used[i]) order.push_back(((uint64_t)s.ts[i] << 32) | (uint64_t)i);
  std::sort(order.begin(), order.end());
  const size_t Sk = order.size();
  F.rank_to_orig.resize(Sk);
  std::vector<uint32_t> rank_of(S, kEmpty);
  for (size_t r = 0; r < Sk; ++r) { F.rank_to_orig[r] = (uint32_t)(order[r] & 0xFFFFFFFFull); rank_of[F.rank_to_orig[r]] = (uint32_t)r; }

  // Steps 3-5 run over ranges of sessions / items on all host cores; the first problem found wins.
  std::mutex err_mu; std::string first_err;
  auto report = [&](const std::string& m) { std::lock_guard<std::mutex> g(err_mu); if (first_err.empty()) first_err = m; };
  auto parallel_for = [&](size_t n, auto&& body) {
    const size_t nt = std::max<size_t>(1, std::min<size_t>(std::thread::hardware_concurrency(), (n + 4095) / 4096));
    std::vector<std::thread> th;
    for (size_t t = 1; t < nt; ++t) th.emplace_back([&, t]() { body(n * t / nt, n * (t + 1) / nt); });
    body(0, n / nt);
    for (auto& x : th) x.join();
  };

  // 3. session -> items (dense, ascending), 16-byte aligned starts
  F.sess_ref.resize(Sk);
  uint64_t P = 0; uint32_t max_len = 0;
  {
    uint64_t start = 0;
    for (size_t r = 0; r < Sk; ++r) {
      const uint32_t o = F.rank_to_orig[r];
      const uint32_t len = (uint32_t)(s.off[o + 1] - s.off[o]);
      if ((start >> 2) > 0xFFFFFFFFull) { *err = "session item array too large"; return false; }
      F.sess_ref[r] = make_uint2((uint32_t)(start >> 2), len);
      start += (len + 3u) & ~3u;
      P += len; max_len = std::max(max_len, len);
    }
    F.sess_items.assign(start, kEmpty);
  }
  parallel_for(Sk, [&](size_t r0, size_t r1) {
    for (size_t r = r0; r < r1; ++r) {
      const uint32_t o = F.rank_to_orig[r];
      const uint2 ref = F.sess_ref[r];
      uint32_t* dst = &F.sess_items[(size_t)ref.x * 4];
      for (uint32_t t = 0; t < ref.y; ++t) {
        dst[t] = host_lookup_item(F, s.items[s.off[o] + t]);
        if (dst[t] == kEmpty) {
          report("session " + std::to_string(o) + " holds item " + std::to_string(s.items[s.off[o] + t]) +
                 ", which has no itemindex record (the reference panics at vmis_index.rs:322)");
          return;
        }
      }
      std::sort(dst, dst + ref.y);
      for (uint32_t t = 1; t < ref.y; ++t) if (dst[t] == dst[t - 1]) { report("duplicate item inside session " + std::to_string(o)); return; }
    }
  });
  if (!first_err.empty()) { *err = first_err; return false; }
  F.n_pairs_kept = P; F.max_len = max_len;

  // 4. postings as time ranks, descending; lists in another order (or with repeats) are normalised
  F.post_ref.resize(I);
  std::vector<std::vector<uint32_t>> lists(I);
  std::atomic<uint64_t> n_reordered{0}, n_dups{0};
  parallel_for(I, [&](size_t d0, size_t d1) {
    for (size_t d = d0; d < d1; ++d) {
      const uint32_t src = by_key[d];
      std::vector<uint32_t>& L = lists[d];
      L.reserve(p.post_off[src + 1] - p.post_off[src]);
      bool sorted = true;
      for (uint64_t e = p.post_off[src]; e < p.post_off[src + 1]; ++e) {
        const uint32_t r = rank_of[p.post_sessions[e]];
        const uint2 ref = F.sess_ref[r];
        const uint32_t* it = &F.sess_items[(size_t)ref.x * 4];
        if (!std::binary_search(it, it + ref.y, (uint32_t)d)) {
          report("the posting list of item " + std::to_string(F.item_key[d]) + " names session " +
                 std::to_string(p.post_sessions[e]) + ", which does not contain the item");
          return;
        }
        if (!L.empty() && r >= L.back()) sorted = false;
        L.push_back(r);
      }
      if (!sorted) {
        ++n_reordered;
        std::sort(L.begin(), L.end(), [](uint32_t a, uint32_t b) { return a > b; });
        const size_t before = L.size();
        L.erase(std::unique(L.begin(), L.end()), L.end());
        n_dups += before - L.size();
      }
    }
  });
  if (!first_err.empty()) { *err = first_err; return false; }
  pi.lists_reordered = n_reordered; pi.duplicate_postings = n_dups;
  std::vector<uint64_t> shard_size(n_shards, 0);
  uint32_t m_build = 0;
  for (size_t d = 0; d < I; ++d) {
    const uint32_t len = (uint32_t)lists[d].size();
    m_build = std::max(m_build, len);
    uint64_t& sz = shard_size[d % n_shards];
    if ((sz >> 2) > 0xFFFFFFFFull) { *err = "posting array too large"; return false; }
    F.post_ref[d] = make_uint2((uint32_t)(sz >> 2), len);
    F.n_postings += len;
    sz += (len + 3u) & ~3u;
  }
  F.m_build = std::max(m_build, 1u);
  F.shard_begin.assign(n_shards + 1, 0);
  for (uint32_t sh = 0; sh < n_shards; ++sh) F.shard_begin[sh + 1] = F.shard_begin[sh] + shard_size[sh];
  F.postings.assign(F.shard_begin[n_shards], kEmpty);
  parallel_for(I, [&](size_t d0, size_t d1) {
    for (size_t d = d0; d < d1; ++d)
      std::copy(lists[d].begin(), lists[d].end(), F.postings.begin() + F.shard_begin[d % n_shards] + (size_t)F.post_ref[d].x * 4);
  });

  // 5. m_carry: the kernel may take the first-match position of mod.rs:133-138 from the merged lists only if a
  //    session of the m-sample is on the list of every evolving item it contains.  That holds for m <= the
  //    shortest list that (a) misses some session containing its item and (b) is the most-recent prefix of them.
  uint32_t m_carry = 0xFFFFFFFFu;
  {
    std::vector<std::atomic<uint32_t>> newer_or_listed(I), df(I);
    for (size_t d = 0; d < I; ++d) { newer_or_listed[d].store(0, std::memory_order_relaxed); df[d].store(0, std::memory_order_relaxed); }
    parallel_for(Sk, [&](size_t r0, size_t r1) {
      for (size_t r = r0; r < r1; ++r) {
        const uint2 ref = F.sess_ref[r];
        const uint32_t* it = &F.sess_items[(size_t)ref.x * 4];
        for (uint32_t t = 0; t < ref.y; ++t) {
          const uint32_t d = it[t];
          df[d].fetch_add(1, std::memory_order_relaxed);
          if (!lists[d].empty() && (uint32_t)r >= lists[d].back()) newer_or_listed[d].fetch_add(1, std::memory_order_relaxed);
        }
      }
    });
    for (size_t d = 0; d < I; ++d) {
      const uint32_t len = (uint32_t)lists[d].size();
      if (len == df[d].load(std::memory_order_relaxed)) continue;      // complete list
      if (newer_or_listed[d].load(std::memory_order_relaxed) != len) { m_carry = 0; break; }   // not a most-recent prefix
      m_carry = std::min(m_carry, len);
    }
  }
  pi.m_carry = m_carry;
  F.m_carry = m_carry;
  if (info) *info = pi;
  return true;
}

}  // namespace vmis
