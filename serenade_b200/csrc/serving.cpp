// serenade_b200/csrc/serving.cpp — the step in front of the hot path online: the evolving-session window of
// GET /v1/recommend (recommend_resource.rs:20-65) over an in-process session store with the semantics of
// RocksDBSessionStore (sessions/mod.rs:8-71, opened with a 30 min TTL at serving.rs:55-56), feeding the
// micro-batcher (batcher.cpp) so that concurrent worker threads share GPU launches.
//
//   key            md5(session_id) as u128 (recommend_resource.rs:27-28) — MD5 implemented below (RFC 1321)
//   get            stored items if the last update is at most max_session_idle_duration (20 min) old, else
//                  empty (sessions/mod.rs:37-57)
//   window         empty → [item]; else append unless it repeats the last item, keep the most recent
//                  max_items_in_session (recommend_resource.rs:39-49; one element is dropped per request)
//   update         items + current epoch seconds (sessions/mod.rs:59-71); entries older than the TTL are
//                  dropped (RocksDB does it at compaction; here a sweep every `sweep_every` updates per stripe)
//   no consent     the store is not touched, the query is [item] (recommend_resource.rs:52-54)
//
// The HTTP layer, istio routing and the on-disk store are out of scope; this is the
// request logic a server thread runs between parsing the query string and writing the JSON array.
#include <chrono>
#include <cstring>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/vmis.h"
#include "vmis_host.h"

namespace {

// ---- MD5 (RFC 1321) ----
struct Md5 {
  static uint32_t rol(uint32_t x, int c) { return (x << c) | (x >> (32 - c)); }
  static void digest(const uint8_t* msg, size_t len, uint8_t out[16]) {
    static const uint32_t K[64] = {
        0xd76aa478, 0xe8c7b756, 0x242070db, 0xc1bdceee, 0xf57c0faf, 0x4787c62a, 0xa8304613, 0xfd469501, 0x698098d8, 0x8b44f7af,
        0xffff5bb1, 0x895cd7be, 0x6b901122, 0xfd987193, 0xa679438e, 0x49b40821, 0xf61e2562, 0xc040b340, 0x265e5a51, 0xe9b6c7aa,
        0xd62f105d, 0x02441453, 0xd8a1e681, 0xe7d3fbc8, 0x21e1cde6, 0xc33707d6, 0xf4d50d87, 0x455a14ed, 0xa9e3e905, 0xfcefa3f8,
        0x676f02d9, 0x8d2a4c8a, 0xfffa3942, 0x8771f681, 0x6d9d6122, 0xfde5380c, 0xa4beea44, 0x4bdecfa9, 0xf6bb4b60, 0xbebfbc70,
        0x289b7ec6, 0xeaa127fa, 0xd4ef3085, 0x04881d05, 0xd9d4d039, 0xe6db99e5, 0x1fa27cf8, 0xc4ac5665, 0xf4292244, 0x432aff97,
        0xab9423a7, 0xfc93a039, 0x655b59c3, 0x8f0ccc92, 0xffeff47d, 0x85845dd1, 0x6fa87e4f, 0xfe2ce6e0, 0xa3014314, 0x4e0811a1,
        0xf7537e82, 0xbd3af235, 0x2ad7d2bb, 0xeb86d391};
    static const int R[64] = {7, 12, 17, 22, 7, 12, 17, 22, 7, 12, 17, 22, 7, 12, 17, 22, 5, 9,  14, 20, 5, 9,
                              14, 20, 5, 9,  14, 20, 5, 9,  14, 20, 4, 11, 16, 23, 4, 11, 16, 23, 4, 11, 16, 23,
                              4, 11, 16, 23, 6, 10, 15, 21, 6, 10, 15, 21, 6, 10, 15, 21, 6, 10, 15, 21};
    uint32_t h0 = 0x67452301, h1 = 0xefcdab89, h2 = 0x98badcfe, h3 = 0x10325476;
    std::vector<uint8_t> m(msg, msg + len);
    m.push_back(0x80);
    while (m.size() % 64 != 56) m.push_back(0);
    const uint64_t bits = (uint64_t)len * 8;
    for (int i = 0; i < 8; ++i) m.push_back((uint8_t)(bits >> (8 * i)));
    for (size_t off = 0; off < m.size(); off += 64) {
      uint32_t w[16];
      for (int i = 0; i < 16; ++i) std::memcpy(&w[i], &m[off + 4 * i], 4);
      uint32_t a = h0, b = h1, c = h2, d = h3;
      for (int i = 0; i < 64; ++i) {
        uint32_t f; int g;
        if (i < 16) { f = (b & c) | (~b & d); g = i; }
        else if (i < 32) { f = (d & b) | (~d & c); g = (5 * i + 1) % 16; }
        else if (i < 48) { f = b ^ c ^ d; g = (3 * i + 5) % 16; }
        else { f = c ^ (b | ~d); g = (7 * i) % 16; }
        const uint32_t t = d; d = c; c = b;
        b = b + rol(a + f + K[i] + w[g], R[i]);
        a = t;
      }
      h0 += a; h1 += b; h2 += c; h3 += d;
    }
    std::memcpy(out, &h0, 4); std::memcpy(out + 4, &h1, 4); std::memcpy(out + 8, &h2, 4); std::memcpy(out + 12, &h3, 4);
  }
};

struct Key128 {
  uint64_t hi, lo;                      // Builder::from_bytes(digest).build().as_u128(): big-endian bytes
  bool operator==(const Key128& o) const { return hi == o.hi && lo == o.lo; }
};
struct KeyHash { size_t operator()(const Key128& k) const { return (size_t)(k.hi ^ (k.lo * 0x9E3779B97F4A7C15ull)); } };

struct DBValue { std::vector<uint64_t> session_items; uint64_t epoch_secs; };      // sessions/mod.rs:12-16

constexpr int kStripes = 64;

}  // namespace

struct vmis_server {
  vmis_batcher_t* batcher = nullptr;
  uint32_t how_many = 0, max_items_in_session = 0;
  uint64_t ttl_secs = 0, idle_secs = 0;
  uint64_t fixed_now = 0;               // tests: 0 = system clock
  uint32_t sweep_every = 4096;
  struct Stripe { std::mutex mu; std::unordered_map<Key128, DBValue, KeyHash> map; uint32_t since_sweep = 0; };
  Stripe stripes[kStripes];

  uint64_t now() const {
    if (fixed_now) return fixed_now;
    return (uint64_t)std::chrono::duration_cast<std::chrono::seconds>(std::chrono::system_clock::now().time_since_epoch()).count();
  }
  static Key128 key_of(const char* session_id) {
    uint8_t d[16];
    Md5::digest(reinterpret_cast<const uint8_t*>(session_id), std::strlen(session_id), d);
    Key128 k{0, 0};
    for (int i = 0; i < 8; ++i) { k.hi = (k.hi << 8) | d[i]; k.lo = (k.lo << 8) | d[8 + i]; }
    return k;
  }
  Stripe& stripe_of(const Key128& k) { return stripes[(k.lo ^ k.hi) % kStripes]; }

  // recommend_resource.rs:39-54 on one stripe lock: get → window → update; returns the evolving session
  std::vector<uint64_t> advance(const Key128& key, uint64_t item) {
    Stripe& st = stripe_of(key);
    const uint64_t t = now();
    std::lock_guard<std::mutex> g(st.mu);
    std::vector<uint64_t> items;
    auto it = st.map.find(key);
    if (it != st.map.end() && t - it->second.epoch_secs <= idle_secs && t - it->second.epoch_secs <= ttl_secs)   // sessions/mod.rs:45-52
      items = it->second.session_items;
    if (items.empty()) items.push_back(item);                                        // recommend_resource.rs:41-42
    else if (items.back() != item) {                                                 // :43
      items.push_back(item);
      if (items.size() > max_items_in_session) items.erase(items.begin());           // :45-48 drain(0..1)
    }
    DBValue& v = st.map[key];                                                        // sessions/mod.rs:59-71
    v.session_items = items; v.epoch_secs = t;
    if (++st.since_sweep >= sweep_every) {                                           // TTL (serving.rs:55-56)
      st.since_sweep = 0;
      for (auto e = st.map.begin(); e != st.map.end();) e = (t - e->second.epoch_secs > ttl_secs) ? st.map.erase(e) : std::next(e);
    }
    return items;
  }
};

extern "C" {

vmis_server_t* vmis_server_create(const vmis_index_t* index, uint32_t k, uint32_t m, uint32_t how_many,
                                  uint32_t max_items_in_session, int enable_business_logic, uint32_t max_batch,
                                  uint32_t max_wait_us, uint64_t session_ttl_secs, uint64_t max_session_idle_secs) {
  if (!index || max_items_in_session == 0) { vmis::set_last_error(VMIS_ERR_ARG, "vmis_server_create: NULL index or max_items_in_session == 0"); return nullptr; }
  // a window longer than the kernel takes would make every request of a long visit fail (and, before requests were
  // validated one by one, its batch mates too): refuse the configuration instead
  if (max_items_in_session > VMIS_MAX_SESSION_LEN) {
    vmis::set_last_error(VMIS_ERR_LIMIT, "max_items_in_session exceeds VMIS_MAX_SESSION_LEN (128)");
    return nullptr;
  }
  vmis_server* s = new vmis_server();
  s->batcher = vmis_batcher_create(index, k, m, how_many, enable_business_logic, max_batch ? max_batch : 4096, max_wait_us);
  if (!s->batcher) { delete s; return nullptr; }
  s->how_many = how_many; s->max_items_in_session = max_items_in_session;
  s->ttl_secs = session_ttl_secs ? session_ttl_secs : 30 * 60;                 // serving.rs:55
  s->idle_secs = max_session_idle_secs ? max_session_idle_secs : 20 * 60;       // sessions/mod.rs:34
  return s;
}

int vmis_server_session_window(vmis_server_t* s, const char* session_id, uint64_t item_id, int user_consent,
                               uint64_t* out_items, size_t cap) {
  if (!s || !session_id || (!out_items && cap)) return VMIS_ERR_ARG;
  std::vector<uint64_t> items;
  if (user_consent) items = s->advance(vmis_server::key_of(session_id), item_id);
  else items.push_back(item_id);                                                // recommend_resource.rs:52-54
  for (size_t i = 0; i < items.size() && i < cap; ++i) out_items[i] = items[i];
  return (int)items.size();
}

int vmis_server_recommend(vmis_server_t* s, const char* session_id, uint64_t item_id, int user_consent,
                          uint64_t* out_ids, double* out_scores_or_null) {
  if (!s || !session_id || (!out_ids && s->how_many)) return VMIS_ERR_ARG;
  std::vector<uint64_t> items;
  if (user_consent) items = s->advance(vmis_server::key_of(session_id), item_id);
  else items.push_back(item_id);
  std::vector<double> scores(s->how_many ? s->how_many : 1);
  return vmis_batcher_predict(s->batcher, items.data(), items.size(), out_ids, out_scores_or_null ? out_scores_or_null : scores.data());
}

int vmis_server_stored_items(vmis_server_t* s, const char* session_id, uint64_t* out_items, size_t cap) {
  if (!s || !session_id) return VMIS_ERR_ARG;
  const Key128 key = vmis_server::key_of(session_id);
  vmis_server::Stripe& st = s->stripe_of(key);
  const uint64_t t = s->now();
  std::lock_guard<std::mutex> g(st.mu);
  auto it = st.map.find(key);
  if (it == st.map.end() || t - it->second.epoch_secs > s->idle_secs || t - it->second.epoch_secs > s->ttl_secs) return 0;   // sessions/mod.rs:37-57
  const auto& v = it->second.session_items;
  for (size_t i = 0; i < v.size() && i < cap; ++i) out_items[i] = v[i];
  return (int)v.size();
}

int vmis_server_set_clock(vmis_server_t* s, uint64_t epoch_secs) {
  if (!s) return VMIS_ERR_ARG;
  s->fixed_now = epoch_secs;
  const uint64_t t = s->now();                                                   // a clock jump also runs the TTL sweep
  for (auto& st : s->stripes) {
    std::lock_guard<std::mutex> g(st.mu);
    for (auto e = st.map.begin(); e != st.map.end();) e = (t - e->second.epoch_secs > s->ttl_secs) ? st.map.erase(e) : std::next(e);
  }
  return VMIS_OK;
}

int vmis_server_stats(vmis_server_t* s, uint64_t* n_sessions, uint64_t* n_batches, uint64_t* n_requests) {
  if (!s) return VMIS_ERR_ARG;
  if (n_sessions) {
    uint64_t n = 0;
    for (auto& st : s->stripes) { std::lock_guard<std::mutex> g(st.mu); n += st.map.size(); }
    *n_sessions = n;
  }
  return vmis_batcher_stats(s->batcher, n_batches, n_requests);
}

void vmis_md5(const void* data, size_t len, uint8_t out16[16]) { Md5::digest(static_cast<const uint8_t*>(data), len, out16); }

void vmis_server_destroy(vmis_server_t* s) {
  if (!s) return;
  vmis_batcher_destroy(s->batcher);
  delete s;
}

}  // extern "C"
