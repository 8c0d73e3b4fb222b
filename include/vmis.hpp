// include/vmis.hpp — header-only C++17 host mirror of the reference's Rust surface over the C ABI of vmis.h.
//
// Same names, argument meaning and error behaviour as the reference (file:line under the reference tree):
//   vmis::VMISIndex::new_from_csv(path, m_most_recent_sessions, idf_weighting)   src/vmisknn/vmis_index.rs:38
//   vmis::VMISIndex::new_(base_path)   [`new` is a C++ keyword]                 src/vmisknn/vmis_index.rs:85
//   trait SimilarityComputationNew { items_for_session, idf, find_neighbors, find_attributes }
//                                                                                src/vmisknn/similarity_indexed.rs:8-24
//   vmis::predict(index, evolving_session, k, m, how_many, enable_business_logic) src/vmisknn/mod.rs:118-125
// Where the reference panics (unknown item in idf(), I/O errors in the constructor) this wrapper throws
// vmis::Error; results come back in `into_sorted_vec()` order (score descending).
#pragma once
#include <cstdint>
#include <optional>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "vmis.h"

namespace vmis {

struct Error : std::runtime_error {
  int code;
  Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

struct SessionScore { uint32_t id; double score; };          // mod.rs:15-19
struct ItemScore { uint64_t id; double score; };             // mod.rs:45-49
struct ProductAttributes { bool is_adult; bool is_for_sale; };   // vmis_index.rs:23-26

class VMISIndex {
 public:
  // VMISIndex::new_from_csv — vmis_index.rs:38 (device: CUDA ordinal, VMIS_DEVICE_NONE for a host-only handle)
  static VMISIndex new_from_csv(const std::string& path_to_training, size_t m_most_recent_sessions, double idf_weighting,
                                int device = 0) {
    return VMISIndex(vmis_index_from_csv(path_to_training.c_str(), m_most_recent_sessions, idf_weighting, device));
  }
  // VMISIndex::new — vmis_index.rs:85: <base_path>/itemindex/*.avro + <base_path>/sessionindex/*.avro
  static VMISIndex new_(const std::string& base_path, int device = 0) {
    return VMISIndex(vmis_index_from_avro(base_path.c_str(), device));
  }
  // prepare_hashmap + struct assembly — vmis_index.rs:422, :75-82
  static VMISIndex from_sessions(const std::vector<std::vector<uint64_t>>& historical_sessions,
                                 const std::vector<uint32_t>& timestamps, size_t m_most_recent_sessions,
                                 size_t max_training_session_length, double idf_weighting, int device = 0) {
    std::vector<uint64_t> items, off{0};
    for (const auto& s : historical_sessions) { items.insert(items.end(), s.begin(), s.end()); off.push_back(items.size()); }
    return VMISIndex(vmis_index_from_sessions(items.data(), off.data(), timestamps.data(), timestamps.size(),
                                              m_most_recent_sessions, max_training_session_length, idf_weighting, device));
  }
  VMISIndex(VMISIndex&& o) noexcept : h_(o.h_) { o.h_ = nullptr; }
  VMISIndex& operator=(VMISIndex&& o) noexcept { std::swap(h_, o.h_); return *this; }
  VMISIndex(const VMISIndex&) = delete;
  VMISIndex& operator=(const VMISIndex&) = delete;
  ~VMISIndex() { vmis_index_free(h_); }

  // ---- trait SimilarityComputationNew (similarity_indexed.rs:8-24) ----
  std::pair<const uint64_t*, size_t> items_for_session(uint32_t session_idx) const {       // vmis_index.rs:317-319
    size_t n = 0;
    const uint64_t* p = vmis_items_for_session(h_, session_idx, &n);
    if (!p && n == 0 && vmis_last_error_code() != 0) fail();
    return {p, n};
  }
  double idf(uint64_t item_id) const {                                                     // vmis_index.rs:321-323
    double v = 0;
    if (vmis_idf(h_, item_id, &v) != 0) fail();                                            // the reference panics
    return v;
  }
  std::vector<SessionScore> find_neighbors(const std::vector<uint64_t>& evolving_session, size_t k, size_t m) const {
    const uint32_t off[2] = {0u, (uint32_t)evolving_session.size()};                      // vmis_index.rs:325-415
    std::vector<uint32_t> s(k ? k : 1); std::vector<double> sim(k ? k : 1); uint32_t cnt = 0;
    if (vmis_find_neighbors_batch(h_, evolving_session.data(), off, 1, (uint32_t)k, (uint32_t)m, s.data(), sim.data(), &cnt,
                                  nullptr) != 0) fail();
    std::vector<SessionScore> out(cnt);
    for (uint32_t i = 0; i < cnt; ++i) out[i] = SessionScore{s[i], sim[i]};
    return out;
  }
  std::optional<ProductAttributes> find_attributes(uint64_t item_id) const {               // vmis_index.rs:417-419
    const int a = vmis_find_attributes(h_, item_id);
    if (!(a & VMIS_ATTR_EXISTS)) return std::nullopt;
    return ProductAttributes{(a & VMIS_ATTR_ADULT) != 0, (a & VMIS_ATTR_FOR_SALE) != 0};
  }

  vmis_stats_t stats() const { vmis_stats_t st{}; if (vmis_index_stats(h_, &st) != 0) fail(); return st; }
  const vmis_index_t* handle() const { return h_; }

 private:
  explicit VMISIndex(vmis_index_t* h) : h_(h) { if (!h_) fail(); }
  [[noreturn]] static void fail() { throw Error(vmis_last_error_code(), vmis_last_error()); }
  vmis_index_t* h_;
  friend std::vector<ItemScore> predict(const VMISIndex&, const std::vector<uint64_t>&, size_t, size_t, size_t, bool);
};

// vmisknn::predict — mod.rs:118-125; order of `into_sorted_vec()` (recommend_resource.rs:58-62)
inline std::vector<ItemScore> predict(const VMISIndex& index, const std::vector<uint64_t>& evolving_session, size_t k, size_t m,
                                      size_t how_many, bool enable_business_logic) {
  std::vector<uint64_t> ids(how_many ? how_many : 1); std::vector<double> sc(how_many ? how_many : 1);
  const int n = vmis_predict(index.h_, evolving_session.data(), evolving_session.size(), k, m, how_many,
                             enable_business_logic ? 1 : 0, ids.data(), sc.data());
  if (n < 0) VMISIndex::fail();
  std::vector<ItemScore> out((size_t)n);
  for (int i = 0; i < n; ++i) out[(size_t)i] = ItemScore{ids[(size_t)i], sc[(size_t)i]};
  return out;
}

// batched predict over CSR queries (evaluator.rs / objective.rs replay): rows of `how_many`
struct BatchResult { std::vector<uint64_t> ids; std::vector<double> scores; std::vector<uint32_t> counts; };
inline BatchResult predict_batch(const VMISIndex& index, const std::vector<uint64_t>& q_items, const std::vector<uint32_t>& q_off,
                                 size_t k, size_t m, size_t how_many, bool enable_business_logic) {
  const uint32_t n_q = (uint32_t)(q_off.size() - 1);
  BatchResult r; r.ids.assign((size_t)n_q * how_many, 0); r.scores.assign((size_t)n_q * how_many, 0.0); r.counts.assign(n_q, 0);
  if (vmis_predict_batch(index.handle(), q_items.data(), q_off.data(), n_q, (uint32_t)k, (uint32_t)m, (uint32_t)how_many,
                         enable_business_logic ? 1 : 0, r.ids.data(), r.scores.data(), r.counts.data(), nullptr) != 0)
    throw Error(vmis_last_error_code(), vmis_last_error());
  return r;
}

// GET /v1/recommend without the HTTP layer (recommend_resource.rs:20-65): evolving-session window over an in-process
// store with RocksDBSessionStore semantics (sessions/mod.rs), then predict through a micro-batcher.  Thread safe.
class Server {
 public:
  Server(const VMISIndex& index, size_t k, size_t m, size_t how_many, size_t max_items_in_session,
         bool enable_business_logic = false, uint32_t max_batch = 4096, uint32_t max_wait_us = 200,
         uint64_t session_ttl_secs = 30 * 60, uint64_t max_session_idle_secs = 20 * 60)
      : how_many_(how_many), cap_(max_items_in_session),
        s_(vmis_server_create(index.handle(), (uint32_t)k, (uint32_t)m, (uint32_t)how_many, (uint32_t)max_items_in_session,
                              enable_business_logic ? 1 : 0, max_batch, max_wait_us, session_ttl_secs, max_session_idle_secs)) {
    if (!s_) throw Error(VMIS_ERR_ARG, "vmis_server_create failed");
  }
  Server(const Server&) = delete;
  Server& operator=(const Server&) = delete;
  ~Server() { vmis_server_destroy(s_); }
  // v1_recommend: the recommended item ids, best first
  std::vector<uint64_t> recommend(const std::string& session_id, uint64_t item_id, bool user_consent) {
    std::vector<uint64_t> ids(how_many_ ? how_many_ : 1);
    const int n = vmis_server_recommend(s_, session_id.c_str(), item_id, user_consent ? 1 : 0, ids.data(), nullptr);
    if (n < 0) throw Error(n, vmis_last_error());
    ids.resize((size_t)n);
    return ids;
  }
  // the window step alone (recommend_resource.rs:39-54)
  std::vector<uint64_t> session_window(const std::string& session_id, uint64_t item_id, bool user_consent) {
    std::vector<uint64_t> w(cap_ + 1);
    const int n = vmis_server_session_window(s_, session_id.c_str(), item_id, user_consent ? 1 : 0, w.data(), w.size());
    if (n < 0) throw Error(n, "vmis_server_session_window failed");
    w.resize((size_t)n);
    return w;
  }
  void set_clock(uint64_t epoch_secs) { vmis_server_set_clock(s_, epoch_secs); }

 private:
  size_t how_many_, cap_;
  vmis_server_t* s_;
};

}  // namespace vmis
