// serenade_b200/csrc/synth.cpp — deterministic synthetic click-log generator for
// BASELINE.json configs 2-5 (the reference ships no generator; spec in SURVEY.md §8d).
// Counter-based (stateless) randomness: every session is a pure function of
// (seed, session number), so generation can be split across threads or devices
// without changing the data.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <thread>
#include <vector>

#include "../../include/vmis.h"

#include "synth_common.h"

using namespace vmis_synth;

extern "C" {

int vmis_synth_sessions(uint64_t seed, uint64_t n_items, uint64_t n_sessions, uint64_t* items, uint64_t* sess_off,
                        uint32_t* sess_ts, uint64_t* n_interactions) {
  if (n_items == 0 || n_sessions == 0 || n_sessions >= 0xFFFFFFF0ull) return VMIS_ERR_ARG;
  const double log_n1 = std::log((double)n_items + 1.0);
  const unsigned nt = std::max(1u, std::min(32u, std::thread::hardware_concurrency()));
  if (!items) {   // sizing pass: lengths only
    std::vector<uint64_t> part(nt, 0);
    std::vector<std::thread> th;
    for (unsigned t = 0; t < nt; ++t) th.emplace_back([&, t]() {
      uint64_t s = 0;
      for (uint64_t sn = t; sn < n_sessions; sn += nt) {
        s += session_length_of(seed, sn, n_items);
      }
      part[t] = s; });
    for (auto& x : th) x.join();
    uint64_t tot = 0; for (auto v : part) tot += v;
    if (n_interactions) *n_interactions = tot;
    return VMIS_OK;
  }
  if (!sess_off || !sess_ts) return VMIS_ERR_ARG;
  // offsets first (serial prefix over lengths), then items in parallel
  sess_off[0] = 0;
  for (uint64_t sn = 0; sn < n_sessions; ++sn) {
    sess_off[sn + 1] = sess_off[sn] + session_length_of(seed, sn, n_items);
  }
  if (n_interactions) *n_interactions = sess_off[n_sessions];
  const Perm perm(n_sessions, splitmix(seed ^ 0x7157ull));
  std::vector<std::thread> th;
  for (unsigned t = 0; t < nt; ++t) th.emplace_back([&, t]() {
    uint64_t tmp[64];
    const uint64_t lo = n_sessions * t / nt, hi = n_sessions * (t + 1) / nt;
    for (uint64_t sn = lo; sn < hi; ++sn) {
      const uint32_t len = gen_session(seed, sn, n_items, log_n1, tmp);
      for (uint32_t i = 0; i < len; ++i) tmp[i] = external_id(tmp[i]);
      std::sort(tmp, tmp + len);
      std::copy(tmp, tmp + len, items + sess_off[sn]);
      sess_ts[sn] = (uint32_t)(kTsBase + perm(sn));
    } });
  for (auto& x : th) x.join();
  return VMIS_OK;
}

int vmis_synth_queries(uint64_t seed, uint64_t n_items, uint32_t n_q, uint32_t max_items_in_session,
                       uint64_t* q_items, uint32_t* q_off) {
  if (n_items == 0 || !q_items || !q_off || max_items_in_session == 0) return VMIS_ERR_ARG;
  const double log_n1 = std::log((double)n_items + 1.0);
  uint64_t tmp[64];
  q_off[0] = 0;
  for (uint32_t q = 0; q < n_q; ++q) {
    const uint32_t len = gen_session(seed, q, n_items, log_n1, tmp);          // held-out session, click order
    const uint64_t r = splitmix(splitmix(seed ^ 0xABCDull) + q);
    const uint32_t prefix = 1 + (uint32_t)(r % len);                            // evaluator.rs:47 session_state
    const uint32_t L = std::min(prefix, max_items_in_session);                  // evaluator.rs:49-56
    uint32_t o = q_off[q];
    for (uint32_t i = prefix - L; i < prefix; ++i) q_items[o++] = external_id(tmp[i]);
    q_off[q + 1] = o;
  }
  return VMIS_OK;
}

}  // extern "C"
