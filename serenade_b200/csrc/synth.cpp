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

namespace {

inline uint64_t splitmix(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull; x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull; x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
inline double u01(uint64_t r) { return (double)(r >> 11) * (1.0 / 9007199254740992.0); }

// Session-length quantile function through the reference's empirical percentiles
// (vmis_index.rs:116-126: p5=2 p25=2 p50=3 p75=6 p90=10 p95=14 p99=27 p99.5=34), capped at 34.
inline uint32_t session_length(double u) {
  static const double q[] = {0.0, 0.05, 0.25, 0.50, 0.75, 0.90, 0.95, 0.99, 0.995, 1.0};
  static const double v[] = {2, 2, 2, 3, 6, 10, 14, 27, 34, 34};
  int i = 0; while (i < 8 && u > q[i + 1]) ++i;
  const double t = (u - q[i]) / (q[i + 1] - q[i]);
  const long len = std::lround(v[i] + t * (v[i + 1] - v[i]));
  return (uint32_t)std::min(34l, std::max(1l, len));
}

// popularity rank (0-based) with P(r) ~ 1/(r+1): log-uniform inverse CDF
inline uint64_t zipf_rank(double u, uint64_t n_items, double log_n1) {
  uint64_t r = (uint64_t)std::exp(u * log_n1);   // in [1, n_items + 1)
  if (r < 1) r = 1; if (r > n_items) r = n_items;
  return r - 1;
}
// bijection rank -> sparse external id (odd multiplier modulo 2^48), exercises the id map
inline uint64_t external_id(uint64_t rank) { return ((rank + 1) * 0x9E3779B97F4Bull) & 0xFFFFFFFFFFFFull; }

// Feistel permutation of [0, n) by cycle walking: unique pseudo-random timestamps
struct Perm {
  uint64_t n, seed; unsigned half; uint64_t hmask;
  Perm(uint64_t n_, uint64_t seed_) : n(n_), seed(seed_) {
    unsigned bits = 2; while ((1ull << bits) < n) bits += 2;
    half = bits / 2; hmask = (1ull << half) - 1;
  }
  uint64_t operator()(uint64_t x) const {
    do {
      uint64_t l = x >> half, r = x & hmask;
      for (int round = 0; round < 4; ++round) {
        const uint64_t f = splitmix(r ^ (seed + (uint64_t)round * 0x1234567ull)) & hmask;
        const uint64_t nl = r; r = l ^ f; l = nl;
      }
      x = (l << half) | r;
    } while (x >= n);
    return x;
  }
};

// items of session number `sn` under `seed`: distinct ranks, returns count
inline uint32_t gen_session(uint64_t seed, uint64_t sn, uint64_t n_items, double log_n1, uint64_t* out) {
  const uint64_t base = splitmix(seed ^ splitmix(sn));
  uint32_t len = session_length(u01(splitmix(base)));
  if ((uint64_t)len > n_items) len = (uint32_t)n_items;
  uint32_t cnt = 0; uint64_t ctr = 1;
  while (cnt < len) {
    const uint64_t r = zipf_rank(u01(splitmix(base + (ctr++) * 0x632BE59BD9B4E019ull)), n_items, log_n1);
    bool dup = false;
    for (uint32_t i = 0; i < cnt; ++i) if (out[i] == r) { dup = true; break; }
    if (!dup) out[cnt++] = r;
    if (ctr > 4096) { // pathological tiny catalogues: fill with the first unused ranks
      for (uint64_t c = 0; cnt < len; ++c) { bool d = false; for (uint32_t i = 0; i < cnt; ++i) if (out[i] == c) d = true; if (!d) out[cnt++] = c; }
    }
  }
  return len;
}

}  // namespace

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
        uint32_t len = session_length(u01(splitmix(splitmix(seed ^ splitmix(sn)))));
        if ((uint64_t)len > n_items) len = (uint32_t)n_items;
        s += len;
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
    uint32_t len = session_length(u01(splitmix(splitmix(seed ^ splitmix(sn)))));
    if ((uint64_t)len > n_items) len = (uint32_t)n_items;
    sess_off[sn + 1] = sess_off[sn] + len;
  }
  if (n_interactions) *n_interactions = sess_off[n_sessions];
  const Perm perm(n_sessions, splitmix(seed ^ 0x7157ull));
  const uint64_t ts_base = 1500000000ull;   // spread over [base, base + n_sessions): unique per session
  std::vector<std::thread> th;
  for (unsigned t = 0; t < nt; ++t) th.emplace_back([&, t]() {
    uint64_t tmp[64];
    const uint64_t lo = n_sessions * t / nt, hi = n_sessions * (t + 1) / nt;
    for (uint64_t sn = lo; sn < hi; ++sn) {
      const uint32_t len = gen_session(seed, sn, n_items, log_n1, tmp);
      for (uint32_t i = 0; i < len; ++i) tmp[i] = external_id(tmp[i]);
      std::sort(tmp, tmp + len);
      std::copy(tmp, tmp + len, items + sess_off[sn]);
      sess_ts[sn] = (uint32_t)(ts_base + perm(sn));
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
