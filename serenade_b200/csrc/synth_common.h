// serenade_b200/csrc/synth_common.h — the synthetic click-log generator's pure functions, shared by the host
// generator (synth.cpp) and the on-device generator (build_sm100.cu).  Spec: SURVEY.md §8d.
#pragma once
#include <cmath>
#include <cstdint>

#ifdef __CUDACC__
#define VMIS_HD __host__ __device__ __forceinline__
#else
#define VMIS_HD inline
#endif

namespace vmis_synth {

VMIS_HD uint64_t splitmix(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull; x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull; x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
VMIS_HD double u01(uint64_t r) { return (double)(r >> 11) * (1.0 / 9007199254740992.0); }

// Session-length quantile function through the reference's empirical percentiles
// (vmis_index.rs:116-126: p5=2 p25=2 p50=3 p75=6 p90=10 p95=14 p99=27 p99.5=34), capped at 34.
VMIS_HD uint32_t session_length(double u) {
  const double q[] = {0.0, 0.05, 0.25, 0.50, 0.75, 0.90, 0.95, 0.99, 0.995, 1.0};
  const double v[] = {2, 2, 2, 3, 6, 10, 14, 27, 34, 34};
  int i = 0; while (i < 8 && u > q[i + 1]) ++i;
  const double t = (u - q[i]) / (q[i + 1] - q[i]);
  long long len = llround(v[i] + t * (v[i + 1] - v[i]));
  if (len < 1) len = 1; if (len > 34) len = 34;
  return (uint32_t)len;
}
VMIS_HD uint32_t session_length_of(uint64_t seed, uint64_t sn, uint64_t n_items) {
  uint32_t len = session_length(u01(splitmix(splitmix(seed ^ splitmix(sn)))));
  if ((uint64_t)len > n_items) len = (uint32_t)n_items;
  return len;
}

// popularity rank (0-based) with P(r) ~ 1/(r+1): log-uniform inverse CDF
VMIS_HD uint64_t zipf_rank(double u, uint64_t n_items, double log_n1) {
  uint64_t r = (uint64_t)exp(u * log_n1);   // in [1, n_items + 1)
  if (r < 1) r = 1; if (r > n_items) r = n_items;
  return r - 1;
}
// bijection rank -> sparse external id (odd multiplier modulo 2^48), exercises the id map
VMIS_HD uint64_t external_id(uint64_t rank) { return ((rank + 1) * 0x9E3779B97F4Bull) & 0xFFFFFFFFFFFFull; }

// Feistel permutation of [0, n) by cycle walking: unique pseudo-random timestamps
struct Perm {
  uint64_t n, seed; unsigned half; uint64_t hmask;
  VMIS_HD Perm(uint64_t n_, uint64_t seed_) : n(n_), seed(seed_) {
    unsigned bits = 2; while ((1ull << bits) < n) bits += 2;
    half = bits / 2; hmask = (1ull << half) - 1;
  }
  VMIS_HD uint64_t operator()(uint64_t x) const {
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
constexpr uint64_t kTsBase = 1500000000ull;   // timestamps spread over [base, base + n_sessions): unique per session

// item ranks of session number `sn` under `seed` in click order: distinct ranks, returns count (<= 34)
VMIS_HD uint32_t gen_session(uint64_t seed, uint64_t sn, uint64_t n_items, double log_n1, uint64_t* out) {
  const uint64_t base = splitmix(seed ^ splitmix(sn));
  uint32_t len = session_length(u01(splitmix(base)));
  if ((uint64_t)len > n_items) len = (uint32_t)n_items;
  uint32_t cnt = 0; uint64_t ctr = 1;
  while (cnt < len) {
    const uint64_t r = zipf_rank(u01(splitmix(base + (ctr++) * 0x632BE59BD9B4E019ull)), n_items, log_n1);
    bool dup = false;
    for (uint32_t i = 0; i < cnt; ++i) if (out[i] == r) { dup = true; break; }
    if (!dup) out[cnt++] = r;
    if (ctr > 4096) {   // pathological tiny catalogues: fill with the first unused ranks
      for (uint64_t c = 0; cnt < len; ++c) { bool d = false; for (uint32_t i = 0; i < cnt; ++i) if (out[i] == c) d = true; if (!d) out[cnt++] = c; }
    }
  }
  return len;
}

}  // namespace vmis_synth
