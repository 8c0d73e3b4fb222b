// serenade_b200/csrc/build_sm100.cu — on-device index build (the prepare_hashmap equivalent,
// vmis_index.rs:422-528) and on-device synthetic click-log generation for the large configs.
//
// The build produces exactly the arrays of IndexView (vmis_device.h) from training sessions that already live
// in HBM: time-rank the kept sessions (radix sort of (timestamp, session idx)), build the item dictionary
// (hash set of the external ids, radix sort of the distinct ones, device hash table), emit the session→items lists (dense, ascending,
// 16-byte aligned) and the item→sessions posting lists (stable radix sort of the rank-ordered (item, rank) pairs on
// the item bits, truncate to m, shard by item).  idf needs `ln`: the document frequencies go to the host and idf is computed there with the same libm
// as the CPU oracle, so scores stay bit-identical to the host-built index.  Sorting uses CUB's device radix sort
// (a plain library sort, not a hot-path kernel); every other step is a kernel in this file.
#include <cub/cub.cuh>

#include <algorithm>
#include <cmath>
#include <string>
#include <thread>
#include <vector>

#include "build_device.h"
#include "synth_common.h"

namespace vmis {
namespace {

#define CU_OK(expr)                                                                              \
  do {                                                                                           \
    cudaError_t e__ = (expr);                                                                    \
    if (e__ != cudaSuccess) { *err = std::string(#expr) + ": " + cudaGetErrorString(e__); return false; } \
  } while (0)

struct DevBuf {                       // RAII device allocation
  void* p = nullptr;
  ~DevBuf() { if (p) cudaFree(p); }
  cudaError_t alloc(size_t bytes) { if (p) { cudaFree(p); p = nullptr; } return cudaMalloc(&p, std::max<size_t>(bytes, 16)); }
  void release_to(void** out) { *out = p; p = nullptr; }
  void free_now() { if (p) { cudaFree(p); p = nullptr; } }
  template <class T> T* as() { return static_cast<T*>(p); }
};

constexpr int kB = 256;
inline unsigned blocks_for(uint64_t n) { return (unsigned)std::min<uint64_t>((n + kB - 1) / kB, 1u << 30); }
constexpr uint32_t kMaxLenDevice = 128;   // per-thread staging of one session in k_emit_sessions

__device__ __forceinline__ uint32_t hash_u64_dev(uint64_t x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
  return (uint32_t)x;
}

// (timestamp, session idx) sort key of every kept session; pruned / empty sessions sort last
__global__ void k_session_keys(const uint64_t* off, const uint32_t* ts, uint64_t S, uint64_t max_len, uint64_t* keys,
                               unsigned long long* n_kept, unsigned long long* n_pairs) {
  const uint64_t s = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint64_t len = 0; bool kept = false;
  if (s < S) { len = off[s + 1] - off[s]; kept = len > 0 && len <= max_len; keys[s] = kept ? (((uint64_t)ts[s] << 32) | s) : ~0ull; }
  // block-level counts
  const unsigned km = __ballot_sync(0xFFFFFFFFu, kept);
  unsigned long long l = kept ? len : 0;
  for (int d = 16; d > 0; d >>= 1) l += __shfl_xor_sync(0xFFFFFFFFu, l, d);
  if ((threadIdx.x & 31) == 0 && km) { atomicAdd(n_kept, (unsigned long long)__popc(km)); atomicAdd(n_pairs, l); }
}

__global__ void k_rank_lengths(const uint64_t* sorted_keys, const uint64_t* off, uint64_t Sk, uint32_t* rank_to_orig,
                               uint64_t* elen, uint64_t* plen) {
  const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= Sk) return;
  const uint32_t o = (uint32_t)sorted_keys[r];
  rank_to_orig[r] = o;
  const uint64_t len = off[o + 1] - off[o];
  elen[r] = len; plen[r] = (len + 3) & ~3ull;
}

// ---- item dictionary by hashing: every external id of the kept sessions goes into an open-addressing SET (8-byte
// slots, all ones = empty); most probes end on a plain load because the id is already there (60 M interactions hold
// 1.7 M distinct items), so there are few atomics and no 64-bit sort of all interactions (that sort was 4.4 of the
// build's 13.6 ms in round 1).  The distinct ids are then collected and only THEY are sorted.
constexpr unsigned long long kNoKey = ~0ull;
__global__ void k_dedup_insert(const uint64_t* items, const uint64_t* off, const uint32_t* rank_to_orig, uint64_t Sk,
                               unsigned long long* set, uint64_t mask, unsigned long long* n_distinct, unsigned int* flags) {
  const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= Sk) return;
  const uint32_t o = rank_to_orig[r];
  const uint64_t b = off[o], n = off[o + 1] - b;
  for (uint64_t t = 0; t < n; ++t) {
    const unsigned long long k = items[b + t];
    if (k == kNoKey) { flags[0] = 1u; continue; }                      // the one id that collides with "empty"
    uint64_t h = hash_u64_dev(k) & mask;
    for (;;) {
      unsigned long long cur = set[h];
      if (cur == k) break;
      if (cur == kNoKey) {
        cur = atomicCAS(&set[h], kNoKey, k);
        if (cur == kNoKey) { if (atomicAdd(n_distinct, 1ull) + 1 > (mask + 1) / 2) flags[1] = 1u; break; }   // over half full: retry bigger
        if (cur == k) break;
      }
      h = (h + 1) & mask;
    }
    if (flags[1]) return;
  }
}
__global__ void k_dedup_collect(const unsigned long long* set, uint64_t cap, uint64_t* uniq, unsigned long long* cursor) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned long long k = i < cap ? set[i] : kNoKey;
  const bool have = k != kNoKey;
  const unsigned m = __ballot_sync(0xFFFFFFFFu, have);
  if (!m) return;
  unsigned long long base = 0;
  const int lane = threadIdx.x & 31, leader = __ffs((int)m) - 1;
  if (lane == leader) base = atomicAdd(cursor, (unsigned long long)__popc(m));
  base = __shfl_sync(0xFFFFFFFFu, base, leader);
  if (have) uniq[base + __popc(m & ((1u << lane) - 1u))] = k;
}

__global__ void k_hash_clear(ItemHashEntry* tab, uint64_t cap) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < cap) { tab[i].key = 0; tab[i].val = kEmpty; tab[i].pad = 0; }
}
__global__ void k_hash_insert(const uint64_t* item_key, uint32_t I, ItemHashEntry* tab, uint32_t mask) {
  const uint32_t d = blockIdx.x * blockDim.x + threadIdx.x;
  if (d >= I) return;
  const uint64_t key = item_key[d];
  uint32_t h = hash_u64_dev(key) & mask;
  for (;;) {
    if (atomicCAS(&tab[h].val, kEmpty, d) == kEmpty) { tab[h].key = key; return; }
    h = (h + 1) & mask;
  }
}
__device__ __forceinline__ uint32_t hash_lookup(const ItemHashEntry* tab, uint32_t mask, uint64_t key) {
  uint32_t h = hash_u64_dev(key) & mask;
  for (;;) {
    const ItemHashEntry e = tab[h];
    if (e.val == kEmpty) return kEmpty;
    if (e.key == key) return e.val;
    h = (h + 1) & mask;
  }
}

// one thread per kept session (by rank): dense item indices, ascending, 16-byte aligned + padded; also the
// (item, ~rank) keys whose sort yields the posting lists
__global__ void k_emit_sessions(const uint64_t* items, const uint64_t* off, const uint32_t* rank_to_orig, const uint64_t* estart,
                                const uint64_t* pstart, uint64_t Sk, const ItemHashEntry* tab, uint32_t mask, uint2* sess_ref,
                                uint32_t* sess_items, uint32_t* post_item, uint32_t* post_rank, unsigned int* error_flag) {
  const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= Sk) return;
  const uint32_t o = rank_to_orig[r];
  const uint64_t b = off[o];
  const uint32_t n = (uint32_t)(off[o + 1] - b);
  uint32_t idx[kMaxLenDevice];
  for (uint32_t t = 0; t < n; ++t) {           // insertion sort while translating
    const uint32_t v = hash_lookup(tab, mask, items[b + t]);
    uint32_t j = t;
    while (j > 0 && idx[j - 1] > v) { idx[j] = idx[j - 1]; --j; }
    idx[j] = v;
  }
  const uint64_t ps = pstart[r], es = estart[r];
  sess_ref[r] = make_uint2((uint32_t)(ps >> 2), n);
  const uint32_t np = (n + 3u) & ~3u;
  for (uint32_t t = 0; t < np; ++t) sess_items[ps + t] = t < n ? idx[t] : kEmpty;
  for (uint32_t t = 0; t < n; ++t) {
    if (t > 0 && idx[t] == idx[t - 1]) atomicExch(error_flag, 1u);       // duplicate item inside a session
    post_item[es + t] = idx[t]; post_rank[es + t] = (uint32_t)r;          // emitted in rank order: a STABLE sort by item keeps it
  }
}

__global__ void k_seg_starts(const uint32_t* sorted_item, uint64_t P, uint32_t I, uint64_t* seg_start) {
  const uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) { if (p == P) seg_start[I] = P; return; }
  const uint32_t d = sorted_item[p];
  if (p == 0 || d != sorted_item[p - 1]) seg_start[d] = p;
}

// per item: df, truncated length, padded length written at its shard-major position
__global__ void k_post_len(const uint64_t* seg_start, uint32_t I, uint64_t m, uint32_t n_shards, uint32_t per_shard,
                           uint32_t* df, uint64_t* padded_t) {
  const uint32_t d = blockIdx.x * blockDim.x + threadIdx.x;
  if (d >= I) return;
  const uint64_t f = seg_start[d + 1] - seg_start[d];
  df[d] = (uint32_t)f;
  const uint64_t len = f < m ? f : m;
  padded_t[(uint64_t)(d % n_shards) * per_shard + d / n_shards] = (len + 3) & ~3ull;
}
__global__ void k_post_ref(const uint64_t* scan_t, const uint32_t* df, uint32_t I, uint64_t m, uint32_t n_shards,
                           uint32_t per_shard, uint2* post_ref) {
  const uint32_t d = blockIdx.x * blockDim.x + threadIdx.x;
  if (d >= I) return;
  const uint32_t sh = d % n_shards;
  const uint64_t rel = scan_t[(uint64_t)sh * per_shard + d / n_shards] - scan_t[(uint64_t)sh * per_shard];
  post_ref[d] = make_uint2((uint32_t)(rel >> 2), (uint32_t)((uint64_t)df[d] < m ? df[d] : m));
}
// an item's segment holds its sessions in rank-ASCENDING order (stable sort of rank-ordered input): the posting list
// is its last min(df, m) entries, most recent first
__global__ void k_fill_postings(const uint32_t* sorted_item, const uint32_t* sorted_rank, uint64_t P, const uint64_t* seg_start,
                                const uint2* post_ref, uint32_t shard, uint32_t n_shards, uint32_t* postings) {
  const uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  const uint32_t d = sorted_item[p];
  if (d % n_shards != shard) return;
  const uint64_t j = seg_start[d + 1] - 1 - p;                   // 0 = the most recent session of the item
  const uint2 ref = post_ref[d];
  if (j < ref.y) postings[(uint64_t)ref.x * 4 + j] = sorted_rank[p];
}
__global__ void k_fill_u8(uint8_t* p, uint64_t n, uint8_t v) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

// ---- synthetic generator (same functions as synth.cpp, see synth_common.h) ----
__global__ void k_synth_len(uint64_t seed, uint64_t n_items, uint64_t S, uint64_t* len) {
  const uint64_t sn = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (sn < S) len[sn] = vmis_synth::session_length_of(seed, sn, n_items);
}
__global__ void k_synth_items(uint64_t seed, uint64_t n_items, double log_n1, uint64_t S, const uint64_t* off, uint64_t* items,
                              uint32_t* ts, vmis_synth::Perm perm) {
  const uint64_t sn = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (sn >= S) return;
  uint64_t tmp[34];
  const uint32_t len = vmis_synth::gen_session(seed, sn, n_items, log_n1, tmp);
  for (uint32_t i = 0; i < len; ++i) {           // external ids, ascending
    const uint64_t v = vmis_synth::external_id(tmp[i]);
    uint32_t j = i;
    while (j > 0 && tmp[j - 1] > v) { tmp[j] = tmp[j - 1]; --j; }
    tmp[j] = v;
  }
  const uint64_t b = off[sn];
  for (uint32_t i = 0; i < len; ++i) items[b + i] = tmp[i];
  ts[sn] = (uint32_t)(vmis_synth::kTsBase + perm(sn));
}

template <class KeyT>
bool radix_sort_keys(KeyT* in, KeyT* out, uint64_t n, int end_bit, cudaStream_t st, std::string* err) {
  size_t tb = 0;
  CU_OK(cub::DeviceRadixSort::SortKeys(nullptr, tb, in, out, (long long)n, 0, end_bit, st));
  DevBuf tmp;
  CU_OK(tmp.alloc(tb));
  CU_OK(cub::DeviceRadixSort::SortKeys(tmp.p, tb, in, out, (long long)n, 0, end_bit, st));
  CU_OK(cudaStreamSynchronize(st));
  return true;
}
bool radix_sort_pairs_u32(uint32_t* k_in, uint32_t* k_out, uint32_t* v_in, uint32_t* v_out, uint64_t n, int end_bit,
                          cudaStream_t st, std::string* err) {
  size_t tb = 0;
  CU_OK(cub::DeviceRadixSort::SortPairs(nullptr, tb, k_in, k_out, v_in, v_out, (long long)n, 0, end_bit, st));
  DevBuf tmp;
  CU_OK(tmp.alloc(tb));
  CU_OK(cub::DeviceRadixSort::SortPairs(tmp.p, tb, k_in, k_out, v_in, v_out, (long long)n, 0, end_bit, st));
  CU_OK(cudaStreamSynchronize(st));
  return true;
}
bool exclusive_scan_u64(const uint64_t* in, uint64_t* out, uint64_t n, cudaStream_t st, std::string* err) {
  size_t tb = 0;
  CU_OK(cub::DeviceScan::ExclusiveSum(nullptr, tb, in, out, (long long)n, st));
  DevBuf tmp;
  CU_OK(tmp.alloc(tb));
  CU_OK(cub::DeviceScan::ExclusiveSum(tmp.p, tb, in, out, (long long)n, st));
  CU_OK(cudaStreamSynchronize(st));
  return true;
}
inline int bits_for(uint64_t v) { int b = 1; while (b < 64 && (v >> b)) ++b; return b; }

}  // namespace

bool synth_sessions_device(uint64_t seed, uint64_t n_items, uint64_t n_sessions, DeviceSessions* out, std::string* err) {
  cudaStream_t st = nullptr;
  DevBuf len, off, items, ts;
  CU_OK(len.alloc((n_sessions + 1) * 8));
  CU_OK(off.alloc((n_sessions + 1) * 8));
  CU_OK(cudaMemsetAsync(len.p, 0, (n_sessions + 1) * 8, st));
  k_synth_len<<<blocks_for(n_sessions), kB, 0, st>>>(seed, n_items, n_sessions, len.as<uint64_t>());
  if (!exclusive_scan_u64(len.as<uint64_t>(), off.as<uint64_t>(), n_sessions + 1, st, err)) return false;
  uint64_t total = 0;
  CU_OK(cudaMemcpy(&total, off.as<uint64_t>() + n_sessions, 8, cudaMemcpyDeviceToHost));
  len.free_now();
  CU_OK(items.alloc(total * 8));
  CU_OK(ts.alloc(n_sessions * 4));
  const vmis_synth::Perm perm(n_sessions, vmis_synth::splitmix(seed ^ 0x7157ull));
  k_synth_items<<<blocks_for(n_sessions), kB, 0, st>>>(seed, n_items, std::log((double)n_items + 1.0), n_sessions,
                                                        off.as<uint64_t>(), items.as<uint64_t>(), ts.as<uint32_t>(), perm);
  CU_OK(cudaGetLastError());
  CU_OK(cudaStreamSynchronize(st));
  out->n_sessions = n_sessions; out->n_entries = total;
  items.release_to((void**)&out->items); off.release_to((void**)&out->off); ts.release_to((void**)&out->ts);
  return true;
}

bool build_index_device(const DeviceSessions& s, uint64_t m, uint64_t max_len, double idf_weighting, uint32_t shard,
                        uint32_t n_shards, DeviceIndexArrays* out, std::string* err) {
  cudaStream_t st = nullptr;
  const uint64_t S = s.n_sessions;
  if (S == 0 || S >= 0xFFFFFFF0ull) { *err = "bad session count"; return false; }
  if (m == 0) { *err = "m must be >= 1"; return false; }
  if (max_len > kMaxLenDevice) { *err = "device build supports max_len <= 128 (use the host builder)"; return false; }
  if (n_shards == 0 || n_shards > (uint32_t)kMaxShards || shard >= n_shards) { *err = "bad shard"; return false; }

  // 1. time-rank the kept sessions
  DevBuf keys, keys2, counters;
  CU_OK(keys.alloc(S * 8)); CU_OK(keys2.alloc(S * 8)); CU_OK(counters.alloc(16));
  CU_OK(cudaMemsetAsync(counters.p, 0, 16, st));
  k_session_keys<<<blocks_for(S), kB, 0, st>>>(s.off, s.ts, S, max_len, keys.as<uint64_t>(),
                                               counters.as<unsigned long long>(), counters.as<unsigned long long>() + 1);
  CU_OK(cudaGetLastError());
  if (!radix_sort_keys(keys.as<uint64_t>(), keys2.as<uint64_t>(), S, 64, st, err)) return false;
  unsigned long long cnt[2];
  CU_OK(cudaMemcpy(cnt, counters.p, 16, cudaMemcpyDeviceToHost));
  const uint64_t Sk = cnt[0], P = cnt[1];
  keys.free_now();
  if (Sk == 0) { *err = "no training session survives max_len"; return false; }

  DevBuf rank_to_orig, elen, plen, estart, pstart;
  CU_OK(rank_to_orig.alloc(Sk * 4)); CU_OK(elen.alloc((Sk + 1) * 8)); CU_OK(plen.alloc((Sk + 1) * 8));
  CU_OK(estart.alloc((Sk + 1) * 8)); CU_OK(pstart.alloc((Sk + 1) * 8));
  CU_OK(cudaMemsetAsync(elen.p, 0, (Sk + 1) * 8, st)); CU_OK(cudaMemsetAsync(plen.p, 0, (Sk + 1) * 8, st));
  k_rank_lengths<<<blocks_for(Sk), kB, 0, st>>>(keys2.as<uint64_t>(), s.off, Sk, rank_to_orig.as<uint32_t>(),
                                                elen.as<uint64_t>(), plen.as<uint64_t>());
  CU_OK(cudaGetLastError());
  if (!exclusive_scan_u64(elen.as<uint64_t>(), estart.as<uint64_t>(), Sk + 1, st, err)) return false;
  if (!exclusive_scan_u64(plen.as<uint64_t>(), pstart.as<uint64_t>(), Sk + 1, st, err)) return false;
  uint64_t Ppad = 0;
  CU_OK(cudaMemcpy(&Ppad, pstart.as<uint64_t>() + Sk, 8, cudaMemcpyDeviceToHost));
  keys2.free_now(); elen.free_now(); plen.free_now();
  if ((Ppad >> 2) > 0xFFFFFFFFull) { *err = "session item array too large"; return false; }

  // 2. item dictionary: distinct external ids through a hash set, then a sort of the distinct ids only
  DevBuf item_key;
  uint32_t I = 0;
  {
    DevBuf set, uniq, uniq_sorted, cnt2;
    CU_OK(cnt2.alloc(32));
    uint64_t cap = 1ull << 22;
    while (cap < P / 8) cap <<= 1;
    for (;;) {
      CU_OK(set.alloc(cap * 8));
      CU_OK(cudaMemsetAsync(set.p, 0xFF, cap * 8, st));
      CU_OK(cudaMemsetAsync(cnt2.p, 0, 32, st));
      k_dedup_insert<<<blocks_for(Sk), kB, 0, st>>>(s.items, s.off, rank_to_orig.as<uint32_t>(), Sk, set.as<unsigned long long>(),
                                                     cap - 1, cnt2.as<unsigned long long>(), cnt2.as<unsigned int>() + 4);
      CU_OK(cudaGetLastError());
      unsigned long long hdr[3];                                     // [0] distinct, [1] collect cursor, [2] flags (2 x u32)
      CU_OK(cudaMemcpy(hdr, cnt2.p, 24, cudaMemcpyDeviceToHost));
      const unsigned int has_max = (unsigned int)(hdr[2] & 0xFFFFFFFFu), overfull = (unsigned int)(hdr[2] >> 32);
      if (overfull) { if (cap >= 4 * P + 1024) { *err = "item dictionary: hash set overflow"; return false; } cap <<= 2; continue; }
      const uint64_t I64 = hdr[0] + (has_max ? 1 : 0);
      if (I64 == 0 || I64 >= 0x7FFFFFFFull) { *err = I64 ? "too many items" : "no items"; return false; }
      I = (uint32_t)I64;
      CU_OK(uniq.alloc((uint64_t)I * 8)); CU_OK(uniq_sorted.alloc((uint64_t)I * 8));
      k_dedup_collect<<<blocks_for(cap), kB, 0, st>>>(set.as<unsigned long long>(), cap, uniq.as<uint64_t>(),
                                                      cnt2.as<unsigned long long>() + 1);
      CU_OK(cudaGetLastError());
      if (has_max) { const unsigned long long mx = kNoKey; CU_OK(cudaMemcpyAsync(uniq.as<uint64_t>() + hdr[0], &mx, 8, cudaMemcpyHostToDevice, st)); }
      set.free_now();
      if (!radix_sort_keys(uniq.as<uint64_t>(), uniq_sorted.as<uint64_t>(), I, 64, st, err)) return false;
      uniq_sorted.release_to(&item_key.p);
      break;
    }
  }

  // 3. device item hash
  uint64_t cap = 16; while (cap < (uint64_t)I * 2) cap <<= 1;
  DevBuf item_hash;
  CU_OK(item_hash.alloc(cap * sizeof(ItemHashEntry)));
  k_hash_clear<<<blocks_for(cap), kB, 0, st>>>(item_hash.as<ItemHashEntry>(), cap);
  k_hash_insert<<<blocks_for(I), kB, 0, st>>>(item_key.as<uint64_t>(), I, item_hash.as<ItemHashEntry>(), (uint32_t)(cap - 1));
  CU_OK(cudaGetLastError());

  // 4. session → items, and the posting sort keys
  DevBuf sess_ref, sess_items, p_item, p_rank, p_item_sorted, p_rank_sorted, eflag;
  CU_OK(sess_ref.alloc(Sk * sizeof(uint2))); CU_OK(sess_items.alloc(Ppad * 4));
  CU_OK(p_item.alloc(P * 4)); CU_OK(p_rank.alloc(P * 4)); CU_OK(eflag.alloc(4));
  CU_OK(cudaMemsetAsync(eflag.p, 0, 4, st));
  k_emit_sessions<<<blocks_for(Sk), kB, 0, st>>>(s.items, s.off, rank_to_orig.as<uint32_t>(), estart.as<uint64_t>(),
                                                 pstart.as<uint64_t>(), Sk, item_hash.as<ItemHashEntry>(), (uint32_t)(cap - 1),
                                                 sess_ref.as<uint2>(), sess_items.as<uint32_t>(), p_item.as<uint32_t>(),
                                                 p_rank.as<uint32_t>(), eflag.as<unsigned int>());
  CU_OK(cudaGetLastError());
  unsigned int ef = 0;
  CU_OK(cudaMemcpy(&ef, eflag.p, 4, cudaMemcpyDeviceToHost));
  if (ef) { *err = "duplicate item inside a training session"; return false; }
  estart.free_now(); pstart.free_now();

  // 5. item → sessions: the (item, rank) pairs were emitted in rank order, so a STABLE radix sort on the item bits
  //    alone (3 digit passes for 21-bit dense ids instead of 7 over 53-bit composite keys) leaves every item's
  //    sessions in rank order; truncate to m, shard by item
  CU_OK(p_item_sorted.alloc(P * 4)); CU_OK(p_rank_sorted.alloc(P * 4));
  if (!radix_sort_pairs_u32(p_item.as<uint32_t>(), p_item_sorted.as<uint32_t>(), p_rank.as<uint32_t>(), p_rank_sorted.as<uint32_t>(),
                            P, bits_for(I), st, err)) return false;
  p_item.free_now(); p_rank.free_now();
  DevBuf seg_start, df, padded_t, scan_t, post_ref, postings;
  const uint32_t per_shard = (I + n_shards - 1) / n_shards;
  const uint64_t nt = (uint64_t)per_shard * n_shards + 1;
  CU_OK(seg_start.alloc(((uint64_t)I + 1) * 8)); CU_OK(df.alloc((uint64_t)I * 4));
  CU_OK(padded_t.alloc(nt * 8)); CU_OK(scan_t.alloc(nt * 8)); CU_OK(post_ref.alloc((uint64_t)I * sizeof(uint2)));
  CU_OK(cudaMemsetAsync(padded_t.p, 0, nt * 8, st));
  k_seg_starts<<<blocks_for(P + 1), kB, 0, st>>>(p_item_sorted.as<uint32_t>(), P, I, seg_start.as<uint64_t>());
  k_post_len<<<blocks_for(I), kB, 0, st>>>(seg_start.as<uint64_t>(), I, m, n_shards, per_shard, df.as<uint32_t>(),
                                           padded_t.as<uint64_t>());
  CU_OK(cudaGetLastError());
  if (!exclusive_scan_u64(padded_t.as<uint64_t>(), scan_t.as<uint64_t>(), nt, st, err)) return false;
  k_post_ref<<<blocks_for(I), kB, 0, st>>>(scan_t.as<uint64_t>(), df.as<uint32_t>(), I, m, n_shards, per_shard,
                                           post_ref.as<uint2>());
  CU_OK(cudaGetLastError());
  uint64_t sh_lo = 0, sh_hi = 0;
  CU_OK(cudaMemcpy(&sh_lo, scan_t.as<uint64_t>() + (uint64_t)shard * per_shard, 8, cudaMemcpyDeviceToHost));
  CU_OK(cudaMemcpy(&sh_hi, scan_t.as<uint64_t>() + (uint64_t)(shard + 1) * per_shard, 8, cudaMemcpyDeviceToHost));
  const uint64_t shard_entries = sh_hi - sh_lo;
  if ((shard_entries >> 2) > 0xFFFFFFFFull) { *err = "posting array too large"; return false; }
  CU_OK(postings.alloc(shard_entries * 4));
  CU_OK(cudaMemsetAsync(postings.p, 0xFF, std::max<uint64_t>(shard_entries * 4, 16), st));
  k_fill_postings<<<blocks_for(P), kB, 0, st>>>(p_item_sorted.as<uint32_t>(), p_rank_sorted.as<uint32_t>(), P, seg_start.as<uint64_t>(),
                                                post_ref.as<uint2>(), shard, n_shards, postings.as<uint32_t>());
  CU_OK(cudaGetLastError());
  CU_OK(cudaStreamSynchronize(st));
  p_item_sorted.free_now(); p_rank_sorted.free_now(); seg_start.free_now(); padded_t.free_now(); scan_t.free_now();

  // 6. idf on the host (same libm as the oracle), attributes, host copies for the accessors
  std::vector<uint32_t> h_df(I);
  CU_OK(cudaMemcpy(h_df.data(), df.p, (uint64_t)I * 4, cudaMemcpyDeviceToHost));
  df.free_now();
  out->host_idf.resize(I);
  uint64_t n_post = 0;
  {
    // ln on the host (the oracle's libm), spread over the cores: 1.7 M logarithms are 15 ms on one thread
    const unsigned nt = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
    std::vector<uint64_t> part(nt, 0);
    std::vector<std::thread> th;
    for (unsigned t = 0; t < nt; ++t) th.emplace_back([&, t]() {
      const uint32_t lo = (uint32_t)((uint64_t)I * t / nt), hi = (uint32_t)((uint64_t)I * (t + 1) / nt);
      uint64_t np = 0;
      for (uint32_t d = lo; d < hi; ++d) {
        out->host_idf[d] = std::log((double)P / (double)h_df[d]) * idf_weighting;      // vmis_index.rs:509-513
        np += std::min<uint64_t>(h_df[d], m);
      }
      part[t] = np; });
    for (auto& x : th) x.join();
    for (uint64_t v : part) n_post += v;
  }
  DevBuf idf, attr;
  CU_OK(idf.alloc((uint64_t)I * 8)); CU_OK(attr.alloc(I));
  CU_OK(cudaMemcpy(idf.p, out->host_idf.data(), (uint64_t)I * 8, cudaMemcpyHostToDevice));
  k_fill_u8<<<blocks_for(I), kB, 0, st>>>(attr.as<uint8_t>(), I, (uint8_t)(VMIS_ATTR_EXISTS | VMIS_ATTR_FOR_SALE));
  CU_OK(cudaGetLastError());
  out->host_item_key.resize(I);                 // the sorted dictionary doubles as the host-side lookup structure
  CU_OK(cudaMemcpy(out->host_item_key.data(), item_key.p, (uint64_t)I * 8, cudaMemcpyDeviceToHost));
  CU_OK(cudaStreamSynchronize(st));

  out->n_items = I; out->n_kept = Sk; out->n_pairs_kept = P; out->n_postings = n_post;
  out->item_hash_cap = cap; out->shard_entries = shard_entries; out->sess_items_entries = Ppad;
  item_key.release_to((void**)&out->item_key); item_hash.release_to((void**)&out->item_hash);
  post_ref.release_to((void**)&out->post_ref); postings.release_to((void**)&out->postings);
  sess_ref.release_to((void**)&out->sess_ref); sess_items.release_to((void**)&out->sess_items);
  idf.release_to((void**)&out->idf); attr.release_to((void**)&out->attr);
  rank_to_orig.release_to((void**)&out->rank_to_orig);
  return true;
}

}  // namespace vmis
