// serenade_b200/csrc/vmis_host.h — host-side index: the mirror that serves the
// SimilarityComputationNew accessors (similarity_indexed.rs:8-24) and the flat
// CSR arrays that are uploaded to HBM.  Internal header.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "vmis_device.h"

namespace vmis {

// Training sessions exactly as the reference holds them after read_from_file /
// Avro load: `session_to_items_sorted` (un-pruned, vmis_index.rs:79) and
// `session_to_max_time_stamp` (vmis_index.rs:77).
struct Sessions {
  std::vector<uint64_t> items;     // concatenated item ids
  std::vector<uint64_t> off;       // n_sessions + 1
  std::vector<uint32_t> ts;        // n_sessions
  size_t size() const { return ts.size(); }
};

// Flat arrays in the exact layout of IndexView (vmis_device.h).
struct FlatIndex {
  std::vector<uint64_t> item_key;
  std::vector<ItemHashEntry> item_hash;
  std::vector<uint2> post_ref;
  std::vector<uint32_t> postings;          // all shards back to back; shard s starts at shard_begin[s]
  std::vector<uint64_t> shard_begin;       // n_shards + 1 offsets into postings (post_ref offsets are shard relative)
  uint32_t n_shards = 1;
  std::vector<uint2> sess_ref;
  std::vector<uint32_t> sess_items;
  std::vector<double> idf;
  std::vector<uint8_t> attr;
  std::vector<uint32_t> rank_to_orig;
  uint64_t n_pairs_kept = 0;
  uint64_t n_postings = 0;
  uint32_t m_build = 0;
  uint32_t m_carry = 0;                    // largest m for which first-match positions come from the merged lists
  uint32_t max_len = 0;
  double idf_weighting = 0;
};

// read_from_file (vmis_index.rs:591-752).  Returns false and sets err on I/O failure.
bool read_sessions_from_csv(const std::string& path, Sessions* out, std::string* err);

// qty_events_p99_5 (vmis_index.rs:693-716): t-digest (size 100) estimate of the 99.5th
// percentile of session lengths, rounded.
size_t session_length_p99_5(const Sessions& s);

// prepare_hashmap (vmis_index.rs:422-528) → flat CSR arrays.
bool build_flat_index(const Sessions& s, size_t m, size_t max_len, double idf_weighting, uint32_t n_shards,
                      FlatIndex* out, std::string* err);

// vmis_last_error() / vmis_last_error_code() of the calling thread (capi.cu); code VMIS_OK clears them
void set_last_error(int code, const char* msg);

// dense index of an external item id, kEmpty if unknown
uint32_t host_lookup_item(const FlatIndex& f, uint64_t item);

}  // namespace vmis
