// serenade_b200/csrc/build_device.h — interface of the on-device index build and synthetic generator
// (build_sm100.cu).  Internal header.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "vmis_device.h"

namespace vmis {

// training sessions resident in HBM: items[off[s] .. off[s+1]) are the external item ids of session s
struct DeviceSessions {
  const uint64_t* items = nullptr;
  const uint64_t* off = nullptr;      // n_sessions + 1
  const uint32_t* ts = nullptr;
  uint64_t n_sessions = 0;
  uint64_t n_entries = 0;
};

// device arrays of a freshly built index (cudaMalloc'ed, ownership passes to the caller) + host copies that the
// trait accessors need
struct DeviceIndexArrays {
  uint64_t* item_key = nullptr; ItemHashEntry* item_hash = nullptr; uint64_t item_hash_cap = 0;
  uint2* post_ref = nullptr; uint32_t* postings = nullptr; uint64_t shard_entries = 0;
  uint2* sess_ref = nullptr; uint32_t* sess_items = nullptr; uint64_t sess_items_entries = 0;
  double* idf = nullptr; uint8_t* attr = nullptr; uint32_t* rank_to_orig = nullptr;
  uint64_t n_items = 0, n_kept = 0, n_pairs_kept = 0, n_postings = 0;
  std::vector<uint64_t> host_item_key; std::vector<ItemHashEntry> host_item_hash; std::vector<double> host_idf;
};

// prepare_hashmap (vmis_index.rs:422-528) on the device; max_len must be explicit (<= 128)
bool build_index_device(const DeviceSessions& s, uint64_t m, uint64_t max_len, double idf_weighting, uint32_t shard,
                        uint32_t n_shards, DeviceIndexArrays* out, std::string* err);

// the generator of synth.cpp, run on the device (arrays cudaMalloc'ed, owned by the caller)
bool synth_sessions_device(uint64_t seed, uint64_t n_items, uint64_t n_sessions, DeviceSessions* out, std::string* err);

}  // namespace vmis
