// serenade_b200/csrc/avro_reader.h — reader for the production on-disk format of the VMIS index:
// a directory with itemindex/*.avro and sessionindex/*.avro (VMISIndex::new, vmis_index.rs:85-313; record
// layouts at :184-192 and :249-255).  Host only, no third-party Avro library: object-container framing, the
// null / deflate / snappy block codecs and schema-driven decoding are implemented in avro_reader.cpp.
// Internal header.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "vmis_host.h"

namespace vmis {

// A pre-computed index exactly as VMISIndex::new holds it after the Avro load: posting lists, idf and
// attributes per item are TAKEN from the files, not derived from the sessions.
struct PrebuiltIndex {
  std::vector<uint64_t> item_ids;        // ItemId                          (:187)
  std::vector<uint64_t> post_off;        // n_items + 1
  std::vector<uint32_t> post_sessions;   // session_indices_time_ordered    (:188), SessionIndex values
  std::vector<double> idf;               // idf                             (:189)
  std::vector<uint8_t> attr;             // VMIS_ATTR_* from ForSale/IsAdult (:190-191), EXISTS always set
  Sessions sessions;                     // dense by SessionIndex (:252); unused positions are empty with ts 0
};

struct AvroLoadInfo {
  uint64_t item_files = 0, session_files = 0, item_records = 0, session_records = 0;
};

// Reads <base_path>/itemindex/*.avro and <base_path>/sessionindex/*.avro (files in name order; a later
// record of the same ItemId / SessionIndex replaces an earlier one like the HashMap / Vec stores of
// :214-226 and :289-290).  Returns false and sets err on any I/O, framing, codec or schema problem.
bool read_index_from_avro(const std::string& base_path, PrebuiltIndex* out, AvroLoadInfo* info, std::string* err);

// The reverse direction: writes `p` as <base_path>/itemindex/part-NNNNN.avro and <base_path>/sessionindex/part-NNNNN.avro
// with the record layouts of vmis_index.rs:184-192 / :249-255 (what the reference's offline Spark job produces), so an
// index built on the GPU can be served by the reference itself.  codec: "null" or "deflate".  Sessions without items
// are not written.
bool write_index_to_avro(const std::string& base_path, const PrebuiltIndex& p, const std::string& codec, size_t n_files,
                         std::string* err);

// Outcome of checking / normalising a pre-computed index (build_flat_index_prebuilt).
struct PrebuiltInfo {
  uint64_t lists_reordered = 0;    // posting lists that were not in (timestamp desc, session idx desc) order
  uint64_t duplicate_postings = 0; // repeated session ids dropped from posting lists
  uint32_t m_carry = 0;            // largest m for which the first-match position can be carried from the lists
};

// Pre-computed parts → flat CSR arrays of vmis_device.h.  Fails (err) if a posting names a session that is
// missing or does not contain the item, or if a session holds an item without an itemindex record — the
// reference would panic on such data at query time (mod.rs:138, vmis_index.rs:322).
bool build_flat_index_prebuilt(const PrebuiltIndex& p, uint32_t n_shards, FlatIndex* out, PrebuiltInfo* info,
                               std::string* err);

}  // namespace vmis
