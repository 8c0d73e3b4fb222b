// serenade_b200/csrc/vmis_device.h — device-side view of the VMIS index and the
// launch interface of the sm_100a predict kernel.  Internal header (the public
// boundary is include/vmis.h).
#pragma once
#include <cstdint>
#include <string>
#include <cuda_runtime.h>

#include "../../include/vmis.h"

namespace vmis {

constexpr uint32_t kEmpty = 0xFFFFFFFFu;
#ifndef VMIS_THREADS
#define VMIS_THREADS 256
#endif
#ifndef VMIS_CTAS
#define VMIS_CTAS 5
#endif
constexpr int kThreads = VMIS_THREADS;   // one CTA per evolving session
constexpr int kCtasPerSm = VMIS_CTAS;    // resident CTAs per SM the kernel is compiled (register-bounded) for
constexpr int kWarps = kThreads / 32;
constexpr int kMaxSessionLen = VMIS_MAX_SESSION_LEN;   // evolving-session length limit of the kernel (reference HPO grid: <= 100)
constexpr uint32_t kMaxK = 2048;         // keeps the int32 item numerators exact
constexpr uint32_t kMaxM = 8192;         // shared-memory bound of the m-sample buffers
constexpr int kMaxShards = 8;            // item-sharded postings: one shard per GPU of a box

// ext item id -> dense item index; open addressing, val == kEmpty marks a free slot
struct alignas(16) ItemHashEntry {
  uint64_t key;
  uint32_t val;
  uint32_t pad;
};

// HBM layout (all arrays immutable after build):
//   item dictionary   item_key[I] ascending external ids; dense index = rank, so
//                     (score desc, dense idx asc) == (score desc, item_id asc)
//   postings          (optionally sharded by item over the GPUs of a box, see post_shard)
//                     per item: kept-session TIME RANKS, descending (most recent first),
//                     truncated to m_build, list start aligned to 16 B
//                     (replaces item_to_top_sessions_ordered + the session_to_max_time_stamp
//                     gather of vmis_index.rs:359: rank order == (ts, session idx) order)
//   sess_items        per kept session (by time rank): dense item indices ascending,
//                     list start aligned to 16 B (replaces session_to_items_sorted)
//   idf[I], attr[I]   item_to_idf_score / item_to_product_attributes; g32[I] fp32 image of the idf weight
//   rank_to_orig[Sk]  time rank -> reference session index (find_neighbors output only)
struct IndexView {
  const uint64_t* item_key;
  const ItemHashEntry* item_hash;
  uint32_t item_hash_mask;
  const uint2* post_ref;      // {offset in units of 4 entries, length}
  const uint32_t* post_shard[kMaxShards];   // posting arrays; shard s holds the lists of items with dense idx % n_shards == s
  uint32_t n_shards;                        // 1 = everything local; >1 = peers mapped over NVLink (CUDA IPC)
  const uint2* sess_ref;      // {offset in units of 4 entries, length}
  const uint32_t* sess_items;
  const double* idf;
  const uint8_t* attr;
  const uint32_t* rank_to_orig;
  uint32_t n_items;
  uint32_t n_kept;
  uint32_t m_build;           // longest posting list (staging capacity)
  uint32_t m_carry;           // m <= m_carry: every session of the m-sample is on the list of each evolving item it holds
  uint32_t max_len;
  const float* g32;           // g(idf[i]) = (idf > 0 ? (float)idf : 1) as fp32: the coarse pass of the top-n selection reads 4 bytes per item
};

struct PredictArgs {
  const uint64_t* q_items;    // device
  const uint32_t* q_off;      // device, n_q + 1
  uint32_t q_item_base;       // q_off values are relative to q_items - q_item_base (chunks of a larger CSR batch)
  uint32_t n_q;
  uint32_t k, m, how_many;
  int biz;
  int host_flags;             // predict mode, outputs in mapped host memory: out_counts[q] is written last, after a system-wide fence
  // outputs (device); predict mode
  uint64_t* out_ids;
  double* out_scores;
  uint32_t* out_counts;
  vmis_query_stats_t* out_stats;  // optional
  // find_neighbors mode (out_sess != nullptr): rows of k
  uint32_t* out_sess;
  double* out_sim;
};

// Persistent-CTA workspace owned by the caller: a work counter plus one global
// overflow score table per resident CTA.  Invariant: between launches every table
// slot is {kEmpty, 0} (init_workspace() establishes it, the kernel restores it).
struct Workspace {
  uint32_t* counter;          // [0] work counter, [1] exit counter, [2] lock of the huge table; zero between launches
  unsigned long long* gtab;   // grid × gtab_cap score-table slots {key : 32 | value : 32}, all ones = empty
  uint32_t* gtab_occ;         // grid × gtab_cap / 2: occupied-slot lists
  unsigned long long* huge;   // ONE table of gtab_huge slots for the rare query that outgrows its CTA's table (taken
  uint32_t* huge_occ;         // under the lock); its occupied-slot list
  uint32_t gtab_cap;          // per-CTA overflow table: power of two, min(2 * k * max_len, 2^15) slots
  uint32_t gtab_huge;         // 0, or the power of two >= 2 * k * max_len when that exceeds gtab_cap
  uint32_t grid;
};

struct LaunchPlan {
  uint32_t grid;
  uint32_t smem_bytes;
  uint32_t tab_cap;           // shared score table slots (power of two)
  uint32_t occ_cap;           // occupancy budget of the shared table (distinct items)
  uint32_t gran_cap;          // capacity of the flat granule -> neighbour map (16-byte granules of the neighbours' item lists)
  uint32_t m_eff;             // acc buffer capacity
  uint32_t list_cap;          // posting staging capacity
  uint32_t gtab_cap;
  uint32_t gtab_huge;
};

// Computes launch geometry; returns VMIS_OK or VMIS_ERR_LIMIT (then *why names the limit that was hit and the remedy).
int plan_launch(const IndexView& ix, uint32_t k, uint32_t m, int sm_count, LaunchPlan* plan, std::string* why);
size_t workspace_bytes(const LaunchPlan& plan);
// carve a raw device allocation of workspace_bytes() into a Workspace
Workspace carve_workspace(void* base, const LaunchPlan& plan);
// One-time initialisation of a freshly carved workspace (enqueued on `stream`).
cudaError_t init_workspace(const Workspace& ws, cudaStream_t stream);
// Enqueues the predict kernel on `stream` (the workspace must have been initialised once).
cudaError_t launch_predict(const IndexView& ix, const PredictArgs& args, const LaunchPlan& plan, const Workspace& ws,
                           cudaStream_t stream);
// g32[i] = idf[i] > 0 ? (float)idf[i] : 1 (IndexView::g32), computed on the device from the uploaded idf array
cudaError_t build_g32(const double* idf, float* g32, uint32_t n_items, cudaStream_t stream);
// number of kernels launch_predict enqueues (bench "gpu_launches")
constexpr int kLaunchesPerBatch = 1;

}  // namespace vmis
