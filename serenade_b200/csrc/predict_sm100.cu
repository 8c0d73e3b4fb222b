// serenade_b200/csrc/predict_sm100.cu — the VMIS-kNN predict_next kernel for sm_100a.
//
// Persistent CTAs (5 per SM, 256 threads, ~44 KB shared memory) run one query at a time, end to end in shared memory:
// CTA b starts with query b, every further one comes from a global work counter (the next two indices are always in
// flight):
//
//   phase 0  de-duplicate the evolving session (vmis_index.rs:335-348), translate external item ids through the
//            HBM item hash.  For sessions of <= 32 items the LAST WARP does this for the NEXT query while the other
//            warps insert (phase 2b hands its rounds out dynamically, so nobody waits for it)
//   phase 1  m-sample: the time-descending posting lists of the distinct items (local HBM, or a peer GPU's HBM over
//            NVLink when the index is item-sharded) are streamed into shared memory by TMA bulk copies
//            (cp.async.bulk + mbarrier; list j+1 is requested when the fold of list j ends) and folded one by one with
//            a block-wide merge-path merge that de-duplicates, sums the integer similarity numerators, keeps the first-match position and
//            truncates to the m most recent sessions (closed form of the heap procedure of vmis_index.rs:344-391)
//   phase 1b top-k neighbours by (numerator desc, recency desc): packed / ballot histogram or binary search for the
//            threshold numerator, one ordered scan for the ties (vmis_index.rs:394-412)
//   phase 2a neighbour directory: item-list refs, integer weight 10·linear_score·numerator (mod.rs:133-142,
//            :110-116), granule -> neighbour map (a granule = 16 bytes = up to four items of one item list)
//   phase 2b A[item] += weight for every item of every neighbour (mod.rs:144-153) into a 4096-slot shared table of
//            64-bit slots {item | A}: a claim is ONE ATOMS.CAS.64, a hit an add to the low word; rounds of 32
//            granules per warp from a block-wide counter; the four first probes of a lane are in flight together;
//            collided items are re-dealt one per lane through a staging row.  Rare overflow -> a table in HBM
//   phase 3  drop the current item (mod.rs:157-160), business rules (:162-182), top-n by (score desc, item id asc)
//            (mod.rs:185-214) straight from the table: fp32 coarse keys for all slots, a block-wide lower bound
//            from the best thread maxima (REDUX), the few survivors rescored exactly (f64 g(idf)·A/(10·u)) by the
//            threads that found them and ranked by counting across all warps; a margin test / exact 96-bit
//            network covers heavy score ties
//
// All session/item arithmetic is integer and order independent; the only floating point that reaches the
// output is one f64 multiply + divide per final candidate, so results are bit-exact against the canonical mode
// of the CPU checker whatever the thread scheduling, batch order or index sharding.
#include "vmis_device.h"

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <string>

namespace vmis {
namespace {

constexpr unsigned kFull = 0xFFFFFFFFu;
constexpr unsigned long long kMaxWorkspaceBytes = 8ull << 30;   // overflow tables of one call context
#ifndef VMIS_VT
#define VMIS_VT 6
#endif
#ifndef VMIS_STAGES
#define VMIS_STAGES 1
#endif
#ifndef VMIS_SELQ
#define VMIS_SELQ 1024
#endif
constexpr int kStages = VMIS_STAGES;         // TMA staging buffers of phase 1.  1: list j+1 is requested when the fold of list j ends; 2: one list earlier.
                                             // Same-box A/B: 1 is +1.5 % (25.73 vs 25.35 M qps) — 6 KB less shared memory per CTA beats the hidden copy
constexpr int kVT = VMIS_VT;                 // merge-path items per thread per tile (A/B on B200: 4: 24.96, 5: 25.1, 6: 26.0, 7: 25.8, 8: 24.2, 9: 25.6, 11: 24.6 M qps — 8 is a bank-conflict pothole)
constexpr int kTile = kThreads * kVT;

struct Elem {          // top-n candidate: order-preserving score bits + dense item idx
  uint64_t s;
  uint32_t id;
};
__device__ __forceinline__ bool better(const Elem& a, const Elem& b) {
  return a.s > b.s || (a.s == b.s && a.id < b.id);
}
__device__ __forceinline__ uint64_t score_bits(double x) {
  uint64_t b = (uint64_t)__double_as_longlong(x);
  return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double bits_score(uint64_t s) {
  uint64_t b = (s >> 63) ? (s & 0x7FFFFFFFFFFFFFFFull) : ~s;
  return __longlong_as_double((long long)b);
}
__device__ __forceinline__ Elem shfl_elem(const Elem& e, int src) {
  Elem r;
  r.s = __shfl_sync(kFull, e.s, src);
  r.id = __shfl_sync(kFull, e.id, src);
  return r;
}
__device__ __forceinline__ Elem shfl_xor_elem(const Elem& e, int mask) {
  Elem r;
  r.s = __shfl_xor_sync(kFull, e.s, mask);
  r.id = __shfl_xor_sync(kFull, e.id, mask);
  return r;
}
// one compare-exchange stage of a descending bitonic network across the 32 lanes
__device__ __forceinline__ Elem bitonic_step(const Elem& mine, int lane, int j, bool up) {
  Elem other = shfl_xor_elem(mine, j);
  bool want_better = (((lane & j) == 0) == up);
  bool mine_better = better(mine, other);
  return (want_better == mine_better) ? mine : other;
}
__device__ __forceinline__ Elem warp_sort_desc(Elem e, int lane) {
#pragma unroll
  for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) e = bitonic_step(e, lane, j, (lane & k) == 0);
  }
  return e;
}
// top (sorted desc over lanes) ← best 32 of top ∪ cand (cand sorted desc over lanes)
__device__ __forceinline__ Elem warp_merge_top(const Elem& top, const Elem& cand_sorted, int lane) {
  Elem rev = shfl_elem(cand_sorted, 31 - lane);
  Elem e = better(top, rev) ? top : rev;
#pragma unroll
  for (int j = 16; j > 0; j >>= 1) e = bitonic_step(e, lane, j, true);
  return e;
}

__device__ __forceinline__ int warp_incl_scan(int v, int lane) {
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    int n = __shfl_up_sync(kFull, v, d);
    if (lane >= d) v += n;
  }
  return v;
}
// Inclusive warp scan of SMALL values (v < 2^kBits) without a shuffle chain: one ballot per bit, popcounts of the lanes at
// or below — kBits independent votes instead of five dependent shuffles (merge tiles and the re-deal of collided
// items: +1.0 %, same-box A/B).
template <int kBits>
__device__ __forceinline__ int warp_incl_scan_small(int v, int lane, int& warp_total) {
  const uint32_t le = kFull >> (31 - lane);
  int inc = 0, tot = 0;
#pragma unroll
  for (int b = 0; b < kBits; ++b) {
    const uint32_t m = __ballot_sync(kFull, (v >> b) & 1);
    inc += __popc(m & le) << b;
    tot += __popc(m) << b;
  }
  warp_total = tot;
  return inc;
}
// Block-wide exclusive scan with ONE barrier: warp totals go to a double-buffered scratch row and every warp adds up
// the totals of the warps before it (two 16-byte broadcast reads, no second shuffle scan).  `par` alternates the row;
// a row is only rewritten two calls later, i.e. after another barrier, so no trailing barrier is needed.  All threads
// must call.
struct ScanScratch { alignas(16) int row[2][kWarps]; };
__device__ __forceinline__ int block_excl_scan(int v, ScanScratch& sc, uint32_t& par, int& total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int* row = sc.row[par & 1u];
  par ^= 1u;
  const int inc = warp_incl_scan(v, lane);
  if (lane == 31) row[warp] = inc;
  __syncthreads();
  int before = 0, all = 0;
  if (kWarps == 8) {
    const int4 a = *reinterpret_cast<const int4*>(row), b = *reinterpret_cast<const int4*>(row + 4);
    const int t[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
    for (int w = 0; w < 8; ++w) { all += t[w]; if (w < warp) before += t[w]; }
  } else {
    for (int w = 0; w < kWarps; ++w) { const int t = row[w]; all += t; if (w < warp) before += t; }
  }
  total = all;
  return before + inc - v;
}
// the same for small values (v < 2^kBits): ballot scan inside the warps
template <int kBits>
__device__ __forceinline__ int block_excl_scan_small(int v, ScanScratch& sc, uint32_t& par, int& total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int* row = sc.row[par & 1u];
  par ^= 1u;
  int wt;
  const int inc = warp_incl_scan_small<kBits>(v, lane, wt);
  if (lane == 0) row[warp] = wt;
  __syncthreads();
  int before = 0, all = 0;
  if (kWarps == 8) {
    const int4 a = *reinterpret_cast<const int4*>(row), b = *reinterpret_cast<const int4*>(row + 4);
    const int t[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
    for (int w = 0; w < 8; ++w) { all += t[w]; if (w < warp) before += t[w]; }
  } else {
    for (int w = 0; w < kWarps; ++w) { const int t = row[w]; all += t; if (w < warp) before += t; }
  }
  total = all;
  return before + inc - v;
}
__device__ __forceinline__ int block_sum(int v, ScanScratch& sc, uint32_t& par) {
  int total;
  block_excl_scan(v, sc, par, total);
  return total;
}

__device__ __forceinline__ uint32_t hash_u64(uint64_t x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
  return (uint32_t)x;
}
__device__ __forceinline__ uint32_t lookup_item(const IndexView& ix, uint64_t item) {
  uint32_t h = hash_u64(item) & ix.item_hash_mask;
  for (;;) {
    ItemHashEntry e = ix.item_hash[h];
    if (e.val == kEmpty) return kEmpty;
    if (e.key == item) return e.val;
    h = (h + 1) & ix.item_hash_mask;
  }
}

// rules of mod.rs:162-182 on packed attribute bytes
__device__ __forceinline__ bool passes_business_rules(uint32_t cur, uint32_t reco) {
  if (!(reco & VMIS_ATTR_EXISTS)) return false;
  if (reco & VMIS_ATTR_FOR_SALE) {
    if (reco & VMIS_ATTR_ADULT) return (cur & VMIS_ATTR_EXISTS) && (cur & VMIS_ATTR_ADULT);
    return true;
  }
  return false;
}

// ---- TMA (bulk async copy) + mbarrier helpers: posting lists are contiguous, 16-byte aligned rows in HBM,
// so one elected thread streams a whole list into shared memory with a single cp.async.bulk while the
// block merges the previous list.
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  } while (!done);
}
__device__ __forceinline__ void bulk_load(void* dst_smem, const void* src_gmem, uint32_t bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
// L2 prefetch of a neighbour's session ref as soon as the neighbour is known (A/B: +1.2 %)
#define PF_SESSREF(p) prefetch_l2(p)
__device__ __forceinline__ void fence_async_proxy() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

#ifdef VMIS_PHASE_CLOCKS
#define VMIS_CLK_FIELDS long long clk[16]; uint32_t nclk;
#define VMIS_CLK_RESET(S) do { (S).nclk = 0; } while (0)
#define VMIS_CLK(S) do { if (threadIdx.x == 0 && (S).nclk < 16) (S).clk[(S).nclk++] = clock64(); } while (0)
#define VMIS_CLK_AFTER(S, v) do { asm volatile("" ::"r"(v)); VMIS_CLK(S); } while (0)   // probe once the value has arrived
#else
#define VMIS_CLK_AFTER(S, v) do {} while (0)
#define VMIS_CLK_FIELDS
#define VMIS_CLK_RESET(S) do {} while (0)
#define VMIS_CLK(S) do {} while (0)
#endif
constexpr uint32_t kRefCache = 8;     // items whose post_ref is fetched ahead of phase 1 (sessions of up to 8 distinct items)
struct SmemLayout {
  VMIS_CLK_FIELDS
  // fixed part
  uint32_t d_idx[kMaxSessionLen];    // distinct known items, most recent first
  uint8_t d_pos[kMaxSessionLen];
  uint2 d_ref[kRefCache];            // posting-list refs of the first distinct items, loaded while the previous query inserts
  ScanScratch scan;
  uint32_t nd;
  // the query this CTA runs next (double buffered): thread 0 publishes `q` while the current query is in phase 1;
  // during phase 2b the last warp runs its phase 0 (short sessions) and leaves nd / u / L here with ok = 1
  struct Next { uint32_t q, ok, nd, u, L; } nx[2];
  uint32_t n_occ;                    // occupied score-table slots of the current query
  uint32_t round;                    // phase 2b: next round of 32 granules to hand to a warp
  uint32_t overflow;                 // shared table over its occupancy budget → redo on the global table
  uint32_t sel_ok, sel_count;
  uint32_t bound[kWarps];            // per-warp lower bounds of the n-th best coarse score
  alignas(16) uint32_t top4[kWarps][4];   // phase 3: the four best thread candidates of every warp
  uint32_t qcount;                   // phase 3: entries in the block-wide candidate queue
  unsigned long long bar[2];         // mbarriers of the two posting-list staging buffers (TMA bulk copies)
  unsigned long long stage[kWarps][32];   // phase 2b: per-warp staging of the items whose first probe collided
};
// Scratch that aliases the neighbour arrays (dead or not yet written when it is live): the per-warp numerator
// histogram of phase 1b and the cross-warp candidate buffers of phase 3.
union Scratch {
  uint32_t hist[kWarps][32];
  struct { uint64_t s[kWarps * 32]; uint32_t id[kWarps * 32]; } top;      // exact path: per-warp top-32 lists
  struct { uint64_t s[32]; uint32_t id[32]; } ex;                         // phase 3: exact elements of the first 32 queue entries
};
constexpr uint32_t kSelQ = VMIS_SELQ;     // block-wide queue of top-n candidates (phase 3); more: exact scan

// bytes of the neighbour arrays (3 x K + 1 words), never smaller than the scratch that aliases them
__host__ __device__ constexpr size_t nbr_bytes_min() { return sizeof(Scratch); }
__host__ __device__ inline size_t nbr_bytes(uint32_t k) {
  size_t b = (size_t(k) * 3 + 1) * 4;
  if (b < nbr_bytes_min()) b = nbr_bytes_min();
  return (b + 15) & ~size_t(15);
}
// bytes of the granule -> neighbour map, never smaller than the per-warp candidate queues of phase 3 that alias it
__host__ __device__ inline size_t gran_bytes(uint32_t gran_cap) {
  size_t b = size_t(gran_cap) * 2;
  if (b < size_t(kSelQ) * 4) b = size_t(kSelQ) * 4;
  return (b + 15) & ~size_t(15);
}

// start of an item's posting list: local HBM, or a peer GPU's HBM over NVLink when the index is item-sharded
__device__ __forceinline__ const uint32_t* posting_list(const IndexView& ix, uint32_t item_idx, uint32_t off4) {
  const uint32_t s = ix.n_shards > 1 ? item_idx % ix.n_shards : 0u;
  return ix.post_shard[s] + (size_t)off4 * 4;
}

struct QueryCtx {
  uint32_t q, u, last_idx, cur_attr;
};

constexpr uint32_t kNumMask = 0x00FFFFFFu;   // low word of an m-sample entry: [first-match pos : 8 | numerator : 24]
// 10 x linear_score(pos + 1) (mod.rs:110-116,140-142) times the similarity numerator
__device__ __forceinline__ int32_t session_weight10(uint32_t low) {
  const uint32_t p = (low >> 24) + 1;
  const int32_t w10 = p < 100 ? 10 - (int32_t)p : 0;
  return w10 * (int32_t)(low & kNumMask);
}

// ---- score table: open addressing, one 64-bit slot {item : 32 | A : 32} per item, all ones = empty.  A claim is ONE
// 64-bit compare-and-swap that deposits key and first weight together; a later hit adds to the low word.
typedef unsigned long long Slot;
constexpr Slot kEmptySlot = ~0ull;
__device__ __forceinline__ uint32_t slot_key(Slot s) { return (uint32_t)(s >> 32); }
__device__ __forceinline__ int32_t slot_val(Slot s) { return (int32_t)(uint32_t)s; }

__device__ __forceinline__ Slot make_slot(uint32_t item, int32_t w) { return ((Slot)item << 32) | (uint32_t)w; }
// power-of-two capacity
__device__ __forceinline__ uint32_t home_slot(uint32_t item, uint32_t cap) { return ((item * 0x9E3779B1u) >> 7) & (cap - 1u); }
__device__ __forceinline__ uint32_t hash_stride(uint32_t item) { return ((item * 0x9E3779B1u) >> 20) | 1u; }
__device__ __forceinline__ uint32_t next_slot(uint32_t h, uint32_t stride, uint32_t cap) { return (h + stride) & (cap - 1u); }

// phase 2b: A[item] += w for every item of every neighbour session (mod.rs:144-153).
//
// The unit of work is one 16-byte GRANULE of a neighbour's item list (lists start 16-byte aligned and are padded with
// kEmpty): one LDG.128 brings up to four items that share one weight.  insert_granule() is warp-wide, one granule per
// lane:
//   1. the FIRST probe of all four items is straight-line code — four independent 64-bit compare-and-swaps in flight
//      per lane, no loop, no vote; a claim deposits key and weight at once, a hit adds to the low word (~80 % of the
//      items are placed here at the table's load factors)
//   2. the items whose first probe hit a foreign key are finished by a short divergent loop (double hashing: an odd
//      stride visits every slot of the power-of-two table and avoids the primary clustering of linear probing)
// No bookkeeping of claimed slots: phase 3 scans the table.  The most recent item of the evolving session is never
// inserted: it is dropped from the result anyway (mod.rs:157-160) and would be the hottest slot of the table.
// `nclaim` counts the slots this lane claimed; S.overflow is raised if a probe sequence wrapped (full table).
__device__ __forceinline__ void insert_granule(SmemLayout& S, const uint4 it, int32_t w, uint32_t last_idx, Slot* tab,
                                               uint32_t cap, uint32_t& nclaim) {
  // stage A: four independent compare-and-swaps in flight (no result is looked at before all are issued)
  const bool v0 = it.x != kEmpty && it.x != last_idx, v1 = it.y != kEmpty && it.y != last_idx;
  const bool v2 = it.z != kEmpty && it.z != last_idx, v3 = it.w != kEmpty && it.w != last_idx;
  const uint32_t h0 = home_slot(it.x, cap), h1 = home_slot(it.y, cap), h2 = home_slot(it.z, cap), h3 = home_slot(it.w, cap);
  Slot o0 = 0, o1 = 0, o2 = 0, o3 = 0;                   // key 0 of an unissued probe is never looked at
  if (v0) o0 = atomicCAS(&tab[h0], kEmptySlot, make_slot(it.x, w));
  if (v1) o1 = atomicCAS(&tab[h1], kEmptySlot, make_slot(it.y, w));
  if (v2) o2 = atomicCAS(&tab[h2], kEmptySlot, make_slot(it.z, w));
  if (v3) o3 = atomicCAS(&tab[h3], kEmptySlot, make_slot(it.w, w));
  // a claim is done, a hit adds to the low word (little endian: low word = A), the rest goes on probing
  const uint32_t k0 = slot_key(o0), k1 = slot_key(o1), k2 = slot_key(o2), k3 = slot_key(o3);
  if (v0 && k0 == it.x) atomicAdd(reinterpret_cast<int*>(&tab[h0]), w);
  if (v1 && k1 == it.y) atomicAdd(reinterpret_cast<int*>(&tab[h1]), w);
  if (v2 && k2 == it.z) atomicAdd(reinterpret_cast<int*>(&tab[h2]), w);
  if (v3 && k3 == it.w) atomicAdd(reinterpret_cast<int*>(&tab[h3]), w);
  nclaim += (uint32_t)(v0 && k0 == kEmpty) + (uint32_t)(v1 && k1 == kEmpty) + (uint32_t)(v2 && k2 == kEmpty) + (uint32_t)(v3 && k3 == kEmpty);
  uint32_t pend = (v0 && k0 != kEmpty && k0 != it.x ? 1u : 0u) | (v1 && k1 != kEmpty && k1 != it.y ? 2u : 0u) |
                  (v2 && k2 != kEmpty && k2 != it.z ? 4u : 0u) | (v3 && k3 != kEmpty && k3 != it.w ? 8u : 0u);
  // stage B: the collided items (about one in five) are dealt out again, ONE per lane, through a 32-entry staging row
  // of the warp, and every lane walks the rest of its item's probe sequence — a couple of trips with most lanes busy
  // instead of four mostly idle slots per lane
  const uint32_t lane = threadIdx.x & 31u;
  unsigned long long* stage = S.stage[threadIdx.x >> 5];
  for (;;) {
    const int cnt = __popc(pend);
    if (!__any_sync(kFull, cnt != 0)) break;
    int wt;
    const int incl = warp_incl_scan_small<3>(cnt, (int)lane, wt);      // cnt <= 4
    const uint32_t total = (uint32_t)wt;
    uint32_t pos = (uint32_t)(incl - cnt);
    while (pend != 0u && pos < 32u) {
      const uint32_t item = (pend & 1u) ? it.x : (pend & 2u) ? it.y : (pend & 4u) ? it.z : it.w;
      pend &= pend - 1u;
      stage[pos++] = make_slot(item, w);
    }
    __syncwarp();
    if (lane < min(total, 32u)) {
      const Slot e = stage[lane];
      const uint32_t item = slot_key(e);
      const int32_t wi = slot_val(e);
      const uint32_t stride = hash_stride(item);
      uint32_t h = home_slot(item, cap);
      uint32_t tries = cap - 1u;                         // every other slot once
      for (;;) {
        h = next_slot(h, stride, cap);
        const uint32_t ok = slot_key(atomicCAS(&tab[h], kEmptySlot, e));
        if (ok == kEmpty) { ++nclaim; break; }
        if (ok == item) { atomicAdd(reinterpret_cast<int*>(&tab[h]), wi); break; }
        if (--tries == 0u) { S.overflow = 1u; break; }   // unreachable while the table has a free slot
      }
    }
    __syncwarp();
  }
}

// Walks the item lists of the nn neighbours into the table.  kFlat: phase 2a flattened the lists into G granules with
// a granule -> neighbour map (one granule per lane, rounds of 256 granules per block); otherwise (the map did not fit:
// very long sessions) every warp takes whole neighbours and its lanes stride over that neighbour's granules.
// With `guard` (the neighbours hold more items than the table's occupancy budget) every warp publishes its claims
// after each round and stops once the budget is exceeded; the query is then redone on the global table.
struct NeighbourLists {
  const uint32_t* goff;      // [nn] kFlat: (item list offset in 16-byte units) - (first granule); else the plain offset
  const uint32_t* len;       // [nn] item list length
  const int32_t* w;          // [nn] weight 10*linear_score*numerator
  const uint16_t* gran_nbr;  // [G]  kFlat: neighbour owning the granule
};
template <bool kFlat>
__device__ __forceinline__ void accumulate(const IndexView& ix, SmemLayout& S, const NeighbourLists nl, uint32_t nn, uint32_t G,
                                           uint32_t last_idx, Slot* tab, uint32_t cap, bool guard, uint32_t occ_cap) {
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  const uint4* lists = reinterpret_cast<const uint4*>(ix.sess_items);
  const uint4 none = make_uint4(kEmpty, kEmpty, kEmpty, kEmpty);
  uint32_t nclaim = 0;
  auto over_budget = [&]() -> bool {                    // warp-uniform; only called with guard
    const uint32_t c = __reduce_add_sync(kFull, nclaim);
    nclaim = 0;
    uint32_t tot = 0;
    if (lane == 0) tot = atomicAdd(&S.n_occ, c) + c;
    return __shfl_sync(kFull, tot, 0) > occ_cap;
  };
  if (kFlat) {
    auto fetch = [&](uint32_t g, uint4& it, int32_t& w) {
      it = none; w = 0;
      if (g < G) { const uint32_t i = nl.gran_nbr[g]; w = nl.w[i]; it = __ldg(lists + (nl.goff[i] + g)); }
    };
    // rounds of 32 granules are handed out by a block-wide counter: the warps finish within a round of each other
    // whatever their luck with the probe sequences
    auto grab = [&]() -> uint32_t {
      uint32_t r = 0;
      // one lane, one atomic: written in PTX so that the compiler does not wrap it in its warp-aggregation sequence
      // (vote, find-leader, popc, shuffle — 19 instructions for a single active lane)
      if (lane == 0) asm volatile("atom.shared.add.u32 %0, [%1], 1;" : "=r"(r) : "r"(smem_u32(&S.round)) : "memory");
      return __shfl_sync(kFull, r, 0) * 32u;
    };
    uint4 it; int32_t w;
    uint32_t base = grab();
    fetch(base + lane, it, w);
    VMIS_CLK_AFTER(S, it.x);
#ifdef VMIS_PHASE_CLOCKS
    bool first_round = true;
#endif
    while (base < G) {
      const uint32_t nbase = grab();
      uint4 nit; int32_t nw;                            // next round's granule travels while this one is inserted
      fetch(nbase + lane, nit, nw);
      insert_granule(S, it, w, last_idx, tab, cap, nclaim);
#ifdef VMIS_PHASE_CLOCKS
      if (first_round) { VMIS_CLK(S); first_round = false; }
#endif
      if (guard && over_budget()) break;
      it = nit; w = nw; base = nbase;
    }
#ifdef VMIS_PHASE_CLOCKS
    if (first_round) VMIS_CLK(S);
    VMIS_CLK(S);
#endif
  } else {
    bool stop = false;
    for (uint32_t i = warp; i < nn && !stop; i += kWarps) {
      const uint32_t ng = (nl.len[i] + 3u) >> 2;
      const int32_t w = nl.w[i];
      const uint4* p = lists + nl.goff[i];
      for (uint32_t g0 = 0; g0 < ng && !stop; g0 += 32u) {
        insert_granule(S, g0 + lane < ng ? __ldg(p + g0 + lane) : none, w, last_idx, tab, cap, nclaim);
        if (guard && over_budget()) stop = true;
      }
    }
  }
}

// compaction of the occupied slots into `occ` (any order): every thread inspects the slots tid, tid+256, ...
// (coalesced), one block scan places its finds.  Result count in S.n_occ; sets S.overflow if the list is too small.
template <typename OccT>
__device__ __forceinline__ void compact_slots(SmemLayout& S, uint32_t& par, const Slot* tab, uint32_t cap, OccT* occ,
                                              uint32_t occ_cap) {
  const uint32_t tid = threadIdx.x;
  uint32_t run_total = 0;
  for (uint32_t chunk = 0; chunk < cap; chunk += kThreads * 32) {          // 32 slots per thread per pass
    uint32_t used = 0;
#pragma unroll 8
    for (uint32_t j = 0; j < 32; ++j) {
      const uint32_t slot = chunk + j * kThreads + tid;
      if (slot < cap && slot_key(tab[slot]) != kEmpty) used |= 1u << j;
    }
    int total;
    uint32_t pos = run_total + (uint32_t)block_excl_scan(__popc(used), S.scan, par, total);
    while (used) {
      const uint32_t j = (uint32_t)__ffs((int)used) - 1u;
      used &= used - 1u;
      if (pos < occ_cap) occ[pos] = (OccT)(chunk + j * kThreads + tid);
      ++pos;
    }
    run_total += (uint32_t)total;
  }
  if (tid == 0) { S.n_occ = min(run_total, occ_cap); if (run_total > occ_cap) S.overflow = 1u; }
}

constexpr int kIdxBits = 13;                                 // entry index bits inside a coarse key
constexpr uint32_t kIdxMask = (1u << kIdxBits) - 1;

__device__ __forceinline__ uint32_t u32_step(uint32_t x, int lane, int j, bool up) {
  const uint32_t o = __shfl_xor_sync(kFull, x, j);
  return (((lane & j) == 0) == up) ? max(x, o) : min(x, o);
}
__device__ __forceinline__ uint32_t u32_sort_desc(uint32_t x, int lane) {
#pragma unroll
  for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) x = u32_step(x, lane, j, (lane & k) == 0);
  }
  return x;
}
__device__ __forceinline__ uint32_t u32_merge_top(uint32_t top, uint32_t cand_sorted, int lane) {
  uint32_t x = max(top, __shfl_sync(kFull, cand_sorted, 31 - lane));
#pragma unroll
  for (int j = 16; j > 0; j >>= 1) x = u32_step(x, lane, j, true);
  return x;
}

// exact candidate of one occupied slot; sentinel if filtered (current item, business rules)
__device__ __forceinline__ Elem exact_elem(const IndexView& ix, const PredictArgs& a, const QueryCtx& c, uint32_t key,
                                           int32_t A, double denom) {
  Elem e; e.s = 0; e.id = kEmpty;
  if (key == kEmpty || key == c.last_idx) return e;                    // mod.rs:157-160
  if (a.biz && !passes_business_rules(c.cur_attr, ix.attr[key])) return e;
  const double idf = ix.idf[key];
  const double g = idf > 0.0 ? idf : 1.0;                              // mod.rs:145-152
  e.s = score_bits(g * (double)A / denom); e.id = key;
  return e;
}

// phase 3, exact path: top-n by (score desc, item asc) with the exact 96-bit (f64 score bits, dense idx) network,
// rounds of 32 over every entry.  Used when how_many > 31, when the coarse pass could not prove its candidate set
// (heavy score ties) and for the global table.  Entries: kIdentity — every slot of the shared table (empty slots
// yield sentinels); otherwise the compacted occupied-slot list `occ` of the global table.
template <bool kIdentity, typename OccT>
__device__ __forceinline__ uint32_t select_exact(const IndexView& ix, const PredictArgs& a, SmemLayout& S, Scratch& X,
                                                 const QueryCtx& c, const Slot* tab, const OccT* occ, uint32_t n_entries) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t N = a.how_many, q = c.q;
  const double denom = (double)(10u * c.u);
  uint32_t written = 0;
  Elem bound; bound.s = ~0ull; bound.id = 0;                      // exclusive upper bound of the current round
  bool first_round = true;
  while (written < N) {
    Elem top; top.s = 0; top.id = kEmpty;
    for (uint32_t base = (uint32_t)warp * 32u; base < n_entries; base += kThreads) {
      const uint32_t e = base + lane;
      Elem x; x.s = 0; x.id = kEmpty;
      if (e < n_entries) {
        const Slot sl = tab[kIdentity ? e : (uint32_t)occ[e]];
        x = exact_elem(ix, a, c, slot_key(sl), slot_val(sl), denom);
        if (!first_round && !better(bound, x)) { x.s = 0; x.id = kEmpty; }
      }
      const Elem worst = shfl_elem(top, 31);
      if (__any_sync(kFull, better(x, worst))) top = warp_merge_top(top, warp_sort_desc(x, lane), lane);
    }
    X.top.s[warp * 32 + lane] = top.s; X.top.id[warp * 32 + lane] = top.id;
    __syncthreads();
    if (warp == 0) {
      Elem best; best.s = X.top.s[lane]; best.id = X.top.id[lane];
      for (int w = 1; w < kWarps; ++w) {
        Elem o; o.s = X.top.s[w * 32 + lane]; o.id = X.top.id[w * 32 + lane];
        best = warp_merge_top(best, o, lane);
      }
      const uint32_t valid = __popc(__ballot_sync(kFull, best.id != kEmpty));
      const uint32_t take = min(valid, N - written);
      if ((uint32_t)lane < take) {
        a.out_ids[(size_t)q * N + written + lane] = ix.item_key[best.id];
        a.out_scores[(size_t)q * N + written + lane] = bits_score(best.s);
      }
      if (lane == 0) S.sel_count = take;
      if (take > 0) { const Elem lastE = shfl_elem(best, (int)take - 1); if (lane == 0) { X.top.s[0] = lastE.s; X.top.id[0] = lastE.id; } }
    }
    __syncthreads();
    const uint32_t emitted = S.sel_count;
    if (emitted > 0) { bound.s = X.top.s[0]; bound.id = X.top.id[0]; }
    written += emitted;
    first_round = false;
    __syncthreads();
    if (emitted < 32) break;
  }
  return written;
}

// phase 3 on the shared table: top-n by (score desc, item asc) straight from the table, no list of occupied slots
// and as few dependent steps as possible.
//
// A monotone 19-bit coarse key (fp32 image of the score: A * g32[item] / (10 u)) packed with the slot index stands
// for a slot; exact order can only disagree with coarse order by one coarse unit.
//   1. every thread owns 16 slots per 4096 (16-byte reads, conflict free) and scores them with all its 4-byte gathers
//      in flight at once; the keys stay in registers
//   2. the four best thread maxima of every warp (four REDUX steps) go to shared memory; after one barrier every warp
//      sorts these 32 keys: their N-th largest is a lower bound B of the query's N-th best key — close to the true
//      one, because the best items mostly sit with different threads
//   3. every key reaching B minus one unit goes to ONE block-wide queue (warp scan + one shared atomic per warp);
//      typically a few more than N survive
//   4. warp 0: up to 32 survivors are rescored exactly (f64 g(idf) * A / (10 u)) and sorted once with the 96-bit
//      network; more than 32 are first cut to the 32 best coarse keys, with the margin test proving that cut; if it
//      cannot (heavy score ties) or the queue overflowed, select_exact() scans the table.
template <bool kBiz>     // business rules on / off: compiled twice, so the plain call carries no trace of the filter
__device__ __forceinline__ uint32_t select_table(const IndexView& ix, const PredictArgs& a, SmemLayout& S, Scratch& X,
                                                 uint32_t* queue, const QueryCtx& c, const Slot* tab, uint32_t tab_cap) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t N = a.how_many, q = c.q;
  const double denom = (double)(10u * c.u);
  if (N <= 31) {
    // coarse key of a slot (0 = empty / filtered): 19-bit monotone image of an fp32 APPROXIMATION of the score
    // (relative error < 2^-21, far below the 2^-10 coarse unit): every exact top-n element has coarse >= (n-th largest
    // coarse) - 1.  The current item is never in the table (accumulate drops it, mod.rs:157-160).
    const float rdenom = 1.0f / (float)denom;
    auto coarse = [&](uint32_t slot, uint32_t key, uint32_t A) -> uint32_t {
      bool valid = key != kEmpty;
      float g = 0.0f;
      if (valid) g = __ldg(ix.g32 + key);                                  // mod.rs:145-152
      if (kBiz) { if (valid && !passes_business_rules(c.cur_attr, ix.attr[key])) valid = false; }
      const uint32_t fb = __float_as_uint((float)(int32_t)A * g * rdenom);
      const uint32_t mono = fb ^ ((uint32_t)((int32_t)fb >> 31) | 0x80000000u);
      return valid ? ((mono & ~kIdxMask) | slot) : 0u;
    };
    const uint4* t2 = reinterpret_cast<const uint4*>(tab);                 // {A0, key0, A1, key1}
    const uint32_t passes = tab_cap / (16u * kThreads);                    // tab_cap is a multiple of 4096
    uint32_t thr = 1u;                                                     // keys at or above stay in the race
    if (tid == 0) S.qcount = 0;
    for (uint32_t ps = 0; ps < passes; ++ps) {
      uint32_t cc[16];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const uint32_t p = (ps * 8u + (uint32_t)j) * kThreads + (uint32_t)tid;
        const uint4 u = t2[p];
        cc[2 * j] = coarse(2u * p, u.y, u.x);
        cc[2 * j + 1] = coarse(2u * p + 1u, u.w, u.z);
      }
      if (ps == 0) {
        uint32_t m = cc[0];
#pragma unroll
        for (int j = 1; j < 16; ++j) m = max(m, cc[j]);
        // the warp's four best thread maxima (keys are unique: they carry the slot)
        const uint32_t t0 = __reduce_max_sync(kFull, m); if (m == t0) m = 0u;
        const uint32_t t1 = __reduce_max_sync(kFull, m); if (m == t1) m = 0u;
        const uint32_t t2_ = __reduce_max_sync(kFull, m); if (m == t2_) m = 0u;
        const uint32_t t3 = __reduce_max_sync(kFull, m);
        if (lane == 0) *reinterpret_cast<uint4*>(S.top4[warp]) = make_uint4(t0, t1, t2_, t3);
        VMIS_CLK(S);
        __syncthreads();
        VMIS_CLK(S);
        const uint32_t s32 = u32_sort_desc((lane >> 2) < kWarps ? S.top4[lane >> 2][lane & 3] : 0u, lane);
        const uint32_t bound = __shfl_sync(kFull, s32, (int)N - 1) >> kIdxBits;   // 0: fewer than N candidates so far
        thr = bound > 1u ? (bound - 1u) << kIdxBits : 1u;
      }
      uint32_t keep = 0;
#pragma unroll
      for (int j = 0; j < 16; ++j) keep |= (cc[j] >= thr ? 1u : 0u) << j;
      if (keep != 0u) {                                                   // a handful of lanes per query
        uint32_t pos = atomicAdd(&S.qcount, (uint32_t)__popc(keep));
        // The queue holds slot numbers.  Its first 32 entries are rescored exactly right here, by the thread that
        // queued them (all warps in parallel, off the serial tail): f64 g(idf) * A / (10 u), mod.rs:145-152
        do {
          const uint32_t j = (uint32_t)__ffs((int)keep) - 1u;
          keep &= keep - 1u;
          const uint32_t slot = 2u * ((ps * 8u + (j >> 1)) * kThreads + (uint32_t)tid) + (j & 1u);
          if (pos < kSelQ) queue[pos] = slot;
          if (pos < 32u) {
            const Slot sl = tab[slot];
            asm volatile("prefetch.global.L2 [%0];" ::"l"(ix.item_key + slot_key(sl)));   // read by the tail
            const Elem e = exact_elem(ix, a, c, slot_key(sl), slot_val(sl), denom);
            X.ex.s[pos] = e.s; X.ex.id[pos] = e.id;
          }
          ++pos;
        } while (keep != 0u);
      }
    }
    VMIS_CLK(S);
    __syncthreads();
    VMIS_CLK(S);
    const uint32_t n = S.qcount;
    if (n <= 32u) {
      // Rank by counting, all warps at once: the exact order is total (ids are unique), so the ranks are a
      // permutation.  A group of 8 lanes owns one candidate (4 per warp), each lane compares it with 4 others, three
      // shuffle steps add up the group — no serial tail, and every thread knows the outcome without a broadcast.
      for (uint32_t c0 = (uint32_t)warp * 4u; c0 < n; c0 += (uint32_t)kWarps * 4u) {     // one trip with 8 warps
        const uint32_t cnd = c0 + ((uint32_t)lane >> 3);
        Elem my; my.s = 0; my.id = kEmpty;
        if (cnd < n) { my.s = X.ex.s[cnd]; my.id = X.ex.id[cnd]; }
        uint32_t rank = 0;
#pragma unroll
        for (uint32_t k = 0; k < 4u; ++k) {
          const uint32_t j = ((uint32_t)lane & 7u) + 8u * k;
          if (j < n) { Elem o; o.s = X.ex.s[j]; o.id = X.ex.id[j]; rank += better(o, my) ? 1u : 0u; }
        }
        rank += __shfl_xor_sync(kFull, rank, 1);
        rank += __shfl_xor_sync(kFull, rank, 2);
        rank += __shfl_xor_sync(kFull, rank, 4);
        if (((uint32_t)lane & 7u) == 0u && cnd < n && rank < N) {
          a.out_ids[(size_t)q * N + rank] = ix.item_key[my.id];
          a.out_scores[(size_t)q * N + rank] = bits_score(my.s);
        }
      }
      return min(n, N);
    }
    if (warp == 0) {
      bool ok = n <= kSelQ;
      uint32_t take = 0;
      if (ok) {
        // cut to the 32 best coarse keys; the cut is proven if the last one is more than a unit below the N-th
        uint32_t best = 0;
        for (uint32_t b = 0; b < n; b += 32u) {
          uint32_t key = 0u;
          if (b + lane < n) { const uint32_t slot = queue[b + lane]; const Slot sl = tab[slot]; key = coarse(slot, slot_key(sl), (uint32_t)slot_val(sl)); }
          best = u32_merge_top(best, u32_sort_desc(key, lane), lane);
        }
        ok = (__shfl_sync(kFull, best, 31) >> kIdxBits) + 1u < (__shfl_sync(kFull, best, (int)N - 1) >> kIdxBits);
        if (ok) {
          take = N;
          const Slot sl = tab[best & kIdxMask];
          asm volatile("prefetch.global.L2 [%0];" ::"l"(ix.item_key + slot_key(sl)));
          Elem x = exact_elem(ix, a, c, slot_key(sl), slot_val(sl), denom);
          x = warp_sort_desc(x, lane);
          if ((uint32_t)lane < take) {
            a.out_ids[(size_t)q * N + lane] = ix.item_key[x.id];
            a.out_scores[(size_t)q * N + lane] = bits_score(x.s);
          }
        }
      }
      if (lane == 0) { S.sel_ok = ok ? 1u : 0u; S.sel_count = take; }
    }
    __syncthreads();     // sel_ok / sel_count are not rewritten before several later barriers: no trailing barrier
    const uint32_t ok = S.sel_ok, cnt = S.sel_count;
    if (ok) return cnt;
  }
  return select_exact<true, uint16_t>(ix, a, S, X, c, tab, nullptr, tab_cap);
}


// Phase 0 of the NEXT query, run by one warp while the others insert (phase 2b hands its rounds out dynamically, so
// nobody waits for this warp): de-duplicate the evolving session (vmis_index.rs:335-348: a later duplicate of an item
// does not count; unique items include unknown ones), translate ids through the HBM item hash, compact the distinct
// known items in position order.  Sessions of more than 32 items are left to the block-wide path at the loop top.
// d_idx / d_pos of the current query are dead after phase 2a, so the results go straight into them.
__device__ __forceinline__ void phase0_next(const IndexView& ix, const PredictArgs& a, SmemLayout& S, SmemLayout::Next& nx) {
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t q = nx.q;
  if (q >= a.n_q) return;
  const uint32_t qo = a.q_off[q];
  const uint32_t L = a.q_off[q + 1] - qo;
  if (L > 32u) return;                                             // nx.ok stays 0
  const uint32_t qb = qo - a.q_item_base;
  const uint32_t valid = L == 32u ? kFull : (1u << L) - 1u;
  // lane t holds the item at position t from the end; idle lanes get values that cannot match a valid lane's mask
  const uint64_t it = lane < L ? a.q_items[qb + (L - 1u - lane)] : 0ull;
  const uint32_t same = __match_any_sync(kFull, it) & valid;
  const bool distinct = lane < L && (uint32_t)__ffs((int)same) - 1u == lane;   // the most recent occurrence counts
  uint32_t my_idx = kEmpty;
  if (distinct) my_idx = lookup_item(ix, it);
  const uint32_t known = __ballot_sync(kFull, my_idx != kEmpty);
  const uint32_t uniq = __ballot_sync(kFull, distinct);
  if (my_idx != kEmpty) {
    const uint32_t pos = (uint32_t)__popc(known & ((1u << lane) - 1u));
    S.d_idx[pos] = my_idx; S.d_pos[pos] = (uint8_t)lane;
  }
  if (my_idx != kEmpty) {
    // the posting-list ref of the item: one dependent global load less at the head of phase 1 and of every fold
    const uint32_t pos = (uint32_t)__popc(known & ((1u << lane) - 1u));
    if (pos < kRefCache) {
      const uint2 ref = ix.post_ref[my_idx];
      S.d_ref[pos] = ref;
    }
  }
  if (lane == 0) { nx.nd = (uint32_t)__popc(known); nx.u = (uint32_t)__popc(uniq); nx.L = L; nx.ok = 1u; }
}

__global__ void __launch_bounds__(kThreads, kCtasPerSm)
vmis_predict_kernel(const IndexView ix, const PredictArgs a, const LaunchPlan plan, const Workspace ws) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SmemLayout& S = *reinterpret_cast<SmemLayout*>(smem_raw);
  unsigned char* dyn = smem_raw + ((sizeof(SmemLayout) + 15) & ~size_t(15));
  // neighbour arrays
  Scratch& X = *reinterpret_cast<Scratch*>(dyn);
  uint32_t* nbr_goff = reinterpret_cast<uint32_t*>(dyn);           // [K]   item list offset (16-byte units) - first granule (phase 2)
  uint32_t* nbr_sid = nbr_goff + a.k;                              // [K+1] time rank of the neighbour session ...
  uint32_t* nbr_len = nbr_sid;                                     //       ... later its item list length
  uint32_t* nbr_low = nbr_sid + a.k + 1;                           // [K]   pos|numerator, later the weight w
  unsigned char* region = dyn + nbr_bytes(a.k);
  // phase-0 view of the region
  uint64_t* q_item = reinterpret_cast<uint64_t*>(region);          // evolving session reversed: [pos]
  // phase-1 view of the region
  uint64_t* acc0 = reinterpret_cast<uint64_t*>(region);
  uint64_t* acc1 = acc0 + plan.m_eff;
  uint32_t* listbuf = reinterpret_cast<uint32_t*>(acc1 + plan.m_eff);
  // phase-2/3 view of the region (aliases phase 1)
  Slot* stab = reinterpret_cast<Slot*>(region);                    // score table {item | A}
  uint16_t* gran_nbr = reinterpret_cast<uint16_t*>(stab + plan.tab_cap);        // granule -> neighbour; phase 3: candidate queues

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t K = a.k, M = a.m, N = a.how_many;
  const bool neighbors_mode = a.out_sess != nullptr;
  // With m <= m_carry (= m_build for an index built here; checked at load for a pre-computed one) a session of
  // the m-sample is on the (truncated) posting list of EVERY evolving item it contains, so the first-match
  // position of mod.rs:133-138 is the position of the first list it came from.  Otherwise the item lists are
  // scanned as the reference does.
  const bool pos_from_lists = M <= ix.m_carry;
  uint32_t par = 0;                                                // scan scratch parity (block-uniform)
  uint32_t bar_parity = 0;                                         // bit b: phase parity of S.bar[b]
  if (tid == 0) {
    mbar_init(&S.bar[0], 1); mbar_init(&S.bar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }

  // thread 0 always has the work item after the next in flight: the global atomic's round trip overlaps a query
  uint32_t next_q = 0;
  uint32_t nb = 0;                                                 // S.nx[nb]: the query to run now
  // The first query of CTA b is query b (the grid never exceeds the batch): no round trip to the counter before the
  // first phase 0.  The counter hands out the queries from gridDim.x on.
  if (tid == 0) {
    S.nx[0].q = blockIdx.x; S.nx[0].ok = 0u;
    next_q = gridDim.x + atomicAdd(ws.counter, 1u);
  }
  for (;; nb ^= 1u) {
    __syncthreads();                                               // the previous query is finished; S.nx[nb] is complete
    const uint32_t q = S.nx[nb].q;
    if (q >= a.n_q) break;
    if (tid == 0) { S.nx[nb ^ 1u].q = next_q; S.nx[nb ^ 1u].ok = 0u; next_q = gridDim.x + atomicAdd(ws.counter, 1u); }

    if (tid == 0) VMIS_CLK_RESET(S);
    VMIS_CLK(S);
    // ------------------------------------------------------------------ phase 0
    uint32_t nd, u, L;
    bool too_long = false;
    if (S.nx[nb].ok) {
      // done by the last warp during the previous query's phase 2b (phase0_next)
      nd = S.nx[nb].nd; u = S.nx[nb].u; L = S.nx[nb].L;
    } else {
      const uint32_t qo = a.q_off[q];
      const uint32_t Lfull = a.q_off[q + 1] - qo;
      const uint32_t qb = qo - a.q_item_base;
      too_long = Lfull > (uint32_t)kMaxSessionLen;                   // rejected up front by the host API; flagged per query here
      L = too_long ? 0u : Lfull;
      if (tid < (int)L) q_item[tid] = a.q_items[qb + (L - 1 - tid)];
      __syncthreads();
      uint32_t my_idx = kEmpty;
      bool distinct = false;
      if (tid < (int)L) {
        const uint64_t it = q_item[tid];
        distinct = true;
        for (int t = 0; t < tid; ++t) if (q_item[t] == it) { distinct = false; break; }
        if (distinct) my_idx = lookup_item(ix, it);
      }
      // one packed scan: low half compacts the distinct KNOWN items in position order, high half counts the
      // unique items including unknown ones (vmis_index.rs:335-339)
      {
        const int flag = (my_idx != kEmpty) ? 1 : 0;
        int total;
        const int pos = block_excl_scan(flag | ((distinct ? 1 : 0) << 16), S.scan, par, total) & 0xFFFF;
        if (flag) {
          S.d_idx[pos] = my_idx; S.d_pos[pos] = (uint8_t)tid;
          if ((uint32_t)pos < kRefCache) S.d_ref[pos] = ix.post_ref[my_idx];
        }
        if (tid == 0) S.nd = (uint32_t)total & 0xFFFFu;
        u = (uint32_t)total >> 16;
      }
      __syncthreads();
      nd = S.nd;
    }
    // most recent item: removed from the result (mod.rs:157-160); attributes for the adult rule (:186)
    const uint32_t last_idx = (nd > 0 && S.d_pos[0] == 0) ? S.d_idx[0] : kEmpty;
    const uint32_t cur_attr = (a.biz && last_idx != kEmpty) ? ix.attr[last_idx] : 0u;

    uint32_t nn = 0;                 // number of neighbours
    uint32_t postings_visited = 0;

    VMIS_CLK(S);
    if (nd > 0 && K > 0 && M > 0 && (N > 0 || neighbors_mode)) {
      // ---------------------------------------------------------------- phase 1
      const uint2 ref0 = S.d_ref[0];
      const uint32_t n0 = min(ref0.y, M);
      const uint32_t* P0 = posting_list(ix, S.d_idx[0], ref0.x);
      const uint32_t low0 = (S.d_pos[0] << 24) | (L - S.d_pos[0]);
      postings_visited = n0;
      if (nd == 1) {
        // single distinct known item: S = first m postings, all similarities equal → N = first k
        nn = min(n0, K);
        for (uint32_t i = tid; i < nn; i += kThreads) { const uint32_t sid = P0[i]; nbr_sid[i] = sid; nbr_low[i] = low0; PF_SESSREF(ix.sess_ref + sid); }
      } else {
        uint64_t* acc = acc0;
        uint64_t* out = acc1;
        // TMA: list 1 (with two staging buffers also list 2) starts streaming in while list 0 is converted
        auto issue = [&](uint32_t j, uint32_t b) {
          const uint2 ref = j < kRefCache ? S.d_ref[j] : ix.post_ref[S.d_idx[j]];
          const uint32_t bytes = ((min(min(ref.y, M), plan.list_cap) + 3u) & ~3u) * 4u;
          mbar_expect_tx(&S.bar[b], bytes);
          bulk_load(listbuf + (size_t)b * plan.list_cap, posting_list(ix, S.d_idx[j], ref.x), bytes, &S.bar[b]);
        };
        if (tid == 0) { fence_async_proxy(); issue(1, 0); if (kStages > 1 && nd > 2) issue(2, 1); }
        for (uint32_t i = tid; i < n0; i += kThreads) acc[i] = ((uint64_t)P0[i] << 32) | low0;
        if (tid == 0) acc[n0] = ~0ull;                                        // sentinel (see the merge steps)
        uint32_t na = n0;
        for (uint32_t j = 1; j < nd; ++j) {
          const uint32_t bsel = kStages > 1 ? (j - 1) & 1u : 0u;
          const uint32_t* lst = listbuf + (size_t)bsel * plan.list_cap;
          const uint2 ref = j < kRefCache ? S.d_ref[j] : ix.post_ref[S.d_idx[j]];
          const uint32_t nb = min(min(ref.y, M), plan.list_cap);
          const uint32_t cj = L - S.d_pos[j];
          const uint32_t lowj = (S.d_pos[j] << 24) | cj;
          postings_visited += nb;
          mbar_wait(&S.bar[bsel], (bar_parity >> bsel) & 1u);
          bar_parity ^= 1u << bsel;
          if (tid == 0) const_cast<uint32_t*>(lst)[nb] = kEmpty;              // sentinel behind the staged list
          __syncthreads();
          // merge-path fold: out ← first M distinct of acc ∪ B, numerators summed, first position kept
          const uint32_t T = na + nb;
          uint32_t out_count = 0;
          for (uint32_t base = 0; base < T && out_count < M; base += kTile) {
            const uint32_t d0 = min(base + (uint32_t)tid * kVT, T);
            const uint32_t d1 = min(d0 + kVT, T);
            uint32_t lo = d0 > nb ? d0 - nb : 0, hi = min(d0, na);
            while (lo < hi) {
              const uint32_t mid = (lo + hi) >> 1;
              if ((uint32_t)(acc[mid] >> 32) >= lst[d0 - 1 - mid]) lo = mid + 1; else hi = mid;
            }
            uint32_t ai = lo, bi = d0 - lo;
            uint32_t rh[kVT], rl[kVT];                    // merged element s: session rank, pos|numerator
            uint32_t vmask = 0;
            // Heads of both runs and the key of the last consumed A element stay in registers: one shared-memory
            // load per merged element.  Both runs end in a kEmpty sentinel, which compares below every rank as a
            // signed number, so the steps need no bounds checks.
            uint64_t av = acc[ai];
            uint32_t bk = lst[bi];
            uint32_t pak = ai > 0 ? (uint32_t)(acc[ai - 1] >> 32) : kEmpty;   // kEmpty is never a session rank
            const uint32_t steps = d1 - d0;
#pragma unroll
            for (int s = 0; s < kVT; ++s) {
              const uint32_t ak = (uint32_t)(av >> 32);
              const bool takeA = (int32_t)ak >= (int32_t)bk;                  // equal ranks: A first, B is then a duplicate
              const bool emit = (uint32_t)s < steps && (takeA || pak != bk);
              rh[s] = takeA ? ak : bk;
              rl[s] = takeA ? (uint32_t)av + (ak == bk ? cj : 0u) : lowj;
              vmask |= (emit ? 1u : 0u) << s;
              // no step past the thread's share: the heads would run beyond the sentinels, into the staging buffer a
              // TMA copy may be filling (harmless values, but a read under an in-flight asynchronous write)
              if ((uint32_t)s < steps) { if (takeA) { pak = ak; ++ai; av = acc[ai]; } else { ++bi; bk = lst[bi]; } }
            }
            int total;
            static_assert(kVT < 16, "merge tile counts must fit the 4-bit ballot scan");
            uint32_t p = out_count + (uint32_t)block_excl_scan_small<(kVT < 8 ? 3 : 4)>(__popc(vmask), S.scan, par, total);
#pragma unroll
            for (int s = 0; s < kVT; ++s) {
              if ((vmask >> s) & 1u) { if (p < M) out[p] = ((uint64_t)rh[s] << 32) | rl[s]; ++p; }
            }
            out_count = min(M, out_count + (uint32_t)total);
          }
          if (tid == 0) out[out_count] = ~0ull;                               // sentinel for the next fold
          __syncthreads();
          if (tid == 0 && j + kStages < nd) { fence_async_proxy(); issue(j + kStages, bsel); }
          uint64_t* t = acc; acc = out; out = t;
          na = out_count;
        }
        // -------------------------------------------------------------- phase 1b
        if (na <= K) {
          nn = na;
          for (uint32_t i = tid; i < na; i += kThreads) {
            const uint64_t e = acc[i];
            nbr_sid[i] = (uint32_t)(e >> 32); nbr_low[i] = (uint32_t)e; PF_SESSREF(ix.sess_ref + (uint32_t)(e >> 32));
          }
        } else {
          // v* = max v with count(num >= v) >= K  (numerators are >= 1)
          // contiguous chunk per thread keeps recency order; an ODD chunk length makes the 64-bit reads of a warp
          // (stride 2 E words) conflict free
          const uint32_t E = ((na + kThreads - 1) / kThreads) | 1u;
          const uint32_t e0 = min((uint32_t)tid * E, na), e1 = min(e0 + E, na);
          const uint32_t vmax = L * (L + 1) / 2;
          uint32_t vstar, tot_g;
          if (vmax <= 10 && na < 4096) {
            // evolving sessions of <= 4 items: per-thread packed histogram (12-bit fields, counts <= m < 4096),
            // one 64-bit shuffle reduction per half, then a 10-step suffix sum
            unsigned long long h0 = 0, h1 = 0;                  // h0: numerators 1..5, h1: 6..10
            for (uint32_t i = e0; i < e1; ++i) {
              const uint32_t nm = (uint32_t)acc[i] & kNumMask;
              if (nm <= 5) h0 += 1ull << (12 * (nm - 1)); else h1 += 1ull << (12 * (nm - 6));
            }
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) { h0 += __shfl_xor_sync(kFull, h0, d); h1 += __shfl_xor_sync(kFull, h1, d); }
            unsigned long long* hw = reinterpret_cast<unsigned long long*>(X.hist);
            if (lane == 0) { hw[warp * 2] = h0; hw[warp * 2 + 1] = h1; }
            __syncthreads();
            h0 = 0; h1 = 0;
#pragma unroll
            for (int w = 0; w < kWarps; ++w) { h0 += hw[w * 2]; h1 += hw[w * 2 + 1]; }
            uint32_t ge = 0;                                    // count(num >= v), v descending
            vstar = 1; tot_g = 0;
            bool found = false;
            for (int v = 10; v >= 1; --v) {
              const uint32_t eq = (uint32_t)(((v <= 5 ? h0 >> (12 * (v - 1)) : h1 >> (12 * (v - 6)))) & 0xFFFull);
              if (!found && ge + eq >= K) { vstar = (uint32_t)v; tot_g = ge; found = true; }
              ge += eq;
            }
          } else if (vmax <= 31) {
            // one pass: lane v of every warp counts numerators >= v through ballots
            uint32_t cnt = 0;
            for (uint32_t t = 0; t < E; ++t) {
              const uint32_t i = e0 + t;
              const uint32_t nm = i < e1 ? ((uint32_t)acc[i] & kNumMask) : 0u;
              for (uint32_t v = 1; v <= vmax; ++v) {
                const uint32_t b = __ballot_sync(kFull, nm >= v);
                if ((uint32_t)lane == v) cnt += __popc(b);
              }
            }
            X.hist[warp][lane] = cnt;
            __syncthreads();
            uint32_t c = 0;
#pragma unroll
            for (int w = 0; w < kWarps; ++w) c += X.hist[w][lane];
            const uint32_t okm = __ballot_sync(kFull, lane >= 1 && (uint32_t)lane <= vmax && c >= K);
            vstar = 31u - (uint32_t)__clz((int)okm);             // bit 1 is always set: count(num >= 1) = na > K
            tot_g = vstar < 31u ? __shfl_sync(kFull, c, (int)vstar + 1) : 0u;
            if (vstar + 1 > vmax) tot_g = 0u;
          } else {
            uint32_t vlo = 1, vhi = vmax;
            while (vlo < vhi) {
              const uint32_t v = (vlo + vhi + 1) >> 1;
              int c = 0;
              for (uint32_t i = e0; i < e1; ++i) c += (((uint32_t)acc[i] & kNumMask) >= v);
              if ((uint32_t)block_sum(c, S.scan, par) >= K) vlo = v; else vhi = v - 1;
            }
            vstar = vlo;
            int c = 0;
            for (uint32_t i = e0; i < e1; ++i) c += (((uint32_t)acc[i] & kNumMask) > vstar);
            tot_g = (uint32_t)block_sum(c, S.scan, par);
          }
          int cg = 0, ce = 0;
          for (uint32_t i = e0; i < e1; ++i) { const uint32_t nm = (uint32_t)acc[i] & kNumMask; cg += nm > vstar; ce += nm == vstar; }
          int tot_packed;
          const int pre = block_excl_scan(cg | (ce << 16), S.scan, par, tot_packed);
          uint32_t gpos = (uint32_t)pre & 0xFFFFu, pre_e = (uint32_t)pre >> 16;
          const uint32_t quota = K - tot_g;                      // ties at v*: the `quota` most recent win
          for (uint32_t i = e0; i < e1; ++i) {
            const uint64_t e = acc[i];
            const uint32_t nm = (uint32_t)e & kNumMask;
            if (nm > vstar) {
              nbr_sid[gpos] = (uint32_t)(e >> 32); nbr_low[gpos] = (uint32_t)e; ++gpos; PF_SESSREF(ix.sess_ref + (uint32_t)(e >> 32));
            } else if (nm == vstar) {
              if (pre_e < quota) { nbr_sid[tot_g + pre_e] = (uint32_t)(e >> 32); nbr_low[tot_g + pre_e] = (uint32_t)e; PF_SESSREF(ix.sess_ref + (uint32_t)(e >> 32)); }
              ++pre_e;
            }
          }
          nn = K;
        }
      }
    }
    __syncthreads();

    if (neighbors_mode) {
      // find_neighbors output: canonical order (num desc, recency desc); O(nn^2) ranking, not a hot path
      for (uint32_t i = tid; i < nn; i += kThreads) {
        const uint32_t sid = nbr_sid[i]; const uint32_t nm = nbr_low[i] & kNumMask;
        uint32_t rank = 0;
        for (uint32_t t = 0; t < nn; ++t) {
          const uint32_t on = nbr_low[t] & kNumMask;
          rank += (on > nm) || (on == nm && nbr_sid[t] > sid);
        }
        a.out_sess[(size_t)q * K + rank] = ix.rank_to_orig[sid];
        a.out_sim[(size_t)q * K + rank] = (double)nm / (double)u;
      }
      for (uint32_t i = nn + tid; i < K; i += kThreads) {            // deterministic padding
        a.out_sess[(size_t)q * K + i] = 0; a.out_sim[(size_t)q * K + i] = 0.0;
      }
      if (tid == 0) a.out_counts[q] = nn;
      continue;
    }

    VMIS_CLK(S);
    // ------------------------------------------------------------------ phase 2a: neighbour directory
    // the first neighbour's item list ref is requested before the table is cleared: the clear hides its latency
    // Every thread owns a CONTIGUOUS run of neighbours, so the granules are numbered in neighbour order: the warps of
    // phase 2b then read the per-neighbour arrays at (nearly) consecutive indices (a strided assignment put
    // neighbours t and t + 256 next to each other — same bank — and cost 2.7 wavefronts per read).
    const uint32_t n_own = (nn + kThreads - 1) / kThreads;
    const uint32_t i0 = min((uint32_t)tid * n_own, nn), i1 = min(i0 + n_own, nn);
    uint2 r_first = make_uint2(0u, 0u);
    if (i0 < i1) r_first = ix.sess_ref[nbr_sid[i0]];
    {
      // all ones = empty; the capacity is a multiple of 4096: 16-byte stores
      uint4* t4 = reinterpret_cast<uint4*>(stab);
      const uint32_t n4 = plan.tab_cap >> 1;
      for (uint32_t i = tid; i < n4; i += kThreads) t4[i] = make_uint4(kEmpty, kEmpty, kEmpty, kEmpty);
    }
    if (tid == 0) { S.n_occ = 0; S.overflow = 0; S.round = 0; }
    // item list refs, weights, granule counts of this thread's neighbours
    uint32_t my_g = 0, my_len = 0;
    for (uint32_t i = i0; i < i1; ++i) {
      const uint2 r = i == i0 ? r_first : ix.sess_ref[nbr_sid[i]];
      const uint32_t* items = ix.sess_items + (size_t)r.x * 4;
      nbr_goff[i] = r.x; nbr_len[i] = r.y; my_g += (r.y + 3u) >> 2; my_len += r.y;
      uint32_t low = nbr_low[i];
      if (!pos_from_lists) {                                        // reference scan (mod.rs:133-138)
        uint32_t pmin = 0xFFu;
        for (uint32_t t = 0; t < r.y; ++t) {
          const uint32_t it = items[t];
          for (uint32_t j = 0; j < nd; ++j) if (S.d_idx[j] == it) pmin = min(pmin, (uint32_t)S.d_pos[j]);
        }
        low = (pmin << 24) | (low & kNumMask);
      }
      nbr_low[i] = (uint32_t)session_weight10(low);
    }
    // one packed scan for the granule offsets and the item count (two when the counts may not fit 16 bits each)
    uint32_t run, G, total_items;
    if (K * ix.max_len <= 0xFFFFu) {
      int tot;
      run = (uint32_t)block_excl_scan((int)(my_g | (my_len << 16)), S.scan, par, tot) & 0xFFFFu;
      G = (uint32_t)tot & 0xFFFFu; total_items = (uint32_t)tot >> 16;
    } else {
      int tot;
      run = (uint32_t)block_excl_scan((int)my_g, S.scan, par, tot);
      G = (uint32_t)tot;
      total_items = (uint32_t)block_sum((int)my_len, S.scan, par);
    }
    const bool flat = G <= plan.gran_cap;
    if (flat) {
      for (uint32_t i = i0; i < i1; ++i) {
        const uint32_t ng = (nbr_len[i] + 3u) >> 2;
        nbr_goff[i] -= run;
        for (uint32_t g = run; g < run + ng; ++g) gran_nbr[g] = (uint16_t)i;
        run += ng;
      }
    }
    __syncthreads();

    VMIS_CLK(S);
    // ------------------------------------------------------------------ phase 2b + 3
    QueryCtx c;
    c.q = q; c.u = u; c.last_idx = last_idx; c.cur_attr = cur_attr;
    uint32_t written;
    // shared-memory score table first; the rare query whose neighbours hold more distinct items than its
    // occupancy budget is redone on this CTA's global table
    NeighbourLists nl;
    nl.goff = nbr_goff; nl.len = nbr_len; nl.w = reinterpret_cast<const int32_t*>(nbr_low); nl.gran_nbr = gran_nbr;
    if (nn == 0 || N == 0) {
      written = 0;
    } else {
      const bool guard = total_items > plan.occ_cap;                // the neighbours may hold more distinct items than the budget
      // d_idx / d_pos of this query are dead from here on: the last warp prepares the next query before it joins in
      if (warp == kWarps - 1) phase0_next(ix, a, S, S.nx[nb ^ 1u]);
      if (flat) accumulate<true>(ix, S, nl, nn, G, last_idx, stab, plan.tab_cap, guard, plan.occ_cap);
      else accumulate<false>(ix, S, nl, nn, G, last_idx, stab, plan.tab_cap, guard, plan.occ_cap);
      __syncthreads();
      VMIS_CLK(S);
      if (!S.overflow && S.n_occ <= plan.occ_cap) {
        written = a.biz ? select_table<true>(ix, a, S, X, reinterpret_cast<uint32_t*>(gran_nbr), c, stab, plan.tab_cap)
                        : select_table<false>(ix, a, S, X, reinterpret_cast<uint32_t*>(gran_nbr), c, stab, plan.tab_cap);
      } else {
        // redo on a table in HBM sized for THIS query (every item of every neighbour distinct): the CTA's own
        // overflow table, or — a neighbourhood of very long sessions — the one huge table, taken under a lock.
        // Either is all-empty between uses and cleaned through its occupied-slot list.
        uint32_t need = 1024u;
        while (need < 2u * total_items && need < 0x80000000u) need <<= 1;
        const bool huge = need > ws.gtab_cap;                          // plan_launch sized ws.gtab_huge for the worst case
        Slot* gtab = huge ? ws.huge : ws.gtab + (size_t)blockIdx.x * ws.gtab_cap;
        uint32_t* gocc = huge ? ws.huge_occ : ws.gtab_occ + (size_t)blockIdx.x * (ws.gtab_cap / 2);
        __syncthreads();
        if (tid == 0) {
          S.n_occ = 0; S.overflow = 0; S.round = 0;
          if (huge) { while (atomicCAS(ws.counter + 2, 0u, 1u) != 0u) __nanosleep(500); __threadfence(); }
        }
        __syncthreads();
        if (flat) accumulate<true>(ix, S, nl, nn, G, last_idx, gtab, need, false, 0u);
        else accumulate<false>(ix, S, nl, nn, G, last_idx, gtab, need, false, 0u);
        __syncthreads();
        compact_slots<uint32_t>(S, par, gtab, need, gocc, need / 2);
        __syncthreads();
        const uint32_t n_occ = S.n_occ;
        written = select_exact<false, uint32_t>(ix, a, S, X, c, gtab, gocc, n_occ);
        for (uint32_t e = tid; e < n_occ; e += kThreads) gtab[gocc[e]] = kEmptySlot;
        if (huge) {
          __threadfence();
          __syncthreads();
          if (tid == 0) atomicExch(ws.counter + 2, 0u);
        }
      }
    }
    for (uint32_t i = written + tid; i < N; i += kThreads) {           // deterministic padding
      a.out_ids[(size_t)q * N + i] = 0; a.out_scores[(size_t)q * N + i] = 0.0;
    }
    // zero-copy call (outputs in mapped host memory): the count doubles as the completion flag the host polls, so the
    // row must be visible system-wide before it
    if (a.host_flags) { __threadfence_system(); __syncthreads(); }
    if (tid == 0) {
      a.out_counts[q] = too_long ? VMIS_COUNT_TOO_LONG : written;
      if (a.out_stats) {
        vmis_query_stats_t st; st.postings_visited = postings_visited; st.n_neighbors = nn;
        st.neighbor_items = total_items; st.n_out = written;
#ifdef VMIS_PHASE_CLOCKS
        // tuning build: the cycle counts between the probes replace the first ids of the row
        VMIS_CLK(S);
        for (uint32_t i = 0; i + 1 < S.nclk && i < N; ++i) a.out_ids[(size_t)q * N + i] = (uint64_t)(S.clk[i + 1] - S.clk[i]);
        for (uint32_t i = S.nclk > 0 ? S.nclk - 1 : 0; i < N; ++i) a.out_ids[(size_t)q * N + i] = 0;
        if (N > 13) a.out_ids[(size_t)q * N + 13] = S.qcount;
#endif
        a.out_stats[q] = st;
      }
    }
  }
  // the last CTA to leave re-arms the work counter for the next launch on this workspace (no memset per launch)
  if (tid == 0) {
    __threadfence();
    if (atomicAdd(ws.counter + 1, 1u) == gridDim.x - 1) { ws.counter[0] = 0u; ws.counter[1] = 0u; __threadfence(); }
  }
}

uint32_t next_pow2(uint32_t x) { uint32_t p = 1; while (p < x) p <<= 1; return p; }

}  // namespace

int plan_launch(const IndexView& ix, uint32_t k, uint32_t m, int sm_count, LaunchPlan* plan, std::string* why) {
  auto refuse = [&](const std::string& msg) { if (why) *why = msg; return VMIS_ERR_LIMIT; };
  if (k > kMaxK) return refuse("k = " + std::to_string(k) + " exceeds the kernel limit of " + std::to_string(kMaxK));
  if (m > kMaxM) return refuse("m = " + std::to_string(m) + " exceeds the kernel limit of " + std::to_string(kMaxM));
  LaunchPlan p{};
  // one extra entry each: the merge steps read a sentinel behind both runs
  p.m_eff = (std::max(m, 1u) + 1u + 3u) & ~3u;
  p.list_cap = (std::min(std::max(m, 1u), std::max(ix.m_build, 1u)) + 1u + 3u) & ~3u;
  // score table: >= 4096 slots so that the occupancy budget plus one round of inserts (4 items per thread) never
  // fills it (insert_granule relies on a free slot being reachable)
  const uint32_t tab = next_pow2(std::max(k, 1u) * 12u);
  p.tab_cap = std::min(std::max(tab, 4096u), 8192u);
#ifdef VMIS_EXPERIMENT_TAB
  p.tab_cap = VMIS_EXPERIMENT_TAB;   // occupancy experiments only: breaks the guard invariant below
#endif
  p.occ_cap = p.tab_cap / 2 + p.tab_cap / 8;                        // 62.5 % of the slots
  const size_t fixed = (sizeof(SmemLayout) + 15) & ~size_t(15);
  const size_t nbr = nbr_bytes(k);
  const size_t r1 = size_t(p.m_eff) * 16 + size_t(p.list_cap) * 4 * kStages;   // two m-sample buffers + the TMA staging buffers
  // granule -> neighbour map behind the table: the worst case (every neighbour as long as the longest session) if it
  // leaves room for the full CTA count, else what fits; queries beyond it walk their neighbours list by list
  const uint64_t want = (uint64_t)std::max(k, 1u) * ((std::max(ix.max_len, 1u) + 3u) / 4u);
  const size_t per_cta = (227 * 1024) / kCtasPerSm - 1024;           // 1 KB per CTA is reserved by the driver
  const size_t base = fixed + nbr + 16 + size_t(p.tab_cap) * 8;
  uint64_t room = per_cta > base ? (per_cta - base) / 2 : 0;
  if (fixed + nbr + 16 + r1 > base) room = std::max<uint64_t>(room, (fixed + nbr + 16 + r1 - base) / 2);   // free under phase 1
  p.gran_cap = (uint32_t)std::min<uint64_t>(want, std::max<uint64_t>(room, 1280));
  p.gran_cap &= ~7u;
  const size_t r2 = size_t(p.tab_cap) * 8 + gran_bytes(p.gran_cap);
  const size_t total = fixed + nbr + std::max(std::max(r1, r2), size_t(kMaxSessionLen) * 8) + 16;
  if (total > 227 * 1024)
    return refuse("k = " + std::to_string(k) + ", m = " + std::to_string(m) + " need " + std::to_string(total) +
                  " bytes of shared memory per query (limit 232448): lower k or m");
  p.smem_bytes = (uint32_t)total;
  int per_sm = (int)std::min<size_t>(kCtasPerSm, (227 * 1024) / (total + 1024));
  if (const char* e = std::getenv("VMIS_CTAS_PER_SM")) per_sm = std::min(per_sm, std::max(1, std::atoi(e)));   // tuning knob
  if (per_sm < 1) per_sm = 1;
  p.grid = (uint32_t)(sm_count * per_sm);
  // overflow score tables in HBM: every resident CTA owns a small one (up to 2^15 slots, enough for a neighbourhood
  // of 16 k items); the rare query beyond that takes the single table sized for the worst case (k sessions of max_len
  // items, all distinct) under a lock
  const uint64_t gt = 2ull * std::max(k, 1u) * std::max(ix.max_len, 1u);
  const uint64_t worst = gt > (1ull << 31) ? 0 : next_pow2(std::max<uint32_t>((uint32_t)gt, 1024u));
  p.gtab_cap = (uint32_t)std::min<uint64_t>(worst ? worst : (1u << 15), 1u << 15);
  p.gtab_huge = worst > p.gtab_cap ? (uint32_t)worst : 0u;
  const uint64_t bytes = 10ull * p.gtab_cap * p.grid + 10ull * p.gtab_huge;
  if (worst == 0 || bytes > kMaxWorkspaceBytes)
    return refuse("the longest indexed session has " + std::to_string(ix.max_len) + " items: k = " + std::to_string(k) +
                  " neighbours of that length need an overflow score table of " + std::to_string(gt) + " slots (" +
                  std::to_string((worst ? bytes : gt * 10ull) >> 20) + " MB, limit " + std::to_string(kMaxWorkspaceBytes >> 20) +
                  " MB); prune long sessions when loading the index (max_session_len of vmis_index_from_avro_ex / "
                  "vmis_index_from_parts_ex; the reference's CSV path prunes at the 99.5th length percentile, "
                  "vmis_index.rs:67,452) or lower k");
  *plan = p;
  return VMIS_OK;
}

size_t workspace_bytes(const LaunchPlan& plan) {
  return 256 + (size_t(plan.grid) * plan.gtab_cap + plan.gtab_huge) * (8 + 2);
}

Workspace carve_workspace(void* base, const LaunchPlan& plan) {
  Workspace ws{};
  unsigned char* b = static_cast<unsigned char*>(base);
  const size_t n = size_t(plan.grid) * plan.gtab_cap;
  ws.counter = reinterpret_cast<uint32_t*>(b);
  ws.gtab = reinterpret_cast<unsigned long long*>(b + 256);
  ws.huge = ws.gtab + n;
  ws.gtab_occ = reinterpret_cast<uint32_t*>(b + 256 + (n + plan.gtab_huge) * 8);
  ws.huge_occ = ws.gtab_occ + n / 2;
  ws.gtab_cap = plan.gtab_cap;
  ws.gtab_huge = plan.gtab_huge;
  ws.grid = plan.grid;
  return ws;
}

cudaError_t init_workspace(const Workspace& ws, cudaStream_t stream) {
  const size_t n = size_t(ws.grid) * ws.gtab_cap + ws.gtab_huge;
  cudaError_t e = cudaMemsetAsync(ws.counter, 0, 4 * sizeof(uint32_t), stream);   // work counter, exit counter, lock
  if (e != cudaSuccess) return e;
  return cudaMemsetAsync(ws.gtab, 0xFF, n * 8, stream);                            // all ones = empty slot
}

namespace {
__global__ void g32_kernel(const double* __restrict__ idf, float* __restrict__ g32, uint32_t n) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { const double x = idf[i]; g32[i] = x > 0.0 ? (float)x : 1.0f; }     // mod.rs:145-152
}
}  // namespace

cudaError_t build_g32(const double* idf, float* g32, uint32_t n_items, cudaStream_t stream) {
  if (n_items == 0) return cudaSuccess;
  g32_kernel<<<(n_items + 255) / 256, 256, 0, stream>>>(idf, g32, n_items);
  return cudaGetLastError();
}

cudaError_t launch_predict(const IndexView& ix, const PredictArgs& args, const LaunchPlan& plan, const Workspace& ws,
                           cudaStream_t stream) {
  if (args.n_q == 0) return cudaSuccess;
  // the dynamic shared memory opt-in is per device and sticky: raise it only when a larger plan shows up
  static std::atomic<int> smem_opt_in[64];
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  std::atomic<int>& cur = smem_opt_in[dev & 63];
  if ((int)plan.smem_bytes > cur.load(std::memory_order_relaxed)) {
    e = cudaFuncSetAttribute(vmis_predict_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan.smem_bytes);
    if (e != cudaSuccess) return e;
    cur.store((int)plan.smem_bytes, std::memory_order_relaxed);
  }
  const uint32_t grid = std::min(plan.grid, args.n_q);
  vmis_predict_kernel<<<grid, kThreads, plan.smem_bytes, stream>>>(ix, args, plan, ws);
  return cudaGetLastError();
}

}  // namespace vmis
