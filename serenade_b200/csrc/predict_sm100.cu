// serenade_b200/csrc/predict_sm100.cu — the VMIS-kNN predict_next kernel for sm_100a.
//
// One persistent CTA per SM slot pulls evolving sessions from a global work
// counter and runs the whole query in shared memory:
//
//   phase 0  de-duplicate the evolving session (vmis_index.rs:335-348), translate
//            external item ids through the HBM item hash
//   phase 1  m-sample: fold the time-descending posting lists of the distinct items
//            with a block-wide merge-path merge that de-duplicates, sums the integer
//            similarity numerators and truncates to the m most recent sessions
//            (closed form of the heap procedure of vmis_index.rs:344-391)
//   phase 1b top-k neighbours by (numerator desc, recency desc) — threshold search
//            + ordered prefix scan (vmis_index.rs:394-412)
//   phase 2  per neighbour: first-match position → linear session weight
//            (mod.rs:133-142, :110-116); integer score numerators are accumulated
//            per item in a shared-memory hash table (mod.rs:144-153)
//   phase 3  drop the current item (mod.rs:157-160), business rules (:162-182),
//            f64 score = g(idf)·A/(10·u), warp-bitonic top-n (mod.rs:185-214)
//
// All item/session arithmetic is integer and order independent; the only floating
// point is one f64 multiply + divide per candidate item, so results are bit-exact
// against oracle/vmis_oracle.cpp canonical mode.
#include "vmis_device.h"

#include <algorithm>
#include <cstdio>

namespace vmis {
namespace {

constexpr unsigned kFull = 0xFFFFFFFFu;
constexpr int kVT = 8;                       // merge-path items per thread per tile
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
// block-wide exclusive scan; all threads must call.  s_scan has kWarps + 1 ints.
__device__ __forceinline__ int block_excl_scan(int v, int* s_scan, int& total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int inc = warp_incl_scan(v, lane);
  if (lane == 31) s_scan[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    int w = lane < kWarps ? s_scan[lane] : 0;
    int wi = warp_incl_scan(w, lane);
    if (lane < kWarps) s_scan[lane] = wi - w;
    if (lane == kWarps - 1) s_scan[kWarps] = wi;
  }
  __syncthreads();
  int res = s_scan[warp] + inc - v;
  total = s_scan[kWarps];
  __syncthreads();
  return res;
}
__device__ __forceinline__ int block_sum(int v, int* s_scan) {
  int total;
  block_excl_scan(v, s_scan, total);
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

__device__ __forceinline__ void table_add(uint32_t* keys, int32_t* vals, uint32_t mask, uint32_t idx, int32_t v) {
  uint32_t h = (idx * 0x9E3779B1u) >> 7 & mask;
  for (;;) {
    uint32_t old = atomicCAS(&keys[h], kEmpty, idx);
    if (old == kEmpty || old == idx) { atomicAdd(&vals[h], v); return; }
    h = (h + 1) & mask;
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

struct SmemLayout {
  // fixed part
  uint64_t q_item[kMaxSessionLen];   // evolving session reversed: [pos]
  uint32_t d_idx[kMaxSessionLen];    // distinct known items, most recent first
  uint32_t d_pos[kMaxSessionLen];
  int scan[kWarps + 1];
  uint32_t q;                        // current query
  uint32_t nbr_count;
  uint32_t nd;
  Elem topbuf[kWarps * 32];
};

__global__ void __launch_bounds__(kThreads, 4)
vmis_predict_kernel(const IndexView ix, const PredictArgs a, const LaunchPlan plan, const Workspace ws) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SmemLayout& S = *reinterpret_cast<SmemLayout*>(smem_raw);
  unsigned char* dyn = smem_raw + ((sizeof(SmemLayout) + 15) & ~size_t(15));
  // neighbour arrays (k entries each)
  uint32_t* nbr_sid = reinterpret_cast<uint32_t*>(dyn);
  int32_t* nbr_val = reinterpret_cast<int32_t*>(nbr_sid + a.k);
  uint2* nbr_ref = reinterpret_cast<uint2*>(dyn + ((size_t(a.k) * 8 + 15) & ~size_t(15)));
  unsigned char* region = reinterpret_cast<unsigned char*>(nbr_ref + a.k);
  region = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(region) + 15) & ~uintptr_t(15));
  // phase-1 view of the region
  uint64_t* acc0 = reinterpret_cast<uint64_t*>(region);
  uint64_t* acc1 = acc0 + plan.m_eff;
  uint32_t* listbuf = reinterpret_cast<uint32_t*>(acc1 + plan.m_eff);
  // phase-2/3 view of the region (aliases phase 1)
  uint32_t* stab_keys = reinterpret_cast<uint32_t*>(region);
  int32_t* stab_vals = reinterpret_cast<int32_t*>(stab_keys + plan.tab_cap);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t K = a.k, M = a.m, N = a.how_many;
  const bool neighbors_mode = a.out_sess != nullptr;

  for (;;) {
    __syncthreads();
    if (tid == 0) S.q = atomicAdd(ws.counter, 1u);
    __syncthreads();
    const uint32_t q = S.q;
    if (q >= a.n_q) break;

    // ------------------------------------------------------------------ phase 0
    const uint32_t qb = a.q_off[q];
    const uint32_t Lfull = a.q_off[q + 1] - qb;
    const uint32_t L = Lfull > (uint32_t)kMaxSessionLen ? 0u : Lfull;   // over-long sessions are rejected host-side
    if (tid < (int)L) S.q_item[tid] = a.q_items[qb + (L - 1 - tid)];
    __syncthreads();
    uint32_t my_idx = kEmpty;
    bool distinct = false;
    if (tid < (int)L) {
      const uint64_t it = S.q_item[tid];
      distinct = true;
      for (int t = 0; t < tid; ++t) if (S.q_item[t] == it) { distinct = false; break; }
      if (distinct) my_idx = lookup_item(ix, it);
    }
    const uint32_t u = (uint32_t)__syncthreads_count(distinct);          // unique items incl. unknown (:335-339)
    // compact distinct known items in position order
    {
      int flag = (my_idx != kEmpty) ? 1 : 0, total;
      int pos = block_excl_scan(flag, S.scan, total);
      if (flag) { S.d_idx[pos] = my_idx; S.d_pos[pos] = (uint32_t)tid; }
      if (tid == 0) S.nd = (uint32_t)total;
    }
    __syncthreads();
    const uint32_t nd = S.nd;
    // most recent item: removed from the result (mod.rs:157-160); attributes for the adult rule (:186)
    const uint32_t last_idx = (nd > 0 && S.d_pos[0] == 0) ? S.d_idx[0] : kEmpty;
    const uint32_t cur_attr = (a.biz && last_idx != kEmpty) ? ix.attr[last_idx] : 0u;

    uint32_t nn = 0;                 // number of neighbours
    uint32_t postings_visited = 0;

    if (nd > 0 && K > 0 && M > 0 && (N > 0 || neighbors_mode)) {
      // ---------------------------------------------------------------- phase 1
      const uint2 ref0 = ix.post_ref[S.d_idx[0]];
      const uint32_t n0 = min(ref0.y, M);
      const uint32_t* P0 = ix.postings + (size_t)ref0.x * 4;
      const uint32_t c0 = L - S.d_pos[0];
      postings_visited = n0;
      if (nd == 1) {
        // single distinct known item: S = first m postings, all similarities equal → N = first k
        nn = min(n0, K);
        for (uint32_t i = tid; i < nn; i += kThreads) { nbr_sid[i] = P0[i]; nbr_val[i] = (int32_t)c0; }
      } else {
        uint64_t* acc = acc0;
        uint64_t* out = acc1;
        for (uint32_t i = tid; i < n0; i += kThreads) acc[i] = ((uint64_t)P0[i] << 32) | c0;
        uint32_t na = n0;
        for (uint32_t j = 1; j < nd; ++j) {
          const uint2 ref = ix.post_ref[S.d_idx[j]];
          const uint32_t nb = min(min(ref.y, M), plan.list_cap);
          const uint32_t* Pj = ix.postings + (size_t)ref.x * 4;
          const uint32_t cj = L - S.d_pos[j];
          postings_visited += nb;
          for (uint32_t i = tid; i < nb; i += kThreads) listbuf[i] = Pj[i];
          __syncthreads();
          // merge-path fold: out ← first M distinct of acc ∪ B, numerators summed
          const uint32_t T = na + nb;
          uint32_t out_count = 0;
          for (uint32_t base = 0; base < T && out_count < M; base += kTile) {
            const uint32_t d0 = min(base + (uint32_t)tid * kVT, T);
            const uint32_t d1 = min(d0 + kVT, T);
            uint32_t lo = d0 > nb ? d0 - nb : 0, hi = min(d0, na);
            while (lo < hi) {
              const uint32_t mid = (lo + hi) >> 1;
              if ((uint32_t)(acc[mid] >> 32) >= listbuf[d0 - 1 - mid]) lo = mid + 1; else hi = mid;
            }
            uint32_t ai = lo, bi = d0 - lo;
            uint64_t r[kVT];
            uint32_t vmask = 0;
#pragma unroll
            for (int s = 0; s < kVT; ++s) {
              r[s] = 0;
              if (d0 + s < d1) {
                const uint64_t av = ai < na ? acc[ai] : 0ull;
                const uint32_t ak = (uint32_t)(av >> 32);
                const uint32_t bk = bi < nb ? listbuf[bi] : 0u;
                const bool takeA = (ai < na) && (bi >= nb || ak >= bk);
                if (takeA) {
                  r[s] = av + ((bi < nb && bk == ak) ? cj : 0u);
                  vmask |= 1u << s; ++ai;
                } else {
                  const bool dup = ai > 0 && (uint32_t)(acc[ai - 1] >> 32) == bk;
                  if (!dup) { r[s] = ((uint64_t)bk << 32) | cj; vmask |= 1u << s; }
                  ++bi;
                }
              }
            }
            int total;
            uint32_t p = out_count + (uint32_t)block_excl_scan(__popc(vmask), S.scan, total);
#pragma unroll
            for (int s = 0; s < kVT; ++s) {
              if ((vmask >> s) & 1u) { if (p < M) out[p] = r[s]; ++p; }
            }
            out_count = min(M, out_count + (uint32_t)total);
          }
          __syncthreads();
          uint64_t* t = acc; acc = out; out = t;
          na = out_count;
        }
        // -------------------------------------------------------------- phase 1b
        if (na <= K) {
          nn = na;
          for (uint32_t i = tid; i < na; i += kThreads) {
            const uint64_t e = acc[i];
            nbr_sid[i] = (uint32_t)(e >> 32); nbr_val[i] = (int32_t)(uint32_t)e;
          }
        } else {
          // v* = max v with count(num >= v) >= K  (numerators are >= 1)
          const uint32_t E = (na + kThreads - 1) / kThreads;     // contiguous chunk per thread keeps recency order
          const uint32_t e0 = min((uint32_t)tid * E, na), e1 = min(e0 + E, na);
          uint32_t vlo = 1, vhi = L * (L + 1) / 2;
          while (vlo < vhi) {
            const uint32_t v = (vlo + vhi + 1) >> 1;
            int c = 0;
            for (uint32_t i = e0; i < e1; ++i) c += ((uint32_t)acc[i] >= v);
            if ((uint32_t)block_sum(c, S.scan) >= K) vlo = v; else vhi = v - 1;
          }
          const uint32_t vstar = vlo;
          int cg = 0, ce = 0;
          for (uint32_t i = e0; i < e1; ++i) { const uint32_t nm = (uint32_t)acc[i]; cg += nm > vstar; ce += nm == vstar; }
          int tot_g, tot_e;
          const int pre_g = block_excl_scan(cg, S.scan, tot_g);
          int pre_e = block_excl_scan(ce, S.scan, tot_e);
          const uint32_t quota = K - (uint32_t)tot_g;            // ties at v*: the `quota` most recent win
          uint32_t gpos = (uint32_t)pre_g;
          for (uint32_t i = e0; i < e1; ++i) {
            const uint64_t e = acc[i];
            const uint32_t nm = (uint32_t)e;
            if (nm > vstar) {
              nbr_sid[gpos] = (uint32_t)(e >> 32); nbr_val[gpos] = (int32_t)nm; ++gpos;
            } else if (nm == vstar) {
              if ((uint32_t)pre_e < quota) {
                const uint32_t p = (uint32_t)tot_g + (uint32_t)pre_e;
                nbr_sid[p] = (uint32_t)(e >> 32); nbr_val[p] = (int32_t)nm;
              }
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
        const uint32_t sid = nbr_sid[i]; const int32_t nm = nbr_val[i];
        uint32_t rank = 0;
        for (uint32_t t = 0; t < nn; ++t) {
          const int32_t on = nbr_val[t];
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

    // ------------------------------------------------------------------ phase 2
    int my_len = 0;
    for (uint32_t i = tid; i < nn; i += kThreads) {
      const uint2 r = ix.sess_ref[nbr_sid[i]];
      nbr_ref[i] = r;
      my_len += (int)r.y;
    }
    const uint32_t total_items = (uint32_t)block_sum(my_len, S.scan);   // also orders phase-1 reads before table init
    uint32_t* tkeys; int32_t* tvals; uint32_t tcap;
    if (total_items * 2 <= plan.tab_cap) { tkeys = stab_keys; tvals = stab_vals; tcap = plan.tab_cap; }
    else {
      tkeys = ws.gtab_keys + (size_t)blockIdx.x * ws.gtab_cap; tvals = ws.gtab_vals + (size_t)blockIdx.x * ws.gtab_cap;
      tcap = ws.gtab_cap;
    }
    const uint32_t tmask = tcap - 1;
    for (uint32_t i = tid; i < tcap; i += kThreads) { tkeys[i] = kEmpty; tvals[i] = 0; }
    __syncthreads();
    for (uint32_t i = tid; i < nn; i += kThreads) {
      const uint2 r = nbr_ref[i];
      const uint32_t* items = ix.sess_items + (size_t)r.x * 4;
      uint32_t pmin = 0xFFFFFFFFu;                                 // first match, most recent first (mod.rs:133-138)
      for (uint32_t t = 0; t < r.y; ++t) {
        const uint32_t it = items[t];
        for (uint32_t j = 0; j < nd; ++j) if (S.d_idx[j] == it) pmin = min(pmin, S.d_pos[j]);
      }
      const uint32_t p = pmin + 1;                                 // 1-based position (mod.rs:140)
      const int32_t w10 = (pmin != 0xFFFFFFFFu && p < 100) ? 10 - (int32_t)p : 0;   // linear_score ×10 (mod.rs:110-116)
      const int32_t v = w10 * nbr_val[i];
      for (uint32_t t = 0; t < r.y; ++t) table_add(tkeys, tvals, tmask, items[t], v);
    }
    __syncthreads();

    // ------------------------------------------------------------------ phase 3
    const double denom = (double)(10u * u);
    uint32_t written = 0;
    Elem bound; bound.s = ~0ull; bound.id = 0;                      // exclusive upper bound of the current round
    bool first_round = true;
    while (written < N) {
      Elem top; top.s = 0; top.id = kEmpty;
      for (uint32_t base = warp * 32; base < tcap; base += kWarps * 32) {
        const uint32_t slot = base + lane;
        Elem c; c.s = 0; c.id = kEmpty;
        const uint32_t key = tkeys[slot];
        if (key != kEmpty && key != last_idx) {
          bool ok = true;
          if (a.biz) ok = passes_business_rules(cur_attr, ix.attr[key]);
          if (ok) {
            const double idf = ix.idf[key];
            const double g = idf > 0.0 ? idf : 1.0;                // mod.rs:145-152
            const double score = g * (double)tvals[slot] / denom;
            c.s = score_bits(score); c.id = key;
            if (!first_round && !better(bound, c)) { c.s = 0; c.id = kEmpty; }
          }
        }
        const Elem worst = shfl_elem(top, 31);
        if (__any_sync(kFull, better(c, worst))) {
          c = warp_sort_desc(c, lane);
          top = warp_merge_top(top, c, lane);
        }
      }
      S.topbuf[warp * 32 + lane] = top;
      __syncthreads();
      uint32_t emitted = 0;
      if (warp == 0) {
        Elem best = S.topbuf[lane];
        for (int w = 1; w < kWarps; ++w) best = warp_merge_top(best, S.topbuf[w * 32 + lane], lane);
        const uint32_t valid = __popc(__ballot_sync(kFull, best.id != kEmpty));
        const uint32_t take = min(valid, N - written);
        if ((uint32_t)lane < take) {
          a.out_ids[(size_t)q * N + written + lane] = ix.item_key[best.id];
          a.out_scores[(size_t)q * N + written + lane] = bits_score(best.s);
        }
        if (lane == 0) { S.nbr_count = take; }
        if (take > 0) { const Elem lastE = shfl_elem(best, (int)take - 1); if (lane == 0) S.topbuf[0] = lastE; }
      }
      __syncthreads();
      emitted = S.nbr_count;
      if (emitted > 0) bound = S.topbuf[0];
      written += emitted;
      first_round = false;
      __syncthreads();
      if (emitted < 32) break;
    }
    for (uint32_t i = written + tid; i < N; i += kThreads) {           // deterministic padding
      a.out_ids[(size_t)q * N + i] = 0; a.out_scores[(size_t)q * N + i] = 0.0;
    }
    if (tid == 0) {
      a.out_counts[q] = written;
      if (a.out_stats) {
        vmis_query_stats_t st; st.postings_visited = postings_visited; st.n_neighbors = nn;
        st.neighbor_items = total_items; st.n_out = written;
        a.out_stats[q] = st;
      }
    }
  }
}

uint32_t next_pow2(uint32_t x) { uint32_t p = 1; while (p < x) p <<= 1; return p; }

}  // namespace

int plan_launch(const IndexView& ix, uint32_t k, uint32_t m, int sm_count, LaunchPlan* plan) {
  if (k > kMaxK || m > kMaxM) return VMIS_ERR_LIMIT;
  LaunchPlan p{};
  p.m_eff = (std::max(m, 1u) + 3u) & ~3u;
  p.list_cap = (std::min(std::max(m, 1u), std::max(ix.m_build, 1u)) + 3u) & ~3u;
  uint32_t tab = next_pow2(std::max(k, 1u) * 12u);
  p.tab_cap = std::min(std::max(tab, 1024u), 8192u);
  const size_t fixed = (sizeof(SmemLayout) + 15) & ~size_t(15);
  const size_t nbr = ((size_t(k) * 8 + 15) & ~size_t(15)) + size_t(k) * 8 + 16;
  const size_t r1 = size_t(p.m_eff) * 16 + size_t(p.list_cap) * 4;
  const size_t r2 = size_t(p.tab_cap) * 8;
  const size_t total = fixed + nbr + std::max(r1, r2) + 16;
  if (total > 227 * 1024) return VMIS_ERR_LIMIT;
  p.smem_bytes = (uint32_t)total;
  int per_sm = (int)std::min<size_t>(4, (227 * 1024) / total);
  if (per_sm < 1) per_sm = 1;
  p.grid = (uint32_t)(sm_count * per_sm);
  p.gtab_cap = next_pow2(std::max(2u * std::max(k, 1u) * std::max(ix.max_len, 1u), 1024u));
  *plan = p;
  return VMIS_OK;
}

size_t workspace_bytes(const LaunchPlan& plan) {
  return 256 + size_t(plan.grid) * plan.gtab_cap * 8;
}

Workspace carve_workspace(void* base, const LaunchPlan& plan) {
  Workspace ws{};
  unsigned char* b = static_cast<unsigned char*>(base);
  ws.counter = reinterpret_cast<uint32_t*>(b);
  ws.gtab_keys = reinterpret_cast<uint32_t*>(b + 256);
  ws.gtab_vals = reinterpret_cast<int32_t*>(b + 256 + size_t(plan.grid) * plan.gtab_cap * 4);
  ws.gtab_cap = plan.gtab_cap;
  ws.grid = plan.grid;
  return ws;
}

cudaError_t launch_predict(const IndexView& ix, const PredictArgs& args, const LaunchPlan& plan, const Workspace& ws,
                           cudaStream_t stream) {
  if (args.n_q == 0) return cudaSuccess;
  cudaError_t e = cudaFuncSetAttribute(vmis_predict_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)plan.smem_bytes);
  if (e != cudaSuccess) return e;
  e = cudaMemsetAsync(ws.counter, 0, sizeof(uint32_t), stream);
  if (e != cudaSuccess) return e;
  const uint32_t grid = std::min(plan.grid, args.n_q);
  vmis_predict_kernel<<<grid, kThreads, plan.smem_bytes, stream>>>(ix, args, plan, ws);
  return cudaGetLastError();
}

}  // namespace vmis
