// serenade_b200/csrc/capi.cu — the C ABI of include/vmis.h: index handle (host
// mirror + HBM arrays), pooled per-call contexts (stream, staging buffers, kernel
// workspace) and the batched entry points.  No CPU fallback anywhere: every query
// entry point ends in launch_predict() or fails.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "avro_reader.h"
#include "build_device.h"
#include "vmis_host.h"

namespace {

thread_local std::string g_err;
thread_local int g_err_code = 0;

int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
  for (char* c = buf; *c; ++c) if ((unsigned char)*c < 0x20 || (unsigned char)*c > 0x7E) *c = '?';   // messages may quote file content
  g_err = buf;
  g_err_code = code;
  return code;
}
#define CU_TRY(expr)                                                                     \
  do {                                                                                   \
    cudaError_t e__ = (expr);                                                            \
    if (e__ != cudaSuccess) return fail(VMIS_ERR_CUDA, "%s: %s", #expr, cudaGetErrorString(e__)); \
  } while (0)

struct CallCtx {
  cudaStream_t stream = nullptr;
  cudaEvent_t done = nullptr;      // last use of ws / buf
  void* buf = nullptr; size_t buf_cap = 0;
  void* ws = nullptr; size_t ws_cap = 0;
  uint32_t ws_grid = 0, ws_gtab_cap = 0, ws_gtab_huge = 0;   // geometry the workspace tables were initialised for
  void* pinned = nullptr; size_t pinned_cap = 0;   // host staging of small batches (one copy each way)
  void* pinned_dev = nullptr;                      // device view of `pinned` (mapped: tiny batches skip the copies)
  ~CallCtx() {
    if (buf) cudaFree(buf);
    if (ws) cudaFree(ws);
    if (pinned) cudaFreeHost(pinned);
    if (done) cudaEventDestroy(done);
    if (stream) cudaStreamDestroy(stream);
  }
};

}  // namespace

namespace vmis {
void set_last_error(int code, const char* msg) {
  if (code == VMIS_OK) { g_err.clear(); g_err_code = 0; }
  else fail(code, "%s", msg ? msg : "");
}
}  // namespace vmis

struct vmis_index {
  int device = 0;
  int sm_count = 0;
  vmis::Sessions sessions;          // host mirror: session_to_items_sorted / session_to_max_time_stamp
  vmis::FlatIndex flat;             // item dictionary, idf, attr stay resident for the accessors
  vmis::IndexView view{};
  std::vector<void*> dev_allocs;
  uint64_t device_bytes = 0;
  uint64_t n_sessions_kept = 0;
  uint32_t shard = 0;                        // item-sharded postings: the shard this handle owns
  uint64_t synth_interactions = 0;           // vmis_index_synth: interactions generated on the device
  uint64_t n_post_entries = 0;               // u32 entries of this handle's posting shard (incl. padding)
  uint64_t n_sess_item_entries = 0;          // u32 entries of sess_items (incl. padding)
  std::vector<void*> ipc_mapped;             // peer shards opened through CUDA IPC
  vmis_prebuilt_info_t prebuilt{};           // pre-computed (Avro / parts) index: load and normalisation report
  std::mutex mu;
  std::vector<std::unique_ptr<CallCtx>> pool;
};

namespace {

// IndexView::g32 from the uploaded idf array (one small kernel); owned by the handle like the other arrays
int attach_g32(vmis_index* ix);

template <class T>
int upload(vmis_index* ix, const std::vector<T>& v, const T** out) {
  void* d = nullptr;
  const size_t bytes = std::max<size_t>(v.size() * sizeof(T), 16);
  CU_TRY(cudaMalloc(&d, bytes));
  ix->dev_allocs.push_back(d);
  ix->device_bytes += bytes;
  if (!v.empty()) CU_TRY(cudaMemcpy(d, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
  *out = static_cast<const T*>(d);
  return VMIS_OK;
}

int select_device(int device, int* sm_count) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0)
    return fail(VMIS_ERR_CUDA, "no CUDA device available (%s); this library has no CPU fallback",
                e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
  if (device < 0 || device >= n) return fail(VMIS_ERR_ARG, "device %d out of range [0,%d)", device, n);
  cudaDeviceProp p;
  CU_TRY(cudaGetDeviceProperties(&p, device));
  if (p.major < 10) return fail(VMIS_ERR_CUDA, "device %d is sm_%d%d; kernels are built for sm_100a only", device, p.major, p.minor);
  CU_TRY(cudaSetDevice(device));
  *sm_count = p.multiProcessorCount;
  return VMIS_OK;
}

vmis_index* adopt_device_index(std::unique_ptr<vmis_index> ix, vmis::DeviceIndexArrays& A, size_t m, size_t max_len,
                               double idf_w, int device, uint32_t shard, uint32_t n_shards);
vmis_index* upload_flat(std::unique_ptr<vmis_index> ix, int device, uint32_t shard, uint32_t n_shards);

// Sessions held on the host → index.  With a device the build itself runs there (build_sm100.cu: ~0.1 s for 60 M
// interactions instead of ~10 s on one host core; the reference rebuilds the index for every HPO trial,
// objective.rs:17); the host builder serves VMIS_DEVICE_NONE handles, max_len > 128 and VMIS_BUILD=host.
vmis_index* finish_index(std::unique_ptr<vmis_index> ix, size_t m, size_t max_len, double idf_w, int device,
                         uint32_t shard = 0, uint32_t n_shards = 1) {
  std::string err;
  if (n_shards == 0 || shard >= n_shards) { fail(VMIS_ERR_ARG, "shard %u out of range [0,%u)", shard, n_shards); return nullptr; }
  if (max_len == 0) max_len = vmis::session_length_p99_5(ix->sessions);
  const char* how = std::getenv("VMIS_BUILD");
  if (device != VMIS_DEVICE_NONE && max_len <= 128 && m >= 1 && !(how && !std::strcmp(how, "host")) && ix->sessions.size() > 0) {
    if (select_device(device, &ix->sm_count) != VMIS_OK) return nullptr;
    const vmis::Sessions& S = ix->sessions;
    void *d_items = nullptr, *d_off = nullptr, *d_ts = nullptr;
    bool ok = cudaMalloc(&d_items, std::max<size_t>(S.items.size() * 8, 16)) == cudaSuccess &&
              cudaMalloc(&d_off, S.off.size() * 8) == cudaSuccess && cudaMalloc(&d_ts, std::max<size_t>(S.ts.size() * 4, 16)) == cudaSuccess &&
              cudaMemcpy(d_items, S.items.data(), S.items.size() * 8, cudaMemcpyHostToDevice) == cudaSuccess &&
              cudaMemcpy(d_off, S.off.data(), S.off.size() * 8, cudaMemcpyHostToDevice) == cudaSuccess &&
              cudaMemcpy(d_ts, S.ts.data(), S.ts.size() * 4, cudaMemcpyHostToDevice) == cudaSuccess;
    vmis::DeviceIndexArrays A;
    if (ok) {
      vmis::DeviceSessions ds; ds.items = (const uint64_t*)d_items; ds.off = (const uint64_t*)d_off; ds.ts = (const uint32_t*)d_ts;
      ds.n_sessions = S.size(); ds.n_entries = S.items.size();
      ok = vmis::build_index_device(ds, m, max_len, idf_w, shard, n_shards, &A, &err);
    } else err = cudaGetErrorString(cudaGetLastError());
    cudaFree(d_items); cudaFree(d_off); cudaFree(d_ts);
    if (ok) return adopt_device_index(std::move(ix), A, m, max_len, idf_w, device, shard, n_shards);
    if (err.find("duplicate item") != std::string::npos || err.find("no training session") != std::string::npos) {
      fail(VMIS_ERR_ARG, "%s", err.c_str());
      return nullptr;
    }
    cudaGetLastError();   // e.g. out of memory during the sorts: fall through to the host builder
  }
  if (!vmis::build_flat_index(ix->sessions, m, max_len, idf_w, n_shards, &ix->flat, &err)) { fail(VMIS_ERR_ARG, "%s", err.c_str()); return nullptr; }
  return upload_flat(std::move(ix), device, shard, n_shards);
}

// FlatIndex (host) → HBM; VMIS_DEVICE_NONE keeps a host-only handle
vmis_index* upload_flat(std::unique_ptr<vmis_index> ix, int device, uint32_t shard, uint32_t n_shards) {
  ix->shard = shard;
  ix->n_sessions_kept = ix->flat.rank_to_orig.size();
  ix->device = device;
  if (device == VMIS_DEVICE_NONE) return ix.release();   // host-only handle: accessors work, queries fail
  if (select_device(device, &ix->sm_count) != VMIS_OK) return nullptr;
  vmis::FlatIndex& F = ix->flat;
  vmis::IndexView& V = ix->view;
  // only this handle's shard of the postings goes to its HBM; peers are attached later over NVLink
  std::vector<uint32_t> own(F.postings.begin() + F.shard_begin[shard], F.postings.begin() + F.shard_begin[shard + 1]);
  const uint32_t* own_dev = nullptr;
  for (int s2 = 0; s2 < vmis::kMaxShards; ++s2) V.post_shard[s2] = nullptr;
  V.n_shards = n_shards;
  if (upload(ix.get(), F.item_key, &V.item_key) || upload(ix.get(), F.item_hash, &V.item_hash) ||
      upload(ix.get(), F.post_ref, &V.post_ref) || upload(ix.get(), own, &own_dev) ||
      upload(ix.get(), F.sess_ref, &V.sess_ref) || upload(ix.get(), F.sess_items, &V.sess_items) ||
      upload(ix.get(), F.idf, &V.idf) || upload(ix.get(), F.attr, &V.attr) ||
      upload(ix.get(), F.rank_to_orig, &V.rank_to_orig)) {
    for (void* d : ix->dev_allocs) cudaFree(d);
    return nullptr;
  }
  V.post_shard[shard] = own_dev;
  V.item_hash_mask = (uint32_t)F.item_hash.size() - 1;
  V.n_items = (uint32_t)F.item_key.size();
  V.n_kept = (uint32_t)F.rank_to_orig.size();
  V.m_build = F.m_build;
  V.m_carry = F.m_carry;
  V.max_len = F.max_len;
  if (attach_g32(ix.get())) { for (void* d : ix->dev_allocs) cudaFree(d); return nullptr; }
  ix->n_post_entries = own.size();
  ix->n_sess_item_entries = F.sess_items.size();
  // the big CSR arrays now live in HBM only
  std::vector<uint32_t>().swap(F.postings);
  std::vector<uint32_t>().swap(F.sess_items);
  std::vector<uint2>().swap(F.sess_ref);
  return ix.release();
}

// wrap the arrays of an on-device build into a handle (ix->sessions, if any, stays as the host mirror)
vmis_index* adopt_device_index(std::unique_ptr<vmis_index> ix, vmis::DeviceIndexArrays& A, size_t m, size_t max_len,
                               double idf_w, int device, uint32_t shard, uint32_t n_shards) {
  vmis::FlatIndex& F = ix->flat;
  F.item_key.swap(A.host_item_key); F.item_hash.swap(A.host_item_hash); F.idf.swap(A.host_idf);
  F.attr.assign(A.n_items, (uint8_t)(VMIS_ATTR_EXISTS | VMIS_ATTR_FOR_SALE));
  F.n_pairs_kept = A.n_pairs_kept; F.n_postings = A.n_postings; F.n_shards = n_shards;
  F.m_build = (uint32_t)std::min<size_t>(m, 0xFFFFFFFFu); F.m_carry = F.m_build; F.max_len = (uint32_t)max_len; F.idf_weighting = idf_w;
  ix->n_sessions_kept = A.n_kept; ix->device = device; ix->shard = shard;
  vmis::IndexView& V = ix->view;
  for (int s2 = 0; s2 < vmis::kMaxShards; ++s2) V.post_shard[s2] = nullptr;
  V.item_key = A.item_key; V.item_hash = A.item_hash; V.item_hash_mask = (uint32_t)(A.item_hash_cap - 1);
  V.post_ref = A.post_ref; V.post_shard[shard] = A.postings; V.n_shards = n_shards;
  V.sess_ref = A.sess_ref; V.sess_items = A.sess_items; V.idf = A.idf; V.attr = A.attr; V.rank_to_orig = A.rank_to_orig;
  V.n_items = (uint32_t)A.n_items; V.n_kept = (uint32_t)A.n_kept; V.m_build = F.m_build; V.m_carry = F.m_carry; V.max_len = F.max_len;
  void* owned[] = {A.item_key, A.item_hash, A.post_ref, A.postings, A.sess_ref, A.sess_items, A.idf, A.attr, A.rank_to_orig};
  for (void* p : owned) ix->dev_allocs.push_back(p);
  ix->n_post_entries = A.shard_entries;
  ix->n_sess_item_entries = A.sess_items_entries;
  ix->device_bytes = A.n_items * 8 + A.item_hash_cap * sizeof(vmis::ItemHashEntry) + A.n_items * 8 + A.shard_entries * 4 +
                     A.n_kept * 8 + A.sess_items_entries * 4 + A.n_items * 9 + A.n_kept * 4;
  if (attach_g32(ix.get())) { for (void* d : ix->dev_allocs) cudaFree(d); return nullptr; }
  return ix.release();
}

int attach_g32(vmis_index* ix) {
  void* d = nullptr;
  const size_t n = ix->view.n_items;
  CU_TRY(cudaMalloc(&d, std::max<size_t>(n * sizeof(float), 16)));
  ix->dev_allocs.push_back(d);
  ix->device_bytes += n * sizeof(float);
  CU_TRY(vmis::build_g32(ix->view.idf, static_cast<float*>(d), (uint32_t)n, nullptr));
  CU_TRY(cudaDeviceSynchronize());
  ix->view.g32 = static_cast<const float*>(d);
  return VMIS_OK;
}

// borrow a call context; its previous work (possibly on a caller stream) is ordered before ours
int acquire_ctx(vmis_index* ix, std::unique_ptr<CallCtx>* out) {
  {
    std::lock_guard<std::mutex> g(ix->mu);
    if (!ix->pool.empty()) { *out = std::move(ix->pool.back()); ix->pool.pop_back(); return VMIS_OK; }
  }
  std::unique_ptr<CallCtx> c(new CallCtx());
  CU_TRY(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  CU_TRY(cudaEventCreateWithFlags(&c->done, cudaEventDisableTiming));
  *out = std::move(c);
  return VMIS_OK;
}
void release_ctx(vmis_index* ix, std::unique_ptr<CallCtx> c) {
  std::lock_guard<std::mutex> g(ix->mu);
  ix->pool.push_back(std::move(c));
}
int ensure(void** p, size_t* cap, size_t need) {
  if (*cap >= need) return VMIS_OK;
  if (*p) { CU_TRY(cudaFree(*p)); *p = nullptr; *cap = 0; }
  need = (need + (size_t(1) << 20)) & ~((size_t(1) << 20) - 1);
  CU_TRY(cudaMalloc(p, need));
  *cap = need;
  return VMIS_OK;
}

int check_common(const vmis_index* ix, uint32_t k, uint32_t m, vmis::LaunchPlan* plan) {
  if (!ix) return fail(VMIS_ERR_ARG, "index is NULL");
  if (ix->device == VMIS_DEVICE_NONE)
    return fail(VMIS_ERR_CUDA, "host-only index (VMIS_DEVICE_NONE): queries need a B200; there is no CPU fallback");
  for (uint32_t s2 = 0; s2 < ix->view.n_shards; ++s2)
    if (!ix->view.post_shard[s2]) return fail(VMIS_ERR_ARG, "posting shard %u of %u is not attached", s2, ix->view.n_shards);
  if (ix->view.n_kept >= 0x80000000u)     // the list merges compare time ranks as signed numbers (kEmpty sentinel = -1)
    return fail(VMIS_ERR_LIMIT, "%u kept sessions; the kernel handles fewer than 2^31", ix->view.n_kept);
  std::string why;
  const int rc = vmis::plan_launch(ix->view, k, m, ix->sm_count, plan, &why);
  if (rc != VMIS_OK) return fail(rc, "%s", why.c_str());
  return VMIS_OK;
}

// Enqueue one batch whose buffers are all on the device.
int run_device(vmis_index* ix, CallCtx* c, const vmis::PredictArgs& args, const vmis::LaunchPlan& plan, cudaStream_t stream) {
  const size_t old_cap = c->ws_cap;
  const int rc = ensure(&c->ws, &c->ws_cap, vmis::workspace_bytes(plan));
  if (rc) return rc;
  CU_TRY(cudaStreamWaitEvent(stream, c->done, 0));
  const vmis::Workspace ws = vmis::carve_workspace(c->ws, plan);
  if (c->ws_cap != old_cap || c->ws_grid != plan.grid || c->ws_gtab_cap != plan.gtab_cap || c->ws_gtab_huge != plan.gtab_huge) {
    CU_TRY(vmis::init_workspace(ws, stream));
    c->ws_grid = plan.grid; c->ws_gtab_cap = plan.gtab_cap; c->ws_gtab_huge = plan.gtab_huge;
  }
  CU_TRY(vmis::launch_predict(ix->view, args, plan, ws, stream));
  CU_TRY(cudaEventRecord(c->done, stream));
  return VMIS_OK;
}

int host_batch(const vmis_index* cix, const uint64_t* q_items, const uint32_t* q_off, uint32_t n_q, uint32_t k, uint32_t m,
               uint32_t how_many, int biz, uint64_t* out_ids, double* out_scores, uint32_t* out_counts,
               uint32_t* out_sess, double* out_sim, void* stream_) {
  vmis_index* ix = const_cast<vmis_index*>(cix);
  vmis::LaunchPlan plan;
  int rc = check_common(ix, k, m, &plan);
  if (rc) return rc;
  if (n_q == 0) return VMIS_OK;
  if (!q_items || !q_off || !out_counts) return fail(VMIS_ERR_ARG, "NULL query / output buffer");
  const bool nb_mode = out_sess != nullptr;
  if (!nb_mode && how_many > 0 && (!out_ids || !out_scores)) return fail(VMIS_ERR_ARG, "NULL output buffer");
  for (uint32_t q = 0; q < n_q; ++q) {
    if (q_off[q + 1] < q_off[q]) return fail(VMIS_ERR_ARG, "q_off not monotone at %u", q);
    if (q_off[q + 1] - q_off[q] > (uint32_t)vmis::kMaxSessionLen)
      return fail(VMIS_ERR_LIMIT, "evolving session %u has %u items; kernel limit is %d", q, q_off[q + 1] - q_off[q], vmis::kMaxSessionLen);
  }
  CU_TRY(cudaSetDevice(ix->device));
  // The batch is cut into chunks that flow through a small ring of call contexts (own stream, staging buffers and
  // kernel workspace each): the H2D copy of chunk i+1 and the D2H copy of chunk i-1 overlap the kernel of chunk i.
  // With a caller-provided stream everything is enqueued there, in order.
  // chunk of the pipeline (tuning knob VMIS_CHUNK_LOG2, default 2^15; same-box e2e at 2^17 / 2^16 / 2^15: 24.33 / 24.54 / 24.65 M qps): smaller chunks shorten the fill / drain of the
  // copy-kernel-copy pipeline, larger ones amortise the launches
  static const uint32_t kChunk = [] {
    const char* e = std::getenv("VMIS_CHUNK_LOG2");
    const int l = e ? std::atoi(e) : 15;
    return 1u << (l < 10 ? 10 : l > 20 ? 20 : l);
  }();
  // streams in the ring (tuning knob VMIS_PIPE, default 3)
  static const size_t kPipe = [] { const char* e = std::getenv("VMIS_PIPE"); const int p = e ? std::atoi(e) : 3; return (size_t)(p < 1 ? 1 : p > 8 ? 8 : p); }();
  const uint32_t n_chunks = (n_q + kChunk - 1) / kChunk;
  const size_t n_ctx = stream_ ? 1 : std::min<size_t>(kPipe, n_chunks);
  const size_t width = nb_mode ? k : how_many;
  auto al = [](size_t x) { return (x + 255) & ~size_t(255); };
  size_t max_items = 0;
  for (uint32_t ch = 0; ch < n_chunks; ++ch) {
    const uint32_t c0 = ch * kChunk, c1 = std::min(n_q, c0 + kChunk);
    max_items = std::max<size_t>(max_items, q_off[c1] - q_off[c0]);
  }
  const uint32_t max_q = std::min(n_q, kChunk);
  const size_t o_items = 0, o_off = o_items + al(max_items * 8), o_ids = o_off + al((size_t(max_q) + 1) * 4);
  const size_t o_sc = o_ids + al(size_t(max_q) * width * 8), o_cnt = o_sc + al(size_t(max_q) * width * 8);
  const size_t total = o_cnt + al(size_t(max_q) * 4);
  std::vector<std::unique_ptr<CallCtx>> ring(n_ctx);
  auto body = [&]() -> int {
    for (auto& c : ring) {
      int r = acquire_ctx(ix, &c);
      if (r) return r;
      r = ensure(&c->buf, &c->buf_cap, total);
      if (r) return r;
    }
    // latency path: a small batch goes through one mapped pinned buffer.  The kernel writes every result row straight
    // into it (posted writes over PCIe, no copy command) and raises the row's count last, behind a system-wide fence;
    // the caller polls the counts and copies finished rows to the output buffers WHILE the kernel is still running, so
    // neither a device-to-host copy nor a stream synchronisation nor the final memcpy is on the critical path.
    // A handful of sessions (the reference's own call shape, mod.rs:118-125) are also READ from the mapped buffer.
    if (!stream_ && n_chunks == 1 && total <= (size_t(1) << 20)) {
      CallCtx* c = ring[0].get();
      if (c->pinned_cap < total) {
        if (c->pinned) { CU_TRY(cudaFreeHost(c->pinned)); c->pinned = nullptr; c->pinned_cap = 0; }
        CU_TRY(cudaHostAlloc(&c->pinned, size_t(1) << 20, cudaHostAllocMapped));
        c->pinned_cap = size_t(1) << 20;
        CU_TRY(cudaHostGetDevicePointer(&c->pinned_dev, c->pinned, 0));
      }
      unsigned char* b = static_cast<unsigned char*>(c->buf);
      unsigned char* p = static_cast<unsigned char*>(c->pinned);
      unsigned char* pd = static_cast<unsigned char*>(c->pinned_dev);
      const size_t ni = q_off[n_q];
      if (ni) std::memcpy(p + o_items, q_items, ni * 8);
      std::memcpy(p + o_off, q_off, (size_t(n_q) + 1) * 4);
      const bool mapped = pd != nullptr;
      // more sessions: one H2D copy instead of two PCIe reads per CTA (same-box A/B on launches of 1024: 16 / 256 / 4096
      // sessions as the limit give 9.45 / 9.46 / 8.97 M qps end to end; the lone call is indifferent)
      const bool in_mapped = mapped && n_q <= 16;
      const bool flags = mapped && !nb_mode;                        // find_neighbors rows: plain copy back
      constexpr uint32_t kPending = 0xFFFFFFFEu;                    // never a count (VMIS_COUNT_TOO_LONG is all ones)
      volatile uint32_t* done_flag = reinterpret_cast<volatile uint32_t*>(p + o_cnt);
      if (flags) for (uint32_t q = 0; q < n_q; ++q) done_flag[q] = kPending;
      if (!in_mapped) {
        CU_TRY(cudaStreamWaitEvent(c->stream, c->done, 0));
        CU_TRY(cudaMemcpyAsync(b, p, o_ids, cudaMemcpyHostToDevice, c->stream));
      }
      unsigned char* in = in_mapped ? pd : b;
      unsigned char* out = flags ? pd : b;
      vmis::PredictArgs a{};
      a.q_items = reinterpret_cast<const uint64_t*>(in + o_items);
      a.q_off = reinterpret_cast<const uint32_t*>(in + o_off);
      a.n_q = n_q; a.k = k; a.m = m; a.how_many = how_many; a.biz = biz;
      a.host_flags = flags ? 1 : 0;
      a.out_counts = reinterpret_cast<uint32_t*>(out + o_cnt);
      if (nb_mode) { a.out_sess = reinterpret_cast<uint32_t*>(out + o_ids); a.out_sim = reinterpret_cast<double*>(out + o_sc); }
      else { a.out_ids = reinterpret_cast<uint64_t*>(out + o_ids); a.out_scores = reinterpret_cast<double*>(out + o_sc); }
      int r = run_device(ix, c, a, plan, c->stream);               // records c->done behind the kernel
      if (r) return r;
      if (!flags) {
        CU_TRY(cudaMemcpyAsync(p + o_ids, b + o_ids, total - o_ids, cudaMemcpyDeviceToHost, c->stream));
        CU_TRY(cudaEventRecord(c->done, c->stream));
        CU_TRY(cudaStreamSynchronize(c->stream));
        if (k) { std::memcpy(out_sess, p + o_ids, size_t(n_q) * k * 4); std::memcpy(out_sim, p + o_sc, size_t(n_q) * k * 8); }
        std::memcpy(out_counts, p + o_cnt, size_t(n_q) * 4);
        return VMIS_OK;
      }
      // rows [copied, n_q) are still to be handed over
      uint32_t copied = 0;
      auto hand_over = [&](uint32_t upto) {
        std::atomic_thread_fence(std::memory_order_acquire);
        const size_t w = how_many;
        if (w) {
          std::memcpy(out_ids + size_t(copied) * w, p + o_ids + size_t(copied) * w * 8, size_t(upto - copied) * w * 8);
          std::memcpy(out_scores + size_t(copied) * w, p + o_sc + size_t(copied) * w * 8, size_t(upto - copied) * w * 8);
        }
        for (uint32_t q = copied; q < upto; ++q) out_counts[q] = done_flag[q];
        copied = upto;
      };
      // Poll for a while (a lone query takes ~10 us on the device, 1024 of them ~75 us); a kernel that takes longer, or
      // one that failed and will never raise its flags, is picked up by the stream synchronisation below.
      const auto t_end = std::chrono::steady_clock::now() + std::chrono::microseconds(500);
      for (uint32_t spin = 0; copied < n_q; ++spin) {
        uint32_t q = copied;
        while (q < n_q && done_flag[q] != kPending) ++q;
        if (q == n_q || q - copied >= 128u) { hand_over(q); continue; }
        if ((spin & 63u) == 63u && std::chrono::steady_clock::now() > t_end) break;
#if defined(__x86_64__) || defined(__i386__)
        __builtin_ia32_pause();
#endif
      }
      if (copied < n_q) {
        CU_TRY(cudaStreamSynchronize(c->stream));
        hand_over(n_q);
      }
      return VMIS_OK;
    }
    for (uint32_t ch = 0; ch < n_chunks; ++ch) {
      CallCtx* c = ring[ch % n_ctx].get();
      cudaStream_t stream = stream_ ? static_cast<cudaStream_t>(stream_) : c->stream;
      const uint32_t c0 = ch * kChunk, c1 = std::min(n_q, c0 + kChunk), nq = c1 - c0;
      const size_t i0 = q_off[c0], ni = q_off[c1] - i0;
      unsigned char* b = static_cast<unsigned char*>(c->buf);
      CU_TRY(cudaStreamWaitEvent(stream, c->done, 0));
      if (ni) CU_TRY(cudaMemcpyAsync(b + o_items, q_items + i0, ni * 8, cudaMemcpyHostToDevice, stream));
      CU_TRY(cudaMemcpyAsync(b + o_off, q_off + c0, (size_t(nq) + 1) * 4, cudaMemcpyHostToDevice, stream));
      vmis::PredictArgs a{};
      a.q_items = reinterpret_cast<const uint64_t*>(b + o_items);
      a.q_off = reinterpret_cast<const uint32_t*>(b + o_off);
      a.q_item_base = (uint32_t)i0;
      a.n_q = nq; a.k = k; a.m = m; a.how_many = how_many; a.biz = biz;
      a.out_counts = reinterpret_cast<uint32_t*>(b + o_cnt);
      if (nb_mode) { a.out_sess = reinterpret_cast<uint32_t*>(b + o_ids); a.out_sim = reinterpret_cast<double*>(b + o_sc); }
      else { a.out_ids = reinterpret_cast<uint64_t*>(b + o_ids); a.out_scores = reinterpret_cast<double*>(b + o_sc); }
      int r = run_device(ix, c, a, plan, stream);
      if (r) return r;
      if (nb_mode) {
        if (k) {
          CU_TRY(cudaMemcpyAsync(out_sess + size_t(c0) * k, b + o_ids, size_t(nq) * k * 4, cudaMemcpyDeviceToHost, stream));
          CU_TRY(cudaMemcpyAsync(out_sim + size_t(c0) * k, b + o_sc, size_t(nq) * k * 8, cudaMemcpyDeviceToHost, stream));
        }
      } else if (how_many) {
        CU_TRY(cudaMemcpyAsync(out_ids + size_t(c0) * how_many, b + o_ids, size_t(nq) * how_many * 8, cudaMemcpyDeviceToHost, stream));
        CU_TRY(cudaMemcpyAsync(out_scores + size_t(c0) * how_many, b + o_sc, size_t(nq) * how_many * 8, cudaMemcpyDeviceToHost, stream));
      }
      CU_TRY(cudaMemcpyAsync(out_counts + c0, b + o_cnt, size_t(nq) * 4, cudaMemcpyDeviceToHost, stream));
      CU_TRY(cudaEventRecord(c->done, stream));
    }
    for (auto& c : ring) {
      cudaStream_t stream = stream_ ? static_cast<cudaStream_t>(stream_) : c->stream;
      CU_TRY(cudaStreamSynchronize(stream));
    }
    return VMIS_OK;
  };
  rc = body();
  if (rc) cudaDeviceSynchronize();               // drain whatever was enqueued before the contexts go back to the pool
  for (auto& c : ring) if (c) release_ctx(ix, std::move(c));
  return rc;
}

}  // namespace

namespace {
struct BlobHeader {
  char magic[8];                 // "VMISB200"
  uint32_t version, n_shards, shard, m_build, max_len, m_carry;   // m_carry: version >= 2 (version 1 blobs: = m_build)
  uint64_t n_items, hash_cap, n_kept, n_post_entries, n_sess_item_entries, n_pairs_kept, n_postings;
  double idf_weighting;
};
template <class T>
bool dump_dev(FILE* f, const T* dev, uint64_t n) {
  std::vector<T> h(n);
  if (n && cudaMemcpy(h.data(), dev, n * sizeof(T), cudaMemcpyDeviceToHost) != cudaSuccess) return false;
  return std::fwrite(h.data(), sizeof(T), n, f) == n;
}
template <class T>
bool slurp(FILE* f, std::vector<T>* v, uint64_t n) { v->resize(n); return std::fread(v->data(), sizeof(T), n, f) == n; }
// FNV-1a (64 bit, 8 bytes at a time) over the arrays of a blob in file order
struct Fnv {
  uint64_t h = 0xcbf29ce484222325ull;
  void add(const void* p, size_t bytes) {
    const unsigned char* b = static_cast<const unsigned char*>(p);
    size_t i = 0;
    for (; i + 8 <= bytes; i += 8) { uint64_t w; std::memcpy(&w, b + i, 8); h = (h ^ w) * 0x100000001b3ull; }
    for (; i < bytes; ++i) h = (h ^ b[i]) * 0x100000001b3ull;
  }
  template <class T> void add(const std::vector<T>& v) { add(v.data(), v.size() * sizeof(T)); }
};
uint64_t blob_checksum(const vmis::FlatIndex& F, const std::vector<uint32_t>& own) {
  Fnv c;
  c.add(F.item_key); c.add(F.item_hash); c.add(F.post_ref); c.add(own); c.add(F.sess_ref); c.add(F.sess_items);
  c.add(F.idf); c.add(F.attr); c.add(F.rank_to_orig);
  return c.h;
}
template <class T>
bool dump_dev_sum(FILE* f, const T* dev, uint64_t n, Fnv* c) {
  std::vector<T> h(n);
  if (n && cudaMemcpy(h.data(), dev, n * sizeof(T), cudaMemcpyDeviceToHost) != cudaSuccess) return false;
  c->add(h);
  return std::fwrite(h.data(), sizeof(T), n, f) == n;
}
}  // namespace

extern "C" {

const char* vmis_last_error(void) { return g_err.c_str(); }
int vmis_last_error_code(void) { return g_err_code; }
const char* vmis_version(void) { return "serenade_b200 0.1 (sm_100a)"; }

vmis_index_t* vmis_index_from_sessions(const uint64_t* items, const uint64_t* sess_off, const uint32_t* sess_ts,
                                       size_t n_sessions, size_t m, size_t max_len, double idf_weighting, int device) {
  g_err_code = 0;
  if (!sess_off || !sess_ts || (!items && n_sessions && sess_off[n_sessions] > 0)) { fail(VMIS_ERR_ARG, "NULL session arrays"); return nullptr; }
  std::unique_ptr<vmis_index> ix(new vmis_index());
  ix->sessions.off.assign(sess_off, sess_off + n_sessions + 1);
  ix->sessions.ts.assign(sess_ts, sess_ts + n_sessions);
  ix->sessions.items.assign(items, items + sess_off[n_sessions]);
  return finish_index(std::move(ix), m, max_len, idf_weighting, device);
}

vmis_index_t* vmis_index_from_sessions_attrs(const uint64_t* items, const uint64_t* sess_off, const uint32_t* sess_ts,
                                             size_t n_sessions, size_t m, size_t max_len, double idf_weighting,
                                             const uint64_t* attr_items, const uint8_t* attr_flags, size_t n_attrs,
                                             int device) {
  if (n_attrs && (!attr_items || !attr_flags)) { fail(VMIS_ERR_ARG, "NULL attribute arrays"); return nullptr; }
  vmis_index_t* ix = vmis_index_from_sessions(items, sess_off, sess_ts, n_sessions, m, max_len, idf_weighting, device);
  if (ix && n_attrs && vmis_index_set_attributes(ix, attr_items, attr_flags, n_attrs) != VMIS_OK) { vmis_index_free(ix); return nullptr; }
  return ix;
}

vmis_index_t* vmis_index_from_sessions_sharded(const uint64_t* items, const uint64_t* sess_off, const uint32_t* sess_ts,
                                               size_t n_sessions, size_t m, size_t max_len, double idf_weighting,
                                               int device, uint32_t shard, uint32_t n_shards) {
  g_err_code = 0;
  if (!sess_off || !sess_ts || (!items && n_sessions && sess_off[n_sessions] > 0)) { fail(VMIS_ERR_ARG, "NULL session arrays"); return nullptr; }
  std::unique_ptr<vmis_index> ix(new vmis_index());
  ix->sessions.off.assign(sess_off, sess_off + n_sessions + 1);
  ix->sessions.ts.assign(sess_ts, sess_ts + n_sessions);
  ix->sessions.items.assign(items, items + sess_off[n_sessions]);
  return finish_index(std::move(ix), m, max_len, idf_weighting, device, shard, n_shards);
}

vmis_index_t* vmis_index_from_device_sessions(const uint64_t* d_items, const uint64_t* d_sess_off, const uint32_t* d_sess_ts,
                                              size_t n_sessions, size_t m, size_t max_len, double idf_weighting, int device,
                                              uint32_t shard, uint32_t n_shards) {
  g_err_code = 0;
  if (!d_items || !d_sess_off || !d_sess_ts) { fail(VMIS_ERR_ARG, "NULL device session arrays"); return nullptr; }
  if (max_len == 0) { fail(VMIS_ERR_ARG, "the device build needs an explicit max_len"); return nullptr; }
  std::unique_ptr<vmis_index> ix(new vmis_index());
  if (select_device(device, &ix->sm_count) != VMIS_OK) return nullptr;
  vmis::DeviceSessions s; s.items = d_items; s.off = d_sess_off; s.ts = d_sess_ts; s.n_sessions = n_sessions;
  vmis::DeviceIndexArrays A; std::string err;
  if (!vmis::build_index_device(s, m, max_len, idf_weighting, shard, n_shards, &A, &err)) { fail(VMIS_ERR_CUDA, "%s", err.c_str()); return nullptr; }
  return adopt_device_index(std::move(ix), A, m, max_len, idf_weighting, device, shard, n_shards);
}

vmis_index_t* vmis_index_synth(uint64_t seed, uint64_t n_items, uint64_t n_sessions, size_t m, size_t max_len,
                               double idf_weighting, int device, uint32_t shard, uint32_t n_shards) {
  g_err_code = 0;
  if (max_len == 0) max_len = 34;
  std::unique_ptr<vmis_index> ix(new vmis_index());
  if (select_device(device, &ix->sm_count) != VMIS_OK) return nullptr;
  vmis::DeviceSessions s; std::string err;
  if (!vmis::synth_sessions_device(seed, n_items, n_sessions, &s, &err)) { fail(VMIS_ERR_CUDA, "%s", err.c_str()); return nullptr; }
  vmis::DeviceIndexArrays A;
  const bool ok = vmis::build_index_device(s, m, max_len, idf_weighting, shard, n_shards, &A, &err);
  cudaFree(const_cast<uint64_t*>(s.items)); cudaFree(const_cast<uint64_t*>(s.off)); cudaFree(const_cast<uint32_t*>(s.ts));
  if (!ok) { fail(VMIS_ERR_CUDA, "%s", err.c_str()); return nullptr; }
  vmis_index* r = adopt_device_index(std::move(ix), A, m, max_len, idf_weighting, device, shard, n_shards);
  if (r) r->synth_interactions = s.n_entries;
  return r;
}

// ---- pre-computed index (VMISIndex::new, vmis_index.rs:85-313): posting lists, idf and attributes are taken as given ----
static vmis_index_t* finish_prebuilt(std::unique_ptr<vmis_index> ix, vmis::PrebuiltIndex& P, int device, uint32_t shard, uint32_t n_shards,
                                     size_t max_session_len = 0) {
  if (n_shards == 0 || shard >= n_shards) { fail(VMIS_ERR_ARG, "shard %u out of range [0,%u)", shard, n_shards); return nullptr; }
  if (max_session_len) {
    // drop the sessions above the bound from the posting lists, as prepare_hashmap does for a CSV index
    // (vmis_index.rs:452); they stay in the host mirror (items_for_session) but can no longer become neighbours
    const vmis::Sessions& S = P.sessions;
    std::vector<uint64_t> off(1, 0);
    size_t w = 0;
    for (size_t d = 0; d + 1 < P.post_off.size(); ++d) {
      for (uint64_t e = P.post_off[d]; e < P.post_off[d + 1]; ++e) {
        const uint32_t sid = P.post_sessions[e];
        if (sid < S.size() && S.off[sid + 1] - S.off[sid] > max_session_len) { ++ix->prebuilt.pruned_postings; continue; }
        P.post_sessions[w++] = sid;
      }
      off.push_back(w);
    }
    P.post_sessions.resize(w);
    P.post_off.swap(off);
  }
  vmis::PrebuiltInfo pi; std::string err;
  if (!vmis::build_flat_index_prebuilt(P, n_shards, &ix->flat, &pi, &err)) { fail(VMIS_ERR_ARG, "%s", err.c_str()); return nullptr; }
  ix->sessions.items.swap(P.sessions.items); ix->sessions.off.swap(P.sessions.off); ix->sessions.ts.swap(P.sessions.ts);
  ix->prebuilt.prebuilt = 1; ix->prebuilt.lists_reordered = pi.lists_reordered;
  ix->prebuilt.duplicate_postings = pi.duplicate_postings; ix->prebuilt.m_carry = pi.m_carry;
  return upload_flat(std::move(ix), device, shard, n_shards);
}

vmis_index_t* vmis_index_from_avro_ex(const char* base_path, int device, uint32_t shard, uint32_t n_shards, size_t max_session_len) {
  g_err_code = 0;
  if (!base_path) { fail(VMIS_ERR_ARG, "path is NULL"); return nullptr; }
  std::unique_ptr<vmis_index> ix(new vmis_index());
  vmis::PrebuiltIndex P; vmis::AvroLoadInfo li; std::string err;
  if (!vmis::read_index_from_avro(base_path, &P, &li, &err)) { fail(VMIS_ERR_IO, "%s", err.c_str()); return nullptr; }
  ix->prebuilt.item_files = li.item_files; ix->prebuilt.session_files = li.session_files;
  ix->prebuilt.item_records = li.item_records; ix->prebuilt.session_records = li.session_records;
  return finish_prebuilt(std::move(ix), P, device, shard, n_shards, max_session_len);
}

vmis_index_t* vmis_index_from_avro_sharded(const char* base_path, int device, uint32_t shard, uint32_t n_shards) {
  return vmis_index_from_avro_ex(base_path, device, shard, n_shards, 0);
}

vmis_index_t* vmis_index_from_avro(const char* base_path, int device) { return vmis_index_from_avro_ex(base_path, device, 0, 1, 0); }

vmis_index_t* vmis_index_from_parts(const uint64_t* item_ids, const uint64_t* post_off, const uint32_t* post_sessions,
                                    const double* idf, const uint8_t* attr_or_null, size_t n_items, const uint64_t* items,
                                    const uint64_t* sess_off, const uint32_t* sess_ts, size_t n_sessions, int device,
                                    uint32_t shard, uint32_t n_shards) {
  return vmis_index_from_parts_ex(item_ids, post_off, post_sessions, idf, attr_or_null, n_items, items, sess_off, sess_ts, n_sessions,
                                  device, shard, n_shards, 0);
}

vmis_index_t* vmis_index_from_parts_ex(const uint64_t* item_ids, const uint64_t* post_off, const uint32_t* post_sessions,
                                       const double* idf, const uint8_t* attr_or_null, size_t n_items, const uint64_t* items,
                                       const uint64_t* sess_off, const uint32_t* sess_ts, size_t n_sessions, int device,
                                       uint32_t shard, uint32_t n_shards, size_t max_session_len) {
  g_err_code = 0;
  if (!item_ids || !post_off || !idf || !sess_off || !sess_ts || (!post_sessions && n_items && post_off[n_items] > 0) ||
      (!items && n_sessions && sess_off[n_sessions] > 0)) { fail(VMIS_ERR_ARG, "NULL index arrays"); return nullptr; }
  std::unique_ptr<vmis_index> ix(new vmis_index());
  vmis::PrebuiltIndex P;
  P.item_ids.assign(item_ids, item_ids + n_items);
  P.post_off.assign(post_off, post_off + n_items + 1);
  P.post_sessions.assign(post_sessions, post_sessions + post_off[n_items]);
  P.idf.assign(idf, idf + n_items);
  if (attr_or_null) { P.attr.assign(attr_or_null, attr_or_null + n_items); for (auto& a : P.attr) a = a ? (uint8_t)(a | VMIS_ATTR_EXISTS) : 0; }
  else P.attr.assign(n_items, (uint8_t)(VMIS_ATTR_EXISTS | VMIS_ATTR_FOR_SALE));
  P.sessions.off.assign(sess_off, sess_off + n_sessions + 1);
  P.sessions.ts.assign(sess_ts, sess_ts + n_sessions);
  P.sessions.items.assign(items, items + sess_off[n_sessions]);
  return finish_prebuilt(std::move(ix), P, device, shard, n_shards, max_session_len);
}

int vmis_index_prebuilt_info(const vmis_index_t* ix, vmis_prebuilt_info_t* out) {
  g_err_code = 0;
  if (!ix || !out) return fail(VMIS_ERR_ARG, "NULL argument");
  *out = ix->prebuilt;
  if (!ix->prebuilt.prebuilt) out->m_carry = ix->flat.m_carry;
  return VMIS_OK;
}

// The index in the production on-disk format (the reference computes it offline with Spark; vmis_index.rs:85-313 reads
// it back): posting lists as reference session indices, idf, attributes, and the host mirror of the sessions.
int vmis_index_to_avro(const vmis_index_t* ix, const char* base_path, const char* codec, uint32_t n_files) {
  g_err_code = 0;
  if (!ix || !base_path) return fail(VMIS_ERR_ARG, "NULL argument");
  if (ix->sessions.size() == 0) return fail(VMIS_ERR_ARG, "this handle has no host mirror of the sessions (blob-loaded or device-generated)");
  if (ix->flat.n_shards != 1) return fail(VMIS_ERR_ARG, "export an unsharded handle");
  const vmis::FlatIndex& F = ix->flat;
  const size_t I = F.item_key.size();
  std::vector<uint2> post_ref = F.post_ref;
  std::vector<uint32_t> postings = F.postings, rank_to_orig = F.rank_to_orig;
  if (post_ref.empty() || postings.empty() || rank_to_orig.empty()) {                // the arrays live in HBM only
    if (ix->device == VMIS_DEVICE_NONE) return fail(VMIS_ERR_ARG, "index arrays are missing");
    CU_TRY(cudaSetDevice(ix->device));
    CU_TRY(cudaDeviceSynchronize());
    post_ref.resize(I); postings.resize(ix->n_post_entries); rank_to_orig.resize(ix->view.n_kept);
    CU_TRY(cudaMemcpy(post_ref.data(), ix->view.post_ref, I * sizeof(uint2), cudaMemcpyDeviceToHost));
    if (!postings.empty()) CU_TRY(cudaMemcpy(postings.data(), ix->view.post_shard[0], postings.size() * 4, cudaMemcpyDeviceToHost));
    if (!rank_to_orig.empty()) CU_TRY(cudaMemcpy(rank_to_orig.data(), ix->view.rank_to_orig, rank_to_orig.size() * 4, cudaMemcpyDeviceToHost));
  }
  vmis::PrebuiltIndex P;
  P.item_ids = F.item_key; P.idf = F.idf; P.attr = F.attr;
  P.post_off.assign(1, 0);
  for (size_t d = 0; d < I; ++d) {
    const uint2 ref = post_ref[d];
    for (uint32_t i = 0; i < ref.y; ++i) P.post_sessions.push_back(rank_to_orig[postings[(size_t)ref.x * 4 + i]]);
    P.post_off.push_back(P.post_sessions.size());
  }
  P.sessions = ix->sessions;
  std::string err;
  if (!vmis::write_index_to_avro(base_path, P, codec ? codec : "deflate", n_files ? n_files : 1, &err)) return fail(VMIS_ERR_IO, "%s", err.c_str());
  return VMIS_OK;
}

// ---- serialised index blob: the "checkpoint" of this path (the reference rebuilds or re-reads Avro at start-up,
// serving.rs:37-52; loading the flat arrays is a plain read + upload) ----
int vmis_index_save(const vmis_index_t* ix, const char* path) {
  g_err_code = 0;
  if (!ix || !path) return fail(VMIS_ERR_ARG, "NULL argument");
  if (ix->device == VMIS_DEVICE_NONE) return fail(VMIS_ERR_ARG, "only device-resident indexes can be saved");
  CU_TRY(cudaSetDevice(ix->device));
  CU_TRY(cudaDeviceSynchronize());
  FILE* f = std::fopen(path, "wb");
  if (!f) return fail(VMIS_ERR_IO, "cannot create %s", path);
  const vmis::IndexView& V = ix->view;
  BlobHeader h{};
  std::memcpy(h.magic, "VMISB200", 8);
  h.version = 3; h.n_shards = V.n_shards; h.shard = ix->shard; h.m_build = V.m_build; h.max_len = V.max_len; h.m_carry = V.m_carry;
  h.n_items = V.n_items; h.hash_cap = (uint64_t)V.item_hash_mask + 1; h.n_kept = V.n_kept;
  h.n_post_entries = ix->n_post_entries; h.n_sess_item_entries = ix->n_sess_item_entries;
  h.n_pairs_kept = ix->flat.n_pairs_kept; h.n_postings = ix->flat.n_postings; h.idf_weighting = ix->flat.idf_weighting;
  Fnv sum;
  bool ok = std::fwrite(&h, sizeof h, 1, f) == 1 && dump_dev_sum(f, V.item_key, h.n_items, &sum) && dump_dev_sum(f, V.item_hash, h.hash_cap, &sum) &&
            dump_dev_sum(f, V.post_ref, h.n_items, &sum) && dump_dev_sum(f, V.post_shard[ix->shard], h.n_post_entries, &sum) &&
            dump_dev_sum(f, V.sess_ref, h.n_kept, &sum) && dump_dev_sum(f, V.sess_items, h.n_sess_item_entries, &sum) &&
            dump_dev_sum(f, V.idf, h.n_items, &sum) && dump_dev_sum(f, V.attr, h.n_items, &sum) && dump_dev_sum(f, V.rank_to_orig, h.n_kept, &sum);
  ok = ok && std::fwrite(&sum.h, 8, 1, f) == 1;
  ok = (std::fclose(f) == 0) && ok;
  if (!ok) return fail(VMIS_ERR_IO, "short write to %s", path);
  return VMIS_OK;
}

// Loads a blob written by vmis_index_save.  Nothing in the file is trusted: the header is checked against the file
// size before anything is allocated, every offset / rank / item index is checked against the array it points into
// (the kernel does no bounds checks), and version 3 blobs carry a checksum over the arrays.
vmis_index_t* vmis_index_load(const char* path, int device) {
  g_err_code = 0;
  if (!path) { fail(VMIS_ERR_ARG, "path is NULL"); return nullptr; }
  FILE* f = std::fopen(path, "rb");
  if (!f) { fail(VMIS_ERR_IO, "cannot open %s", path); return nullptr; }
  std::unique_ptr<vmis_index> ix(new vmis_index());
  BlobHeader h{};
  std::vector<uint32_t> own;
  const char* why = nullptr;
  try {
    vmis::FlatIndex& F = ix->flat;
    std::fseek(f, 0, SEEK_END);
    const long long fsize = std::ftell(f);
    std::fseek(f, 0, SEEK_SET);
    bool ok = std::fread(&h, sizeof h, 1, f) == 1 && !std::memcmp(h.magic, "VMISB200", 8) && h.version >= 1 && h.version <= 3;
    if (!ok) why = "not a VMIS index blob";
    if (ok && !(h.n_shards >= 1 && h.n_shards <= (uint32_t)vmis::kMaxShards && h.shard < h.n_shards)) { ok = false; why = "bad shard fields"; }
    if (ok && !(h.n_items < 0xFFFFFFFFull && h.n_kept < 0x80000000ull && h.hash_cap <= (1ull << 33) && h.n_post_entries < (1ull << 40) &&
                h.n_sess_item_entries < (1ull << 40))) { ok = false; why = "counts out of range"; }
    if (ok && !((h.hash_cap & (h.hash_cap - 1)) == 0 && h.hash_cap > h.n_items)) { ok = false; why = "item hash capacity is not a power of two above the item count"; }
    if (ok && (h.version >= 2 ? h.m_carry : h.m_build) > h.m_build) { ok = false; why = "m_carry above m_build"; }
    const unsigned long long expect = sizeof h + h.n_items * 8 + h.hash_cap * sizeof(vmis::ItemHashEntry) + h.n_items * sizeof(uint2) +
                                      h.n_post_entries * 4 + h.n_kept * sizeof(uint2) + h.n_sess_item_entries * 4 + h.n_items * 8 +
                                      h.n_items + h.n_kept * 4 + (h.version >= 3 ? 8 : 0);
    if (ok && (unsigned long long)fsize != expect) { ok = false; why = "file size does not match the header (truncated or corrupt)"; }
    ok = ok && slurp(f, &F.item_key, h.n_items) && slurp(f, &F.item_hash, h.hash_cap) && slurp(f, &F.post_ref, h.n_items) &&
         slurp(f, &own, h.n_post_entries) && slurp(f, &F.sess_ref, h.n_kept) && slurp(f, &F.sess_items, h.n_sess_item_entries) &&
         slurp(f, &F.idf, h.n_items) && slurp(f, &F.attr, h.n_items) && slurp(f, &F.rank_to_orig, h.n_kept);
    if (ok && h.version >= 3) {
      uint64_t stored = 0;
      ok = std::fread(&stored, 8, 1, f) == 1;
      if (ok && stored != blob_checksum(F, own)) { ok = false; why = "checksum mismatch"; }
    }
    if (ok) {
      // structure: everything the kernel dereferences stays inside its array
      for (size_t d = 1; d < F.item_key.size() && ok; ++d) ok = F.item_key[d - 1] < F.item_key[d];
      if (!ok) why = "item dictionary is not strictly ascending";
      for (size_t i = 0; i < F.item_hash.size() && ok; ++i) ok = F.item_hash[i].val == vmis::kEmpty || F.item_hash[i].val < h.n_items;
      if (!ok && !why) why = "item hash entry out of range";
      for (size_t d = 0; d < F.post_ref.size() && ok; ++d) {
        ok = F.post_ref[d].y <= h.m_build;
        if (ok && d % h.n_shards == h.shard) ok = (uint64_t)F.post_ref[d].x * 4 + F.post_ref[d].y <= h.n_post_entries;
      }
      if (!ok && !why) why = "posting list reference out of range";
      for (size_t i = 0; i < own.size() && ok; ++i) ok = own[i] == vmis::kEmpty || own[i] < h.n_kept;
      if (!ok && !why) why = "posting (session rank) out of range";
      for (size_t r = 0; r < F.sess_ref.size() && ok; ++r)
        ok = F.sess_ref[r].y <= h.max_len && (uint64_t)F.sess_ref[r].x * 4 + F.sess_ref[r].y <= h.n_sess_item_entries;
      if (!ok && !why) why = "session item list reference out of range";
      for (size_t i = 0; i < F.sess_items.size() && ok; ++i) ok = F.sess_items[i] == vmis::kEmpty || F.sess_items[i] < h.n_items;
      if (!ok && !why) why = "session item out of range";
      if (ok && (F.sess_items.size() & 3)) { ok = false; why = "session item array is not padded to 16 bytes"; }
    }
    std::fclose(f); f = nullptr;
    if (!ok) { fail(VMIS_ERR_IO, "%s: %s", path, why ? why : "truncated"); return nullptr; }
    if (h.version == 1) h.m_carry = h.m_build;
    F.n_pairs_kept = h.n_pairs_kept; F.n_postings = h.n_postings; F.m_build = h.m_build; F.m_carry = h.m_carry; F.max_len = h.max_len;
    F.idf_weighting = h.idf_weighting; F.n_shards = h.n_shards;
  } catch (const std::exception& e) {
    if (f) std::fclose(f);
    fail(VMIS_ERR_IO, "%s: %s", path, e.what());
    return nullptr;
  }
  vmis::FlatIndex& F = ix->flat;
  ix->shard = h.shard; ix->n_sessions_kept = h.n_kept; ix->device = device;
  if (select_device(device, &ix->sm_count) != VMIS_OK) return nullptr;
  vmis::IndexView& V = ix->view;
  const uint32_t* own_dev = nullptr;
  for (int s2 = 0; s2 < vmis::kMaxShards; ++s2) V.post_shard[s2] = nullptr;
  if (upload(ix.get(), F.item_key, &V.item_key) || upload(ix.get(), F.item_hash, &V.item_hash) ||
      upload(ix.get(), F.post_ref, &V.post_ref) || upload(ix.get(), own, &own_dev) ||
      upload(ix.get(), F.sess_ref, &V.sess_ref) || upload(ix.get(), F.sess_items, &V.sess_items) ||
      upload(ix.get(), F.idf, &V.idf) || upload(ix.get(), F.attr, &V.attr) || upload(ix.get(), F.rank_to_orig, &V.rank_to_orig)) {
    for (void* d : ix->dev_allocs) cudaFree(d);
    return nullptr;
  }
  V.post_shard[h.shard] = own_dev; V.n_shards = h.n_shards;
  V.item_hash_mask = (uint32_t)(h.hash_cap - 1); V.n_items = (uint32_t)h.n_items; V.n_kept = (uint32_t)h.n_kept;
  V.m_build = h.m_build; V.m_carry = h.m_carry; V.max_len = h.max_len;
  if (attach_g32(ix.get())) { for (void* d : ix->dev_allocs) cudaFree(d); return nullptr; }
  ix->n_post_entries = h.n_post_entries; ix->n_sess_item_entries = h.n_sess_item_entries;
  std::vector<uint32_t>().swap(F.sess_items); std::vector<uint2>().swap(F.sess_ref); std::vector<uint2>().swap(F.post_ref);
  return ix.release();
}

int vmis_index_export_shard(const vmis_index_t* ix, void* handle64) {
  g_err_code = 0;
  if (!ix || !handle64 || ix->device == VMIS_DEVICE_NONE) return fail(VMIS_ERR_ARG, "no device shard to export");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  CU_TRY(cudaSetDevice(ix->device));
  cudaIpcMemHandle_t h;
  CU_TRY(cudaIpcGetMemHandle(&h, const_cast<uint32_t*>(ix->view.post_shard[ix->shard])));
  std::memcpy(handle64, &h, 64);
  return VMIS_OK;
}

int vmis_index_attach_shard(vmis_index_t* ix, uint32_t shard, const void* handle64) {
  g_err_code = 0;
  if (!ix || !handle64 || ix->device == VMIS_DEVICE_NONE) return fail(VMIS_ERR_ARG, "NULL argument");
  if (shard >= ix->view.n_shards || shard == ix->shard) return fail(VMIS_ERR_ARG, "cannot attach shard %u", shard);
  CU_TRY(cudaSetDevice(ix->device));
  cudaIpcMemHandle_t h;
  std::memcpy(&h, handle64, 64);
  void* p = nullptr;
  CU_TRY(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  ix->ipc_mapped.push_back(p);
  ix->view.post_shard[shard] = static_cast<const uint32_t*>(p);
  return VMIS_OK;
}

int vmis_index_attach_shard_ptr(vmis_index_t* ix, uint32_t shard, const void* device_ptr) {
  g_err_code = 0;
  if (!ix || !device_ptr || ix->device == VMIS_DEVICE_NONE) return fail(VMIS_ERR_ARG, "NULL argument");
  if (shard >= ix->view.n_shards || shard == ix->shard) return fail(VMIS_ERR_ARG, "cannot attach shard %u", shard);
  ix->view.post_shard[shard] = static_cast<const uint32_t*>(device_ptr);
  return VMIS_OK;
}

const void* vmis_index_shard_ptr(const vmis_index_t* ix) {
  if (!ix || ix->device == VMIS_DEVICE_NONE) return nullptr;
  return ix->view.post_shard[ix->shard];
}

// read_from_file (vmis_index.rs:591-752) on its own: the parsed training sessions, so that callers which build many
// indexes from one file (the HPO objective rebuilds it for every trial, objective.rs:17) parse it once.
struct vmis_sessions { vmis::Sessions s; };

vmis_sessions_t* vmis_sessions_from_csv(const char* path) {
  g_err_code = 0;
  if (!path) { fail(VMIS_ERR_ARG, "path is NULL"); return nullptr; }
  std::unique_ptr<vmis_sessions> h(new vmis_sessions());
  std::string err;
  if (!vmis::read_sessions_from_csv(path, &h->s, &err)) { fail(VMIS_ERR_IO, "%s", err.c_str()); return nullptr; }
  return h.release();
}

int vmis_sessions_view(const vmis_sessions_t* h, const uint64_t** items, const uint64_t** sess_off, const uint32_t** sess_ts,
                       size_t* n_sessions) {
  g_err_code = 0;
  if (!h || !items || !sess_off || !sess_ts || !n_sessions) return fail(VMIS_ERR_ARG, "NULL argument");
  *items = h->s.items.data(); *sess_off = h->s.off.data(); *sess_ts = h->s.ts.data(); *n_sessions = h->s.size();
  return VMIS_OK;
}

void vmis_sessions_free(vmis_sessions_t* h) { delete h; }

vmis_index_t* vmis_index_from_csv_ex(const char* path, size_t m, double idf_weighting, size_t max_len, int device) {
  g_err_code = 0;
  if (!path) { fail(VMIS_ERR_ARG, "path is NULL"); return nullptr; }
  std::unique_ptr<vmis_index> ix(new vmis_index());
  std::string err;
  if (!vmis::read_sessions_from_csv(path, &ix->sessions, &err)) { fail(VMIS_ERR_IO, "%s", err.c_str()); return nullptr; }
  return finish_index(std::move(ix), m, max_len, idf_weighting, device);
}

vmis_index_t* vmis_index_from_csv(const char* path, size_t m, double idf_weighting, int device) {
  g_err_code = 0;
  return vmis_index_from_csv_ex(path, m, idf_weighting, 0, device);
}

int vmis_index_set_attributes(vmis_index_t* ix, const uint64_t* items, const uint8_t* flags, size_t n) {
  g_err_code = 0;
  if (!ix || (n && (!items || !flags))) return fail(VMIS_ERR_ARG, "NULL argument");
  for (size_t i = 0; i < n; ++i) {
    const uint32_t d = vmis::host_lookup_item(ix->flat, items[i]);
    if (d != vmis::kEmpty) ix->flat.attr[d] = flags[i] ? (uint8_t)(flags[i] | VMIS_ATTR_EXISTS) : 0;
  }
  if (ix->device == VMIS_DEVICE_NONE) return VMIS_OK;
  CU_TRY(cudaSetDevice(ix->device));
  CU_TRY(cudaDeviceSynchronize());
  if (!ix->flat.attr.empty())
    CU_TRY(cudaMemcpy(const_cast<uint8_t*>(ix->view.attr), ix->flat.attr.data(), ix->flat.attr.size(), cudaMemcpyHostToDevice));
  return VMIS_OK;
}

void vmis_index_free(vmis_index_t* ix) {
  if (!ix) return;
  if (ix->device == VMIS_DEVICE_NONE) { delete ix; return; }
  cudaSetDevice(ix->device);
  cudaDeviceSynchronize();
  ix->pool.clear();
  for (void* p : ix->ipc_mapped) cudaIpcCloseMemHandle(p);
  for (void* d : ix->dev_allocs) cudaFree(d);
  delete ix;
}

int vmis_index_stats(const vmis_index_t* ix, vmis_stats_t* out) {
  g_err_code = 0;
  if (!ix || !out) return fail(VMIS_ERR_ARG, "NULL argument");
  out->n_sessions = ix->sessions.size() ? ix->sessions.size() : ix->n_sessions_kept;
  out->n_sessions_kept = ix->n_sessions_kept;
  out->n_items = ix->flat.item_key.size();
  out->n_pairs_kept = ix->flat.n_pairs_kept;
  out->n_postings = ix->flat.n_postings;
  out->max_len = ix->flat.max_len;
  out->m_build = ix->flat.m_build;
  out->device_bytes = ix->device_bytes;
  out->idf_weighting = ix->flat.idf_weighting;
  return VMIS_OK;
}

int vmis_predict_batch(const vmis_index_t* ix, const uint64_t* q_items, const uint32_t* q_off, uint32_t n_q, uint32_t k,
                       uint32_t m, uint32_t how_many, int biz, uint64_t* out_ids, double* out_scores,
                       uint32_t* out_counts, void* stream) {
  g_err_code = 0;
  return host_batch(ix, q_items, q_off, n_q, k, m, how_many, biz, out_ids, out_scores, out_counts, nullptr, nullptr, stream);
}

int vmis_find_neighbors_batch(const vmis_index_t* ix, const uint64_t* q_items, const uint32_t* q_off, uint32_t n_q,
                              uint32_t k, uint32_t m, uint32_t* out_sess, double* out_sim, uint32_t* out_counts,
                              void* stream) {
  g_err_code = 0;
  if (!out_sess || !out_sim) return fail(VMIS_ERR_ARG, "NULL output buffer");
  return host_batch(ix, q_items, q_off, n_q, k, m, 0, 0, nullptr, nullptr, out_counts, out_sess, out_sim, stream);
}

int vmis_predict_batch_device(const vmis_index_t* cix, const uint64_t* d_q_items, const uint32_t* d_q_off, uint32_t n_q,
                              uint32_t k, uint32_t m, uint32_t how_many, int biz, uint64_t* d_out_ids,
                              double* d_out_scores, uint32_t* d_out_counts, vmis_query_stats_t* d_out_stats, void* stream_) {
  g_err_code = 0;
  vmis_index* ix = const_cast<vmis_index*>(cix);
  vmis::LaunchPlan plan;
  int rc = check_common(ix, k, m, &plan);
  if (rc) return rc;
  if (n_q == 0) return VMIS_OK;
  if (!d_q_items || !d_q_off || !d_out_counts || (how_many && (!d_out_ids || !d_out_scores)))
    return fail(VMIS_ERR_ARG, "NULL device buffer");
  CU_TRY(cudaSetDevice(ix->device));
  std::unique_ptr<CallCtx> c;
  rc = acquire_ctx(ix, &c);
  if (rc) return rc;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);   // NULL = the CUDA default stream
  vmis::PredictArgs a{};
  a.q_items = d_q_items; a.q_off = d_q_off; a.n_q = n_q; a.k = k; a.m = m; a.how_many = how_many; a.biz = biz;
  a.out_ids = d_out_ids; a.out_scores = d_out_scores; a.out_counts = d_out_counts; a.out_stats = d_out_stats;
  rc = run_device(ix, c.get(), a, plan, stream);
  release_ctx(ix, std::move(c));
  return rc;
}

int vmis_predict(const vmis_index_t* ix, const uint64_t* ev, size_t len, size_t k, size_t m, size_t how_many, int biz,
                 uint64_t* out_ids, double* out_scores) {
  g_err_code = 0;
  if (len > 0xFFFFFFFFull || k > 0xFFFFFFFFull || m > 0xFFFFFFFFull || how_many > 0xFFFFFFFFull)
    return fail(VMIS_ERR_ARG, "argument out of range");
  const uint32_t off[2] = {0u, (uint32_t)len};
  uint32_t cnt = 0;
  const int rc = vmis_predict_batch(ix, ev, off, 1, (uint32_t)k, (uint32_t)m, (uint32_t)how_many, biz, out_ids,
                                    out_scores, &cnt, nullptr);
  return rc ? rc : (int)cnt;
}

const uint64_t* vmis_items_for_session(const vmis_index_t* ix, uint32_t session, size_t* len) {
  g_err_code = 0;
  if (!ix || session >= ix->sessions.size()) { fail(VMIS_ERR_ARG, "session %u out of range", session); if (len) *len = 0; return nullptr; }
  if (len) *len = ix->sessions.off[session + 1] - ix->sessions.off[session];
  return ix->sessions.items.data() + ix->sessions.off[session];
}

int vmis_idf(const vmis_index_t* ix, uint64_t item, double* out) {
  g_err_code = 0;
  if (!ix || !out) return fail(VMIS_ERR_ARG, "NULL argument");
  const uint32_t d = vmis::host_lookup_item(ix->flat, item);
  if (d == vmis::kEmpty) return fail(VMIS_ERR_ARG, "unknown item %llu", (unsigned long long)item);
  *out = ix->flat.idf[d];
  return VMIS_OK;
}

int vmis_find_attributes(const vmis_index_t* ix, uint64_t item) {
  g_err_code = 0;
  if (!ix) return 0;
  const uint32_t d = vmis::host_lookup_item(ix->flat, item);
  if (d == vmis::kEmpty) return 0;
  const uint8_t a = ix->flat.attr[d];
  return (a & VMIS_ATTR_EXISTS) ? (int)a : 0;
}

size_t vmis_postings(const vmis_index_t* ix, uint64_t item, uint32_t* out, size_t cap) {
  g_err_code = 0;
  if (!ix) return 0;
  const uint32_t d = vmis::host_lookup_item(ix->flat, item);
  if (d == vmis::kEmpty) return 0;
  if (ix->flat.post_ref.empty()) { fail(VMIS_ERR_ARG, "postings live in HBM only"); return 0; }
  const uint2 ref = ix->flat.post_ref[d];
  if (ix->flat.postings.empty()) { fail(VMIS_ERR_ARG, "postings live in HBM only; use a VMIS_DEVICE_NONE handle to inspect them"); return ref.y; }
  const size_t base = ix->flat.shard_begin[d % ix->flat.n_shards] + (size_t)ref.x * 4;
  for (size_t i = 0; i < ref.y && i < cap; ++i) out[i] = ix->flat.rank_to_orig[ix->flat.postings[base + i]];
  return ref.y;
}

int vmis_session_timestamp(const vmis_index_t* ix, uint32_t session, uint32_t* out) {
  g_err_code = 0;
  if (!ix || !out || session >= ix->sessions.size()) return fail(VMIS_ERR_ARG, "session %u out of range", session);
  *out = ix->sessions.ts[session];
  return VMIS_OK;
}

}  // extern "C"
