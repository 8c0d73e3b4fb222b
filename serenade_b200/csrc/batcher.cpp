// serenade_b200/csrc/batcher.cpp — micro-batching front of vmis_predict_batch for the ONLINE call shape of the
// reference: many worker threads (actix, serving.rs:62-94) each calling predict() for one evolving session
// (recommend_resource.rs:56).  A single GPU launch per request would be latency-bound, so requests are parked
// in a queue and dispatcher threads turn whatever has arrived into one batched call: a dispatcher fires as soon as
// `max_batch` requests are waiting or the oldest has waited `max_wait_us`.  Two dispatchers alternate, so the next
// batch is assembled and launched while the previous one is still on the GPU (vmis_predict_batch is re-entrant:
// pooled streams / staging buffers).  Every request is validated on its own before it is queued and completes
// through its own futex word: one bad request never fails its batch mates, and a finished batch wakes exactly its
// callers.
#include <linux/futex.h>
#include <sys/syscall.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <climits>
#include <condition_variable>
#include <cstring>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/vmis.h"
#include "vmis_host.h"

namespace {
long futex(std::atomic<int>* addr, int op, int val) {
  return syscall(SYS_futex, reinterpret_cast<int*>(addr), op, val, nullptr, nullptr, 0);
}
}  // namespace

struct vmis_batcher {
  struct Req {
    const uint64_t* items; uint32_t len; uint64_t* out_ids; double* out_scores;
    int result = 0;
    const std::string* error = nullptr;          // text of a failed batch call (owned by the dispatcher's batch)
    std::atomic<int> done{0};
  };
  const vmis_index_t* index;
  uint32_t k, m, how_many, max_batch, max_wait_us;
  int biz;
  std::mutex mu;
  std::condition_variable cv_work;
  std::vector<Req*> queue;
  bool stop = false;
  std::atomic<uint64_t> n_batches{0}, n_requests{0};
  std::vector<std::thread> workers;

  void run() {
    std::vector<Req*> batch;
    std::vector<uint64_t> q_items, ids; std::vector<uint32_t> q_off, counts; std::vector<double> scores;
    std::string err;
    for (;;) {
      {
        std::unique_lock<std::mutex> lk(mu);
        cv_work.wait(lk, [&] { return stop || !queue.empty(); });
        if (stop && queue.empty()) return;
        if (queue.size() < max_batch && max_wait_us)      // give concurrent callers a moment to join the batch
          cv_work.wait_for(lk, std::chrono::microseconds(max_wait_us), [&] { return stop || queue.size() >= max_batch; });
        const size_t n = std::min<size_t>(queue.size(), max_batch);
        if (n == 0) continue;                             // the other dispatcher took them
        batch.assign(queue.begin(), queue.begin() + n);
        queue.erase(queue.begin(), queue.begin() + n);
        if (!queue.empty()) cv_work.notify_one();
      }
      q_items.clear(); q_off.assign(1, 0);
      for (Req* r : batch) { q_items.insert(q_items.end(), r->items, r->items + r->len); q_off.push_back((uint32_t)q_items.size()); }
      const uint32_t n_q = (uint32_t)batch.size();
      const size_t width = std::max(how_many, 1u);
      ids.resize((size_t)n_q * width); scores.resize(ids.size()); counts.assign(n_q, 0);
      if (q_items.empty()) q_items.push_back(0);
      const int rc = vmis_predict_batch(index, q_items.data(), q_off.data(), n_q, k, m, how_many, biz, ids.data(),
                                        scores.data(), counts.data(), nullptr);
      if (rc != 0) err = vmis_last_error();               // thread-local to this dispatcher: hand the text to the callers
      n_batches.fetch_add(1, std::memory_order_relaxed); n_requests.fetch_add(n_q, std::memory_order_relaxed);
      for (uint32_t i = 0; i < n_q; ++i) {
        Req* r = batch[i];
        if (rc == 0) {
          std::memcpy(r->out_ids, &ids[(size_t)i * how_many], (size_t)counts[i] * 8);
          if (r->out_scores) std::memcpy(r->out_scores, &scores[(size_t)i * how_many], (size_t)counts[i] * 8);
          r->result = (int)counts[i];
        } else { r->result = rc; r->error = &err; }
      }
      if (rc != 0) {
        // the callers copy the text before this dispatcher may overwrite it: wait until each has acknowledged
        for (Req* r : batch) { r->done.store(1, std::memory_order_release); futex(&r->done, FUTEX_WAKE_PRIVATE, 1); }
        for (Req* r : batch) while (r->done.load(std::memory_order_acquire) != 2) std::this_thread::yield();
        for (Req* r : batch) { r->done.store(3, std::memory_order_release); futex(&r->done, FUTEX_WAKE_PRIVATE, 1); }
      } else {
        for (Req* r : batch) { r->done.store(3, std::memory_order_release); futex(&r->done, FUTEX_WAKE_PRIVATE, 1); }
      }
    }
  }
};

extern "C" {

vmis_batcher_t* vmis_batcher_create(const vmis_index_t* index, uint32_t k, uint32_t m, uint32_t how_many,
                                    int enable_business_logic, uint32_t max_batch, uint32_t max_wait_us) {
  if (!index || max_batch == 0) { vmis::set_last_error(VMIS_ERR_ARG, "vmis_batcher_create: NULL index or max_batch == 0"); return nullptr; }
  vmis_batcher* b = new vmis_batcher();
  b->index = index; b->k = k; b->m = m; b->how_many = how_many; b->biz = enable_business_logic;
  b->max_batch = max_batch; b->max_wait_us = max_wait_us;
  for (int i = 0; i < 2; ++i) b->workers.emplace_back([b] { b->run(); });
  return b;
}

int vmis_batcher_predict(vmis_batcher_t* b, const uint64_t* evolving_session, size_t len, uint64_t* out_ids, double* out_scores) {
  if (!b || (!evolving_session && len) || (!out_ids && b->how_many)) {
    vmis::set_last_error(VMIS_ERR_ARG, "vmis_batcher_predict: NULL argument");
    return VMIS_ERR_ARG;
  }
  // validated per request: an over-long session fails alone instead of failing the batch it would have joined
  if (len > (size_t)VMIS_MAX_SESSION_LEN) {
    vmis::set_last_error(VMIS_ERR_LIMIT, "evolving session longer than the kernel limit (VMIS_MAX_SESSION_LEN)");
    return VMIS_ERR_LIMIT;
  }
  vmis_batcher::Req r{evolving_session, (uint32_t)len, out_ids, out_scores};
  {
    std::lock_guard<std::mutex> lk(b->mu);
    if (b->stop) { vmis::set_last_error(VMIS_ERR_ARG, "batcher is shutting down"); return VMIS_ERR_ARG; }
    b->queue.push_back(&r);
    if (b->queue.size() == 1 || b->queue.size() >= b->max_batch) b->cv_work.notify_one();
  }
  int d;
  while ((d = r.done.load(std::memory_order_acquire)) == 0) futex(&r.done, FUTEX_WAIT_PRIVATE, 0);
  if (d == 1) {                                           // failed batch: take the dispatcher's message, acknowledge
    vmis::set_last_error(r.result, r.error ? r.error->c_str() : "batched call failed");
    r.done.store(2, std::memory_order_release);
    while (r.done.load(std::memory_order_acquire) != 3) futex(&r.done, FUTEX_WAIT_PRIVATE, 2);
  } else {
    vmis::set_last_error(VMIS_OK, "");
  }
  return r.result;
}

int vmis_batcher_stats(vmis_batcher_t* b, uint64_t* n_batches, uint64_t* n_requests) {
  if (!b) return VMIS_ERR_ARG;
  if (n_batches) *n_batches = b->n_batches.load();
  if (n_requests) *n_requests = b->n_requests.load();
  return VMIS_OK;
}

void vmis_batcher_destroy(vmis_batcher_t* b) {
  if (!b) return;
  { std::lock_guard<std::mutex> lk(b->mu); b->stop = true; }
  b->cv_work.notify_all();
  for (auto& w : b->workers) w.join();
  delete b;
}

// Open-loop load generator for the online call shape (bench.py "latency" section; the reference quotes "1000
// predictions/s on 2 vCPU, p90 < 7 ms end to end", README.md:16-17): n_threads caller threads replay the evolving
// sessions of a CSR batch through vmis_batcher_predict at `target_rps` for `duration_ms`.  Request j is DUE at
// t0 + j / target_rps whatever happened to the requests before it, and its latency is measured from that due
// time, so queueing delay under overload is counted (no coordinated omission).  Latencies (microseconds) of up to
// `cap` requests go to lat_us in completion order per thread; returns the number of completed requests or a
// negative error.
long long vmis_batcher_load_test(vmis_batcher_t* b, const uint64_t* q_items, const uint32_t* q_off, uint32_t n_q,
                                 uint32_t n_threads, double target_rps, uint32_t duration_ms, float* lat_us, size_t cap,
                                 double* achieved_rps) {
  if (!b || !q_items || !q_off || n_q == 0 || n_threads == 0 || target_rps <= 0 || duration_ms == 0) return VMIS_ERR_ARG;
  using clock = std::chrono::steady_clock;
  const auto t0 = clock::now() + std::chrono::milliseconds(5);
  const double period_ns = 1e9 / target_rps;
  const uint64_t total = (uint64_t)(target_rps * duration_ms / 1e3);
  std::atomic<uint64_t> next{0}, done{0}, errors{0};
  std::vector<std::thread> th;
  std::atomic<long long> last_ns{0};
  for (uint32_t t = 0; t < n_threads; ++t) th.emplace_back([&, t]() {
    std::vector<uint64_t> ids(std::max(b->how_many, 1u)); std::vector<double> sc(ids.size());
    for (;;) {
      const uint64_t j = next.fetch_add(1, std::memory_order_relaxed);
      if (j >= total) break;
      const auto due = t0 + std::chrono::nanoseconds((long long)(j * period_ns));
      auto now = clock::now();
      if (now < due) {
        if (due - now > std::chrono::microseconds(200)) std::this_thread::sleep_until(due - std::chrono::microseconds(100));
        while (clock::now() < due) {}
      }
      const uint32_t q = (uint32_t)(j % n_q);
      const int rc = vmis_batcher_predict(b, q_items + q_off[q], q_off[q + 1] - q_off[q], ids.data(), sc.data());
      const auto end = clock::now();
      if (rc < 0) errors.fetch_add(1, std::memory_order_relaxed);
      const uint64_t slot = done.fetch_add(1, std::memory_order_relaxed);
      if (slot < cap) lat_us[slot] = (float)(std::chrono::duration<double, std::micro>(end - due).count());
      const long long e = std::chrono::duration_cast<std::chrono::nanoseconds>(end - t0).count();
      long long prev = last_ns.load(std::memory_order_relaxed);
      while (e > prev && !last_ns.compare_exchange_weak(prev, e)) {}
    } });
  for (auto& x : th) x.join();
  if (errors.load()) return VMIS_ERR_CUDA;
  if (achieved_rps) *achieved_rps = last_ns.load() > 0 ? done.load() * 1e9 / (double)last_ns.load() : 0.0;
  return (long long)done.load();
}

// Closed-loop lone caller (bench.py "latency" section): n_calls vmis_predict() calls from THIS thread over the evolving
// sessions of a CSR batch, one session per call — the reference's own call shape (mod.rs:118-125) without a Python
// interpreter between the calls.  Per-call latencies (microseconds) go to lat_us[0..n_calls); returns n_calls or a
// negative VMIS_ERR_*.
long long vmis_predict_latency_test(const vmis_index_t* index, const uint64_t* q_items, const uint32_t* q_off, uint32_t n_q,
                                    uint32_t k, uint32_t m, uint32_t how_many, int enable_business_logic, uint32_t n_calls,
                                    float* lat_us) {
  if (!index || !q_items || !q_off || n_q == 0 || !lat_us) return VMIS_ERR_ARG;
  using clock = std::chrono::steady_clock;
  std::vector<uint64_t> ids(std::max(how_many, 1u)); std::vector<double> sc(ids.size());
  for (uint32_t j = 0; j < n_calls; ++j) {
    const uint32_t q = j % n_q;
    const auto t0 = clock::now();
    const int rc = vmis_predict(index, q_items + q_off[q], q_off[q + 1] - q_off[q], k, m, how_many, enable_business_logic,
                                ids.data(), sc.data());
    const auto t1 = clock::now();
    if (rc < 0) return rc;
    lat_us[j] = (float)std::chrono::duration<double, std::micro>(t1 - t0).count();
  }
  return (long long)n_calls;
}

}  // extern "C"
