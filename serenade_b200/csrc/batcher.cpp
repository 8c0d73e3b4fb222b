// serenade_b200/csrc/batcher.cpp — micro-batching front of vmis_predict_batch for the ONLINE call shape of the
// reference: many worker threads (actix, serving.rs:62-94) each calling predict() for one evolving session
// (recommend_resource.rs:56).  A single GPU launch per request would be latency-bound, so requests are parked
// in a queue and one dispatcher thread turns whatever has arrived into one batched call: it fires as soon as
// `max_batch` requests are waiting or the oldest has waited `max_wait_us`.
#include <chrono>
#include <condition_variable>
#include <cstring>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

#include "../../include/vmis.h"

struct vmis_batcher {
  struct Req {
    const uint64_t* items; uint32_t len; uint64_t* out_ids; double* out_scores;
    int result = 0; bool done = false;
  };
  const vmis_index_t* index;
  uint32_t k, m, how_many, max_batch, max_wait_us;
  int biz;
  std::mutex mu;
  std::condition_variable cv_work, cv_done;
  std::vector<Req*> queue;
  bool stop = false;
  uint64_t n_batches = 0, n_requests = 0;
  std::thread worker;

  void run() {
    std::vector<Req*> batch;
    std::vector<uint64_t> q_items, ids; std::vector<uint32_t> q_off, counts; std::vector<double> scores;
    for (;;) {
      {
        std::unique_lock<std::mutex> lk(mu);
        cv_work.wait(lk, [&] { return stop || !queue.empty(); });
        if (stop && queue.empty()) return;
        if (queue.size() < max_batch && max_wait_us)      // give concurrent callers a moment to join the batch
          cv_work.wait_for(lk, std::chrono::microseconds(max_wait_us), [&] { return stop || queue.size() >= max_batch; });
        const size_t n = std::min<size_t>(queue.size(), max_batch);
        batch.assign(queue.begin(), queue.begin() + n);
        queue.erase(queue.begin(), queue.begin() + n);
      }
      q_items.clear(); q_off.assign(1, 0);
      for (Req* r : batch) { q_items.insert(q_items.end(), r->items, r->items + r->len); q_off.push_back((uint32_t)q_items.size()); }
      const uint32_t n_q = (uint32_t)batch.size();
      ids.assign((size_t)n_q * std::max(how_many, 1u), 0); scores.assign(ids.size(), 0.0); counts.assign(n_q, 0);
      if (q_items.empty()) q_items.push_back(0);
      const int rc = vmis_predict_batch(index, q_items.data(), q_off.data(), n_q, k, m, how_many, biz, ids.data(),
                                        scores.data(), counts.data(), nullptr);
      {
        std::lock_guard<std::mutex> lk(mu);
        for (uint32_t i = 0; i < n_q; ++i) {
          Req* r = batch[i];
          if (rc == 0) {
            std::memcpy(r->out_ids, &ids[(size_t)i * how_many], (size_t)counts[i] * 8);
            std::memcpy(r->out_scores, &scores[(size_t)i * how_many], (size_t)counts[i] * 8);
            r->result = (int)counts[i];
          } else r->result = rc;
          r->done = true;
        }
        ++n_batches; n_requests += n_q;
      }
      cv_done.notify_all();
    }
  }
};

extern "C" {

vmis_batcher_t* vmis_batcher_create(const vmis_index_t* index, uint32_t k, uint32_t m, uint32_t how_many,
                                    int enable_business_logic, uint32_t max_batch, uint32_t max_wait_us) {
  if (!index || max_batch == 0) return nullptr;
  vmis_batcher* b = new vmis_batcher();
  b->index = index; b->k = k; b->m = m; b->how_many = how_many; b->biz = enable_business_logic;
  b->max_batch = max_batch; b->max_wait_us = max_wait_us;
  b->worker = std::thread([b] { b->run(); });
  return b;
}

int vmis_batcher_predict(vmis_batcher_t* b, const uint64_t* evolving_session, size_t len, uint64_t* out_ids, double* out_scores) {
  if (!b || (!evolving_session && len) || len > 0xFFFFFFFFull) return VMIS_ERR_ARG;
  vmis_batcher::Req r{evolving_session, (uint32_t)len, out_ids, out_scores};
  std::unique_lock<std::mutex> lk(b->mu);
  if (b->stop) return VMIS_ERR_ARG;
  b->queue.push_back(&r);
  b->cv_work.notify_one();
  b->cv_done.wait(lk, [&] { return r.done; });
  return r.result;
}

int vmis_batcher_stats(vmis_batcher_t* b, uint64_t* n_batches, uint64_t* n_requests) {
  if (!b) return VMIS_ERR_ARG;
  std::lock_guard<std::mutex> lk(b->mu);
  if (n_batches) *n_batches = b->n_batches;
  if (n_requests) *n_requests = b->n_requests;
  return VMIS_OK;
}

void vmis_batcher_destroy(vmis_batcher_t* b) {
  if (!b) return;
  { std::lock_guard<std::mutex> lk(b->mu); b->stop = true; }
  b->cv_work.notify_all();
  b->worker.join();
  delete b;
}

}  // extern "C"
