// tools/evaluator.cpp — the reference's `evaluator` binary (src/bin/evaluator.rs) on top of include/vmis.hpp:
// builds the index from a training TSV, replays every prefix of every test session (evaluator.rs:46-57) through
// ONE batched predict call, and prints the reference's report: the eight metrics of metrics/evaluation_reporter.rs
// (Mrr, Ndcg, HitRate, Popularity, Precision, Coverage, Recall, F1score @20) and the number of evaluations.
//   usage: evaluator train.txt test.txt [m=500] [k=50] [how_many=21] [max_items_in_session=2] [idf_weighting=1]
// With --kat it runs the reference's known-answer test should_train_and_predict (mod.rs:229-310) instead.
#include <sys/stat.h>
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <set>
#include <string>
#include <unordered_map>

#include "../include/vmis.hpp"

static int kat() {
  // mod.rs:229-310
  std::vector<std::vector<uint64_t>> sessions = {{920006, 920005, 920004}, {920005, 920004, 920003, 920002}};
  auto index = vmis::VMISIndex::from_sessions(sessions, {1, 1}, 5, 5, 1.0);
  auto recs = vmis::predict(index, {920005}, 500, 500, 20, false);
  if (recs.size() != 4 || recs[0].id != 920004) { std::printf("KAT FAILED\n"); return 1; }
  std::printf("KAT ok: %zu recommendations, first %llu score %.9f\n", recs.size(), (unsigned long long)recs[0].id, recs[0].score);
  return 0;
}

int main(int argc, char** argv) {
  try {
    if (argc >= 2 && !std::strcmp(argv[1], "--kat")) return kat();
    if (argc < 3) { std::fprintf(stderr, "usage: %s train.txt test.txt [m k how_many max_items idf]\n", argv[0]); return 2; }
    const size_t m = argc > 3 ? std::strtoull(argv[3], nullptr, 10) : 500, k = argc > 4 ? std::strtoull(argv[4], nullptr, 10) : 50;
    const size_t how_many = argc > 5 ? std::strtoull(argv[5], nullptr, 10) : 21, max_items = argc > 6 ? std::strtoull(argv[6], nullptr, 10) : 2;
    const double idf_w = argc > 7 ? std::strtod(argv[7], nullptr) : 1.0;
    // evaluator.rs:19-35: a directory is the offline-computed Avro index, a file is a training TSV
    struct stat st{};
    if (stat(argv[1], &st) != 0) { std::fprintf(stderr, "Training data file does not exist: %s\n", argv[1]); return 2; }
    auto index = S_ISDIR(st.st_mode) ? vmis::VMISIndex::new_(argv[1]) : vmis::VMISIndex::new_from_csv(argv[1], m, idf_w);
    // io.rs:40-59 read_test_data_evolving
    std::map<uint64_t, std::vector<std::pair<uint64_t, uint64_t>>> by_session;
    FILE* f = std::fopen(argv[2], "rb");
    if (!f) { std::fprintf(stderr, "cannot open %s\n", argv[2]); return 2; }
    char line[4096]; bool header = true;
    while (std::fgets(line, sizeof line, f)) {
      if (header) { header = false; continue; }
      unsigned long long s, i; double t;
      if (std::sscanf(line, "%llu %llu %lf", &s, &i, &t) == 3) by_session[s].push_back({(uint64_t)std::llround(t), i});
    }
    std::fclose(f);
    std::vector<uint64_t> q_items; std::vector<uint32_t> q_off{0};
    std::vector<std::vector<uint64_t>> next_items;                               // evaluator.rs:73: the rest of the session
    for (auto& kv : by_session) {
      auto& ev = kv.second;
      std::stable_sort(ev.begin(), ev.end(), [](auto& a, auto& b) { return a.first < b.first; });
      for (size_t state = 1; state < ev.size(); ++state) {                       // evaluator.rs:47-56
        const size_t start = state > max_items ? state - max_items : 0;
        for (size_t j = start; j < state; ++j) q_items.push_back(ev[j].second);
        q_off.push_back((uint32_t)q_items.size());
        next_items.emplace_back();
        for (size_t j = state; j < ev.size(); ++j) next_items.back().push_back(ev[j].second);
      }
    }
    // Popularity / Coverage are built from the training rows (metrics/popularity.rs:20-37, coverage.rs:17-27)
    std::unordered_map<uint64_t, int> freq; int max_freq = 0;
    if (!S_ISDIR(st.st_mode)) {
      FILE* tf = std::fopen(argv[1], "rb");
      bool hdr = true;
      while (tf && std::fgets(line, sizeof line, tf)) {
        if (hdr) { hdr = false; continue; }
        unsigned long long s2, i2; double t2;
        if (std::sscanf(line, "%llu %llu %lf", &s2, &i2, &t2) == 3) max_freq = std::max(max_freq, ++freq[i2]);
      }
      if (tf) std::fclose(tf);
    }
    auto t0 = std::chrono::steady_clock::now();
    auto r = vmis::predict_batch(index, q_items, q_off, k, m, how_many, false);
    const double us = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count();
    const size_t n = next_items.size(), len = 20;
    double mrr = 0, ndcg = 0, hit = 0, pop = 0, prec = 0, rec = 0;
    std::set<uint64_t> covered;
    auto dcg = [](const std::vector<uint64_t>& top, const std::set<uint64_t>& next_set) {     // metrics/ndcg.rs:13-26
      double v = 0;
      for (size_t i = 0; i < top.size(); ++i) if (next_set.count(top[i])) v += i == 0 ? 1.0 : 1.0 / std::log2((double)i + 1.0);
      return v;
    };
    for (size_t q = 0; q < n; ++q) {
      std::vector<uint64_t> top(r.ids.begin() + q * how_many, r.ids.begin() + q * how_many + std::min<size_t>(r.counts[q], len));
      const std::vector<uint64_t>& nx = next_items[q];
      const std::set<uint64_t> next_set(nx.begin(), nx.end()), top_set(top.begin(), top.end());
      for (size_t j = 0; j < top.size(); ++j) if (top[j] == nx[0]) { mrr += 1.0 / (double)(j + 1); hit += 1; break; }   // mrr.rs, hitrate.rs
      std::vector<uint64_t> ideal(nx.begin(), nx.begin() + std::min(nx.size(), len));
      ndcg += dcg(top, next_set) / dcg(ideal, next_set);                                   // ndcg.rs:44-58
      size_t inter = 0;
      for (uint64_t i : top_set) inter += next_set.count(i);
      prec += (double)inter / (double)len;                                                 // precision.rs:31-40
      rec += (double)inter / (double)nx.size();                                            // recall.rs:33-45
      if (!top_set.empty() && max_freq > 0) {                                              // popularity.rs:41-58
        double sum = 0;
        for (uint64_t i : top_set) { auto it = freq.find(i); if (it != freq.end()) sum += (double)it->second / (double)max_freq; }
        pop += sum / (double)top_set.size();
      }
      covered.insert(top.begin(), top.end());                                              // coverage.rs:31-39
    }
    const double dn = n ? (double)n : 1.0, p = prec / dn, rc = rec / dn;
    const double f1 = (p + rc) > 0 ? 2.0 * p * rc / (p + rc) : 0.0;                        // f1score.rs:29-39
    const double cov = freq.empty() ? 0.0 : (double)covered.size() / (double)freq.size();
    std::printf("===============================================================\n");
    std::printf("===               START EVALUATING TEST FILE               ====\n");
    std::printf("===============================================================\n");
    std::printf("Mrr@20,Ndcg@20,HitRate@20,Popularity@20,Precision@20,Coverage@20,Recall@20,F1score@20\n");
    std::printf("%.4f,%.4f,%.4f,%.4f,%.4f,%.4f,%.4f,%.4f\n", mrr / dn, ndcg / dn, hit / dn, pop / dn, p, cov, rc, f1);
    std::printf("Qty test evaluations: %zu\n", n);
    std::printf("Prediction latency (whole batch, microseconds): %.0f  (%.2f per evaluation)\n", us, us / (double)n);
    return 0;
  } catch (const vmis::Error& e) {
    std::fprintf(stderr, "vmis error %d: %s\n", e.code, e.what());
    return 1;
  }
}
