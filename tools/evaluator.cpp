// tools/evaluator.cpp — the reference's `evaluator` binary (src/bin/evaluator.rs) on top of include/vmis.hpp:
// builds the index from a training TSV, replays every prefix of every test session (evaluator.rs:46-57) through
// ONE batched predict call, and prints qty evaluations, Mrr@20 and HitRate@20 (metrics/mrr.rs, metrics/hitrate.rs).
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
#include <string>

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
    std::vector<uint64_t> q_items, next_item; std::vector<uint32_t> q_off{0};
    for (auto& kv : by_session) {
      auto& ev = kv.second;
      std::stable_sort(ev.begin(), ev.end(), [](auto& a, auto& b) { return a.first < b.first; });
      for (size_t state = 1; state < ev.size(); ++state) {                       // evaluator.rs:47-56
        const size_t start = state > max_items ? state - max_items : 0;
        for (size_t j = start; j < state; ++j) q_items.push_back(ev[j].second);
        q_off.push_back((uint32_t)q_items.size());
        next_item.push_back(ev[state].second);
      }
    }
    auto t0 = std::chrono::steady_clock::now();
    auto r = vmis::predict_batch(index, q_items, q_off, k, m, how_many, false);
    const double us = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count();
    double rr = 0; size_t hits = 0; const size_t n = next_item.size(), len = std::min<size_t>(20, how_many);
    for (size_t q = 0; q < n; ++q)
      for (size_t j = 0; j < std::min<size_t>(r.counts[q], len); ++j)
        if (r.ids[q * how_many + j] == next_item[q]) { rr += 1.0 / (double)(j + 1); ++hits; break; }
    std::printf("===============================================================\n");
    std::printf("===               START EVALUATING TEST FILE               ====\n");
    std::printf("===============================================================\n");
    std::printf("Mrr@20,HitRate@20\n%.4f,%.4f\n", rr / (double)n, (double)hits / (double)n);
    std::printf("Qty test evaluations: %zu\n", n);
    std::printf("Prediction latency (whole batch, microseconds): %.0f  (%.2f per evaluation)\n", us, us / (double)n);
    return 0;
  } catch (const vmis::Error& e) {
    std::fprintf(stderr, "vmis error %d: %s\n", e.code, e.what());
    return 1;
  }
}
