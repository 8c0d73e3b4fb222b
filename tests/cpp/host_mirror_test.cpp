// CPU-side check of include/vmis.hpp on a host-only handle: the trait accessors behave like the reference's
// (similarity_indexed.rs:8-24) and the query entry points fail loudly without a GPU.
#include <cstdio>
#include <cmath>
#include "../../include/vmis.hpp"

#define CHECK(c) do { if (!(c)) { std::printf("FAILED: %s (line %d)\n", #c, __LINE__); return 1; } } while (0)

int main() {
  // mod.rs:229-310 data
  std::vector<std::vector<uint64_t>> sessions = {{920006, 920005, 920004}, {920005, 920004, 920003, 920002}};
  auto index = vmis::VMISIndex::from_sessions(sessions, {1, 1}, 5, 5, 1.0, VMIS_DEVICE_NONE);
  auto s0 = index.items_for_session(0);
  CHECK(s0.second == 3 && s0.first[0] == 920006);
  CHECK(std::fabs(index.idf(920005) - std::log(3.5)) < 1e-15);       // 7 pairs / df 2
  CHECK(std::fabs(index.idf(920002) - std::log(7.0)) < 1e-15);
  bool threw = false;
  try { index.idf(42); } catch (const vmis::Error& e) { threw = e.code == VMIS_ERR_ARG; }
  CHECK(threw);                                                        // the reference panics (vmis_index.rs:322)
  auto a = index.find_attributes(920004);
  CHECK(a && a->is_for_sale && !a->is_adult);                          // vmis_index.rs:514-517
  CHECK(!index.find_attributes(42));
  CHECK(index.stats().n_pairs_kept == 7 && index.stats().n_items == 5);
  threw = false;
  try { vmis::predict(index, {920005}, 500, 500, 20, false); } catch (const vmis::Error& e) { threw = e.code == VMIS_ERR_CUDA; }
  CHECK(threw);                                                        // no CPU fallback
  threw = false;
  try { vmis::VMISIndex::new_from_csv("/nonexistent/train.txt", 500, 1.0, VMIS_DEVICE_NONE); } catch (const vmis::Error& e) { threw = e.code == VMIS_ERR_IO; }
  CHECK(threw);
  threw = false;
  try { vmis::VMISIndex::new_("/nonexistent/index_dir", VMIS_DEVICE_NONE); } catch (const vmis::Error& e) { threw = e.code == VMIS_ERR_IO; }
  CHECK(threw);                                                        // VMISIndex::new on a missing directory
  {
    // recommend_resource.rs:39-54: window of max_items_in_session items, repeats of the last item are not appended
    vmis::Server srv(index, 5, 5, 5, 2);
    srv.set_clock(1000000);
    CHECK((srv.session_window("s", 7, true) == std::vector<uint64_t>{7}));
    CHECK((srv.session_window("s", 7, true) == std::vector<uint64_t>{7}));
    CHECK((srv.session_window("s", 8, true) == std::vector<uint64_t>{7, 8}));
    CHECK((srv.session_window("s", 9, true) == std::vector<uint64_t>{8, 9}));
    CHECK((srv.session_window("s", 1, false) == std::vector<uint64_t>{1}));
    srv.set_clock(1000000 + 20 * 60 + 1);                               // idle for more than 20 minutes: starts over
    CHECK((srv.session_window("s", 3, true) == std::vector<uint64_t>{3}));
    threw = false;
    try { srv.recommend("s", 920005, true); } catch (const vmis::Error& e) { threw = e.code == VMIS_ERR_CUDA; }
    CHECK(threw);                                                      // host-only handle: no CPU fallback
  }
  std::printf("host mirror ok\n");
  return 0;
}
