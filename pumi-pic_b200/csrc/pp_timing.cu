// pp_timing.cu -- phase timers and NVTX ranges with the reference's labels.
//
// Replaces support/ppTiming.{hpp,cpp}: SetTimingVerbosity / EnableTiming / DisableTiming /
// RecordTime / SummarizeTime (:37-75; accumulation :67-100, table :168-213).  The reference times
// a phase with a host timer around a Kokkos fence; here a phase is bracketed by two CUDA events on
// the stream it runs on (PPTimeScope, pp_internal.cuh), nothing is fenced, and the elapsed times
// are folded into the table when it is read.  Every scope is also an NVTX range of the same name,
// so the labels show up in Nsight Systems timelines whether or not timing is enabled.
// Labels: "pumipic search_mesh" (adjacency.tpp:609), "pumipic search_2d" (adjacency.hpp:1152),
// "Search Mesh 3d" (:553), "<kind> rebuild", "<kind> count active particles", "<kind> SCS specific
// building", "<kind> PSToPs", "<kind> shuffle attempt" (SCS_rebuild.h:166-312), "<kind> particle
// migration" (SCS_migrate.h:218), "gyro scatter".
#include <algorithm>
#include <map>
#include <mutex>
#include <string.h>

#include <nvtx3/nvToolsExt.h>

#include "pp_internal.cuh"

namespace {
struct TimeInfo {
  std::string str;
  double time = 0, timeSq = 0, mn = 1e300, mx = 0;
  long count = 0;
  int order = 0;
};
struct Pending { cudaEvent_t a, b; int index; };
std::mutex g_mu;
std::vector<TimeInfo> g_ops;
std::map<std::string, int> g_index;
std::vector<Pending> g_pending;
std::vector<cudaEvent_t> g_free_events;
int g_enabled = 0, g_verbosity = 0, g_rank = 0;

int op_index(const std::string& s) {
  auto it = g_index.find(s);
  if (it != g_index.end()) return it->second;
  TimeInfo t; t.str = s; t.order = (int)g_ops.size();
  g_ops.push_back(t);
  g_index[s] = t.order;
  return t.order;
}
void add(int i, double seconds) {
  TimeInfo& t = g_ops[(size_t)i];
  t.time += seconds; t.timeSq += seconds * seconds; ++t.count;
  t.mx = std::max(t.mx, seconds); t.mn = std::min(t.mn, seconds);
  if (g_verbosity >= 1) fprintf(stderr, "%d %s (seconds) %f\n", g_rank, t.str.c_str(), seconds);
}
cudaEvent_t get_event() {
  if (!g_free_events.empty()) { cudaEvent_t e = g_free_events.back(); g_free_events.pop_back(); return e; }
  cudaEvent_t e = nullptr;
  cudaEventCreate(&e);
  return e;
}
// fold finished scopes into the table (all == true waits for the unfinished ones)
void resolve(bool all) {
  size_t keep = 0;
  for (size_t i = 0; i < g_pending.size(); ++i) {
    Pending& p = g_pending[i];
    if (all) cudaEventSynchronize(p.b);
    if (all || cudaEventQuery(p.b) == cudaSuccess) {
      float ms = 0;
      if (cudaEventElapsedTime(&ms, p.a, p.b) == cudaSuccess) add(p.index, ms * 1e-3);
      g_free_events.push_back(p.a); g_free_events.push_back(p.b);
    } else {
      g_pending[keep++] = p;
    }
  }
  g_pending.resize(keep);
  cudaGetLastError();
}
}  // namespace

PPTimeScope::PPTimeScope(cudaStream_t s, const char* label) : s_(s), a_(nullptr), index_(-1) {
  nvtxRangePushA(label);
  if (!g_enabled) return;
  std::lock_guard<std::mutex> l(g_mu);
  index_ = op_index(label);
  a_ = get_event();
  cudaEventRecord(a_, s_);
}
PPTimeScope::~PPTimeScope() {
  nvtxRangePop();
  if (index_ < 0) return;
  std::lock_guard<std::mutex> l(g_mu);
  cudaEvent_t b = get_event();
  cudaEventRecord(b, s_);
  g_pending.push_back({a_, b, index_});
  if (g_pending.size() > 4096) resolve(false);
}
const char* pp_kind_name(int kind) {
  switch (kind) {
    case PP_PS_SCS: return "SCS";
    case PP_PS_CSR: return "CSR";
    case PP_PS_DPS: return "DPS";
    case PP_PS_CABM: return "CabM";
    default: return "PS";
  }
}

extern "C" void pp_timing_enable(int32_t on) { g_enabled = on ? 1 : 0; }            // EnableTiming / DisableTiming
extern "C" void pp_timing_set_verbosity(int32_t v) { g_verbosity = v; }             // SetTimingVerbosity
extern "C" void pp_timing_set_rank(int32_t rank) { g_rank = rank; }
extern "C" void pp_timing_record(const char* label, double seconds) {               // RecordTime
  if (!g_enabled || !label || g_verbosity < 0) return;
  std::lock_guard<std::mutex> l(g_mu);
  add(op_index(label), seconds);
}
extern "C" void pp_timing_reset(void) {
  std::lock_guard<std::mutex> l(g_mu);
  resolve(true);
  g_ops.clear(); g_index.clear();
}
extern "C" int32_t pp_timing_count(void) {
  std::lock_guard<std::mutex> l(g_mu);
  resolve(true);
  return (int32_t)g_ops.size();
}
extern "C" pp_status pp_timing_get(int32_t i, char* name, int32_t name_cap, double* total_s, double* min_s,
                                   double* max_s, double* sum_sq, int64_t* calls) {
  std::lock_guard<std::mutex> l(g_mu);
  PP_REQUIRE(i >= 0 && i < (int32_t)g_ops.size(), "no such timing entry");
  const TimeInfo& t = g_ops[(size_t)i];
  if (name && name_cap > 0) { strncpy(name, t.str.c_str(), (size_t)name_cap - 1); name[name_cap - 1] = 0; }
  if (total_s) *total_s = t.time;
  if (min_s) *min_s = t.count ? t.mn : 0;
  if (max_s) *max_s = t.mx;
  if (sum_sq) *sum_sq = t.timeSq;
  if (calls) *calls = t.count;
  return PP_OK;
}
// SummarizeTime (ppTiming.cpp:168-213): sort 0 alphabetical, 1 order of first occurrence, 2 longest
// first, 3 shortest first
extern "C" void pp_timing_summarize(int32_t sort) {
  std::lock_guard<std::mutex> l(g_mu);
  resolve(true);
  if (!g_enabled || g_verbosity < 0) return;
  std::vector<TimeInfo> v = g_ops;
  if (sort == 0) std::sort(v.begin(), v.end(), [](const TimeInfo& a, const TimeInfo& b) { return a.str < b.str; });
  else if (sort == 2) std::sort(v.begin(), v.end(), [](const TimeInfo& a, const TimeInfo& b) { return a.time > b.time; });
  else if (sort == 3) std::sort(v.begin(), v.end(), [](const TimeInfo& a, const TimeInfo& b) { return a.time < b.time; });
  size_t w = strlen("Operation");
  for (const TimeInfo& t : v) w = std::max(w, t.str.size());
  fprintf(stderr, "Timing Summary %d\n%-*s  %-12s  %-12s  %-12s  %-12s  %-10s  %-12s\n", g_rank, (int)w, "Operation",
          "Total Time", "Min Time", "Max Time", "Sqr Average", "Call Count", "Average Time");
  for (const TimeInfo& t : v) {
    const double n = t.count ? (double)t.count : 1.0;
    fprintf(stderr, "%-*s  %-12.6g  %-12.6g  %-12.6g  %-12.6g  %-10ld  %-12.6g\n", (int)w, t.str.c_str(), t.time,
            t.count ? t.mn : 0.0, t.mx, t.timeSq / n, t.count, t.time / n);
  }
}
