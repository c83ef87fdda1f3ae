// Optional per-launch timing with CUDA events on the launching stream, aggregated per kernel class.
// bench.py turns it on for the timed region to report the dominant kernel's achieved FLOP/s / GB/s (roofline block).
#include <stdlib.h>
#include "mitb_internal.h"

namespace mitb {

thread_local Profiler* g_prof = nullptr;

static cudaEvent_t get_event(Profiler& p) {
  if (!p.pool.empty()) { cudaEvent_t e = p.pool.back(); p.pool.pop_back(); return e; }
  cudaEvent_t e; CUDA_OK(cudaEventCreate(&e)); return e;
}

ProfScope::ProfScope(const char* kind, double flops, double bytes, cudaStream_t st, int m, int k, int n) : st_(st) {
  p_ = (g_prof && g_prof->on) ? g_prof : nullptr;
  if (!p_) return;
  Profiler::Rec r; r.kind = kind; r.flops = flops; r.bytes = bytes; r.a = get_event(*p_); r.b = get_event(*p_); r.m = m; r.k = k; r.n = n;
  CUDA_OK(cudaEventRecord(r.a, st_));
  p_->recs.push_back(r);
}
ProfScope::~ProfScope() {
  if (!p_) return;
  cudaEventRecord(p_->recs.back().b, st_);
}

std::string profiler_report(Profiler& p) {
  struct Agg { long n = 0; double ms = 0, flops = 0, bytes = 0; };
  std::map<std::string, Agg> agg;
  std::string per_launch;
  const bool detail = getenv("MITB_PROFILE_LAUNCHES") != nullptr;
  for (auto& r : p.recs) {
    float ms = 0.f;
    cudaEventSynchronize(r.b);
    if (cudaEventElapsedTime(&ms, r.a, r.b) != cudaSuccess) ms = 0.f;
    if (detail && r.m) {
      char buf[160];
      snprintf(buf, sizeof buf, "%s[\"%s\", %d, %d, %d, %.4f]", per_launch.empty() ? "" : ", ", r.kind, r.m, r.k, r.n, ms);
      per_launch += buf;
    }
    Agg& a = agg[r.kind]; a.n++; a.ms += ms; a.flops += r.flops; a.bytes += r.bytes;
    p.pool.push_back(r.a); p.pool.push_back(r.b);
  }
  p.recs.clear();
  std::string s = "{";
  bool first = true;
  for (auto& kv : agg) {
    char buf[256];
    snprintf(buf, sizeof buf, "%s\"%s\": {\"launches\": %ld, \"ms\": %.6f, \"flops\": %.6e, \"bytes\": %.6e}", first ? "" : ", ",
             kv.first.c_str(), kv.second.n, kv.second.ms, kv.second.flops, kv.second.bytes);
    s += buf; first = false;
  }
  if (detail) s += std::string(first ? "" : ", ") + "\"_launches\": [" + per_launch + "]";
  return s + "}";
}

}  // namespace mitb
