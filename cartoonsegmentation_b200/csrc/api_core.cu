// Library-wide state of libcsb200.so: version, thread-local error text, launch counter, per-kernel profiler.
#include <cstdlib>
#include <map>
#include <set>
#include <mutex>
#include <string>
#include <vector>

#include "common.cuh"

namespace csb {
thread_local char g_err[512] = "";
std::atomic<uint64_t> g_launches{0};
std::atomic<int> g_profiling{0};

namespace {
std::mutex g_mu;
std::vector<cudaEvent_t> g_pool;
std::vector<std::pair<const char*, cudaEvent_t>> g_marks;
size_t g_next = 0;
std::set<std::string> g_labels;
}  // namespace

const char* profile_intern(const char* text) {
    std::lock_guard<std::mutex> lk(g_mu);
    return g_labels.insert(text).first->c_str();
}

void profile_mark(const char* what, cudaStream_t st) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (g_next == g_pool.size()) {
        cudaEvent_t e;
        if (cudaEventCreate(&e) != cudaSuccess) return;
        g_pool.push_back(e);
    }
    cudaEvent_t e = g_pool[g_next++];
    cudaEventRecord(e, st);
    g_marks.emplace_back(what, e);
}
}  // namespace csb

extern "C" int csb_version(void) { return 100; }  // 0.1.0
extern "C" const char* csb_last_error(void) { return csb::g_err; }
extern "C" uint64_t csb_launch_count(void) { return csb::g_launches.load(); }

extern "C" int csb_profile_begin(void* stream) {
    {
        std::lock_guard<std::mutex> lk(csb::g_mu);
        csb::g_marks.clear();
        csb::g_next = 0;
    }
    csb::profile_mark("begin", (cudaStream_t) stream);
    const char* detail = getenv("CSB_PROFILE_DETAIL");
    csb::g_profiling.store(detail && detail[0] == '1' ? 2 : 1);
    return CSB_OK;
}

extern "C" int csb_profile_end(char* json, size_t cap) {
    csb::g_profiling.store(0);
    std::lock_guard<std::mutex> lk(csb::g_mu);
    if (!csb::g_marks.empty()) cudaEventSynchronize(csb::g_marks.back().second);
    std::map<std::string, std::pair<double, long>> agg;
    for (size_t i = 1; i < csb::g_marks.size(); ++i) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, csb::g_marks[i - 1].second, csb::g_marks[i].second) != cudaSuccess) continue;
        auto& a = agg[csb::g_marks[i].first];
        a.first += ms;
        a.second += 1;
    }
    std::string out = "{";
    bool first = true;
    for (auto& kv : agg) {
        char buf[256];
        snprintf(buf, sizeof buf, "%s\"%s\": {\"ms\": %.6f, \"count\": %ld}", first ? "" : ", ", kv.first.c_str(), kv.second.first, kv.second.second);
        out += buf;
        first = false;
    }
    out += "}";
    if (json && cap) snprintf(json, cap, "%s", out.c_str());
    csb::g_marks.clear();
    csb::g_next = 0;
    return out.size() < cap ? CSB_OK : CSB_ERR_INVALID;
}
