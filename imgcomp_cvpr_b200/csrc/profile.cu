// Launch accounting: every kernel launch of the library is counted, and -- when
// enabled with ic_profile_enable(1) -- bracketed by CUDA events on the launching
// stream so bench.py can report the live duration of each kernel class.
#include <atomic>
#include <mutex>
#include <vector>

#include "common.cuh"

namespace ic {

namespace {
std::atomic<long long> g_launches{0};
std::atomic<int> g_enabled{0};
struct Span {
    int cls;
    cudaEvent_t a, b;
};
std::mutex g_mu;
std::vector<Span> g_spans;
std::vector<cudaEvent_t> g_free;
long long g_cls_launches[IC_PROF_NUM_CLASSES] = {0};

cudaEvent_t get_event() {
    if (!g_free.empty()) {
        cudaEvent_t e = g_free.back();
        g_free.pop_back();
        return e;
    }
    cudaEvent_t e;
    cudaEventCreate(&e);
    return e;
}
}  // namespace

ProfScope::ProfScope(int cls, cudaStream_t s, int n_launches) : cls_(cls), s_(s), on_(false) {
    g_launches.fetch_add(n_launches, std::memory_order_relaxed);
    if (g_enabled.load(std::memory_order_relaxed)) {
        std::lock_guard<std::mutex> lk(g_mu);
        a_ = get_event();
        b_ = get_event();
        g_cls_launches[cls] += n_launches;
        on_ = true;
        cudaEventRecord(a_, s);
    }
}

ProfScope::~ProfScope() {
    if (on_) {
        cudaEventRecord(b_, s_);
        std::lock_guard<std::mutex> lk(g_mu);
        g_spans.push_back({cls_, a_, b_});
    }
}

}  // namespace ic

using namespace ic;

extern "C" {

long long ic_launch_count(void) { return g_launches.load(); }

void ic_profile_enable(int on) { g_enabled.store(on ? 1 : 0); }

void ic_profile_reset(void) {
    std::lock_guard<std::mutex> lk(g_mu);
    for (auto& s : g_spans) {
        g_free.push_back(s.a);
        g_free.push_back(s.b);
    }
    g_spans.clear();
    for (auto& c : g_cls_launches) c = 0;
}

// Sum of event-to-event durations (ms) and launch count of one class since the last
// reset.  The caller must have synchronised the streams the launches went to.
int ic_profile_get(int cls, double* total_ms, long long* launches) {
    if (cls < 0 || cls >= IC_PROF_NUM_CLASSES || !total_ms || !launches) return IC_ERR_INVALID;
    std::lock_guard<std::mutex> lk(g_mu);
    double t = 0;
    for (auto& s : g_spans)
        if (s.cls == cls) {
            float ms = 0;
            if (cudaEventElapsedTime(&ms, s.a, s.b) != cudaSuccess) {
                cudaGetLastError();
                ic::set_error("ic_profile_get: events not complete (synchronise first)");
                return IC_ERR_STATE;
            }
            t += ms;
        }
    *total_ms = t;
    *launches = g_cls_launches[cls];
    return IC_OK;
}

}  // extern "C"
