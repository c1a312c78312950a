// vf_trace.cu -- section timers and NVTX ranges with the reference's names.
//
// Stands in for MeshFEM's global benchmark timer (3rdParty/MeshFEM/src/lib/MeshFEM/GlobalBenchmark.hh, Timer.hh) and its python
// binding (3rdParty/MeshFEM/src/python_bindings/benchmark.cc:9-13): reset / start_timer_section / stop_timer_section /
// start_timer / stop_timer / report.  Sections nest: a section started while another one runs is reported as "outer:inner"
// (Timer.hh:6-9).  The host control flow of the library opens the sections the reference opens around the same steps
// ("CG Iterations", "Preamble", "V Cycle ...", "OC step", "Bisection", "Build load", ... -- MultigridSolver.hh:1067-1077,
// OptimalityCriterion.hh:52,91, LayerByLayer.hh:243-282) through TraceScope.
//
// The work inside a section is asynchronous GPU work, so a section's wall time only means something if the device is drained
// when the section closes: that is done only while the timers are enabled (vf_benchmark_enable, or VF_BENCHMARK=1), never on
// the benchmarked path.  NVTX ranges of the same names are emitted regardless (they cost nothing without a tool attached).
#include "vf_internal.cuh"
#include <nvtx3/nvToolsExt.h>
#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <sstream>
#include <vector>

namespace vf {

namespace {
struct TimerRec { double total = 0.0; long long calls = 0; bool running = false; std::chrono::steady_clock::time_point t0; };
struct TraceState {
    std::mutex m;
    bool enabled = false;
    std::map<std::string, TimerRec> sections;                    // full path "a:b:c" -> time
    std::map<std::string, std::map<std::string, TimerRec>> timers;   // section path -> plain timers
    std::vector<std::string> stack;                               // open sections (full paths)
    std::vector<std::string> messages;
    TraceState() { const char *e = std::getenv("VF_BENCHMARK"); enabled = e && e[0] == '1'; }
};
TraceState &state() { static TraceState s; return s; }
std::string current_path(TraceState &s) { return s.stack.empty() ? std::string() : s.stack.back(); }
void drain() { cudaDeviceSynchronize(); }
}

bool trace_enabled() { return state().enabled; }

void trace_push(const char *name) {
    nvtxRangePushA(name);
    TraceState &s = state();
    if (!s.enabled) return;
    std::lock_guard<std::mutex> lock(s.m);
    const std::string path = s.stack.empty() ? std::string(name) : s.stack.back() + ":" + name;
    TimerRec &r = s.sections[path];
    r.running = true; ++r.calls; r.t0 = std::chrono::steady_clock::now();
    s.stack.push_back(path);
}
void trace_pop(const char *name) {
    nvtxRangePop();
    TraceState &s = state();
    if (!s.enabled) return;
    drain();
    std::lock_guard<std::mutex> lock(s.m);
    // close the innermost open section of that name (and whatever was left open inside it)
    for (size_t i = s.stack.size(); i-- > 0;) {
        const std::string &p = s.stack[i];
        const size_t pos = p.rfind(':');
        if ((pos == std::string::npos ? p : p.substr(pos + 1)) != name) continue;
        const auto now = std::chrono::steady_clock::now();
        for (size_t k = s.stack.size(); k-- > i;) {
            TimerRec &r = s.sections[s.stack[k]];
            if (r.running) { r.total += std::chrono::duration<double>(now - r.t0).count(); r.running = false; }
        }
        s.stack.resize(i);
        return;
    }
}
static void plain_timer(const char *name, bool start) {
    TraceState &s = state();
    if (!s.enabled) return;
    if (!start) drain();
    std::lock_guard<std::mutex> lock(s.m);
    TimerRec &r = s.timers[current_path(s)][name];
    const auto now = std::chrono::steady_clock::now();
    if (start) { r.running = true; ++r.calls; r.t0 = now; }
    else if (r.running) { r.total += std::chrono::duration<double>(now - r.t0).count(); r.running = false; }
}

static std::string trace_report(bool includeMessages) {
    TraceState &s = state();
    std::lock_guard<std::mutex> lock(s.m);
    std::ostringstream os;
    if (includeMessages) for (const auto &msg : s.messages) os << msg << "\n";
    const auto now = std::chrono::steady_clock::now();
    for (const auto &kv : s.sections) {
        const TimerRec &r = kv.second;
        const double t = r.total + (r.running ? std::chrono::duration<double>(now - r.t0).count() : 0.0);
        const size_t depth = std::count(kv.first.begin(), kv.first.end(), ':');
        os << std::string(4 * depth, ' ') << kv.first << "\t" << t << "\t(" << r.calls << ")\n";
        auto it = s.timers.find(kv.first);
        if (it != s.timers.end()) for (const auto &tv : it->second)
            os << std::string(4 * (depth + 1), ' ') << kv.first << ":" << tv.first << "\t" << tv.second.total << "\t(" << tv.second.calls << ")\n";
    }
    auto top = s.timers.find(std::string());
    if (top != s.timers.end()) for (const auto &tv : top->second) os << tv.first << "\t" << tv.second.total << "\t(" << tv.second.calls << ")\n";
    return os.str();
}

} // namespace vf

extern "C" {
// benchmark.cc:9-13
void vf_benchmark_enable(int on) { vf::state().enabled = on != 0; }
int vf_benchmark_enabled(void) { return vf::state().enabled ? 1 : 0; }
void vf_benchmark_reset(void) {
    vf::TraceState &s = vf::state(); std::lock_guard<std::mutex> lock(s.m);
    s.sections.clear(); s.timers.clear(); s.stack.clear(); s.messages.clear();
}
void vf_benchmark_start_timer_section(const char *name) { vf::trace_push(name); }
void vf_benchmark_stop_timer_section(const char *name) { vf::trace_pop(name); }
void vf_benchmark_start_timer(const char *name) { vf::plain_timer(name, true); }
void vf_benchmark_stop_timer(const char *name) { vf::plain_timer(name, false); }
void vf_benchmark_add_message(const char *msg) { vf::TraceState &s = vf::state(); std::lock_guard<std::mutex> lock(s.m); s.messages.push_back(msg); }
// Writes the report (one line per section / timer: path, seconds, (invocations)) into buf; returns the length it needs
// (excluding the terminating 0), so a caller can size the buffer with a first call (buf = NULL, capacity = 0).
size_t vf_benchmark_report(int include_messages, char *buf, size_t capacity) {
    const std::string r = vf::trace_report(include_messages != 0);
    if (buf && capacity) { const size_t n = std::min(capacity - 1, r.size()); std::memcpy(buf, r.data(), n); buf[n] = 0; }
    return r.size();
}
}
