// Status / device plumbing of the C ABI (include/meteotools_b200.h, section 2).
#include <atomic>
#include <stdarg.h>
#include <string.h>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "common.cuh"

namespace mt {

static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char *what, const char *file, int line)
{
    set_error("CUDA error %d (%s) in %s at %s:%d", (int)e, cudaGetErrorString(e), what, file, line);
    return MT_ECUDA;
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

struct ProfRec {
    const char *name;
    cudaEvent_t e0, e1;
};
static std::atomic<int> g_prof_on{0};
static std::mutex g_prof_mu;
static std::vector<ProfRec> g_prof;

ProfScope::ProfScope(const char *name, mt_stream_t stream) : slot(-1), s((cudaStream_t)stream)
{
    if (!g_prof_on.load(std::memory_order_relaxed)) return;
    ProfRec r;
    r.name = name;
    if (cudaEventCreate(&r.e0) != cudaSuccess || cudaEventCreate(&r.e1) != cudaSuccess) return;
    cudaEventRecord(r.e0, s);
    std::lock_guard<std::mutex> lock(g_prof_mu);
    slot = (int)g_prof.size();
    g_prof.push_back(r);
}

ProfScope::~ProfScope()
{
    if (slot < 0) return;
    std::lock_guard<std::mutex> lock(g_prof_mu);
    cudaEventRecord(g_prof[slot].e1, s);
}

int sm_count()
{
    static thread_local int cached_dev = -1, cached_sms = 148;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (dev != cached_dev) {
        int sms = 0;
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && sms > 0)
            cached_sms = sms;
        cached_dev = dev;
    }
    return cached_sms;
}

}  // namespace mt

extern "C" {

int mt_abi_version(void) { return MT_ABI_VERSION; }
const char *mt_last_error(void) { return mt::g_err; }

int mt_device_count(void)
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) return mt::cuda_fail(e, "cudaGetDeviceCount", __FILE__, __LINE__);
    return n;
}

int mt_set_device(int device)
{
    MT_CUDA(cudaSetDevice(device));
    return MT_OK;
}

int mt_sm_count(void) { return mt::sm_count(); }

int mt_profile_enable(int on)
{
    std::lock_guard<std::mutex> lock(mt::g_prof_mu);
    for (auto &r : mt::g_prof) {
        cudaEventDestroy(r.e0);
        cudaEventDestroy(r.e1);
    }
    mt::g_prof.clear();
    mt::g_prof_on.store(on ? 1 : 0);
    return MT_OK;
}

// Synchronises the device and writes one line per scope name: "name count total_ms\n".
int mt_profile_fetch(char *buf, int64_t buf_len)
{
    MT_REQUIRE(buf && buf_len > 0, "mt_profile_fetch: bad buffer");
    MT_CUDA(cudaDeviceSynchronize());
    std::lock_guard<std::mutex> lock(mt::g_prof_mu);
    std::map<std::string, std::pair<int, double>> agg;
    std::vector<std::string> order;
    for (auto &r : mt::g_prof) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, r.e0, r.e1) != cudaSuccess) continue;
        auto it = agg.find(r.name);
        if (it == agg.end()) {
            agg[r.name] = {1, (double)ms};
            order.push_back(r.name);
        } else {
            it->second.first += 1;
            it->second.second += ms;
        }
    }
    std::string out;
    for (auto &n : order) {
        char line[256];
        snprintf(line, sizeof(line), "%s %d %.6f\n", n.c_str(), agg[n].first, agg[n].second);
        out += line;
    }
    MT_REQUIRE((int64_t)out.size() + 1 <= buf_len, "mt_profile_fetch: buffer too small");
    memcpy(buf, out.c_str(), out.size() + 1);
    return MT_OK;
}
int64_t mt_launch_count(void) { return mt::g_launches.load(); }

}  // extern "C"
