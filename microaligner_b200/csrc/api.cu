// Error plumbing, version, launch counter and per-kernel event profiler of the C ABI
// (include/microaligner_b200.h).
#include <atomic>
#include <mutex>
#include <vector>
#include "common.cuh"

namespace ma {
static thread_local std::string g_last_error;
void set_error(const std::string& s) { g_last_error = s; }
int cuda_fail(cudaError_t e, const char* what) {
    g_last_error = std::string(what) + ": " + cudaGetErrorName(e) + " (" + cudaGetErrorString(e) + ")";
    return MA_ERR_CUDA;
}

static const char* kKernelNames[K_COUNT] = {
    "fb_polyexp", "fb_update", "fb_blur_v", "fb_blur_h", "warp_tiles", "tile_max", "merge_tiles", "pyrdown", "pyrup_flow",
    "minmax", "dog_row", "dog_col", "dog_quant", "nmi_hist", "nmi_entropy", "zmip", "norm_u8", "small", "warp_affine"};

static std::atomic<long long> g_launches{0};
static std::atomic<int> g_options[MA_OPT_COUNT];
int get_option(int option) { return (option >= 0 && option < MA_OPT_COUNT) ? g_options[option].load(std::memory_order_relaxed) : 0; }
static std::atomic<int> g_prof_on{0};
static std::mutex g_prof_mu;
struct Pending { int id; cudaEvent_t a, b; double units; };
static std::vector<Pending> g_pending;
static std::vector<cudaEvent_t> g_pool;
static double g_ms[K_COUNT], g_units[K_COUNT];
static long long g_n[K_COUNT];
static thread_local cudaEvent_t t_start;

static cudaEvent_t get_event() {
    if (!g_pool.empty()) { cudaEvent_t e = g_pool.back(); g_pool.pop_back(); return e; }
    cudaEvent_t e; cudaEventCreate(&e); return e;
}

void prof_begin(int id, cudaStream_t s, double units) {
    g_launches.fetch_add(1, std::memory_order_relaxed);
    if (!g_prof_on.load(std::memory_order_relaxed)) return;
    std::lock_guard<std::mutex> lk(g_prof_mu);
    t_start = get_event();
    cudaEventRecord(t_start, s);
    g_pending.push_back({id, t_start, nullptr, units});
}
void prof_end(int id, cudaStream_t s) {
    if (!g_prof_on.load(std::memory_order_relaxed)) return;
    std::lock_guard<std::mutex> lk(g_prof_mu);
    if (g_pending.empty() || g_pending.back().b) return;
    cudaEvent_t e = get_event();
    cudaEventRecord(e, s);
    g_pending.back().b = e;
}
static void prof_drain() {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    for (auto& p : g_pending) {
        if (!p.b) { g_pool.push_back(p.a); continue; }
        cudaEventSynchronize(p.b);
        float ms = 0;
        if (cudaEventElapsedTime(&ms, p.a, p.b) == cudaSuccess) { g_ms[p.id] += ms; g_n[p.id]++; g_units[p.id] += p.units; }
        g_pool.push_back(p.a); g_pool.push_back(p.b);
    }
    g_pending.clear();
}
}  // namespace ma

using namespace ma;

extern "C" int ma_version(void) { return 100; }
extern "C" int ma_set_option(int option, int value) {
    if (option < 0 || option >= MA_OPT_COUNT) return invalid("ma_set_option: unknown option");
    g_options[option].store(value);
    return MA_OK;
}
extern "C" const char* ma_last_error(void) { return g_last_error.c_str(); }
extern "C" long long ma_launch_count(void) { return g_launches.load(); }
extern "C" int ma_profile_kernels(void) { return K_COUNT; }
extern "C" const char* ma_profile_kernel_name(int id) { return (id >= 0 && id < K_COUNT) ? kKernelNames[id] : ""; }
extern "C" void ma_profile_enable(int on) {
    if (!on) prof_drain();
    g_prof_on.store(on ? 1 : 0);
}
extern "C" void ma_profile_reset(void) {
    prof_drain();
    for (int i = 0; i < K_COUNT; ++i) { g_ms[i] = 0; g_n[i] = 0; g_units[i] = 0; }
}
extern "C" int ma_profile_read(int id, double* total_ms, long long* launches, double* units) {
    if (id < 0 || id >= K_COUNT) return MA_ERR_INVALID;
    prof_drain();
    if (total_ms) *total_ms = g_ms[id];
    if (launches) *launches = g_n[id];
    if (units) *units = g_units[id];
    return MA_OK;
}
