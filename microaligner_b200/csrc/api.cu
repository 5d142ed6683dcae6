// Error plumbing and version of the C ABI (include/microaligner_b200.h).
#include "common.cuh"

namespace ma {
static thread_local std::string g_last_error;
void set_error(const std::string& s) { g_last_error = s; }
int cuda_fail(cudaError_t e, const char* what) {
    g_last_error = std::string(what) + ": " + cudaGetErrorName(e) + " (" + cudaGetErrorString(e) + ")";
    return MA_ERR_CUDA;
}
}  // namespace ma

extern "C" int ma_version(void) { return 100; }
extern "C" const char* ma_last_error(void) { return ma::g_last_error.c_str(); }
