// Shared helpers for the sm_100a kernels of microaligner_b200.
// Compiled with -fmad=false: the compiler never contracts a*b+c; every fused multiply-add in
// this tree is an explicit __fmaf_rn / fma() because parity with OpenCV's CPU arithmetic
// depends on where roundings happen.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include "../../include/microaligner_b200.h"

namespace ma {

void set_error(const std::string& s);
int get_option(int option);   // ma_set_option values (process-wide), 0 when never set
int cuda_fail(cudaError_t e, const char* what);

#define MA_CUDA_CHECK(expr)                                   \
    do {                                                      \
        cudaError_t _e = (expr);                              \
        if (_e != cudaSuccess) return ma::cuda_fail(_e, #expr); \
    } while (0)

#define MA_LAUNCH_CHECK(name)                                        \
    do {                                                             \
        cudaError_t _e = cudaGetLastError();                         \
        if (_e != cudaSuccess) return ma::cuda_fail(_e, name);       \
    } while (0)

// ---- launch accounting + optional per-kernel CUDA-event profiler (ma_profile_* in the C ABI) ----
enum KernelId {
    K_POLYEXP = 0, K_UPDATE0, K_BLUR_V, K_BLUR_H, K_WARP, K_MERGE_MAX, K_MERGE, K_PYRDOWN, K_PYRUP,
    K_MINMAX, K_DOG_ROW, K_DOG_COL, K_DOG_QUANT, K_NMI_HIST, K_NMI_ENTROPY, K_ZMIP, K_NORM_U8, K_SMALL, K_AFFINE, K_COUNT
};
void prof_begin(int id, cudaStream_t s, double units);
void prof_end(int id, cudaStream_t s);
struct KernelScope {  // wraps ONE kernel launch: counts it and, when profiling is on, brackets it with events
    int id; cudaStream_t s;
    KernelScope(int id_, cudaStream_t s_, double units = 0) : id(id_), s(s_) { prof_begin(id, s, units); }
    ~KernelScope() { prof_end(id, s); }
};

inline int invalid(const std::string& s) {
    set_error(s);
    return MA_ERR_INVALID;
}

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// Tile geometry shared by every tiled kernel (slicer.py / stitcher.py as index math).
struct TileGeom {
    int h, w;      // image size
    int Th, Tw;    // tile core size (T x T; = h x w in the untiled branch)
    int ov;        // overlap
    int Sh, Sw;    // window size = core + 2*ov
    int ny, nx;    // tile grid
    unsigned long long magic_h, magic_w;  // ceil(2^40 / Th), ceil(2^40 / Tw): exact n / T for n*T < 2^40
};

// n / Th and n / Tw without an integer division (n < 2^20, T < 2^20)
__host__ __device__ __forceinline__ int div_th(const TileGeom& g, int n) { return (int)(((unsigned long long)(unsigned)n * g.magic_h) >> 40); }
__host__ __device__ __forceinline__ int div_tw(const TileGeom& g, int n) { return (int)(((unsigned long long)(unsigned)n * g.magic_w) >> 40); }

static inline TileGeom make_geom(int h, int w, int T, int ov) {
    TileGeom g;
    g.h = h; g.w = w;
    if (T <= 0) { g.Th = h; g.Tw = w; g.ov = 0; }
    else { g.Th = T; g.Tw = T; g.ov = ov; }
    g.Sh = g.Th + 2 * g.ov;
    g.Sw = g.Tw + 2 * g.ov;
    g.ny = (h + g.Th - 1) / g.Th;
    g.nx = (w + g.Tw - 1) / g.Tw;
    g.magic_h = ((1ull << 40) + g.Th - 1) / g.Th;
    g.magic_w = ((1ull << 40) + g.Tw - 1) / g.Tw;
    return g;
}

__device__ __forceinline__ int reflect101(int i, int n) {
    // single reflection, valid for -n < i < 2n-1
    i = i < 0 ? -i : i;
    return i >= n ? 2 * (n - 1) - i : i;
}

__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// order-preserving float <-> uint key for atomicMin/atomicMax
__device__ __forceinline__ unsigned f2key(float f) {
    unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key2f(unsigned k) {
    unsigned u = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
    return __uint_as_float(u);
}

}  // namespace ma
