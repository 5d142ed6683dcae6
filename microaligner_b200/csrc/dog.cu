// K6 / K9: difference-of-Gaussians prefilter, global min/max, z-max-projection + 8-bit normalise.
//
//   ma_dog_u8             OptFlowRegistrator.dog          (reference optflow_reg/optflow_registrator.py:249-274)
//   ma_minmax             cv.minMaxIdx inside cv.normalize
//   ma_zmip_normalize_u8  read_and_max_project_pages      (reference shared_modules/utils.py:75-95)
//
// dog() = cv.normalize(img, 0, 1, MINMAX, CV_32F) -> GaussianBlur 41x41 sigma 5 and sigma 9 ->
// hs - ls -> cv.normalize(.., 0, 255, MINMAX, CV_8U).  Arithmetic reproduced (see oracle/cv_ops.py):
//   * first normalise: scale = float(1/(max-min)), shift = -float(min*scale), f = fmaf(px, scale, shift)
//   * row filter: s = 0; s = fmaf(x[j-20], k[j], s), j = 0..40 (REFLECT_101 at the image edge)
//   * column filter (symmetric): s = x0*k20; s = fmaf(x[+j] + x[-j], k[20+j], s), j = 1..20
//   * OpenCV's AVX2 separable filter leaves the last (w & 3) columns of the row pass and the last
//     (w & 7) columns of the column pass to scalar code that rounds multiply and add separately;
//     the same columns do so here, so the result is bit-identical to cv2 on an AVX2 host.
//   * second normalise: a = float(255*(1/(dmax-dmin))), b = float(-dmin*scale), u8 = sat(rint(fmaf(d,a,b)))
// Both global reductions stay on the device (ordered-key atomicMin/Max); nothing syncs with the host.
#include <cmath>
#include <atomic>
#include "common.cuh"
#include "packed.cuh"
#include "tma.cuh"

namespace ma {

struct DogTaps {
    float k5[41], k9[41];
    float2 k59[41];          // {k5[j], k9[j]}: both sigmas of the row pass in one packed f32x2 FMA
    float2 c5[21], c9[21];   // k[20 + j] duplicated {k, k}: taps of the symmetric column pass (packed f32x2)
    float2 negzero2;         // {-0.f, -0.f}, see packed.cuh:mul2
};

static void make_dog_taps(DogTaps& t) {
    // cv.getGaussianKernel(41, sigma, CV_32F): exp(-x^2 / (2 sigma^2)) / sum in f64, stored as f32
    for (int which = 0; which < 2; ++which) {
        double sigma = which ? 9.0 : 5.0, k[41], sum = 0;
        double scale2x = -0.5 / (sigma * sigma);
        for (int i = 0; i < 41; ++i) {
            double x = i - 20.0;
            k[i] = std::exp(scale2x * x * x);
            sum += k[i];
        }
        sum = 1. / sum;
        for (int i = 0; i < 41; ++i) {
            (which ? t.k9 : t.k5)[i] = (float)(k[i] * sum);
            (which ? t.k59[i].y : t.k59[i].x) = (float)(k[i] * sum);
        }
        for (int j = 0; j <= 20; ++j) {
            float v = (which ? t.k9 : t.k5)[20 + j];
            (which ? t.c9 : t.c5)[j] = make_float2(v, v);
        }
    }
    t.negzero2 = make_float2(-0.0f, -0.0f);
}

// ------------------------------------------------------------------------------------------------
// min / max
// ------------------------------------------------------------------------------------------------
__global__ void init_minmax_keys(unsigned* keys, int npairs) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < npairs) {
        keys[2 * i] = f2key(INFINITY);
        keys[2 * i + 1] = f2key(-INFINITY);
    }
}

// block-wide min/max, then ONE ordered-key atomic pair per block (same-address atomics serialise in L2)
__device__ __forceinline__ void block_minmax_commit(float lo, float hi, unsigned* keys) {
    __shared__ float s_lo[32], s_hi[32];
    int tid = threadIdx.y * blockDim.x + threadIdx.x, nw = (blockDim.x * blockDim.y + 31) >> 5;
    lo = warp_min(lo);
    hi = warp_max(hi);
    if ((tid & 31) == 0) { s_lo[tid >> 5] = lo; s_hi[tid >> 5] = hi; }
    __syncthreads();
    if (tid < 32) {
        lo = tid < nw ? s_lo[tid] : INFINITY;
        hi = tid < nw ? s_hi[tid] : -INFINITY;
        lo = warp_min(lo);
        hi = warp_max(hi);
        if (tid == 0) {
            atomicMin(&keys[0], f2key(lo));
            atomicMax(&keys[1], f2key(hi));
        }
    }
}

// 16-byte vector loads over the aligned interior of every row, scalar loads for the unaligned head / tail
// (the scalar version moved 2 bytes per thread per request and reached only 13 % of the DRAM rate, ncu)
template <typename T>
__device__ __forceinline__ void minmax_accum16(const uint4 q, float& lo, float& hi) {
    const T* e = reinterpret_cast<const T*>(&q);
#pragma unroll
    for (int i = 0; i < (int)(16 / sizeof(T)); ++i) {
        float v = (float)e[i];
        lo = fminf(lo, v);
        hi = fmaxf(hi, v);
    }
}

template <typename T>
__global__ void __launch_bounds__(256) minmax_kernel(const T* __restrict__ src, size_t pitch, int h, int w, unsigned* keys) {
    constexpr int V = 16 / sizeof(T);
    float lo = INFINITY, hi = -INFINITY;
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, nthr = gridDim.x * blockDim.x;
    for (int y = blockIdx.y; y < h; y += gridDim.y) {
        const T* row = (const T*)((const char*)src + (size_t)y * pitch);
        int head = (int)(((16 - (reinterpret_cast<uintptr_t>(row) & 15)) & 15) / sizeof(T));
        head = min(head, w);
        const int nvec = (w - head) / V, tail0 = head + nvec * V;
        const uint4* vp = reinterpret_cast<const uint4*>(row + head);
        for (int i = tid; i < nvec; i += nthr) minmax_accum16<T>(__ldg(vp + i), lo, hi);
        for (int i = tid; i < head + (w - tail0); i += nthr) {
            float v = (float)__ldg(row + (i < head ? i : tail0 + (i - head)));
            lo = fminf(lo, v);
            hi = fmaxf(hi, v);
        }
    }
    block_minmax_commit(lo, hi, keys);
}

// Default for dense images (pitch == row bytes, 16-byte aligned base; 5.4 vs 3.8 TB/s for the row loop above on a
// 12 000^2 u16 image, profiles/r02_ab_variants.log): the image as one flat array, four independent 16-byte loads in flight per thread and a grid of a few CTAs per
// SM, instead of a row loop that leaves each thread one or two loads per row.  min / max are order-independent.
template <typename T>
__global__ void __launch_bounds__(256) minmax_flat_kernel(const T* __restrict__ src, size_t n, unsigned* keys) {
    constexpr int V = 16 / sizeof(T);
    const size_t nvec = n / V;
    const uint4* vp = reinterpret_cast<const uint4*>(src);
    float lo = INFINITY, hi = -INFINITY;
    const size_t stride = (size_t)gridDim.x * 256;
    size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    for (; i + 3 * stride < nvec; i += 4 * stride) {
        const uint4 a = __ldg(vp + i), b = __ldg(vp + i + stride), c = __ldg(vp + i + 2 * stride), d = __ldg(vp + i + 3 * stride);
        minmax_accum16<T>(a, lo, hi);
        minmax_accum16<T>(b, lo, hi);
        minmax_accum16<T>(c, lo, hi);
        minmax_accum16<T>(d, lo, hi);
    }
    for (; i < nvec; i += stride) minmax_accum16<T>(__ldg(vp + i), lo, hi);
    for (size_t t = nvec * V + (size_t)blockIdx.x * 256 + threadIdx.x; t < n; t += stride) {
        const float v = (float)__ldg(src + t);
        lo = fminf(lo, v);
        hi = fmaxf(hi, v);
    }
    block_minmax_commit(lo, hi, keys);
}

__global__ void keys_to_float_kernel(const unsigned* keys, float* out2) {
    out2[0] = key2f(keys[0]);
    out2[1] = key2f(keys[1]);
}

// scale/shift of cv.normalize(.., 0, 1, NORM_MINMAX, CV_32F)
__device__ __forceinline__ void norm01_coeffs(const float* __restrict__ mm, float& a, float& b) {
    double smin = mm[0], smax = mm[1];
    double scale = (smax - smin > 2.220446049250313e-16) ? __ddiv_rn(1.0, __dsub_rn(smax, smin)) : 0.0;
    a = (float)scale;
    b = __fsub_rn(0.0f, (float)__dmul_rn(smin, (double)a));
}
// scale/shift of cv.normalize(.., 0, 255, NORM_MINMAX, CV_8U)
__device__ __forceinline__ void norm255_coeffs(const float* __restrict__ mm, float& a, float& b) {
    double smin = mm[0], smax = mm[1];
    double scale = __dmul_rn(255.0, (smax - smin > 2.220446049250313e-16) ? __ddiv_rn(1.0, __dsub_rn(smax, smin)) : 0.0);
    double shift = __dsub_rn(0.0, __dmul_rn(smin, scale));
    a = (float)scale;
    b = (float)shift;
}

// ------------------------------------------------------------------------------------------------
// row pass: a block owns one 512-pixel segment and walks down DOG_HROWS rows.  Per row the raw pixels
// (+20 REFLECT_101 halo) are normalised into shared memory; the loads of the NEXT row are issued
// before the current row is convolved, so their latency hides behind the 328 FMAs per thread.
// 4 consecutive outputs per thread, both sigmas from the same 44 staged inputs.
// ------------------------------------------------------------------------------------------------
constexpr int DOG_HT = 128;            // threads per row segment
constexpr int DOG_HW = DOG_HT * 4;     // outputs per row segment
constexpr int DOG_HROWS = 16;          // rows per block
constexpr int DOG_SEG = DOG_HW + 40;   // staged inputs per row
constexpr int DOG_NLD = (DOG_SEG + DOG_HT - 1) / DOG_HT;

template <typename T>
__global__ void __launch_bounds__(DOG_HT) dog_row_kernel(const T* __restrict__ src, size_t pitch, int h, int w,
                                                         const float* __restrict__ src_mm,
                                                         float* __restrict__ A5, float* __restrict__ A9, int wp,
                                                         const __grid_constant__ DogTaps taps, int ybeg, int yend) {
    // A5 / A9 hold image rows [ybeg, yend) only: plane row = y - ybeg
    __shared__ __align__(16) float seg[2][DOG_SEG];
    const int xb = blockIdx.x * DOG_HW;
    const int y_first = ybeg + blockIdx.y * DOG_HROWS;
    float a, b;
    norm01_coeffs(src_mm, a, b);
    int col[DOG_NLD];
#pragma unroll
    for (int q = 0; q < DOG_NLD; ++q) {
        int i = threadIdx.x + q * DOG_HT, xx = xb - 20 + i;
        col[q] = (i < DOG_SEG && xx < w + 20) ? reflect101(xx, w) : -1;
    }
    T raw[DOG_NLD];
    auto fetch = [&](int y) {
        const T* row = (const T*)((const char*)src + (size_t)y * pitch);
#pragma unroll
        for (int q = 0; q < DOG_NLD; ++q) raw[q] = col[q] >= 0 ? __ldg(row + col[q]) : (T)0;
    };
    if (y_first < yend) fetch(y_first);
    const int x0 = xb + threadIdx.x * 4;
    const bool fused = x0 < (w & ~3);  // vector body of OpenCV's row filter; the last w & 3 columns are scalar code
#pragma unroll 1
    for (int r = 0; r < DOG_HROWS; ++r) {
        const int y = y_first + r;
        if (y >= yend) break;
        float* sg = seg[r & 1];
#pragma unroll
        for (int q = 0; q < DOG_NLD; ++q) {
            int i = threadIdx.x + q * DOG_HT;
            if (i < DOG_SEG) sg[i] = col[q] >= 0 ? __fmaf_rn((float)raw[q], a, b) : 0.0f;
        }
        __syncthreads();
        if (y + 1 < yend && r + 1 < DOG_HROWS) fetch(y + 1);
        if (x0 >= w) continue;
        float in[44];
#pragma unroll
        for (int q = 0; q < 11; ++q) {
            float4 v = *reinterpret_cast<const float4*>(&sg[threadIdx.x * 4 + q * 4]);
            in[4 * q] = v.x; in[4 * q + 1] = v.y; in[4 * q + 2] = v.z; in[4 * q + 3] = v.w;
        }
        float s5[4] = {0.f, 0.f, 0.f, 0.f}, s9[4] = {0.f, 0.f, 0.f, 0.f};
        if (fused) {
            // both sigmas of an output share one packed f32x2 accumulator {s5, s9}: input i, broadcast to both halves,
            // meets tap j = i - o of output o -- the same 41 FMAs per sum in the same order, in half the issue slots
            // (the scalar version was issue-bound at 89 %, ncu r01)
            u64 acc[4] = {0ull, 0ull, 0ull, 0ull};
#pragma unroll
            for (int i = 0; i < 44; ++i) {
                const u64 p = pack2(in[i], in[i]);
#pragma unroll
                for (int o = 0; o < 4; ++o) {
                    const int j = i - o;
                    if (j >= 0 && j < 41) acc[o] = fma2(p, *reinterpret_cast<const u64*>(&taps.k59[j]), acc[o]);
                }
            }
#pragma unroll
            for (int o = 0; o < 4; ++o) {
                const float2 v = unpack2(acc[o]);
                s5[o] = v.x;
                s9[o] = v.y;
            }
        } else {  // scalar tail: multiply and add rounded separately
#pragma unroll
            for (int j = 0; j < 41; ++j)
#pragma unroll
                for (int o = 0; o < 4; ++o) {
                    s5[o] = __fadd_rn(s5[o], __fmul_rn(in[o + j], taps.k5[j]));
                    s9[o] = __fadd_rn(s9[o], __fmul_rn(in[o + j], taps.k9[j]));
                }
        }
        const size_t o = (size_t)(y - ybeg) * wp + x0;
        if (x0 + 3 < w) {
            *reinterpret_cast<float4*>(A5 + o) = make_float4(s5[0], s5[1], s5[2], s5[3]);
            *reinterpret_cast<float4*>(A9 + o) = make_float4(s9[0], s9[1], s9[2], s9[3]);
        } else {
            for (int q = 0; q < 4 && x0 + q < w; ++q) { A5[o + q] = s5[q]; A9[o + q] = s9[q]; }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// column pass: lane <-> column, 8 consecutive rows per thread; d = blur9 - blur5; min/max of d
// ------------------------------------------------------------------------------------------------

// ------------------------------------------------------------------------------------------------
// column pass: CTA = 64 columns x 64 rows.  The two row-filtered planes arrive as two TMA boxes
// (64 x 104 floats each, zero fill outside the plane); REFLECT_101 rows are mirrored in shared memory.
// A thread owns the column pair (2*lane, 2*lane+1) as one packed f32x2 value and 8 consecutive rows,
// with two 8-deep sliding windows in registers (2 LDS.64 per 8 x (FADD2, FFMA2)).
// d = blur9 - blur5 goes to D; min/max of d to the ordered-key atomics.
// ------------------------------------------------------------------------------------------------
constexpr int DC_OUT = 64, DC_ROWS = DC_OUT + 40, DC_ROWU = 32;

template <bool FUSED>
__device__ __forceinline__ void col_conv8x2(const u64* __restrict__ centre, const float2* __restrict__ k2, u64 nz, u64 (&acc)[8]) {
    u64 wp[8], wm[8];
    const u64 k0 = *reinterpret_cast<const u64*>(&k2[0]);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        u64 c = centre[j * DC_ROWU];
        wp[j] = c;
        wm[j] = c;
        acc[j] = mul2(c, k0, nz);
    }
    // step i = 1..20 (fully unrolled: static register rotation, see farneback.cu:conv8x2)
#pragma unroll
    for (int i = 1; i <= 20; ++i) {
        const int s = (i - 1) & 7;
        wp[s] = centre[(7 + i) * DC_ROWU];
        wm[(63 - s) & 7] = centre[-i * DC_ROWU];
        const u64 kk = *reinterpret_cast<const u64*>(&k2[i]);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            u64 pr = add2(wp[(j + s + 1) & 7], wm[(j + 63 - s) & 7]);
            acc[j] = FUSED ? fma2(pr, kk, acc[j]) : add2(acc[j], mul2(pr, kk, nz));
        }
    }
}

__global__ void __launch_bounds__(256) dog_col_kernel(const __grid_constant__ CUtensorMap mapA, int wp, int h, int w,
                                                      float* __restrict__ D, unsigned* keys_out,
                                                      const __grid_constant__ DogTaps taps, int ybeg, int yend, int plane_row0) {
    // the TMA planes hold image rows from plane_row0 on; D holds rows [ybeg, yend) (row = y - ybeg)
    extern __shared__ __align__(128) float smem[];
    __shared__ __align__(8) uint64_t bar;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int x0 = blockIdx.x * 64, y0 = ybeg + blockIdx.y * DC_OUT;
    const int v_first = y0 - 20;
    float* buf5 = smem;
    float* buf9 = smem + DC_ROWS * 64;
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        fence_mbar_init();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_expect_tx(&bar, 2 * DC_ROWS * 64 * sizeof(float));
        tma_load_3d(buf5, &mapA, x0, v_first - plane_row0, 0, &bar);
        tma_load_3d(buf9, &mapA, x0, v_first - plane_row0, 1, &bar);
    }
    mbar_wait(&bar, 0);
    if (v_first < 0 || v_first + DC_ROWS > h) {  // CTA-uniform: mirror the rows outside the image
        for (int p = threadIdx.x; p < DC_ROWS * 64; p += 256) {
            int r = p >> 6, c = p & 63, v = v_first + r;
            if (v < 0 || v >= h) {
                int rr = min(max(reflect101(min(v, 2 * h - 2), h) - v_first, 0), DC_ROWS - 1);
                buf5[p] = buf5[rr * 64 + c];
                buf9[p] = buf9[rr * 64 + c];
            }
        }
        __syncthreads();
    }
    float lo = INFINITY, hi = -INFINITY;
    const int x = x0 + 2 * lane, o0 = warp * 8;
    if (x < w && y0 + o0 < yend) {
        const u64 nz = *reinterpret_cast<const u64*>(&taps.negzero2);
        u64 s5[8], s9[8];
        const u64* c5 = reinterpret_cast<const u64*>(buf5) + (o0 + 20) * DC_ROWU + lane;
        const u64* c9 = reinterpret_cast<const u64*>(buf9) + (o0 + 20) * DC_ROWU + lane;
        if (x < (w & ~7)) {
            col_conv8x2<true>(c5, taps.c5, nz, s5);
            col_conv8x2<true>(c9, taps.c9, nz, s9);
        } else {
            col_conv8x2<false>(c5, taps.c5, nz, s5);
            col_conv8x2<false>(c9, taps.c9, nz, s9);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int y = y0 + o0 + j;
            if (y < yend) {
                float2 d = unpack2(sub2(s9[j], s5[j]));
                float* out = D + (size_t)(y - ybeg) * wp + x;
                if (x + 1 < w) {
                    *reinterpret_cast<float2*>(out) = d;
                    lo = fminf(lo, fminf(d.x, d.y));
                    hi = fmaxf(hi, fmaxf(d.x, d.y));
                } else {
                    out[0] = d.x;
                    lo = fminf(lo, d.x);
                    hi = fmaxf(hi, d.x);
                }
            }
        }
    }
    block_minmax_commit(lo, hi, keys_out);
}

__global__ void __launch_bounds__(256) dog_quant_kernel(const float* __restrict__ D, int wp, int h, int w,
                                                        const float* __restrict__ mm, uint8_t* __restrict__ dst, size_t dp, int ybeg,
                                                        int d_row0) {
    float a, b;
    norm255_coeffs(mm, a, b);
    int x = (blockIdx.x * blockDim.x + threadIdx.x) * 4, y = ybeg + blockIdx.y;
    if (x >= w) return;
    const float* row = D + (size_t)(y - d_row0) * wp;
    uint8_t* out = dst + (size_t)y * dp;
    auto q = [&](float d) { return (uint8_t)max(0, min(255, __float2int_rn(__fmaf_rn(d, a, b)))); };
    if (x + 3 < w && ((dp & 3) == 0) && ((((uintptr_t)dst) & 3) == 0)) {
        float4 v = *reinterpret_cast<const float4*>(row + x);
        uchar4 u = make_uchar4(q(v.x), q(v.y), q(v.z), q(v.w));
        *reinterpret_cast<uchar4*>(out + x) = u;
    } else {
        for (int i = 0; i < 4 && x + i < w; ++i) out[x + i] = q(row[x + i]);
    }
}

// ------------------------------------------------------------------------------------------------
// z max-projection + normalise to u8
// ------------------------------------------------------------------------------------------------
struct PagePtrs {
    const void* p[64];
};

template <typename T>
__global__ void __launch_bounds__(256) zmip_kernel(PagePtrs pages, int n, size_t pitch, int h, int w, T* __restrict__ mip, int mp,
                                                   unsigned* keys) {
    float lo = INFINITY, hi = -INFINITY;
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x < w) {
        for (int y = blockIdx.y; y < h; y += gridDim.y) {
            T m = 0;
            for (int i = 0; i < n; ++i) {
                T v = __ldg((const T*)((const char*)pages.p[i] + (size_t)y * pitch) + x);
                m = v > m ? v : m;
            }
            mip[(size_t)y * mp + x] = m;
            lo = fminf(lo, (float)m);
            hi = fmaxf(hi, (float)m);
        }
    }
    block_minmax_commit(lo, hi, keys);
}

template <typename T>
__global__ void __launch_bounds__(256) norm_u8_kernel(const T* __restrict__ mip, int mp, int h, int w, const unsigned* __restrict__ keys,
                                                      uint8_t* __restrict__ dst, size_t dp) {
    float a, b;
    float mm[2] = {key2f(keys[0]), key2f(keys[1])};
    norm255_coeffs(mm, a, b);
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= w) return;
    float v = (float)mip[(size_t)y * mp + x];
    dst[(size_t)y * dp + x] = (uint8_t)max(0, min(255, __float2int_rn(__fmaf_rn(v, a, b))));
}

static inline int pad4(int w) { return (w + 3) / 4 * 4; }

}  // namespace ma

using namespace ma;

extern "C" int ma_minmax(const void* src, size_t pitch, int dtype, int h, int w, float* out2, void* stream) {
    if (!src || !out2 || h <= 0 || w <= 0) return invalid("ma_minmax: bad argument");
    cudaStream_t s = (cudaStream_t)stream;
    // out2 doubles as the key scratch (2 x 4 bytes), converted in place at the end
    unsigned* keys = (unsigned*)out2;
    { KernelScope ks(K_SMALL, s); init_minmax_keys<<<1, 32, 0, s>>>(keys, 1); }
    dim3 grid(std::min(ceil_div(w, 256), 8), std::min(h, 1184));
    const size_t esz = dtype == MA_U8 ? 1 : dtype == MA_U16 ? 2 : 4;
    if ((dtype == MA_U8 || dtype == MA_U16 || dtype == MA_F32) &&
        pitch == (size_t)w * esz && (reinterpret_cast<uintptr_t>(src) & 15) == 0) {
        const size_t n = (size_t)h * w;
        const int blocks = (int)std::max<size_t>(1, std::min<size_t>(148 * 8, (n * esz / 16 + 1023) / 1024));
        { KernelScope ks(K_MINMAX, s, (double)n);
        if (dtype == MA_U8) minmax_flat_kernel<uint8_t><<<blocks, 256, 0, s>>>((const uint8_t*)src, n, keys);
        else if (dtype == MA_U16) minmax_flat_kernel<uint16_t><<<blocks, 256, 0, s>>>((const uint16_t*)src, n, keys);
        else minmax_flat_kernel<float><<<blocks, 256, 0, s>>>((const float*)src, n, keys); }
        { KernelScope ks(K_SMALL, s); keys_to_float_kernel<<<1, 1, 0, s>>>(keys, out2); }
        MA_LAUNCH_CHECK("minmax_flat_kernel");
        return MA_OK;
    }
    { KernelScope ks(K_MINMAX, s, (double)h * w);
    if (dtype == MA_U8) minmax_kernel<uint8_t><<<grid, 256, 0, s>>>((const uint8_t*)src, pitch, h, w, keys);
    else if (dtype == MA_U16) minmax_kernel<uint16_t><<<grid, 256, 0, s>>>((const uint16_t*)src, pitch, h, w, keys);
    else if (dtype == MA_F32) minmax_kernel<float><<<grid, 256, 0, s>>>((const float*)src, pitch, h, w, keys);
    else return invalid("ma_minmax: bad dtype"); }
    { KernelScope ks(K_SMALL, s); keys_to_float_kernel<<<1, 1, 0, s>>>(keys, out2); }
    MA_LAUNCH_CHECK("minmax_kernel");
    return MA_OK;
}

extern "C" size_t ma_dog_workspace_bytes(int h, int w) {
    if (h <= 0 || w <= 0) return 0;
    return 256 + 3 * (size_t)h * pad4(w) * sizeof(float);
}

extern "C" size_t ma_dog_diff_pitch_floats(int w) { return (size_t)pad4(w); }

// scratch of ma_dog_diff_rows for a band of `nrows` rows: keys + two row-filtered planes of nrows + 40 rows
extern "C" size_t ma_dog_band_workspace_bytes(int w, int nrows) {
    if (w <= 0 || nrows < 0) return 0;
    return 256 + 2 * (size_t)(nrows + 40) * pad4(w) * sizeof(float);
}

// phase 1: rows [row_begin, row_end) of d = blur9(f) - blur5(f), f = normalised src, plus min/max of those rows
extern "C" int ma_dog_diff_rows(const void* src, size_t src_pitch, int dtype, int h, int w, const float* src_minmax,
                                int row_begin, int row_end, float* diff, float* diff_minmax, void* workspace, void* stream) {
    if (!src || !src_minmax || !diff || !diff_minmax || !workspace) return invalid("ma_dog_diff_rows: null pointer");
    if (h < 21 || w < 21) return invalid("ma_dog_u8: image must be at least 21 x 21 (single-reflection border)");
    if (dtype != MA_U8 && dtype != MA_U16) return invalid("ma_dog_u8: dtype must be MA_U8 or MA_U16");
    if (row_begin < 0 || row_end > h || row_begin > row_end) return invalid("ma_dog_diff_rows: bad row range");
    cudaStream_t s = (cudaStream_t)stream;
    unsigned* keys = (unsigned*)workspace;
    int wp = pad4(w);
    const int ra = std::max(row_begin - 20, 0), rb = std::min(row_end + 20, h);   // rows the row pass produces
    float* A5 = (float*)((char*)workspace + 256);
    float* A9 = A5 + (size_t)(rb - ra) * wp;
    static const DogTaps taps = [] { DogTaps t; make_dog_taps(t); return t; }();   // thread-safe one-time init
    { KernelScope ks(K_SMALL, s); init_minmax_keys<<<1, 32, 0, s>>>(keys, 1); }
    if (row_begin < row_end) {
        double px = (double)(row_end - row_begin) * w;
        dim3 rg(ceil_div(w, DOG_HW), ceil_div(rb - ra, DOG_HROWS)), rbk(DOG_HT);
        { KernelScope ks(K_DOG_ROW, s, px);
        if (dtype == MA_U8) dog_row_kernel<uint8_t><<<rg, rbk, 0, s>>>((const uint8_t*)src, src_pitch, h, w, src_minmax, A5, A9, wp, taps, ra, rb);
        else dog_row_kernel<uint16_t><<<rg, rbk, 0, s>>>((const uint16_t*)src, src_pitch, h, w, src_minmax, A5, A9, wp, taps, ra, rb); }
        { KernelScope ks(K_DOG_COL, s, px);
        CUtensorMap mapA;
        if (!make_plane_map(&mapA, A5, (uint64_t)w, (uint64_t)(rb - ra), 2, (uint64_t)wp * 4, (uint64_t)(rb - ra) * wp * 4, 64, DC_ROWS)) {
            set_error("ma_dog_diff_rows: cuTensorMapEncodeTiled failed");
            return MA_ERR_CUDA;
        }
        static std::atomic<bool> attr_set[64];  // per device; setting the attribute twice is harmless
        int dev_id = 0;
        cudaGetDevice(&dev_id);
        if (dev_id < 0 || dev_id >= 64 || !attr_set[dev_id]) {
            MA_CUDA_CHECK(cudaFuncSetAttribute(dog_col_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * DC_ROWS * 64 * 4));
            if (dev_id >= 0 && dev_id < 64) attr_set[dev_id] = true;
        }
        dog_col_kernel<<<dim3(ceil_div(w, 64), ceil_div(row_end - row_begin, DC_OUT)), 256, 2 * DC_ROWS * 64 * 4, s>>>(
            mapA, wp, h, w, diff, keys, taps, row_begin, row_end, ra); }
    }
    { KernelScope ks(K_SMALL, s); keys_to_float_kernel<<<1, 1, 0, s>>>(keys, diff_minmax); }
    MA_LAUNCH_CHECK("dog diff kernels");
    return MA_OK;
}

// phase 2: rows [row_begin, row_end) of the u8 result from d and the GLOBAL min/max of d
extern "C" int ma_dog_quantize_rows(const float* diff, int diff_row0, int h, int w, const float* diff_minmax, int row_begin,
                                    int row_end, uint8_t* dst, size_t dst_pitch, void* stream) {
    if (!diff || !diff_minmax || !dst) return invalid("ma_dog_quantize_rows: null pointer");
    if (row_begin < diff_row0 || row_end > h || row_begin > row_end) return invalid("ma_dog_quantize_rows: bad row range");
    if (row_begin == row_end) return MA_OK;
    cudaStream_t s = (cudaStream_t)stream;
    KernelScope ks(K_DOG_QUANT, s, (double)(row_end - row_begin) * w);
    dog_quant_kernel<<<dim3(ceil_div(ceil_div(w, 4), 256), row_end - row_begin), 256, 0, s>>>(diff, pad4(w), h, w, diff_minmax, dst, dst_pitch, row_begin, diff_row0);
    MA_LAUNCH_CHECK("dog_quant_kernel");
    return MA_OK;
}

extern "C" int ma_dog_u8(const void* src, size_t src_pitch, int dtype, int h, int w,
                         uint8_t* dst, size_t dst_pitch, void* workspace, void* stream) {
    if (!src || !dst || !workspace) return invalid("ma_dog_u8: null pointer");
    if (h < 21 || w < 21) return invalid("ma_dog_u8: image must be at least 21 x 21 (single-reflection border)");
    float* mm = (float*)((char*)workspace + 64);       // [0,1] source min/max, [2,3] diff min/max
    float* D = (float*)((char*)workspace + 256) + 2 * (size_t)h * pad4(w);
    int rc = ma_minmax(src, src_pitch, dtype, h, w, mm, stream);
    if (rc) return rc;
    rc = ma_dog_diff_rows(src, src_pitch, dtype, h, w, mm, 0, h, D, mm + 2, workspace, stream);
    if (rc) return rc;
    return ma_dog_quantize_rows(D, 0, h, w, mm + 2, 0, h, dst, dst_pitch, stream);
}

extern "C" size_t ma_zmip_workspace_bytes(int h, int w, int dtype) {
    if (h <= 0 || w <= 0) return 0;
    return 256 + (size_t)h * pad4(w) * (dtype == MA_U8 ? 1 : 2);
}

extern "C" int ma_zmip_normalize_u8(const void* const* pages_host, int n_pages, size_t pitch, int dtype,
                                    int h, int w, uint8_t* dst, size_t dst_pitch, void* workspace, void* stream) {
    if (!pages_host || !dst || !workspace || h <= 0 || w <= 0) return invalid("ma_zmip_normalize_u8: bad argument");
    if (n_pages < 1 || n_pages > 64) return invalid("ma_zmip_normalize_u8: n_pages must be in [1, 64]");
    if (dtype != MA_U8 && dtype != MA_U16) return invalid("ma_zmip_normalize_u8: dtype must be MA_U8 or MA_U16");
    cudaStream_t s = (cudaStream_t)stream;
    PagePtrs pp;
    for (int i = 0; i < n_pages; ++i) pp.p[i] = pages_host[i];
    unsigned* keys = (unsigned*)workspace;
    void* mip = (char*)workspace + 256;
    int mp = pad4(w);
    { KernelScope ks(K_SMALL, s); init_minmax_keys<<<1, 32, 0, s>>>(keys, 1); }
    dim3 grid(ceil_div(w, 256), h), zgrid(ceil_div(w, 256), std::min(h, 2048));
    if (dtype == MA_U8) {
        { KernelScope ks(K_ZMIP, s, (double)h * w); zmip_kernel<uint8_t><<<zgrid, 256, 0, s>>>(pp, n_pages, pitch, h, w, (uint8_t*)mip, mp, keys); }
        { KernelScope ks(K_NORM_U8, s, (double)h * w); norm_u8_kernel<uint8_t><<<grid, 256, 0, s>>>((const uint8_t*)mip, mp, h, w, keys, dst, dst_pitch); }
    } else {
        { KernelScope ks(K_ZMIP, s, (double)h * w); zmip_kernel<uint16_t><<<zgrid, 256, 0, s>>>(pp, n_pages, pitch, h, w, (uint16_t*)mip, mp, keys); }
        { KernelScope ks(K_NORM_U8, s, (double)h * w); norm_u8_kernel<uint16_t><<<grid, 256, 0, s>>>((const uint16_t*)mip, mp, h, w, keys, dst, dst_pitch); }
    }
    MA_LAUNCH_CHECK("zmip kernels");
    return MA_OK;
}
