// K5a / K5b: image pyramid down-sampling and flow up-sampling.
//
//   ma_pyrdown     cv.pyrDown(img)               (reference optflow_reg/optflow_registrator.py:194)
//   ma_pyrup_flow  cv.pyrUp(flow*k, dstsize)     (reference optflow_reg/optflow_registrator.py:140,150,164,169,212)
//
// pyrDown is integer arithmetic (5x5 binomial, REFLECT_101, (s+128)>>8) -> bit-exact by construction.
// pyrUp reproduces OpenCV's float association: rows first -- interior even (s[i-1] + 6 s[i]) + s[i+1],
// odd (s[i]+s[i+1])*4, left edge 6 s[0] + 2 s[1], right edge s[n-2] + 7 s[n-1] and 8 s[n-1] -- then
// the generic 3-row column form (top reflect-101, bottom replicate) times 1/64.
#include "common.cuh"

namespace ma {

template <typename T>
__global__ void __launch_bounds__(256) pyrdown_kernel(const T* __restrict__ src, size_t sp, int h, int w,
                                                      T* __restrict__ dst, size_t dp, int oh, int ow, int ybeg) {
    int ox = blockIdx.x * blockDim.x + threadIdx.x;
    int oy = ybeg + blockIdx.y * blockDim.y + threadIdx.y;
    if (ox >= ow || oy >= oh) return;
    int xs[5], acc = 0;
#pragma unroll
    for (int d = 0; d < 5; ++d) xs[d] = reflect101(2 * ox + d - 2, w);
    const int k[5] = {1, 4, 6, 4, 1};
#pragma unroll
    for (int r = 0; r < 5; ++r) {
        int yy = reflect101(2 * oy + r - 2, h);
        const T* row = (const T*)((const char*)src + (size_t)yy * sp);
        int s = (int)__ldg(row + xs[0]) + 4 * (int)__ldg(row + xs[1]) + 6 * (int)__ldg(row + xs[2]) +
                4 * (int)__ldg(row + xs[3]) + (int)__ldg(row + xs[4]);
        acc += k[r] * s;
    }
    *((T*)((char*)dst + (size_t)oy * dp) + ox) = (T)((acc + 128) >> 8);
}

// Default: a thread owns one output column and marches down kPdRows output rows.  Every source row is reduced
// horizontally ONCE (the one-output-per-thread kernel above does it 2.5 times) from three aligned pair loads -- pixels
// (2x-2, 2x-1), (2x, 2x+1), (2x+2, 2x+3), consecutive words across the warp -- and kept in a five-deep rolling window
// of registers.  Columns whose taps leave the row take the scalar REFLECT_101 path.  Integer sums: bit-exact.
constexpr int kPdRows = 32, kPdThreads = 128;

template <typename T> struct PairOf;
template <> struct PairOf<uint8_t> { using type = unsigned short; };
template <> struct PairOf<uint16_t> { using type = unsigned int; };

// The column loop, compiled once for interior columns (pair loads, no lane-divergent branch: the loads of several
// source rows are in flight together) and once for columns whose taps leave the row (scalar REFLECT_101 loads).
template <typename T, bool INTERIOR>
__device__ __forceinline__ void pyrdown_march(const T* __restrict__ src, size_t sp, int h, int w, T* __restrict__ dst, size_t dp,
                                              int ox, int oy0, int oy1) {
    using P = typename PairOf<T>::type;
    constexpr int kBits = 8 * (int)sizeof(T);
    constexpr unsigned kMask = (1u << kBits) - 1u;
    const int cx = 2 * ox;
    int xs[5];
#pragma unroll
    for (int d = 0; d < 5; ++d) xs[d] = INTERIOR ? cx + d - 2 : reflect101(cx + d - 2, w);
    auto hrow = [&](int y) -> int {
        const T* row = (const T*)((const char*)src + (size_t)reflect101(y, h) * sp);
        if constexpr (INTERIOR) {
            const bool mid = ((size_t)row & (sizeof(P) - 1)) != 0;      // the row starts in the middle of a pair
            const P* q = (const P*)(row + cx - (mid ? 3 : 2));
            const unsigned a = __ldg(q), b = __ldg(q + 1), c = __ldg(q + 2);
            // aligned: (cx-2, cx-1) (cx, cx+1) (cx+2, cx+3);  mid-pair: (cx-3, cx-2) (cx-1, cx) (cx+1, cx+2)
            return mid ? (int)(a >> kBits) + 4 * (int)(b & kMask) + 6 * (int)(b >> kBits) + 4 * (int)(c & kMask) + (int)(c >> kBits)
                       : (int)(a & kMask) + 4 * (int)(a >> kBits) + 6 * (int)(b & kMask) + 4 * (int)(b >> kBits) + (int)(c & kMask);
        } else {
            return (int)__ldg(row + xs[0]) + 4 * (int)__ldg(row + xs[1]) + 6 * (int)__ldg(row + xs[2]) +
                   4 * (int)__ldg(row + xs[3]) + (int)__ldg(row + xs[4]);
        }
    };
    int r0 = hrow(2 * oy0 - 2), r1 = hrow(2 * oy0 - 1), r2 = hrow(2 * oy0);
#pragma unroll 4
    for (int oy = oy0; oy < oy1; ++oy) {
        const int r3 = hrow(2 * oy + 1), r4 = hrow(2 * oy + 2);
        *((T*)((char*)dst + (size_t)oy * dp) + ox) = (T)((r0 + 4 * r1 + 6 * r2 + 4 * r3 + r4 + 128) >> 8);
        r0 = r2;
        r1 = r3;
        r2 = r4;
    }
}

template <typename T>
__global__ void __launch_bounds__(kPdThreads) pyrdown_march_kernel(const T* __restrict__ src, size_t sp, int h, int w,
                                                                   T* __restrict__ dst, size_t dp, int oh, int ow, int ybeg) {
    const int ox = blockIdx.x * kPdThreads + threadIdx.x;
    const int oy0 = ybeg + blockIdx.y * kPdRows;
    if (ox >= ow || oy0 >= oh) return;
    const int oy1 = min(oy0 + kPdRows, oh);
    if (2 * ox >= 3 && 2 * ox + 3 < w)       // both pair alignments stay inside the row
        pyrdown_march<T, true>(src, sp, h, w, dst, dp, ox, oy0, oy1);
    else
        pyrdown_march<T, false>(src, sp, h, w, dst, dp, ox, oy0, oy1);
}

__device__ __forceinline__ float2 mul2(float2 a, float s) { return make_float2(__fmul_rn(a.x, s), __fmul_rn(a.y, s)); }
__device__ __forceinline__ float2 add2(float2 a, float2 b) { return make_float2(__fadd_rn(a.x, b.x), __fadd_rn(a.y, b.y)); }

// One thread produces the 2 x 2 output block (rows 2r, 2r+1; columns 2c, 2c+1) from the 3 x 3 source
// neighbourhood: 9 float2 loads for 4 outputs.  Horizontal pass per source row (unscaled, weights sum to 8):
//   even column 2c : interior (s[c-1] + 6 s[c]) + s[c+1];  c == 0: 6 s[0] + 2 s[1];  c == n-1: s[n-2] + 7 s[n-1]
//   odd column 2c+1: interior (s[c] + s[c+1]) * 4;         c == n-1: 8 s[n-1];       n == 1: 8 s[0] for both
__device__ __forceinline__ void up_row_pair(const float2* __restrict__ row, int n, int c, float scale, float2& even, float2& odd) {
    float2 sc = mul2(__ldg(row + c), scale);
    if (n == 1) {
        even = odd = mul2(sc, 8.0f);
        return;
    }
    if (c == n - 1) {
        float2 sl = mul2(__ldg(row + c - 1), scale);
        even = add2(sl, mul2(sc, 7.0f));
        odd = mul2(sc, 8.0f);
        return;
    }
    float2 sr = mul2(__ldg(row + c + 1), scale);
    odd = mul2(add2(sc, sr), 4.0f);
    if (c == 0) {
        even = add2(mul2(sc, 6.0f), mul2(sr, 2.0f));
    } else {
        float2 sl = mul2(__ldg(row + c - 1), scale);
        even = add2(add2(sl, mul2(sc, 6.0f)), sr);
    }
}

__global__ void __launch_bounds__(256) pyrup_flow_kernel(const float2* __restrict__ src, int h, int w,
                                                         float2* __restrict__ dst, int dh, int dw, float scale,
                                                         int ybeg, int yend) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;                  // source column
    const int r = (ybeg >> 1) + blockIdx.y * blockDim.y + threadIdx.y;    // source row
    if (c >= w || r >= h || 2 * r >= yend) return;
    const int r0 = h > 1 ? reflect101(r - 1, h) : 0, r2 = min(r + 1, h - 1);
    float2 e0, o0, e1, o1, e2, o2;
    up_row_pair(src + (size_t)r0 * w, w, c, scale, e0, o0);
    up_row_pair(src + (size_t)r * w, w, c, scale, e1, o1);
    up_row_pair(src + (size_t)r2 * w, w, c, scale, e2, o2);
    const int dy = 2 * r, dx = 2 * c;
    const bool has_odd_col = dx + 1 < dw;
    if (dy >= ybeg) {   // even output row: (r0 + 6 r1) + r2, generic 3-row form (top reflect-101, bottom replicate)
        float2* out = dst + (size_t)dy * dw + dx;
        out[0] = mul2(add2(add2(e0, mul2(e1, 6.0f)), e2), 0.015625f);
        if (has_odd_col) out[1] = mul2(add2(add2(o0, mul2(o1, 6.0f)), o2), 0.015625f);
    }
    if (dy + 1 < yend && dy + 1 < dh) {   // odd output row: (r1 + r2) * 4
        float2* out = dst + (size_t)(dy + 1) * dw + dx;
        out[0] = mul2(mul2(add2(e1, e2), 4.0f), 0.015625f);
        if (has_odd_col) out[1] = mul2(mul2(add2(o1, o2), 4.0f), 0.015625f);
    }
}

}  // namespace ma

using namespace ma;

extern "C" int ma_pyrdown_rows(const void* src, size_t src_pitch, int h, int w, int dtype,
                               void* dst, size_t dst_pitch, int row_begin, int row_end, void* stream) {
    if (!src || !dst || h < 3 || w < 3) return invalid("ma_pyrdown: bad argument (need h,w >= 3)");
    int oh = (h + 1) / 2, ow = (w + 1) / 2;
    if (row_begin < 0 || row_end > oh || row_begin > row_end) return invalid("ma_pyrdown: bad row range");
    if (row_begin == row_end) return MA_OK;
    cudaStream_t s = (cudaStream_t)stream;
    if (dtype != MA_U8 && dtype != MA_U16) return invalid("ma_pyrdown: dtype must be MA_U8 or MA_U16");
    KernelScope ks(K_PYRDOWN, s, 4.0 * (row_end - row_begin) * ow);
    static const bool simple = getenv("MA_PYRDOWN_SIMPLE") != nullptr;      // the one-output-per-thread kernel (A/B)
    if (simple) {
        dim3 block(32, 8), grid(ceil_div(ow, 32), ceil_div(row_end - row_begin, 8));
        if (dtype == MA_U8)
            pyrdown_kernel<uint8_t><<<grid, block, 0, s>>>((const uint8_t*)src, src_pitch, h, w, (uint8_t*)dst, dst_pitch, row_end, ow, row_begin);
        else
            pyrdown_kernel<uint16_t><<<grid, block, 0, s>>>((const uint16_t*)src, src_pitch, h, w, (uint16_t*)dst, dst_pitch, row_end, ow, row_begin);
    } else {
        dim3 grid(ceil_div(ow, kPdThreads), ceil_div(row_end - row_begin, kPdRows));
        if (dtype == MA_U8)
            pyrdown_march_kernel<uint8_t><<<grid, kPdThreads, 0, s>>>((const uint8_t*)src, src_pitch, h, w, (uint8_t*)dst, dst_pitch, row_end, ow, row_begin);
        else
            pyrdown_march_kernel<uint16_t><<<grid, kPdThreads, 0, s>>>((const uint16_t*)src, src_pitch, h, w, (uint16_t*)dst, dst_pitch, row_end, ow, row_begin);
    }
    MA_LAUNCH_CHECK("pyrdown_kernel");
    return MA_OK;
}

extern "C" int ma_pyrdown(const void* src, size_t src_pitch, int h, int w, int dtype,
                          void* dst, size_t dst_pitch, void* stream) {
    return ma_pyrdown_rows(src, src_pitch, h, w, dtype, dst, dst_pitch, 0, (h + 1) / 2, stream);
}

extern "C" int ma_pyrup_flow_rows(const float* src, int h, int w, float* dst, int dh, int dw, float scale,
                                  int row_begin, int row_end, void* stream) {
    if (!src || !dst || h <= 0 || w <= 0) return invalid("ma_pyrup_flow: bad argument");
    // cv.pyrUp accepts |dsize - 2*ssize| == dsize % 2; the reference only produces 2n and 2n-1
    if (!((dh == 2 * h || dh == 2 * h - 1) && (dw == 2 * w || dw == 2 * w - 1)) || dh <= 0 || dw <= 0)
        return invalid("ma_pyrup_flow: dstsize must be 2n or 2n-1 per axis");
    if (row_begin < 0 || row_end > dh || row_begin > row_end) return invalid("ma_pyrup_flow: bad row range");
    if (row_begin == row_end) return MA_OK;
    const int src_rows = ((row_end + 1) >> 1) - (row_begin >> 1);   // source rows whose 2-row output block touches the range
    dim3 block(32, 8), grid(ceil_div(w, 32), ceil_div(src_rows, 8));
    KernelScope ks(K_PYRUP, (cudaStream_t)stream, (double)(row_end - row_begin) * dw);
    pyrup_flow_kernel<<<grid, block, 0, (cudaStream_t)stream>>>((const float2*)src, h, w, (float2*)dst, dh, dw, scale, row_begin, row_end);
    MA_LAUNCH_CHECK("pyrup_flow_kernel");
    return MA_OK;
}

extern "C" int ma_pyrup_flow(const float* src, int h, int w, float* dst, int dh, int dw, float scale, void* stream) {
    return ma_pyrup_flow_rows(src, h, w, dst, dh, dw, scale, 0, dh, stream);
}
