// K4 / K8: bilinear remap with OpenCV's INTER_LINEAR fixed-point scheme, tiles as index math.
//
//   ma_warp_tiles        Warper.warp            (reference optflow_reg/warper.py:37-76)
//   ma_merge_flows_tiles merge_two_flows+stitch (reference optflow_reg/optflow_registrator.py:37-47,217-240)
//
// Semantics reproduced (cv::remap, INTER_LINEAR, BORDER_CONSTANT 0, INTER_TAB_SIZE = 32):
//   s = cvRound(map * 32) (round-half-even, "integer indefinite" on overflow), i = sat_s16(s >> 5),
//   a = s & 31; a tap outside the S x S tile window -- or outside the image, where the window is
//   zero padded -- contributes 0.  u8: 15-bit integer weights, (acc + 2^14) >> 15.
//   u16 / f32: float weights (1-ay)(1-ax).. summed left to right, no FMA; u16 rounds half-even.
//
// HBM-bound streaming kernels: one thread per output pixel, 8-byte coalesced flow loads, the four
// taps go through the read-only path (neighbouring threads hit the same 32-byte sectors).
#include "common.cuh"

namespace ma {

__device__ __forceinline__ int cv_round_x32(float v) {
    float p = __fmul_rn(v, 32.0f);
    // cvtss2si: out-of-range / NaN -> 0x80000000
    if (!(fabsf(p) < 2147483648.0f)) return (int)0x80000000;
    return __float2int_rn(p);
}
__device__ __forceinline__ int sat_s16(int v) { return max(-32768, min(32767, v)); }

struct Taps {
    int ix, iy, ax, ay;
};
__device__ __forceinline__ Taps make_taps(float mx, float my) {
    int sx = cv_round_x32(mx), sy = cv_round_x32(my);
    Taps t;
    t.ix = sat_s16(sx >> 5);
    t.iy = sat_s16(sy >> 5);
    t.ax = sx & 31;
    t.ay = sy & 31;
    return t;
}

// one output pixel (x, y) of tile row ti: window geometry along y is passed in (shared by the pixels of a thread)
template <typename T>
__device__ __forceinline__ T warp_pixel(const T* __restrict__ img, size_t img_pitch, const TileGeom& g, float2 f, int x, int ty, int oy) {
    const int tj = div_tw(g, x);
    const int tx = g.ov + (x - tj * g.Tw);  // tile-local pixel
    const int ox = tj * g.Tw - g.ov;        // window origin in the image
    Taps t = make_taps(__fsub_rn((float)tx, f.x), __fsub_rn((float)ty, f.y));
    T v00, v01, v10, v11;
    const int gy0 = oy + t.iy, gx0 = ox + t.ix;
    if ((unsigned)t.ix < (unsigned)(g.Sw - 1) && (unsigned)t.iy < (unsigned)(g.Sh - 1) &&
        (unsigned)gx0 < (unsigned)(g.w - 1) && (unsigned)gy0 < (unsigned)(g.h - 1)) {
        // common case: all four taps inside the tile window and the image
        const T* p0 = (const T*)((const char*)img + (size_t)gy0 * img_pitch) + gx0;
        const T* p1 = (const T*)((const char*)p0 + img_pitch);
        v00 = __ldg(p0); v01 = __ldg(p0 + 1); v10 = __ldg(p1); v11 = __ldg(p1 + 1);
    } else {
        auto tap = [&](int yy, int xx) -> T {
            // inside the S x S window and inside the image (window is zero padded beyond it)
            int gy = oy + yy, gx = ox + xx;
            bool ok = (unsigned)xx < (unsigned)g.Sw && (unsigned)yy < (unsigned)g.Sh &&
                      (unsigned)gx < (unsigned)g.w && (unsigned)gy < (unsigned)g.h;
            return ok ? __ldg((const T*)((const char*)img + (size_t)gy * img_pitch) + gx) : (T)0;
        };
        v00 = tap(t.iy, t.ix); v01 = tap(t.iy, t.ix + 1); v10 = tap(t.iy + 1, t.ix); v11 = tap(t.iy + 1, t.ix + 1);
    }
    if (sizeof(T) == 1) {
        int w00 = (32 - t.ay) * (32 - t.ax) * 32, w01 = (32 - t.ay) * t.ax * 32;
        int w10 = t.ay * (32 - t.ax) * 32, w11 = t.ay * t.ax * 32;
        int acc = (int)v00 * w00 + (int)v01 * w01 + (int)v10 * w10 + (int)v11 * w11;
        return (T)((acc + 16384) >> 15);
    }
    float fx = __fmul_rn((float)t.ax, 0.03125f), fy = __fmul_rn((float)t.ay, 0.03125f);
    float ux = __fsub_rn(1.0f, fx), uy = __fsub_rn(1.0f, fy);
    float w00 = __fmul_rn(uy, ux), w01 = __fmul_rn(uy, fx), w10 = __fmul_rn(fy, ux), w11 = __fmul_rn(fy, fx);
    float acc = __fmul_rn((float)v00, w00);
    acc = __fadd_rn(acc, __fmul_rn((float)v01, w01));
    acc = __fadd_rn(acc, __fmul_rn((float)v10, w10));
    acc = __fadd_rn(acc, __fmul_rn((float)v11, w11));
    int q = __float2int_rn(acc);
    return (T)max(0, min(65535, q));
}

// Four consecutive pixels of a row per thread: the flow arrives as two 128-bit loads and the result leaves as one
// 4- or 8-byte store, so a warp has 1 KB of flow and 16 x 32 taps in flight per row instead of 256 B and 4 x 32
// (the one-pixel-per-thread version ran at 36 % of HBM peak, latency-bound).  VEC = false: pointers / pitches do not
// allow the vector accesses (odd width, unaligned views) -- same arithmetic with scalar loads and stores.
template <typename T, bool VEC>
__global__ void __launch_bounds__(256) warp_tiles_kernel(const T* __restrict__ img, size_t img_pitch,
                                                         const float2* __restrict__ flow, TileGeom g,
                                                         T* __restrict__ out, size_t out_pitch, int ybeg, int yend) {
    const int x0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    const int y = ybeg + blockIdx.y * blockDim.y + threadIdx.y;
    if (x0 >= g.w || y >= yend) return;
    const int ti = div_th(g, y);
    const int ty = g.ov + (y - ti * g.Th), oy = ti * g.Th - g.ov;
    const float2* frow = flow + (size_t)y * g.w + x0;
    T* orow = (T*)((char*)out + (size_t)y * out_pitch) + x0;
    if (VEC && x0 + 4 <= g.w) {
        const float4 fa = __ldg(reinterpret_cast<const float4*>(frow)), fb = __ldg(reinterpret_cast<const float4*>(frow) + 1);
        const T r0 = warp_pixel<T>(img, img_pitch, g, make_float2(fa.x, fa.y), x0, ty, oy);
        const T r1 = warp_pixel<T>(img, img_pitch, g, make_float2(fa.z, fa.w), x0 + 1, ty, oy);
        const T r2 = warp_pixel<T>(img, img_pitch, g, make_float2(fb.x, fb.y), x0 + 2, ty, oy);
        const T r3 = warp_pixel<T>(img, img_pitch, g, make_float2(fb.z, fb.w), x0 + 3, ty, oy);
        if (sizeof(T) == 1) {
            *reinterpret_cast<uchar4*>(orow) = make_uchar4((unsigned char)r0, (unsigned char)r1, (unsigned char)r2, (unsigned char)r3);
        } else {
            *reinterpret_cast<ushort4*>(orow) = make_ushort4((unsigned short)r0, (unsigned short)r1, (unsigned short)r2, (unsigned short)r3);
        }
    } else {
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (x0 + k < g.w) orow[k] = warp_pixel<T>(img, img_pitch, g, __ldg(frow + k), x0 + k, ty, oy);
    }
}

// ---- merge: pass 0 = per-tile signed max of both flows over the full window (zero padding counts)
__global__ void __launch_bounds__(256) tile_max_kernel(const float2* __restrict__ f1, const float2* __restrict__ f2,
                                                       TileGeom g, unsigned* __restrict__ keys, int tile0) {
    // grid: (chunks, ntiles); each block scans a slice of the window rows of one tile
    int tile = tile0 + blockIdx.y;
    int ti = tile / g.nx, tj = tile % g.nx;
    int oy = ti * g.Th - g.ov, ox = tj * g.Tw - g.ov;
    int y0 = max(oy, 0), y1 = min(oy + g.Sh, g.h);
    int x0 = max(ox, 0), x1 = min(ox + g.Sw, g.w);
    bool padded = (oy < 0) || (ox < 0) || (oy + g.Sh > g.h) || (ox + g.Sw > g.w);
    float m1 = padded ? 0.0f : -INFINITY, m2 = m1;
    int nw = x1 - x0, nh = y1 - y0;
    long long total = (long long)nw * nh;
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < total; p += (long long)gridDim.x * blockDim.x) {
        int yy = y0 + (int)(p / nw), xx = x0 + (int)(p % nw);
        float2 a = __ldg(&f1[(size_t)yy * g.w + xx]);
        float2 b = __ldg(&f2[(size_t)yy * g.w + xx]);
        m1 = fmaxf(m1, fmaxf(a.x, a.y));
        m2 = fmaxf(m2, fmaxf(b.x, b.y));
    }
    m1 = warp_max(m1);
    m2 = warp_max(m2);
    if ((threadIdx.x & 31) == 0) {
        atomicMax(&keys[2 * tile], f2key(m1));
        atomicMax(&keys[2 * tile + 1], f2key(m2));
    }
}

__global__ void init_keys_kernel(unsigned* keys, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) keys[i] = f2key(-INFINITY);
}

// ---- merge: pass 1 = per output pixel, branch on the tile's maxima (quirk Q3), sample f2 at the
// ABSOLUTE tile-local coordinate -f1 (quirk Q1: the reference passes map = -flow1 to cv.remap).
__global__ void __launch_bounds__(256) merge_tiles_kernel(const float2* __restrict__ f1, const float2* __restrict__ f2,
                                                          TileGeom g, const unsigned* __restrict__ keys,
                                                          float2* __restrict__ out, int ybeg, int yend) {
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    int y = ybeg + blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= g.w || y >= yend) return;
    int ti = div_th(g, y), tj = div_tw(g, x);
    int tile = ti * g.nx + tj;
    size_t idx = (size_t)y * g.w + x;
    float max1 = key2f(keys[2 * tile]), max2 = key2f(keys[2 * tile + 1]);
    float2 a = __ldg(&f1[idx]);
    float2 r;
    if (max1 == 0.0f) {
        r = __ldg(&f2[idx]);
    } else if (max2 == 0.0f) {
        r = a;
    } else {
        int oy = ti * g.Th - g.ov, ox = tj * g.Tw - g.ov;
        Taps t = make_taps(-a.x, -a.y);
        auto tap = [&](int yy, int xx) -> float2 {
            int gy = oy + yy, gx = ox + xx;
            bool ok = (unsigned)xx < (unsigned)g.Sw && (unsigned)yy < (unsigned)g.Sh &&
                      (unsigned)gx < (unsigned)g.w && (unsigned)gy < (unsigned)g.h;
            return ok ? __ldg(&f2[(size_t)gy * g.w + gx]) : make_float2(0.f, 0.f);
        };
        float2 v00 = tap(t.iy, t.ix), v01 = tap(t.iy, t.ix + 1), v10 = tap(t.iy + 1, t.ix), v11 = tap(t.iy + 1, t.ix + 1);
        float fx = __fmul_rn((float)t.ax, 0.03125f), fy = __fmul_rn((float)t.ay, 0.03125f);
        float ux = __fsub_rn(1.0f, fx), uy = __fsub_rn(1.0f, fy);
        float w00 = __fmul_rn(uy, ux), w01 = __fmul_rn(uy, fx), w10 = __fmul_rn(fy, ux), w11 = __fmul_rn(fy, fx);
        float sx = __fmul_rn(v00.x, w00), sy = __fmul_rn(v00.y, w00);
        sx = __fadd_rn(sx, __fmul_rn(v01.x, w01)); sy = __fadd_rn(sy, __fmul_rn(v01.y, w01));
        sx = __fadd_rn(sx, __fmul_rn(v10.x, w10)); sy = __fadd_rn(sy, __fmul_rn(v10.y, w10));
        sx = __fadd_rn(sx, __fmul_rn(v11.x, w11)); sy = __fadd_rn(sy, __fmul_rn(v11.y, w11));
        r.x = __fadd_rn(a.x, sx);
        r.y = __fadd_rn(a.y, sy);
    }
    out[idx] = r;
}

// ---- opt-in "corrected" composition (NOT the reference's behaviour): the flow that warps
// by f1 first and by f2 second is  m(p) = f2(p) + f1(p - f2(p)), sampled in IMAGE coordinates (no tile windows),
// same 1/32-px fixed-point bilinear scheme, zero outside the image.
__global__ void __launch_bounds__(256) compose_flows_kernel(const float2* __restrict__ f1, const float2* __restrict__ f2,
                                                            int h, int w, float2* __restrict__ out, int ybeg, int yend) {
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    int y = ybeg + blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= w || y >= yend) return;
    size_t idx = (size_t)y * w + x;
    float2 b = __ldg(&f2[idx]);
    int sx = cv_round_x32(__fsub_rn((float)x, b.x)), sy = cv_round_x32(__fsub_rn((float)y, b.y));
    int ix = sx >> 5, iy = sy >> 5, ax = sx & 31, ay = sy & 31;
    auto tap = [&](int yy, int xx) -> float2 {
        bool ok = (unsigned)xx < (unsigned)w && (unsigned)yy < (unsigned)h;
        return ok ? __ldg(&f1[(size_t)yy * w + xx]) : make_float2(0.f, 0.f);
    };
    float2 v00 = tap(iy, ix), v01 = tap(iy, ix + 1), v10 = tap(iy + 1, ix), v11 = tap(iy + 1, ix + 1);
    float fx = __fmul_rn((float)ax, 0.03125f), fy = __fmul_rn((float)ay, 0.03125f);
    float ux = __fsub_rn(1.0f, fx), uy = __fsub_rn(1.0f, fy);
    float w00 = __fmul_rn(uy, ux), w01 = __fmul_rn(uy, fx), w10 = __fmul_rn(fy, ux), w11 = __fmul_rn(fy, fx);
    float qx = __fmul_rn(v00.x, w00), qy = __fmul_rn(v00.y, w00);
    qx = __fadd_rn(qx, __fmul_rn(v01.x, w01)); qy = __fadd_rn(qy, __fmul_rn(v01.y, w01));
    qx = __fadd_rn(qx, __fmul_rn(v10.x, w10)); qy = __fadd_rn(qy, __fmul_rn(v10.y, w10));
    qx = __fadd_rn(qx, __fmul_rn(v11.x, w11)); qy = __fadd_rn(qy, __fmul_rn(v11.y, w11));
    out[idx] = make_float2(__fadd_rn(b.x, qx), __fadd_rn(b.y, qy));
}

// ------------------------------------------------------------------------------------------------
// K10: affine page transform -- transform_img_with_tmat (reference shared_modules/utils.py:98-114):
// pad_to_shape (utils.py:53-66, zero frame around the page) fused with skimage.transform.warp(order=1,
// mode='constant', cval=0, preserve_range=True) by a 3x3 output->input matrix.  Arithmetic as in
// skimage's _warp_fast / bilinear_interpolation for integer images: float64, every product and sum
// rounded separately (left to right), floor / ceil neighbours, zero outside the frame, result truncated
// to the page dtype.  kind: 0 metric (diagonal + translation), 1 affine, 2 projective (divide by z) --
// chosen on the host from the matrix' last row exactly like _warp_fast does.
// HBM-bound gather: 2 B read (four taps share sectors with the neighbours) + 2 B written per pixel.
// ------------------------------------------------------------------------------------------------
struct AffineMat {
    double m[9];
    int kind;
};

template <typename T>
__global__ void __launch_bounds__(256) warp_affine_kernel(const T* __restrict__ src, size_t pitch, int sh, int sw, int top, int left,
                                                          AffineMat A, T* __restrict__ out, size_t out_pitch, int oh, int ow) {
    const int x = blockIdx.x * 64 + threadIdx.x, y = blockIdx.y * 4 + threadIdx.y;
    if (x >= ow || y >= oh) return;
    const double fx = (double)x, fy = (double)y;
    double c, r;
    if (A.kind == 0) {
        c = __dadd_rn(__dmul_rn(A.m[0], fx), A.m[2]);
        r = __dadd_rn(__dmul_rn(A.m[4], fy), A.m[5]);
    } else {
        c = __dadd_rn(__dadd_rn(__dmul_rn(A.m[0], fx), __dmul_rn(A.m[1], fy)), A.m[2]);
        r = __dadd_rn(__dadd_rn(__dmul_rn(A.m[3], fx), __dmul_rn(A.m[4], fy)), A.m[5]);
        if (A.kind == 2) {
            const double z = __dadd_rn(__dadd_rn(__dmul_rn(A.m[6], fx), __dmul_rn(A.m[7], fy)), A.m[8]);
            c = __ddiv_rn(c, z);
            r = __ddiv_rn(r, z);
        }
    }
    T res = 0;
    // outside (-1, oh) x (-1, ow) -- or NaN -- all four neighbours lie outside the frame: the result is 0
    if (r > -1.0 && r < (double)oh && c > -1.0 && c < (double)ow) {
        const double minr = floor(r), minc = floor(c);
        const int r0 = (int)minr, c0 = (int)minc, r1 = (int)ceil(r), c1 = (int)ceil(c);
        const double dr = __dsub_rn(r, minr), dc = __dsub_rn(c, minc);
        auto px = [&](int rr, int cc) -> double {
            const int sy = rr - top, sx = cc - left;   // the page sits at (top, left) of the zero frame
            if ((unsigned)sy >= (unsigned)sh || (unsigned)sx >= (unsigned)sw) return 0.0;
            return (double)__ldg(reinterpret_cast<const T*>(reinterpret_cast<const char*>(src) + (size_t)sy * pitch) + sx);
        };
        const double tl = px(r0, c0), tr = px(r0, c1), bl = px(r1, c0), br = px(r1, c1);
        const double wc = __dsub_rn(1.0, dc), wr = __dsub_rn(1.0, dr);
        const double t = __dadd_rn(__dmul_rn(wc, tl), __dmul_rn(dc, tr));
        const double b = __dadd_rn(__dmul_rn(wc, bl), __dmul_rn(dc, br));
        const double v = __dadd_rn(__dmul_rn(wr, t), __dmul_rn(dr, b));
        res = (T)__double2uint_rz(v);   // .astype(dtype): truncation (v >= 0; the convex combination stays in range)
    }
    reinterpret_cast<T*>(reinterpret_cast<char*>(out) + (size_t)y * out_pitch)[x] = res;
}

}  // namespace ma

using namespace ma;

extern "C" int ma_compose_flows_rows(const float* f1, const float* f2, int h, int w, float* out,
                                     int row_begin, int row_end, void* stream) {
    if (!f1 || !f2 || !out || h <= 0 || w <= 0) return invalid("ma_compose_flows_rows: bad argument");
    if (h > 32767 * 2 || w > 32767 * 2) return invalid("ma_compose_flows_rows: image too large for 1/32-px fixed point");
    if (row_begin < 0 || row_end > h || row_begin > row_end) return invalid("ma_compose_flows_rows: bad row range");
    if (row_begin == row_end) return MA_OK;
    cudaStream_t s = (cudaStream_t)stream;
    dim3 block(64, 4), grid(ceil_div(w, 64), ceil_div(row_end - row_begin, 4));
    KernelScope ks(K_MERGE, s, (double)(row_end - row_begin) * w);
    compose_flows_kernel<<<grid, block, 0, s>>>((const float2*)f1, (const float2*)f2, h, w, (float2*)out, row_begin, row_end);
    MA_LAUNCH_CHECK("compose_flows_kernel");
    return MA_OK;
}

extern "C" int ma_warp_tiles_rows(const void* img, size_t img_pitch, int dtype, const float* flow, int h, int w,
                                  int T, int ov, void* out, size_t out_pitch, int row_begin, int row_end, void* stream) {
    if (!img || !flow || !out || h <= 0 || w <= 0 || T <= 0 || ov < 0) return invalid("ma_warp_tiles: bad argument");
    if (dtype != MA_U8 && dtype != MA_U16) return invalid("ma_warp_tiles: dtype must be MA_U8 or MA_U16");
    if (T + 2 * ov > 32767) return invalid("ma_warp_tiles: tile window exceeds cv.remap's int16 coordinate range");
    if (row_begin < 0 || row_end > h || row_begin > row_end) return invalid("ma_warp_tiles: bad row range");
    if (row_begin == row_end) return MA_OK;
    TileGeom g = make_geom(h, w, T, ov);
    dim3 block(64, 4), grid(ceil_div(w, 256), ceil_div(row_end - row_begin, 4));
    cudaStream_t s = (cudaStream_t)stream;
    KernelScope ks(K_WARP, s, (double)(row_end - row_begin) * w);
    // vector path: every row of the flow starts 16-byte aligned (even width) and every group of four output pixels is
    // naturally aligned
    const size_t esz = dtype == MA_U8 ? 1 : 2;
    const bool vec = (w % 2 == 0) && (reinterpret_cast<uintptr_t>(flow) % 16 == 0) && (out_pitch % (4 * esz) == 0) &&
                     (reinterpret_cast<uintptr_t>(out) % (4 * esz) == 0);
    if (dtype == MA_U8) {
        if (vec) warp_tiles_kernel<uint8_t, true><<<grid, block, 0, s>>>((const uint8_t*)img, img_pitch, (const float2*)flow, g, (uint8_t*)out, out_pitch, row_begin, row_end);
        else warp_tiles_kernel<uint8_t, false><<<grid, block, 0, s>>>((const uint8_t*)img, img_pitch, (const float2*)flow, g, (uint8_t*)out, out_pitch, row_begin, row_end);
    } else {
        if (vec) warp_tiles_kernel<uint16_t, true><<<grid, block, 0, s>>>((const uint16_t*)img, img_pitch, (const float2*)flow, g, (uint16_t*)out, out_pitch, row_begin, row_end);
        else warp_tiles_kernel<uint16_t, false><<<grid, block, 0, s>>>((const uint16_t*)img, img_pitch, (const float2*)flow, g, (uint16_t*)out, out_pitch, row_begin, row_end);
    }
    MA_LAUNCH_CHECK("warp_tiles_kernel");
    return MA_OK;
}

extern "C" int ma_warp_tiles(const void* img, size_t img_pitch, int dtype, const float* flow, int h, int w,
                             int T, int ov, void* out, size_t out_pitch, void* stream) {
    return ma_warp_tiles_rows(img, img_pitch, dtype, flow, h, w, T, ov, out, out_pitch, 0, h, stream);
}

extern "C" size_t ma_merge_workspace_bytes(int h, int w, int T) {
    if (h <= 0 || w <= 0 || T <= 0) return 0;
    TileGeom g = make_geom(h, w, T, 0);
    return (size_t)g.ny * g.nx * 2 * sizeof(unsigned);
}

extern "C" int ma_merge_flows_tile_rows(const float* f1, const float* f2, int h, int w, int T, int ov,
                                        float* out, void* workspace, int tile_row_begin, int tile_row_end, void* stream) {
    if (!f1 || !f2 || !out || !workspace || h <= 0 || w <= 0 || T <= 0 || ov < 0)
        return invalid("ma_merge_flows_tiles: bad argument");
    if (T + 2 * ov > 32767) return invalid("ma_merge_flows_tiles: tile window exceeds int16 coordinate range");
    TileGeom g = make_geom(h, w, T, ov);
    if (tile_row_begin < 0 || tile_row_end > g.ny || tile_row_begin > tile_row_end) return invalid("ma_merge_flows_tiles: bad tile-row range");
    if (tile_row_begin == tile_row_end) return MA_OK;
    cudaStream_t s = (cudaStream_t)stream;
    int tile0 = tile_row_begin * g.nx, ntiles = (tile_row_end - tile_row_begin) * g.nx;
    int ybeg = tile_row_begin * g.Th, yend = std::min(tile_row_end * g.Th, h);
    unsigned* keys = (unsigned*)workspace;
    { KernelScope ks(K_SMALL, s);
    init_keys_kernel<<<ceil_div(2 * ntiles, 256), 256, 0, s>>>(keys + 2 * tile0, 2 * ntiles); }
    int chunks = max(1, min(64, (int)(((long long)g.Sh * g.Sw) / (256 * 16))));
    if (ntiles > 65535) return invalid("ma_merge_flows_tiles: too many tiles");
    { KernelScope ks(K_MERGE_MAX, s, (double)(yend - ybeg) * w);
    tile_max_kernel<<<dim3(chunks, ntiles), 256, 0, s>>>((const float2*)f1, (const float2*)f2, g, keys, tile0); }
    dim3 block(64, 4), grid(ceil_div(w, 64), ceil_div(yend - ybeg, 4));
    KernelScope ks(K_MERGE, s, (double)(yend - ybeg) * w);
    merge_tiles_kernel<<<grid, block, 0, s>>>((const float2*)f1, (const float2*)f2, g, keys, (float2*)out, ybeg, yend);
    MA_LAUNCH_CHECK("merge_tiles_kernel");
    return MA_OK;
}

extern "C" int ma_merge_flows_tiles(const float* f1, const float* f2, int h, int w, int T, int ov,
                                    float* out, void* workspace, void* stream) {
    if (h <= 0 || T <= 0) return invalid("ma_merge_flows_tiles: bad argument");
    return ma_merge_flows_tile_rows(f1, f2, h, w, T, ov, out, workspace, 0, (h + T - 1) / T, stream);
}

extern "C" int ma_warp_affine(const void* img, size_t img_pitch, int dtype, int src_h, int src_w, int pad_top, int pad_left,
                              const double* inv3x3, void* out, size_t out_pitch, int out_h, int out_w, void* stream) {
    if (!img || !out || !inv3x3 || src_h <= 0 || src_w <= 0 || out_h <= 0 || out_w <= 0) return invalid("ma_warp_affine: bad argument");
    if (dtype != MA_U8 && dtype != MA_U16) return invalid("ma_warp_affine: dtype must be MA_U8 or MA_U16");
    if (pad_top < 0 || pad_left < 0 || pad_top + src_h > out_h || pad_left + src_w > out_w)
        return invalid("ma_warp_affine: the page does not fit into the target frame");
    AffineMat A;
    for (int i = 0; i < 9; ++i) A.m[i] = inv3x3[i];
    // _warp_fast's dispatch on the matrix (skimage/transform/_warps_cy.pyx)
    if (A.m[6] == 0 && A.m[7] == 0 && A.m[8] == 1) A.kind = (A.m[1] == 0 && A.m[3] == 0) ? 0 : 1;
    else A.kind = 2;
    cudaStream_t s = (cudaStream_t)stream;
    dim3 block(64, 4), grid(ceil_div(out_w, 64), ceil_div(out_h, 4));
    KernelScope ks(K_AFFINE, s, (double)out_h * out_w);
    if (dtype == MA_U8)
        warp_affine_kernel<uint8_t><<<grid, block, 0, s>>>((const uint8_t*)img, img_pitch, src_h, src_w, pad_top, pad_left, A,
                                                            (uint8_t*)out, out_pitch, out_h, out_w);
    else
        warp_affine_kernel<uint16_t><<<grid, block, 0, s>>>((const uint16_t*)img, img_pitch, src_h, src_w, pad_top, pad_left, A,
                                                             (uint16_t*)out, out_pitch, out_h, out_w);
    MA_LAUNCH_CHECK("warp_affine_kernel");
    return MA_OK;
}
