// K1-K3: tiled single-scale Farneback optical flow, bit-compatible with
//   cv.calcOpticalFlowFarneback(mov, ref, None, 0.5, levels=0, win, iters, poly_n=1, poly_sigma=1.7,
//                               OPTFLOW_FARNEBACK_GAUSSIAN)
// as called per tile by the reference (optflow_reg/flow_calc.py:30-47, 59-98).
//
// Stages (all per S x S tile window; a "tile" is only an index range, never a copy):
//   fb_polyexp_kernel   u8/u16 windows of BOTH images -> 3x3 prefilter (REFLECT_101 at the tile edge) ->
//                       polynomial expansion n=1 (f32 vertical, f64 horizontal) -> R0, R1 (5 planar f32
//                       planes each) and, fused, M = UpdateMatrices(R0, R1, flow = 0)
//   fb_blur_v_kernel    V^T = vertical (2m+1)-tap Gaussian of M, rows clamped at the tile edge, stored
//                       transposed
//   fb_blur_h_kernel    horizontal (2m+1)-tap Gaussian of V, 2x2 solve in f64 -> flow, then either
//                       M = UpdateMatrices(R0, R1, flow) in place (not last iteration) or scatter of
//                       the tile centre into the stitched flow (last iteration).
//
// Arithmetic contract: OpenCV's optflowgf.cpp is built for the SSE3 baseline, i.e. every a*b+c is a
// separately rounded multiply and add.  This file is compiled with -fmad=false and keeps OpenCV's
// association order (s = c*k0; s += (a[+i] + a[-i]) * k[i], i = 1..m), so results are bit-identical.
//
// The two blur kernels are FP32-pipe bound (2 passes x 5 planes x (1 + 3m) flops per pixel), not HBM
// bound.  Inputs arrive by TMA box loads (zero fill outside the plane, edge rows replicated in shared
// memory afterwards); each thread keeps 8 packed (f32x2) outputs and two 8-deep packed sliding windows
// in registers, so shared-memory traffic is 2 LDS.64 per 24 packed FP instructions.
#include <atomic>
#include <cmath>
#include <cstdlib>
#include <vector>
#include "common.cuh"
#include "packed.cuh"
#include "tma.cuh"

namespace ma {

constexpr int kMaxM = 96;         // TMA box rows 64 + 2*m <= 256
constexpr int kStep = 64;         // outputs per CTA along the convolution axis
constexpr int kR = 8;             // outputs per thread along the convolution axis

struct FbConsts {
    float g0, g1, xg1, xxg1;            // polyexp taps: centre g[0], side g[1], x*g, x*x*g at +1
    double ig11, ig03, ig33, ig55;      // inverse Gram constants
    int m;                              // blur half width
    float2 negzero2;                    // {-0.f, -0.f}, see mul2()
    float2 k2[kMaxM + 1];               // blur taps, duplicated {k, k} for the packed f32x2 multiply
};

struct FbBatch {
    TileGeom g;
    int tile0;        // first tile (row-major index) of this batch
    int ntiles;       // tiles in this batch
    int Sp;           // row pitch (floats) of the R0 / R1 / M planes ([y][x])
    int SpT;          // row pitch (floats) of the transposed V planes ([x][y])
    size_t plane;     // floats per plane = max(Sh * Sp, Sw * SpT)
    float* ws;        // workspace base: per slot 22 planes [R0 x5][R1 x5][M x5][V^T x5][flow x2]
};

constexpr int kSlotPlanes = 22;  // R0 x5, R1 x5, M x5, V^T x5, tile flow (float2 = 2 planes)

__device__ __forceinline__ float* slot_plane(const FbBatch& b, int slot, int which /*0 R0,1 R1,2 M,3 V^T,4 flow*/, int c) {
    return b.ws + ((size_t)slot * kSlotPlanes + which * 5 + c) * b.plane;
}

// Output window of one launch of the iteration kernels, in tile-local coordinates.  Only the centre of a tile's
// final flow is stitched (stitcher.py:99-115), so the last iteration needs its H-pass outputs on the centre only, its
// V-pass outputs on the centre rows x (centre +- m) columns, its M on (centre +- m)^2, hence the flow of the previous
// iteration on (centre +- m)^2, and so on: iteration `it` of N works on the centre grown by (N-1-it)*m (H pass,
// UpdateMatrices) resp. by that and m more columns (V pass).  Everything inside such a window is computed from
// inputs inside the previous window, so the stitched flow is bit-identical to the untrimmed computation; what lies
// outside is never read.  The grid covers the window of a full T x T centre from a 32-float aligned origin; tiles
// whose centre is clipped by the image edge drop the blocks beyond their own, smaller window.
struct FbWin {
    int bx0, by0;   // tile-local coordinates of block (0, 0)
    int ex, ey;     // outputs are needed on [ov - ex, ov + cw + ex) x [ov - ey, ov + ch + ey), cw x ch = clipped centre
};

__device__ __forceinline__ bool block_needed(const TileGeom& g, int tile, const FbWin& w, int x0, int xs, int y0, int ys) {
    const int ti = tile / g.nx, tj = tile - ti * g.nx;
    const int cw = min(g.Tw, g.w - tj * g.Tw), ch = min(g.Th, g.h - ti * g.Th);
    return x0 < g.ov + cw + w.ex && x0 + xs > g.ov - w.ex && y0 < g.ov + ch + w.ey && y0 + ys > g.ov - w.ey;
}

// ------------------------------------------------------------------------------------------------
// K1 + K2(first), fallback for windows narrower than 8 pixels (the marching kernel below is the default):
// prefilter + polynomial expansion of BOTH images of a tile, and M0 =
// UpdateMatrices(R0, R1, flow = 0), fused: with a zero flow the bilinear sample of R1 degenerates
// to R1 at the same pixel (weights 1,0,0,0), so M0 is a pointwise function of R0 and R1 and the
// 40 B/px re-read of a separate pass disappears.
//
// CTA = 64 x 16 output pixels, 256 threads x 4 pixels.  Per image: raw window (zero outside the
// image) -> shared memory once, then row prefilter, column prefilter, vertical expansion pass in
// shared memory (all f32), horizontal pass in f64 registers.  Border rules are evaluated in
// tile-local coordinates: REFLECT_101 for the 3x3 prefilter, replicate for the expansion.
// ------------------------------------------------------------------------------------------------
constexpr int PE_BW = 64, PE_BH = 16;
constexpr int PE_RW = PE_BW + 4, PE_RH = PE_BH + 4;   // raw window
constexpr int PE_PW = PE_BW + 2, PE_PH = PE_BH + 2;   // prefiltered window (1-px halo for the expansion)

__device__ __forceinline__ float border_w(int d) {
    // {0.14, 0.14, 0.4472, 0.4472, 0.4472}
    return d < 2 ? 0.14f : 0.4472f;
}

// second half of FarnebackUpdateMatrices: from the sampled / averaged r2..r6 to the five M entries
__device__ __forceinline__ void finish_matrices(float q0, float q1, float r2, float r3, float r4, float r5, float r6,
                                                float dx, float dy, int x, int y, int Sw, int Sh,
                                                float* __restrict__ M, size_t plane, size_t o) {
    r2 = __fmul_rn(__fsub_rn(q0, r2), 0.5f);
    r3 = __fmul_rn(__fsub_rn(q1, r3), 0.5f);
    r2 = __fadd_rn(r2, __fadd_rn(__fmul_rn(r4, dy), __fmul_rn(r6, dx)));
    r3 = __fadd_rn(r3, __fadd_rn(__fmul_rn(r6, dy), __fmul_rn(r5, dx)));
    if ((unsigned)(x - 5) >= (unsigned)(Sw - 10) || (unsigned)(y - 5) >= (unsigned)(Sh - 10)) {
        float scale = (x < 5 ? border_w(x) : 1.0f);
        scale = __fmul_rn(scale, (x >= Sw - 5 ? border_w(Sw - x - 1) : 1.0f));
        scale = __fmul_rn(scale, (y < 5 ? border_w(y) : 1.0f));
        scale = __fmul_rn(scale, (y >= Sh - 5 ? border_w(Sh - y - 1) : 1.0f));
        r2 = __fmul_rn(r2, scale); r3 = __fmul_rn(r3, scale); r4 = __fmul_rn(r4, scale);
        r5 = __fmul_rn(r5, scale); r6 = __fmul_rn(r6, scale);
    }
    M[o] = __fadd_rn(__fmul_rn(r4, r4), __fmul_rn(r6, r6));
    M[plane + o] = __fmul_rn(__fadd_rn(r4, r5), r6);
    M[2 * plane + o] = __fadd_rn(__fmul_rn(r5, r5), __fmul_rn(r6, r6));
    M[3 * plane + o] = __fadd_rn(__fmul_rn(r4, r2), __fmul_rn(r6, r3));
    M[4 * plane + o] = __fadd_rn(__fmul_rn(r6, r2), __fmul_rn(r5, r3));
}

template <typename T>
__global__ void __launch_bounds__(256) fb_polyexp_kernel(const T* __restrict__ mov, const T* __restrict__ ref,
                                                         size_t pitch, FbBatch b, const __grid_constant__ FbConsts cst) {
    __shared__ float raw[PE_RH][PE_RW];
    __shared__ float th[PE_RH][PE_PW];                 // row-prefiltered
    __shared__ float P[PE_PH][PE_PW];                  // prefiltered image at replicated positions
    __shared__ float T0[PE_BH][PE_PW], T1[PE_BH][PE_PW], T2[PE_BH][PE_PW];
    const TileGeom& g = b.g;
    const int Sh = g.Sh, Sw = g.Sw;
    const int slot = blockIdx.z;
    const int tile = b.tile0 + slot;
    const int ti = tile / g.nx, tj = tile % g.nx;
    const int oy = ti * g.Th - g.ov, ox = tj * g.Tw - g.ov;
    const int x0 = blockIdx.x * PE_BW, y0 = blockIdx.y * PE_BH;
    const int tid = threadIdx.x;
    const int tx = tid & 63, ty = tid >> 6;
    float R0v[4][5];
#pragma unroll
    for (int which = 0; which < 2; ++which) {          // 0: prev = moving -> R0, 1: next = reference -> R1
        const T* img = which ? ref : mov;
        __syncthreads();
        // all staging loops walk columns {tx, tx + 64} and rows ty, ty + 4, ...: no div/mod, and the per-column
        // border index math (clamp / REFLECT_101) is hoisted out of the row loops
        for (int c = tx; c < PE_RW; c += 64) {
            const int xx = x0 - 2 + c, gx = ox + xx;
            const bool colok = (unsigned)xx < (unsigned)Sw && (unsigned)gx < (unsigned)g.w;
            for (int r = ty; r < PE_RH; r += 4) {
                const int yy = y0 - 2 + r, gy = oy + yy;
                float v = 0.0f;
                if (colok && (unsigned)yy < (unsigned)Sh && (unsigned)gy < (unsigned)g.h)
                    v = (float)__ldg((const T*)((const char*)img + (size_t)gy * pitch) + gx);
                raw[r][c] = v;
            }
        }
        __syncthreads();
        for (int cu = tx; cu < PE_PW; cu += 64) {                // rows: actual, columns: replicated positions
            const int px = min(max(x0 - 1 + cu, 0), Sw - 1);
            const int ic = px - (x0 - 2), il = reflect101(px - 1, Sw) - (x0 - 2), ir = reflect101(px + 1, Sw) - (x0 - 2);
            for (int r = ty; r < PE_RH; r += 4)
                th[r][cu] = __fadd_rn(__fmul_rn(raw[r][ic], 0.5f), __fmul_rn(__fadd_rn(raw[r][il], raw[r][ir]), 0.25f));
        }
        __syncthreads();
        for (int rv = ty; rv < PE_PH; rv += 4) {
            const int py = min(max(y0 - 1 + rv, 0), Sh - 1);
            const int rc = py - (y0 - 2), ru = reflect101(py - 1, Sh) - (y0 - 2), rd = reflect101(py + 1, Sh) - (y0 - 2);
            for (int cu = tx; cu < PE_PW; cu += 64)
                P[rv][cu] = __fadd_rn(__fmul_rn(th[rc][cu], 0.5f), __fmul_rn(__fadd_rn(th[ru][cu], th[rd][cu]), 0.25f));
        }
        __syncthreads();
        for (int c = tx; c < PE_PW; c += 64) {                   // vertical pass of the expansion (f32)
            for (int r = ty; r < PE_BH; r += 4) {
                float s0 = P[r][c], sc = P[r + 1][c], s1 = P[r + 2][c];
                float pp = __fadd_rn(s0, s1);
                T0[r][c] = __fadd_rn(__fmul_rn(sc, cst.g0), __fmul_rn(cst.g1, pp));
                T1[r][c] = __fadd_rn(0.0f, __fmul_rn(cst.xg1, __fsub_rn(s1, s0)));
                T2[r][c] = __fadd_rn(0.0f, __fmul_rn(cst.xxg1, pp));
            }
        }
        __syncthreads();
        const int x = x0 + tx;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int r = ty + 4 * q, y = y0 + r, c = tx + 1;
            if (x >= Sw || y >= Sh) continue;
            // horizontal pass: float sums/differences, double accumulation (oracle/farneback_np.py:polyexp)
            double b1 = (double)__fmul_rn(T0[r][c], cst.g0);
            double b3 = (double)__fmul_rn(T1[r][c], cst.g0);
            double b5 = (double)__fmul_rn(T2[r][c], cst.g0);
            double tg = (double)__fadd_rn(T0[r][c + 1], T0[r][c - 1]);
            b1 = __dadd_rn(b1, __dmul_rn(tg, (double)cst.g1));
            double b4 = __dmul_rn(tg, (double)cst.xxg1);
            double b2 = (double)__fmul_rn(__fsub_rn(T0[r][c + 1], T0[r][c - 1]), cst.xg1);
            b3 = __dadd_rn(b3, (double)__fmul_rn(__fadd_rn(T1[r][c + 1], T1[r][c - 1]), cst.g1));
            double b6 = (double)__fmul_rn(__fsub_rn(T1[r][c + 1], T1[r][c - 1]), cst.xg1);
            b5 = __dadd_rn(b5, (double)__fmul_rn(__fadd_rn(T2[r][c + 1], T2[r][c - 1]), cst.g1));
            float v[5];
            v[0] = (float)__dmul_rn(b3, cst.ig11);
            v[1] = (float)__dmul_rn(b2, cst.ig11);
            v[2] = (float)__dadd_rn(__dmul_rn(b1, cst.ig03), __dmul_rn(b5, cst.ig33));
            v[3] = (float)__dadd_rn(__dmul_rn(b1, cst.ig03), __dmul_rn(b4, cst.ig33));
            v[4] = (float)__dmul_rn(b6, cst.ig55);
            const size_t o = (size_t)y * b.Sp + x;
            float* Rp = slot_plane(b, slot, which, 0);
#pragma unroll
            for (int k = 0; k < 5; ++k) Rp[k * b.plane + o] = v[k];
            if (which == 0) {
#pragma unroll
                for (int k = 0; k < 5; ++k) R0v[q][k] = v[k];
            } else {
                // UpdateMatrices with flow == 0: sample of R1 = R1 itself where (x, y) has a right/lower neighbour
                float r2, r3, r4, r5, r6;
                if (x < Sw - 1 && y < Sh - 1) {
                    r2 = v[0];
                    r3 = v[1];
                    r4 = __fmul_rn(__fadd_rn(R0v[q][2], v[2]), 0.5f);
                    r5 = __fmul_rn(__fadd_rn(R0v[q][3], v[3]), 0.5f);
                    r6 = __fmul_rn(__fadd_rn(R0v[q][4], v[4]), 0.25f);
                } else {
                    r2 = r3 = 0.0f;
                    r4 = R0v[q][2];
                    r5 = R0v[q][3];
                    r6 = __fmul_rn(R0v[q][4], 0.5f);
                }
                finish_matrices(R0v[q][0], R0v[q][1], r2, r3, r4, r5, r6, 0.0f, 0.0f, x, y, Sw, Sh,
                                slot_plane(b, slot, 2, 0), b.plane, o);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// K1 + K2(first), default: "marching" polynomial expansion without shared memory and without block barriers
// (measured on B200: 0.87 ms vs 1.34 ms per 36 tiles for the staged kernel below, profiles/r02_ab_variants.log).  A warp owns a strip of 28 output columns (lanes 2..29; lanes 0, 1, 30, 31
// carry the halo) and marches down a band of rows: per virtual row v = ya-2 .. yb+1 (actual row
// REFLECT_101(v)) each lane loads one pixel of both images, the row prefilter takes its neighbours by
// shuffle (sources follow REFLECT_101 at the tile edge), and the column prefilter, the vertical expansion
// pass and the row-replication rules work on three-deep rolling registers; the horizontal pass gets T0..T2
// of the neighbouring (replicated) columns by shuffle.  Arithmetic and its order are those of
// fb_polyexp_kernel; tests/emu_polyexp_march.py checks the index logic against the oracle on the CPU.
// 28/32 lanes produce output and a band re-reads 4 of its PM_BAND rows, but the ~620 instructions per pixel
// of the staged kernel (index math, five block-wide phases per image) shrink to ~250.
// ------------------------------------------------------------------------------------------------
constexpr int PM_OUTW = 28;    // output columns per warp
constexpr int PM_BAND = 96;    // output rows per warp task
constexpr int PM_WARPS = 8;

template <typename T>
__global__ void __launch_bounds__(PM_WARPS * 32) fb_polyexp_march_kernel(const T* __restrict__ mov, const T* __restrict__ ref,
                                                                          size_t pitch, FbBatch b, const __grid_constant__ FbConsts cst) {
    const TileGeom& g = b.g;
    const int Sh = g.Sh, Sw = g.Sw;
    const int slot = blockIdx.z;
    const int tile = b.tile0 + slot;
    const int ti = tile / g.nx, tj = tile % g.nx;
    const int oy = ti * g.Th - g.ov, ox = tj * g.Tw - g.ov;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int xs = (blockIdx.x * PM_WARPS + warp) * PM_OUTW;
    if (xs >= Sw) return;                       // warp-uniform; the kernel has no block-wide barrier
    const int ya = blockIdx.y * PM_BAND, yb = min(ya + PM_BAND, Sh);
    const int x = xs - 2 + lane;                // this lane's tile column
    const bool colin = (unsigned)x < (unsigned)Sw;
    const bool colok = colin && (unsigned)(ox + x) < (unsigned)g.w;
    // shuffle sources: prefilter neighbours (REFLECT_101 at the tile edge) and expansion neighbours (replicate)
    int srcL = lane, srcR = lane;
    if (colin) {
        srcL = min(max(reflect101(x - 1, Sw) - (xs - 2), 0), 31);
        srcR = min(max(reflect101(x + 1, Sw) - (xs - 2), 0), 31);
    }
    const int tL = min(max(min(max(x - 1, 0), Sw - 1) - (xs - 2), 0), 31);
    const int tR = min(max(min(max(x + 1, 0), Sw - 1) - (xs - 2), 0), 31);
    const bool writer = lane >= 2 && lane < 2 + PM_OUTW && colin;
    const T* pm = mov + (ox + x);
    const T* pr = ref + (ox + x);
    auto load = [&](int v, float& a, float& c) {
        const int gy = oy + reflect101(v, Sh);
        a = 0.0f;
        c = 0.0f;
        if (colok && (unsigned)gy < (unsigned)g.h) {
            a = (float)__ldg(reinterpret_cast<const T*>(reinterpret_cast<const char*>(pm) + (size_t)gy * pitch));
            c = (float)__ldg(reinterpret_cast<const T*>(reinterpret_cast<const char*>(pr) + (size_t)gy * pitch));
        }
    };
    float th0[2] = {0, 0}, th1[2] = {0, 0}, th2[2] = {0, 0}, P0[2] = {0, 0}, P1[2] = {0, 0}, P2[2] = {0, 0};
    // (deeper load queues -- four rows in flight, as floats or as raw integers -- were measured on B200 and are not
    // faster: 12.2 vs 12.3 ms and 14.4 ms per step at 20 000^2; the kernel is bound by its 60 B/px of stores)
    float nxt[2];
    load(ya - 2, nxt[0], nxt[1]);
    for (int v = ya - 2; v < yb + 2; ++v) {
        float raw[2] = {nxt[0], nxt[1]};
        if (v + 1 < yb + 2) load(v + 1, nxt[0], nxt[1]);    // in flight during this row's arithmetic
#pragma unroll
        for (int im = 0; im < 2; ++im) {
            const float l = __shfl_sync(0xffffffffu, raw[im], srcL), r = __shfl_sync(0xffffffffu, raw[im], srcR);
            th0[im] = th1[im];
            th1[im] = th2[im];
            th2[im] = __fadd_rn(__fmul_rn(raw[im], 0.5f), __fmul_rn(__fadd_rn(l, r), 0.25f));
        }
        if (v < ya) continue;
#pragma unroll
        for (int im = 0; im < 2; ++im) {                     // prefiltered image at virtual row v - 1
            P0[im] = P1[im];
            P1[im] = P2[im];
            P2[im] = __fadd_rn(__fmul_rn(th1[im], 0.5f), __fmul_rn(__fadd_rn(th0[im], th2[im]), 0.25f));
        }
        const int y = v - 2;
        if (y < ya || y >= yb) continue;
        float R[2][5];
#pragma unroll
        for (int im = 0; im < 2; ++im) {
            const float s0 = y == 0 ? P1[im] : P0[im], sc = P1[im], s1 = y == Sh - 1 ? P1[im] : P2[im];
            const float pp = __fadd_rn(s0, s1);
            const float t0 = __fadd_rn(__fmul_rn(sc, cst.g0), __fmul_rn(cst.g1, pp));
            const float t1 = __fadd_rn(0.0f, __fmul_rn(cst.xg1, __fsub_rn(s1, s0)));
            const float t2 = __fadd_rn(0.0f, __fmul_rn(cst.xxg1, pp));
            const float t0l = __shfl_sync(0xffffffffu, t0, tL), t0r = __shfl_sync(0xffffffffu, t0, tR);
            const float t1l = __shfl_sync(0xffffffffu, t1, tL), t1r = __shfl_sync(0xffffffffu, t1, tR);
            const float t2l = __shfl_sync(0xffffffffu, t2, tL), t2r = __shfl_sync(0xffffffffu, t2, tR);
            // horizontal pass: float sums/differences, double accumulation (oracle/farneback_np.py:polyexp)
            double b1 = (double)__fmul_rn(t0, cst.g0);
            double b3 = (double)__fmul_rn(t1, cst.g0);
            double b5 = (double)__fmul_rn(t2, cst.g0);
            const double tg = (double)__fadd_rn(t0r, t0l);
            b1 = __dadd_rn(b1, __dmul_rn(tg, (double)cst.g1));
            const double b4 = __dmul_rn(tg, (double)cst.xxg1);
            const double b2 = (double)__fmul_rn(__fsub_rn(t0r, t0l), cst.xg1);
            b3 = __dadd_rn(b3, (double)__fmul_rn(__fadd_rn(t1r, t1l), cst.g1));
            const double b6 = (double)__fmul_rn(__fsub_rn(t1r, t1l), cst.xg1);
            b5 = __dadd_rn(b5, (double)__fmul_rn(__fadd_rn(t2r, t2l), cst.g1));
            R[im][0] = (float)__dmul_rn(b3, cst.ig11);
            R[im][1] = (float)__dmul_rn(b2, cst.ig11);
            R[im][2] = (float)__dadd_rn(__dmul_rn(b1, cst.ig03), __dmul_rn(b5, cst.ig33));
            R[im][3] = (float)__dadd_rn(__dmul_rn(b1, cst.ig03), __dmul_rn(b4, cst.ig33));
            R[im][4] = (float)__dmul_rn(b6, cst.ig55);
        }
        if (!writer) continue;
        const size_t o = (size_t)y * b.Sp + x;
        float* __restrict__ R0p = slot_plane(b, slot, 0, 0);
        float* __restrict__ R1p = slot_plane(b, slot, 1, 0);
#pragma unroll
        for (int k = 0; k < 5; ++k) {
            R0p[k * b.plane + o] = R[0][k];
            R1p[k * b.plane + o] = R[1][k];
        }
        // UpdateMatrices with flow == 0: the sample of R1 is R1 itself where (x, y) has a right / lower neighbour
        float r2, r3, r4, r5, r6;
        if (x < Sw - 1 && y < Sh - 1) {
            r2 = R[1][0];
            r3 = R[1][1];
            r4 = __fmul_rn(__fadd_rn(R[0][2], R[1][2]), 0.5f);
            r5 = __fmul_rn(__fadd_rn(R[0][3], R[1][3]), 0.5f);
            r6 = __fmul_rn(__fadd_rn(R[0][4], R[1][4]), 0.25f);
        } else {
            r2 = r3 = 0.0f;
            r4 = R[0][2];
            r5 = R[0][3];
            r6 = __fmul_rn(R[0][4], 0.5f);
        }
        finish_matrices(R[0][0], R[0][1], r2, r3, r4, r5, r6, 0.0f, 0.0f, x, y, Sw, Sh, slot_plane(b, slot, 2, 0), b.plane, o);
    }
}


// ------------------------------------------------------------------------------------------------
// K2: UpdateMatrices for one pixel (FarnebackUpdateMatrices, all f32, left-to-right sums)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void update_matrices_px(const float* __restrict__ R0, const float* __restrict__ R1,
                                                   size_t plane, int Sp, int Sw, int Sh, int x, int y,
                                                   float dx, float dy, float* __restrict__ M) {
    size_t o = (size_t)y * Sp + x;
    float fx = __fadd_rn((float)x, dx), fy = __fadd_rn((float)y, dy);
    int x1 = __float2int_rd(fx), y1 = __float2int_rd(fy);
    fx = __fsub_rn(fx, (float)x1);
    fy = __fsub_rn(fy, (float)y1);
    float r2, r3, r4, r5, r6;
    float q2 = __ldg(R0 + 2 * plane + o), q3 = __ldg(R0 + 3 * plane + o), q4 = __ldg(R0 + 4 * plane + o);
    if ((unsigned)x1 < (unsigned)(Sw - 1) && (unsigned)y1 < (unsigned)(Sh - 1)) {
        float ux = __fsub_rn(1.0f, fx), uy = __fsub_rn(1.0f, fy);
        float a00 = __fmul_rn(ux, uy), a01 = __fmul_rn(fx, uy), a10 = __fmul_rn(ux, fy), a11 = __fmul_rn(fx, fy);
        const float* p = R1 + (size_t)y1 * Sp + x1;
        auto samp = [&](int c) {
            const float* q = p + c * plane;
            float s = __fmul_rn(a00, __ldg(q));
            s = __fadd_rn(s, __fmul_rn(a01, __ldg(q + 1)));
            s = __fadd_rn(s, __fmul_rn(a10, __ldg(q + Sp)));
            s = __fadd_rn(s, __fmul_rn(a11, __ldg(q + Sp + 1)));
            return s;
        };
        r2 = samp(0);
        r3 = samp(1);
        r4 = __fmul_rn(__fadd_rn(q2, samp(2)), 0.5f);
        r5 = __fmul_rn(__fadd_rn(q3, samp(3)), 0.5f);
        r6 = __fmul_rn(__fadd_rn(q4, samp(4)), 0.25f);
    } else {
        r2 = r3 = 0.0f;
        r4 = q2;
        r5 = q3;
        r6 = __fmul_rn(q4, 0.5f);
    }
    finish_matrices(__ldg(R0 + o), __ldg(R0 + plane + o), r2, r3, r4, r5, r6, dx, dy, x, y, Sw, Sh, M, plane, o);
}

// ------------------------------------------------------------------------------------------------
// K3 core: symmetric (2m+1)-tap convolution, OpenCV order  s = c*k0; s += (a[+i] + a[-i]) * k[i].
//
// The tile of inputs sits in shared memory as rows of 64 floats (256 B, written by one TMA box load);
// the convolution runs along the row index.  A thread owns the two adjacent columns (2*lane, 2*lane+1)
// as one packed f32x2 value and produces 8 consecutive outputs along the convolution axis, keeping
// two 8-deep sliding windows in registers, so one step costs 2 LDS.64 for 8 x (FADD2, FMUL2, FADD2).
// Packed FADD2 issues at twice the scalar lane rate on sm_100 (measured, scripts/microbench), which
// makes the separately-rounded add/mul/add sequence as cheap as a fused scalar FMA formulation.
// With the loop unrolled by 8 every register index is static.
// ------------------------------------------------------------------------------------------------
constexpr int kRowF = 64;            // floats per shared-memory row (one TMA box row)
constexpr int kRowU = kRowF / 2;     // packed pairs per row

// `centre` -> this thread's pair in the row holding the centre tap of output 0; outputs 0..7 are the
// next rows.  k2[i] = {k[i], k[i]}.
// FUSED = false: OpenCV's separately rounded multiply and add (bit parity, the default).
// FUSED = true : acc = fma(a[+i] + a[-i], k[i], acc) -- opt-in, 1.5x fewer FP32 pipe cycles, ~1e-6 px off.
// one tap of conv8x2 for the 8 outputs of a thread.  The three dependent operations of an output (pair sum, product,
// accumulate) are issued group-wise over kIlp outputs at a time, so that consecutive instructions are independent:
// written output by output, ptxas reuses one temporary register for all chains and serialises them with 4-cycle
// stalls (2.8 issue cycles per packed instruction instead of 2).
#ifndef MA_CONV_ILP
#define MA_CONV_ILP 4
#endif
template <bool FUSED>
__device__ __forceinline__ void conv_step(const u64 (&wp)[kR], const u64 (&wm)[kR], int s, u64 kk, u64 nz, u64 (&acc)[kR]) {
    constexpr int kIlp = MA_CONV_ILP;
#pragma unroll
    for (int j0 = 0; j0 < kR; j0 += kIlp) {
        u64 pr[kIlp];
#pragma unroll
        for (int j = 0; j < kIlp; ++j) pr[j] = add2(wp[(j0 + j + s + 1) & 7], wm[(j0 + j + 63 - s) & 7]);
        if (FUSED) {
#pragma unroll
            for (int j = 0; j < kIlp; ++j) acc[j0 + j] = fma2(pr[j], kk, acc[j0 + j]);
        } else {
#pragma unroll
            for (int j = 0; j < kIlp; ++j) pr[j] = mul2(pr[j], kk, nz);
#pragma unroll
            for (int j = 0; j < kIlp; ++j) acc[j0 + j] = add2(acc[j0 + j], pr[j]);
        }
    }
}

template <bool FUSED>
__device__ __forceinline__ void conv8x2(const u64* __restrict__ centre, int m, const float2* __restrict__ k2, u64 nz, u64 (&acc)[kR]) {
    u64 wp[kR], wm[kR];
    const u64 k0 = *reinterpret_cast<const u64*>(&k2[0]);
#pragma unroll
    for (int j = 0; j < kR; ++j) {
        u64 c = centre[j * kRowU];
        wp[j] = c;
        wm[j] = c;
        acc[j] = mul2(c, k0, nz);
    }
    const u64* pp = centre + kR * kRowU;   // row of in[+7 + i] for i = 1
    const u64* pm = centre - kRowU;        // row of in[-i] for i = 1
    int i = 1;
    // before step i: in[j+i-1] lives in wp[(j+i-1)&7], in[j-i+1] in wm[(j-i+1)&7]; i = 8g+1+s
#pragma unroll 1
    for (; i + kR - 1 <= m; i += kR) {
#pragma unroll
        for (int s = 0; s < kR; ++s) {
            wp[s] = pp[s * kRowU];
            wm[(63 - s) & 7] = pm[-s * kRowU];
            const u64 kk = *reinterpret_cast<const u64*>(&k2[i + s]);
            conv_step<FUSED>(wp, wm, s, kk, nz, acc);
        }
        pp += kR * kRowU;
        pm -= kR * kRowU;
    }
#pragma unroll
    for (int s = 0; s < kR - 1; ++s) {
        if (i + s <= m) {  // warp-uniform tail (m mod 8 steps)
            wp[s] = pp[s * kRowU];
            wm[(63 - s) & 7] = pm[-s * kRowU];
            const u64 kk = *reinterpret_cast<const u64*>(&k2[i + s]);
            conv_step<FUSED>(wp, wm, s, kk, nz, acc);
        }
    }
}

// replicate the first / last valid row of the box into the rows TMA zero-filled outside [0, n)
// (OpenCV clamps rows / replicates columns at the tile edge).  v_first = coordinate of box row 0.
__device__ __forceinline__ void replicate_edges(float* buf, int rows, int v_first, int n) {
    if (v_first >= 0 && v_first + rows <= n) return;   // CTA-uniform
    const int r_lo = -v_first, r_hi = n - 1 - v_first; // box rows of coordinate 0 and n-1
    for (int p = threadIdx.x; p < rows * kRowF; p += blockDim.x) {
        int r = p / kRowF, c = p % kRowF;
        if (r < r_lo) buf[p] = buf[r_lo * kRowF + c];
        else if (r > r_hi) buf[p] = buf[r_hi * kRowF + c];
    }
    fence_proxy_async();
    __syncthreads();
}

// ------------------------------------------------------------------------------------------------
// K3a: vertical pass M -> V^T.  CTA = 64 columns x 64 rows of one plane of one tile.  One TMA box (64 x (64 + 2m)
// floats) lands the inputs; lanes <-> column pairs, warps <-> strips of 8 output rows; every thread writes its 2 x 8
// outputs straight from registers as two 32-byte vectors of V^T (row = x, column = y), so that the horizontal pass
// is the same kernel shape with TMA-friendly rows.  No transpose stage, no barrier after the convolution; 62
// registers -> four CTAs per SM, whose load / convolve / store phases overlap (measured against the 128-row,
// shared-memory-transpose version: 3.62 vs 3.79 ms per 36 tiles x 3 iterations, FMA pipe 80 % vs 76 %).
// grid = (x blocks, y blocks, ntiles*5) over the launch window (FbWin), dynamic smem = (64 + 2m) * 256 B
// ------------------------------------------------------------------------------------------------
constexpr int kVOut = 64;

// V pass: V^T[x][y .. y + 7] for the thread's two columns; the row pitch is a multiple of 32 floats, so a full
// vector never leaves the row even when it runs past Sh (padding, never read back)
__device__ __forceinline__ void store_vt_strip(const FbBatch& b, int slot, int c, int x, int y, int Sw, const u64 (&acc)[kR]) {
    float* __restrict__ dst = slot_plane(b, slot, 3, c) + (size_t)x * b.SpT + y;
    float2 v[kR];
#pragma unroll
    for (int j = 0; j < kR; ++j) v[j] = unpack2(acc[j]);
    if (x < Sw) {
        reinterpret_cast<float4*>(dst)[0] = make_float4(v[0].x, v[1].x, v[2].x, v[3].x);
        reinterpret_cast<float4*>(dst)[1] = make_float4(v[4].x, v[5].x, v[6].x, v[7].x);
    }
    if (x + 1 < Sw) {
        reinterpret_cast<float4*>(dst + b.SpT)[0] = make_float4(v[0].y, v[1].y, v[2].y, v[3].y);
        reinterpret_cast<float4*>(dst + b.SpT)[1] = make_float4(v[4].y, v[5].y, v[6].y, v[7].y);
    }
}

template <bool FUSED>
__global__ void __launch_bounds__(256) fb_blur_v_kernel(const __grid_constant__ CUtensorMap mapM, FbBatch b,
                                                        const __grid_constant__ FbConsts cst, FbWin win) {
    extern __shared__ __align__(128) float smem[];
    __shared__ __align__(8) uint64_t bar;
    const int Sh = b.g.Sh, Sw = b.g.Sw, m = cst.m;
    const int slot = blockIdx.z / 5, c = blockIdx.z % 5;
    const int x0 = win.bx0 + blockIdx.x * kRowF, y0 = win.by0 + blockIdx.y * kVOut;
    if (!block_needed(b.g, b.tile0 + slot, win, x0, kRowF, y0, kVOut)) return;   // CTA-uniform
    const int rows = kVOut + 2 * m;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        fence_mbar_init();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_expect_tx(&bar, rows * kRowF * sizeof(float));
        tma_load_3d(smem, &mapM, x0, y0 - m, slot * kSlotPlanes + 10 + c, &bar);
    }
    mbar_wait(&bar, 0);
    replicate_edges(smem, rows, y0 - m, Sh);
    const u64 nz = *reinterpret_cast<const u64*>(&cst.negzero2);
    const int o0 = warp * kR;
    if (y0 + o0 < Sh) {
        u64 acc[kR];
        conv8x2<FUSED>(reinterpret_cast<const u64*>(smem) + (o0 + m) * kRowU + lane, m, cst.k2, nz, acc);
        store_vt_strip(b, slot, c, x0 + 2 * lane, y0 + o0, Sw, acc);
    }
}

// ------------------------------------------------------------------------------------------------
// K3b: horizontal pass + 2x2 solve + (UpdateMatrices | final scatter).  CTA = 64 rows (y) x 64 columns
// (x) of one tile, all 5 planes.  The planes of V^T stream through two shared-memory buffers (TMA box
// = 64 y-values x (64 + 2m) x-rows; plane c+2 is in flight while plane c is convolved); lanes <-> row
// pairs; the five blurred values of a pixel stay in registers until the f64 solve.  The flow block is
// then transposed through shared memory so that R0 / R1 / M / flow are accessed with lanes <-> x.
// grid = (ceil(Sh/64), ceil(Sw/64), ntiles), dynamic smem = 2 * (64 + 2m) * 256 B
// ------------------------------------------------------------------------------------------------
constexpr int kFlowPitch = 66;  // float2 per x-row of the flow stage (64 + 2: 16-byte aligned rows, few bank conflicts)

template <bool FUSED>
__global__ void __launch_bounds__(256, 2) fb_blur_h_kernel(const __grid_constant__ CUtensorMap mapVT, FbBatch b,
                                                            const __grid_constant__ FbConsts cst, int last_iter,
                                                            float2* __restrict__ flow_out, FbWin win) {
    extern __shared__ __align__(128) float smem[];
    __shared__ __align__(8) uint64_t bars[2];
    const TileGeom& g = b.g;
    const int Sh = g.Sh, Sw = g.Sw, m = cst.m;
    const int slot = blockIdx.z;
    const int y0 = win.by0 + blockIdx.x * kRowF, x0 = win.bx0 + blockIdx.y * kStep;
    if (!block_needed(g, b.tile0 + slot, win, x0, kStep, y0, kRowF)) return;   // CTA-uniform
    const int rows = kStep + 2 * m;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* buf[2] = {smem, smem + rows * kRowF};
    const uint32_t box_bytes = rows * kRowF * sizeof(float);
    if (threadIdx.x == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        fence_mbar_init();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int c = 0; c < 2; ++c) {
            mbar_expect_tx(&bars[c], box_bytes);
            tma_load_3d(buf[c], &mapVT, y0, x0 - m, slot * kSlotPlanes + 15 + c, &bars[c]);
        }
    }
    const u64 nz = *reinterpret_cast<const u64*>(&cst.negzero2);
    u64 acc[5][kR];
    const bool active = x0 + warp * kR < Sw;
#pragma unroll
    for (int c = 0; c < 5; ++c) {
        mbar_wait(&bars[c & 1], (c >> 1) & 1);
        replicate_edges(buf[c & 1], rows, x0 - m, Sw);
        if (active) conv8x2<FUSED>(reinterpret_cast<const u64*>(buf[c & 1]) + (warp * kR + m) * kRowU + lane, m, cst.k2, nz, acc[c]);
        __syncthreads();  // buffer c&1 is free again
        if (c + 2 < 5 && threadIdx.x == 0) {
            mbar_expect_tx(&bars[c & 1], box_bytes);
            tma_load_3d(buf[c & 1], &mapVT, y0, x0 - m, slot * kSlotPlanes + 15 + c + 2, &bars[c & 1]);
        }
    }
    // 2x2 solve in f64 (FarnebackUpdateFlow_GaussianBlur), flow staged as [x][y]
    float2* fl = reinterpret_cast<float2*>(smem);  // both input buffers are free now; host sizes smem >= the stage
    if (active) {
#pragma unroll
        for (int j = 0; j < kR; ++j) {
            float2 G11 = unpack2(acc[0][j]), G12 = unpack2(acc[1][j]), G22 = unpack2(acc[2][j]);
            float2 H1 = unpack2(acc[3][j]), H2 = unpack2(acc[4][j]);
            float4 o;
            {
                double g11 = G11.x, g12 = G12.x, g22 = G22.x, h1 = H1.x, h2 = H2.x;
                double idet = __drcp_rn(__dadd_rn(__dsub_rn(__dmul_rn(g11, g22), __dmul_rn(g12, g12)), 1e-3));
                o.x = (float)__dmul_rn(__dsub_rn(__dmul_rn(g11, h2), __dmul_rn(g12, h1)), idet);
                o.y = (float)__dmul_rn(__dsub_rn(__dmul_rn(g22, h1), __dmul_rn(g12, h2)), idet);
            }
            {
                double g11 = G11.y, g12 = G12.y, g22 = G22.y, h1 = H1.y, h2 = H2.y;
                double idet = __drcp_rn(__dadd_rn(__dsub_rn(__dmul_rn(g11, g22), __dmul_rn(g12, g12)), 1e-3));
                o.z = (float)__dmul_rn(__dsub_rn(__dmul_rn(g11, h2), __dmul_rn(g12, h1)), idet);
                o.w = (float)__dmul_rn(__dsub_rn(__dmul_rn(g22, h1), __dmul_rn(g12, h2)), idet);
            }
            *reinterpret_cast<float4*>(&fl[(warp * kR + j) * kFlowPitch + 2 * lane]) = o;
        }
    }
    __syncthreads();
    // epilogue: lanes <-> x.  Last iteration: scatter the tile centre into the stitched flow.  Otherwise the
    // flow goes to the tile's flow plane and fb_update_kernel recomputes M from it: UpdateMatrices is a
    // latency-bound gather (flow -> address -> 20 R1 values) that a separate, fully occupied pointwise
    // kernel hides far better than the 16-warp convolution CTA can (ncu: 56 % of this kernel's stall
    // samples sat on those loads when it was fused here); the price is 16 B/px of extra traffic.
    const int tile = b.tile0 + slot;
    const int ti = tile / g.nx, tj = tile % g.nx;
    const int cx = threadIdx.x & 63, x = x0 + cx;
    if (x >= Sw) return;
    float2* __restrict__ F = reinterpret_cast<float2*>(slot_plane(b, slot, 4, 0));
    for (int r = threadIdx.x >> 6; r < kRowF; r += 4) {
        int y = y0 + r;
        if (y >= Sh) break;
        float2 f = fl[cx * kFlowPitch + r];
        if (last_iter) {
            int cy = y - g.ov, cxx = x - g.ov;
            if ((unsigned)cy < (unsigned)g.Th && (unsigned)cxx < (unsigned)g.Tw) {
                int gy = ti * g.Th + cy, gx = tj * g.Tw + cxx;
                if (gy < g.h && gx < g.w) flow_out[(size_t)gy * g.w + gx] = f;
            }
        } else {
            F[(size_t)y * b.Sp + x] = f;
        }
    }
}

// K2 (iterations > 0): M = UpdateMatrices(R0, R1, flow), one thread per tile pixel, everything coalesced
// except the bilinear gather of R1 around (x + dx, y + dy).
__global__ void __launch_bounds__(256) fb_update_kernel(FbBatch b, FbWin win) {
    const int bx = win.bx0 + blockIdx.x * 64, by = win.by0 + blockIdx.y * 4;
    const int slot = blockIdx.z;
    if (!block_needed(b.g, b.tile0 + slot, win, bx, 64, by, 4)) return;   // CTA-uniform
    const int x = bx + (threadIdx.x & 63), y = by + (threadIdx.x >> 6);
    if (x >= b.g.Sw || y >= b.g.Sh) return;
    const float2 f = __ldg(reinterpret_cast<const float2*>(slot_plane(b, slot, 4, 0)) + (size_t)y * b.Sp + x);
    update_matrices_px(slot_plane(b, slot, 0, 0), slot_plane(b, slot, 1, 0), b.plane, b.Sp, b.g.Sw, b.g.Sh,
                       x, y, f.x, f.y, slot_plane(b, slot, 2, 0));
}


// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
static void chol_inv6(double A[6][6], double inv[6][6]) {
    // cv::Mat::inv(DECOMP_CHOLESKY): in-place Cholesky storing reciprocal diagonals, then L y = I, Lt x = y
    double L[6][6];
    for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) { L[i][j] = A[i][j]; inv[i][j] = i == j; }
    for (int i = 0; i < 6; ++i) {
        for (int j = 0; j < i; ++j) {
            double s = L[i][j];
            for (int k = 0; k < j; ++k) s -= L[i][k] * L[j][k];
            L[i][j] = s * L[j][j];
        }
        double s = L[i][i];
        for (int k = 0; k < i; ++k) { double t = L[i][k]; s -= t * t; }
        L[i][i] = 1.0 / std::sqrt(s);
    }
    for (int i = 0; i < 6; ++i)
        for (int j = 0; j < 6; ++j) {
            double s = inv[i][j];
            for (int k = 0; k < i; ++k) s -= L[i][k] * inv[k][j];
            inv[i][j] = s * L[i][i];
        }
    for (int i = 5; i >= 0; --i)
        for (int j = 0; j < 6; ++j) {
            double s = inv[i][j];
            for (int k = 5; k > i; --k) s -= L[k][i] * inv[k][j];
            inv[i][j] = s * L[i][i];
        }
}

static void make_consts(int win, FbConsts& c) {
    // FarnebackPrepareGaussian(n = 1, sigma = 1.7)
    const int n = 1;
    const double sigma = 1.7;
    float g[3], xg[3], xxg[3];
    double s = 0.;
    for (int x = -n; x <= n; x++) {
        g[x + n] = (float)std::exp(-x * x / (2 * sigma * sigma));
        s += g[x + n];
    }
    s = 1. / s;
    for (int x = -n; x <= n; x++) {
        g[x + n] = (float)(g[x + n] * s);
        xg[x + n] = (float)(x * g[x + n]);
        xxg[x + n] = (float)(x * x * g[x + n]);
    }
    double G[6][6] = {{0}};
    for (int y = -n; y <= n; y++)
        for (int x = -n; x <= n; x++) {
            volatile float gg = g[y + n] * g[x + n];  // float products, widened on accumulation
            volatile float t1 = gg * x; volatile float t2 = t1 * x;
            volatile float t3 = t2 * x; volatile float t4 = t3 * x;
            volatile float u3 = t2 * y; volatile float u4 = u3 * y;
            G[0][0] += gg;
            G[1][1] += t2;
            G[3][3] += t4;
            G[5][5] += u4;
        }
    G[2][2] = G[0][3] = G[0][4] = G[3][0] = G[4][0] = G[1][1];
    G[4][4] = G[3][3];
    G[3][4] = G[4][3] = G[5][5];
    double inv[6][6];
    chol_inv6(G, inv);
    c.g0 = g[1]; c.g1 = g[2]; c.xg1 = xg[2]; c.xxg1 = xxg[2];
    c.ig11 = inv[1][1]; c.ig03 = inv[0][3]; c.ig33 = inv[3][3]; c.ig55 = inv[5][5];
    // FarnebackUpdateFlow_GaussianBlur taps
    int m = win / 2;
    c.m = m;
    double sg = m * 0.3, sum = 1;
    c.negzero2 = make_float2(-0.0f, -0.0f);
    float k[kMaxM + 1];
    k[0] = 1.0f;
    for (int i = 1; i <= m; i++) {
        float t = (float)std::exp(-i * i / (2 * sg * sg));
        k[i] = t;
        sum += t * 2;
    }
    sum = 1. / sum;
    for (int i = 0; i <= kMaxM; i++) {
        float v = i <= m ? (float)(k[i] * sum) : 0.0f;
        c.k2[i] = make_float2(v, v);
    }
}

static inline int plane_pitch(int Sw) { return (Sw + 31) / 32 * 32; }

}  // namespace ma

using namespace ma;

static inline size_t plane_floats(const TileGeom& g) {
    return std::max((size_t)g.Sh * plane_pitch(g.Sw), (size_t)g.Sw * plane_pitch(g.Sh));
}

extern "C" size_t ma_farneback_workspace_bytes(int h, int w, int T, int ov, int n_batch) {
    if (h <= 0 || w <= 0 || n_batch <= 0) return 0;
    TileGeom g = make_geom(h, w, T, ov);
    return (size_t)n_batch * kSlotPlanes * plane_floats(g) * sizeof(float);
}

// window of one iteration along one axis: centre [ov, ov + core) grown by e, clipped to the tile window [0, S);
// blocks start at a 32-float aligned origin
struct Span { int origin, hi; };
static inline Span grown_centre(int S, int core, int ov, int e, bool full) {
    if (full) return {0, S};
    const int lo = std::max(0, ov - e), hi = std::min(S, ov + core + e);
    return {lo & ~31, hi};
}

extern "C" int ma_farneback_tiles_ex(const void* mov, const void* ref, size_t pitch, int dtype, int h, int w,
                                     int T, int ov, int win, int iters, int tile_begin, int tile_end,
                                     float* flow_out, void* workspace, size_t workspace_bytes, unsigned flags, void* stream) {
    const bool fused = (flags & MA_FB_CONTRACT_FMA) != 0;
    const bool full_windows = (flags & MA_FB_FULL_WINDOWS) != 0;   // A/B and tests: no dependency-cone trimming
    if (flags & ~(MA_FB_CONTRACT_FMA | MA_FB_FULL_WINDOWS)) return invalid("ma_farneback_tiles: unknown flag");
    if (!mov || !ref || !flow_out || !workspace || h <= 0 || w <= 0) return invalid("ma_farneback_tiles: bad argument");
    if (dtype != MA_U8 && dtype != MA_U16) return invalid("ma_farneback_tiles: dtype must be MA_U8 or MA_U16");
    if (iters < 1) return invalid("ma_farneback_tiles: iterations must be >= 1");
    if (win < 1 || win / 2 > kMaxM) return invalid("ma_farneback_tiles: winsize must be in [1, 193]");
    if (T > 0 && ov < 0) return invalid("ma_farneback_tiles: negative overlap");
    TileGeom g = make_geom(h, w, T, ov);
    if (g.Sh < 2 || g.Sw < 2) return invalid("ma_farneback_tiles: window too small");
    int ntot = g.ny * g.nx;
    if (tile_begin < 0 || tile_end > ntot || tile_begin > tile_end) return invalid("ma_farneback_tiles: bad tile range");
    if (tile_begin == tile_end) return MA_OK;
    int Sp = plane_pitch(g.Sw), SpT = plane_pitch(g.Sh);
    size_t plane = plane_floats(g);
    if ((reinterpret_cast<uintptr_t>(workspace) & 127) != 0) return invalid("ma_farneback_tiles: workspace must be 128-byte aligned");
    size_t slot_bytes = kSlotPlanes * plane * sizeof(float);
    int cap = (int)std::min<size_t>(workspace_bytes / slot_bytes, 4096);
    if (cap < 1) {
        set_error("ma_farneback_tiles: workspace smaller than one tile slot");
        return MA_ERR_WORKSPACE;
    }
    FbConsts cst;
    make_consts(win, cst);
    cudaStream_t s = (cudaStream_t)stream;
    const int m = cst.m;
    const size_t v_smem = (size_t)(kVOut + 2 * m) * kRowF * sizeof(float);
    const size_t h_smem = std::max((size_t)2 * (kStep + 2 * m) * kRowF * sizeof(float), (size_t)kStep * kFlowPitch * sizeof(float2));
    static std::atomic<bool> attr_set[64];  // per device; setting the attributes twice is harmless
    int dev_id = 0;
    cudaGetDevice(&dev_id);
    if (dev_id < 0 || dev_id >= 64 || !attr_set[dev_id]) {
        MA_CUDA_CHECK(cudaFuncSetAttribute(fb_blur_v_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 256 * 256));
        MA_CUDA_CHECK(cudaFuncSetAttribute(fb_blur_v_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 256 * 256));
        MA_CUDA_CHECK(cudaFuncSetAttribute(fb_blur_h_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 256 * 256));
        MA_CUDA_CHECK(cudaFuncSetAttribute(fb_blur_h_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 256 * 256));
        if (dev_id >= 0 && dev_id < 64) attr_set[dev_id] = true;
    }
    for (int t0 = tile_begin; t0 < tile_end; t0 += cap) {
        FbBatch b;
        b.g = g; b.tile0 = t0; b.ntiles = std::min(cap, tile_end - t0);
        b.Sp = Sp; b.SpT = SpT; b.plane = plane; b.ws = (float*)workspace;
        const double tpx = (double)b.ntiles * g.Sh * g.Sw;
        { KernelScope ks(K_POLYEXP, s, tpx);
        if (g.Sh >= 8 && g.Sw >= 8) {
            dim3 pg(ceil_div(ceil_div(g.Sw, PM_OUTW), PM_WARPS), ceil_div(g.Sh, PM_BAND), b.ntiles);
            if (dtype == MA_U8)
                fb_polyexp_march_kernel<uint8_t><<<pg, PM_WARPS * 32, 0, s>>>((const uint8_t*)mov, (const uint8_t*)ref, pitch, b, cst);
            else
                fb_polyexp_march_kernel<uint16_t><<<pg, PM_WARPS * 32, 0, s>>>((const uint16_t*)mov, (const uint16_t*)ref, pitch, b, cst);
        } else {
            dim3 pg(ceil_div(g.Sw, PE_BW), ceil_div(g.Sh, PE_BH), b.ntiles);
            if (dtype == MA_U8)
                fb_polyexp_kernel<uint8_t><<<pg, 256, 0, s>>>((const uint8_t*)mov, (const uint8_t*)ref, pitch, b, cst);
            else
                fb_polyexp_kernel<uint16_t><<<pg, 256, 0, s>>>((const uint16_t*)mov, (const uint16_t*)ref, pitch, b, cst);
        } }
        // TMA descriptors over this batch's planes: M as [plane][y][x], V^T as [plane][x][y]
        CUtensorMap mapM, mapVT;
        uint64_t nplanes = (uint64_t)b.ntiles * kSlotPlanes;
        if (!make_plane_map(&mapM, b.ws, g.Sw, g.Sh, nplanes, (uint64_t)Sp * 4, plane * 4, kRowF, kVOut + 2 * m) ||
            !make_plane_map(&mapVT, b.ws, g.Sh, g.Sw, nplanes, (uint64_t)SpT * 4, plane * 4, kRowF, kStep + 2 * m)) {
            set_error("ma_farneback_tiles: cuTensorMapEncodeTiled failed");
            return MA_ERR_CUDA;
        }
        for (int it = 0; it < iters; ++it) {
            const bool last = it == iters - 1;
            // dependency cone of the stitched centre (FbWin): this iteration's flow is needed within e of it
            const int e = (iters - 1 - it) * m;
            const Span hx = grown_centre(g.Sw, g.Tw, g.ov, e, full_windows), hy = grown_centre(g.Sh, g.Th, g.ov, e, full_windows);
            const Span vx = grown_centre(g.Sw, g.Tw, g.ov, e + m, full_windows);
            const int big = 1 << 28;   // "no trimming": every block of the grid is needed
            const FbWin wv = {vx.origin, hy.origin, full_windows ? big : e + m, full_windows ? big : e};
            const FbWin wh = {hx.origin, hy.origin, full_windows ? big : e, full_windows ? big : e};
            { const int nbx = ceil_div(vx.hi - vx.origin, kRowF), nby = ceil_div(hy.hi - hy.origin, kVOut);
            KernelScope ks(K_BLUR_V, s, (double)b.ntiles * nbx * kRowF * nby * kVOut);
            dim3 vg(nbx, nby, b.ntiles * 5);
            if (!fused) fb_blur_v_kernel<false><<<vg, 256, v_smem, s>>>(mapM, b, cst, wv);
            else fb_blur_v_kernel<true><<<vg, 256, v_smem, s>>>(mapM, b, cst, wv); }
            { const int nby = ceil_div(hy.hi - hy.origin, kRowF), nbx = ceil_div(hx.hi - hx.origin, kStep);
            KernelScope ks(K_BLUR_H, s, (double)b.ntiles * nby * kRowF * nbx * kStep);
            dim3 hg(nby, nbx, b.ntiles);
            if (!fused) fb_blur_h_kernel<false><<<hg, 256, h_smem, s>>>(mapVT, b, cst, last, (float2*)flow_out, wh);
            else fb_blur_h_kernel<true><<<hg, 256, h_smem, s>>>(mapVT, b, cst, last, (float2*)flow_out, wh); }
            if (!last) {   // M of the next iteration: pointwise in this iteration's flow, so the same window
                const int nbx = ceil_div(hx.hi - hx.origin, 64), nby = ceil_div(hy.hi - hy.origin, 4);
                KernelScope ks(K_UPDATE0, s, (double)b.ntiles * nbx * 64 * nby * 4);
                fb_update_kernel<<<dim3(nbx, nby, b.ntiles), 256, 0, s>>>(b, wh);
            }
        }
        MA_LAUNCH_CHECK("farneback kernels");
    }
    return MA_OK;
}

extern "C" int ma_farneback_tiles(const void* mov, const void* ref, size_t pitch, int dtype, int h, int w,
                                  int T, int ov, int win, int iters, int tile_begin, int tile_end,
                                  float* flow_out, void* workspace, size_t workspace_bytes, void* stream) {
    return ma_farneback_tiles_ex(mov, ref, pitch, dtype, h, w, T, ov, win, iters, tile_begin, tile_end, flow_out, workspace,
                                 workspace_bytes, 0u, stream);
}
