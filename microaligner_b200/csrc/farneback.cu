// K1-K3: tiled single-scale Farneback optical flow, bit-compatible with
//   cv.calcOpticalFlowFarneback(mov, ref, None, 0.5, levels=0, win, iters, poly_n=1, poly_sigma=1.7,
//                               OPTFLOW_FARNEBACK_GAUSSIAN)
// as called per tile by the reference (optflow_reg/flow_calc.py:30-47, 59-98).
//
// Stages (all per S x S tile window; a "tile" is only an index range, never a copy):
//   fb_polyexp_kernel   u8/u16 window -> 3x3 prefilter (REFLECT_101 at the tile edge) -> polynomial
//                       expansion n=1 (f32 vertical, f64 horizontal) -> R (5 planar f32 planes)
//   fb_update0_kernel   M = UpdateMatrices(R0, R1, flow = 0)
//   fb_blur_v_kernel    V = vertical (2m+1)-tap Gaussian of M, rows replicated at the tile edge
//   fb_blur_h_kernel    horizontal (2m+1)-tap Gaussian of V, 2x2 solve in f64 -> flow, then either
//                       M = UpdateMatrices(R0, R1, flow) in place (not last iteration) or scatter of
//                       the tile centre into the stitched flow (last iteration).
//
// Arithmetic contract: OpenCV's optflowgf.cpp is built for the SSE3 baseline, i.e. every a*b+c is a
// separately rounded multiply and add.  This file is compiled with -fmad=false and keeps OpenCV's
// association order (s = c*k0; s += (a[+i] + a[-i]) * k[i], i = 1..m), so results are bit-identical.
//
// The two blur kernels are FP32-issue bound (2 passes x 5 planes x (1 + 3m) instr per pixel), not
// HBM bound; both sweep a strip with a shared-memory ring so each input element is read from
// global memory exactly once per pass, and every thread keeps 8 outputs + two 8-wide sliding
// windows in registers so shared-memory traffic is 2 loads per 24 FP instructions.
#include <cmath>
#include <vector>
#include "common.cuh"

namespace ma {

constexpr int kMaxM = 96;         // ring capacity 256 >= step 64 + 2*m
constexpr int kRing = 256;        // virtual rows/cols kept in the ring (power of two)
constexpr int kStep = 64;         // outputs per sweep step along the convolution axis
constexpr int kR = 8;             // outputs per thread along the convolution axis
constexpr int kLanes = 32;        // strip width across the convolution axis
constexpr int kHPitch = 33;       // H-pass ring pitch (floats): conflict-free transposed stores

struct FbConsts {
    float g0, g1, xg1, xxg1;            // polyexp taps: centre g[0], side g[1], x*g, x*x*g at +1
    double ig11, ig03, ig33, ig55;      // inverse Gram constants
    int m;                              // blur half width
    float k[kMaxM + 1];                 // blur taps
};

struct FbBatch {
    TileGeom g;
    int tile0;        // first tile (row-major index) of this batch
    int ntiles;       // tiles in this batch
    int Sp;           // plane row pitch in floats
    size_t plane;     // floats per plane (Sh * Sp)
    float* ws;        // workspace base: per slot 20 planes [R0 x5][R1 x5][M x5][V x5]
};

__device__ __forceinline__ float* slot_plane(const FbBatch& b, int slot, int which /*0 R0,1 R1,2 M,3 V*/, int c) {
    return b.ws + ((size_t)slot * 20 + which * 5 + c) * b.plane;
}

// ------------------------------------------------------------------------------------------------
// K1: prefilter + polynomial expansion
// ------------------------------------------------------------------------------------------------
constexpr int PE_BW = 32, PE_BH = 8;

template <typename T>
__device__ __forceinline__ float window_px(const T* __restrict__ img, size_t pitch, const TileGeom& g,
                                           int oy, int ox, int ty, int tx) {
    int gy = oy + ty, gx = ox + tx;  // (ty,tx) is inside the window; zero padding outside the image
    if ((unsigned)gy >= (unsigned)g.h || (unsigned)gx >= (unsigned)g.w) return 0.0f;
    return (float)__ldg((const T*)((const char*)img + (size_t)gy * pitch) + gx);
}

template <typename T>
__global__ void __launch_bounds__(PE_BW* PE_BH) fb_polyexp_kernel(const T* __restrict__ mov, const T* __restrict__ ref,
                                                                   size_t pitch, FbBatch b,
                                                                   const __grid_constant__ FbConsts cst) {
    __shared__ float P[PE_BH + 2][PE_BW + 2];
    __shared__ float T0[PE_BH][PE_BW + 2], T1[PE_BH][PE_BW + 2], T2[PE_BH][PE_BW + 2];
    const TileGeom& g = b.g;
    int slot = blockIdx.z >> 1, which = blockIdx.z & 1;  // 0: prev = moving -> R0, 1: next = reference -> R1
    const T* img = which ? ref : mov;
    int tile = b.tile0 + slot;
    int ti = tile / g.nx, tj = tile % g.nx;
    int oy = ti * g.Th - g.ov, ox = tj * g.Tw - g.ov;
    int x0 = blockIdx.x * PE_BW, y0 = blockIdx.y * PE_BH;
    int tid = threadIdx.y * PE_BW + threadIdx.x;

    // prefiltered image at the (clamped = replicated) positions the expansion will read
    for (int p = tid; p < (PE_BH + 2) * (PE_BW + 2); p += PE_BW * PE_BH) {
        int r = p / (PE_BW + 2), c = p % (PE_BW + 2);
        int py = min(max(y0 - 1 + r, 0), g.Sh - 1), px = min(max(x0 - 1 + c, 0), g.Sw - 1);
        int xl = reflect101(px - 1, g.Sw), xr = reflect101(px + 1, g.Sw);
        int yu = reflect101(py - 1, g.Sh), yd = reflect101(py + 1, g.Sh);
        auto rowf = [&](int yy) {
            float cc = window_px(img, pitch, g, oy, ox, yy, px);
            float l = window_px(img, pitch, g, oy, ox, yy, xl), rr = window_px(img, pitch, g, oy, ox, yy, xr);
            return __fadd_rn(__fmul_rn(cc, 0.5f), __fmul_rn(__fadd_rn(l, rr), 0.25f));
        };
        float tu = rowf(yu), tc = rowf(py), td = rowf(yd);
        P[r][c] = __fadd_rn(__fmul_rn(tc, 0.5f), __fmul_rn(__fadd_rn(tu, td), 0.25f));
    }
    __syncthreads();
    // vertical pass of the expansion (f32): rows r (up), r+1 (centre), r+2 (down)
    for (int p = tid; p < PE_BH * (PE_BW + 2); p += PE_BW * PE_BH) {
        int r = p / (PE_BW + 2), c = p % (PE_BW + 2);
        float s0 = P[r][c], sc = P[r + 1][c], s1 = P[r + 2][c];
        float pp = __fadd_rn(s0, s1);
        T0[r][c] = __fadd_rn(__fmul_rn(sc, cst.g0), __fmul_rn(cst.g1, pp));
        T1[r][c] = __fadd_rn(0.0f, __fmul_rn(cst.xg1, __fsub_rn(s1, s0)));
        T2[r][c] = __fadd_rn(0.0f, __fmul_rn(cst.xxg1, pp));
    }
    __syncthreads();
    int x = x0 + threadIdx.x, y = y0 + threadIdx.y;
    if (x >= g.Sw || y >= g.Sh) return;
    int r = threadIdx.y, c = threadIdx.x + 1;
    // horizontal pass: float sums/differences, double accumulation (see oracle/farneback_np.py:polyexp)
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
    size_t o = (size_t)y * b.Sp + x;
    slot_plane(b, slot, which, 0)[o] = (float)__dmul_rn(b3, cst.ig11);
    slot_plane(b, slot, which, 1)[o] = (float)__dmul_rn(b2, cst.ig11);
    slot_plane(b, slot, which, 2)[o] = (float)__dadd_rn(__dmul_rn(b1, cst.ig03), __dmul_rn(b5, cst.ig33));
    slot_plane(b, slot, which, 3)[o] = (float)__dadd_rn(__dmul_rn(b1, cst.ig03), __dmul_rn(b4, cst.ig33));
    slot_plane(b, slot, which, 4)[o] = (float)__dmul_rn(b6, cst.ig55);
}

// ------------------------------------------------------------------------------------------------
// K2: UpdateMatrices for one pixel (FarnebackUpdateMatrices, all f32, left-to-right sums)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float border_w(int d) {
    // {0.14, 0.14, 0.4472, 0.4472, 0.4472}
    return d < 2 ? 0.14f : 0.4472f;
}

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
    r2 = __fmul_rn(__fsub_rn(__ldg(R0 + o), r2), 0.5f);
    r3 = __fmul_rn(__fsub_rn(__ldg(R0 + plane + o), r3), 0.5f);
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

__global__ void __launch_bounds__(256) fb_update0_kernel(FbBatch b) {
    int x = blockIdx.x * 64 + (threadIdx.x & 63);
    int y = blockIdx.y * 4 + (threadIdx.x >> 6);
    int slot = blockIdx.z;
    if (x >= b.g.Sw || y >= b.g.Sh) return;
    update_matrices_px(slot_plane(b, slot, 0, 0), slot_plane(b, slot, 1, 0), b.plane, b.Sp, b.g.Sw, b.g.Sh,
                       x, y, 0.0f, 0.0f, slot_plane(b, slot, 2, 0));
}

// ------------------------------------------------------------------------------------------------
// K3 core: symmetric (2m+1)-tap convolution of 8 consecutive outputs per thread, OpenCV order.
// `ring` points at this thread's lane; element for virtual index v is ring[(v & (kRing-1)) * pitch].
// Outputs v0 .. v0+7.  Two 8-wide register windows slide in opposite directions; with the loop
// unrolled by 8 all register indices are static.
// ------------------------------------------------------------------------------------------------
template <int PITCH>
__device__ __forceinline__ void conv8_sym(const float* __restrict__ ring, int v0, const FbConsts& cst, float (&acc)[kR]) {
    float wp[kR], wm[kR];
    const float k0 = cst.k[0];
#pragma unroll
    for (int j = 0; j < kR; ++j) {
        float c = ring[((v0 + j) & (kRing - 1)) * PITCH];
        wp[j] = c;
        wm[j] = c;
        acc[j] = __fmul_rn(c, k0);
    }
    const int m = cst.m;
    // invariant before step i: wp[(j+i-1)&7] = in[v0+j+i-1], wm[(j-i+1)&7] = in[v0+j-i+1]
    for (int i0 = 1; i0 <= m; i0 += kR) {
#pragma unroll
        for (int s = 0; s < kR; ++s) {
            int i = i0 + s;
            if (i <= m) {  // warp-uniform
                float ki = cst.k[i];
                // (i0 - 1) is a multiple of 8, so (x + i) & 7 == (x + s + 1) & 7: static indices
                wp[(kR - 1 + s + 1) & 7] = ring[((v0 + kR - 1 + i) & (kRing - 1)) * PITCH];
                wm[(8 * kR - s - 1) & 7] = ring[((v0 - i) & (kRing - 1)) * PITCH];
#pragma unroll
                for (int j = 0; j < kR; ++j) {
                    float t = __fadd_rn(wp[(j + s + 1) & 7], wm[(j + 8 * kR - s - 1) & 7]);
                    acc[j] = __fadd_rn(acc[j], __fmul_rn(t, ki));
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// K3a: vertical pass M -> V.  CTA = strip of 32 columns of one plane of one tile, swept downwards in
// steps of 64 rows (8 warps x 8 rows); lane <-> column (coalesced 128-byte rows, conflict-free smem).
// grid = (ceil(Sw/32), nseg, ntiles*5)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) fb_blur_v_kernel(FbBatch b, const __grid_constant__ FbConsts cst, int rows_per_seg) {
    __shared__ float ring[kRing * kLanes];
    const int Sh = b.g.Sh, Sw = b.g.Sw, Sp = b.Sp, m = cst.m;
    int slot = blockIdx.z / 5, c = blockIdx.z % 5;
    const float* __restrict__ src = slot_plane(b, slot, 2, c);
    float* __restrict__ dst = slot_plane(b, slot, 3, c);
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int x = blockIdx.x * kLanes + lane;
    int xc = min(x, Sw - 1);
    int ybeg = blockIdx.y * rows_per_seg, yend = min(ybeg + rows_per_seg, Sh);
    int loaded = ybeg - m;  // next virtual row to load
    for (int y0 = ybeg; y0 < yend; y0 += kStep) {
        int need = min(y0 + kStep, yend) + m;  // exclusive
        __syncthreads();                       // previous step finished reading the slots we overwrite
        for (int v = loaded + warp; v < need; v += 8) {
            int yy = min(max(v, 0), Sh - 1);
            ring[(v & (kRing - 1)) * kLanes + lane] = __ldg(src + (size_t)yy * Sp + xc);
        }
        loaded = need;
        __syncthreads();
        int v0 = y0 + warp * kR;
        if (v0 < yend) {
            float acc[kR];
            conv8_sym<kLanes>(ring + lane, v0, cst, acc);
            if (x < Sw) {
#pragma unroll
                for (int j = 0; j < kR; ++j)
                    if (v0 + j < yend) dst[(size_t)(v0 + j) * Sp + x] = acc[j];
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// K3b: horizontal pass + solve + (UpdateMatrices | final scatter).  CTA = strip of 32 rows of one
// tile, all 5 planes, swept rightwards in steps of 64 columns (8 warps x 8 columns); lane <-> row
// for the convolution (ring stored transposed, pitch 33), lane <-> column for the global-memory
// phases so every HBM access is coalesced.
// grid = (ceil(Sh/32), nseg, ntiles)
// ------------------------------------------------------------------------------------------------
constexpr size_t kBlurHSmem = (size_t)5 * kRing * kHPitch * sizeof(float) + (size_t)kLanes * kStep * sizeof(float2);

__global__ void __launch_bounds__(256, 1) fb_blur_h_kernel(FbBatch b, const __grid_constant__ FbConsts cst, int cols_per_seg,
                                                            int last_iter, float2* __restrict__ flow_out) {
    extern __shared__ __align__(16) float smem[];
    float* ring = smem;                                             // [5][kRing][33]
    float2* fl = (float2*)(smem + 5 * kRing * kHPitch);             // [32 rows][64 cols]
    const TileGeom& g = b.g;
    const int Sh = g.Sh, Sw = g.Sw, Sp = b.Sp, m = cst.m;
    int slot = blockIdx.z;
    const float* __restrict__ V = slot_plane(b, slot, 3, 0);
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int ybase = blockIdx.x * kLanes;
    int xbeg = blockIdx.y * cols_per_seg, xend = min(xbeg + cols_per_seg, Sw);
    int loaded = xbeg - m;
    int tile = b.tile0 + slot;
    int ti = tile / g.nx, tj = tile % g.nx;
    for (int x0 = xbeg; x0 < xend; x0 += kStep) {
        int need = min(x0 + kStep, xend) + m;
        __syncthreads();
        // load virtual columns [loaded, need) for 32 rows x 5 planes; thread <-> column (coalesced)
        int ncol = need - loaded;
        for (int p = threadIdx.x; p < ncol * kLanes * 5; p += 256) {
            int cidx = p % ncol, rest = p / ncol;
            int r = rest % kLanes, c = rest / kLanes;
            int v = loaded + cidx;
            int xx = min(max(v, 0), Sw - 1), yy = min(ybase + r, Sh - 1);
            ring[(c * kRing + (v & (kRing - 1))) * kHPitch + r] = __ldg(V + c * b.plane + (size_t)yy * Sp + xx);
        }
        loaded = need;
        __syncthreads();
        int v0 = x0 + warp * kR;
        if (v0 < xend) {
            float a0[kR], a1[kR], a2[kR], a3[kR], a4[kR];
            conv8_sym<kHPitch>(ring + 0 * kRing * kHPitch + lane, v0, cst, a0);
            conv8_sym<kHPitch>(ring + 1 * kRing * kHPitch + lane, v0, cst, a1);
            conv8_sym<kHPitch>(ring + 2 * kRing * kHPitch + lane, v0, cst, a2);
            conv8_sym<kHPitch>(ring + 3 * kRing * kHPitch + lane, v0, cst, a3);
            conv8_sym<kHPitch>(ring + 4 * kRing * kHPitch + lane, v0, cst, a4);
#pragma unroll
            for (int j = 0; j < kR; ++j) {
                double g11 = a0[j], g12 = a1[j], g22 = a2[j], h1 = a3[j], h2 = a4[j];
                double det = __dadd_rn(__dsub_rn(__dmul_rn(g11, g22), __dmul_rn(g12, g12)), 1e-3);
                double idet = __ddiv_rn(1.0, det);
                float fx = (float)__dmul_rn(__dsub_rn(__dmul_rn(g11, h2), __dmul_rn(g12, h1)), idet);
                float fy = (float)__dmul_rn(__dsub_rn(__dmul_rn(g22, h1), __dmul_rn(g12, h2)), idet);
                fl[lane * kStep + warp * kR + j] = make_float2(fx, fy);
            }
        }
        __syncthreads();
        // epilogue: thread <-> column
        int cx = threadIdx.x & 63;
        int x = x0 + cx;
        if (x < xend) {
            for (int r = threadIdx.x >> 6; r < kLanes; r += 4) {
                int y = ybase + r;
                if (y >= Sh) break;
                float2 f = fl[r * kStep + cx];
                if (last_iter) {
                    int cy = y - g.ov, cxx = x - g.ov;
                    if ((unsigned)cy < (unsigned)g.Th && (unsigned)cxx < (unsigned)g.Tw) {
                        int gy = ti * g.Th + cy, gx = tj * g.Tw + cxx;
                        if (gy < g.h && gx < g.w) flow_out[(size_t)gy * g.w + gx] = f;
                    }
                } else {
                    update_matrices_px(slot_plane(b, slot, 0, 0), slot_plane(b, slot, 1, 0), b.plane, Sp, Sw, Sh,
                                       x, y, f.x, f.y, slot_plane(b, slot, 2, 0));
                }
            }
        }
    }
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
    c.k[0] = 1.0f;
    for (int i = 1; i <= m; i++) {
        float t = (float)std::exp(-i * i / (2 * sg * sg));
        c.k[i] = t;
        sum += t * 2;
    }
    sum = 1. / sum;
    for (int i = 0; i <= m; i++) c.k[i] = (float)(c.k[i] * sum);
    for (int i = m + 1; i <= kMaxM; i++) c.k[i] = 0.0f;
}

static inline int plane_pitch(int Sw) { return (Sw + 31) / 32 * 32; }

}  // namespace ma

using namespace ma;

extern "C" size_t ma_farneback_workspace_bytes(int h, int w, int T, int ov, int n_batch) {
    if (h <= 0 || w <= 0 || n_batch <= 0) return 0;
    TileGeom g = make_geom(h, w, T, ov);
    return (size_t)n_batch * 20 * (size_t)g.Sh * plane_pitch(g.Sw) * sizeof(float);
}

extern "C" int ma_farneback_tiles(const void* mov, const void* ref, size_t pitch, int dtype, int h, int w,
                                  int T, int ov, int win, int iters, int tile_begin, int tile_end,
                                  float* flow_out, void* workspace, size_t workspace_bytes, void* stream) {
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
    int Sp = plane_pitch(g.Sw);
    size_t plane = (size_t)g.Sh * Sp;
    size_t slot_bytes = 20 * plane * sizeof(float);
    int cap = (int)std::min<size_t>(workspace_bytes / slot_bytes, 4096);
    if (cap < 1) {
        set_error("ma_farneback_tiles: workspace smaller than one tile slot");
        return MA_ERR_WORKSPACE;
    }
    FbConsts cst;
    make_consts(win, cst);
    cudaStream_t s = (cudaStream_t)stream;
    static bool attr_set = false;
    if (!attr_set) {
        MA_CUDA_CHECK(cudaFuncSetAttribute(fb_blur_h_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBlurHSmem));
        attr_set = true;
    }
    for (int t0 = tile_begin; t0 < tile_end; t0 += cap) {
        FbBatch b;
        b.g = g; b.tile0 = t0; b.ntiles = std::min(cap, tile_end - t0);
        b.Sp = Sp; b.plane = plane; b.ws = (float*)workspace;
        dim3 pg(ceil_div(g.Sw, PE_BW), ceil_div(g.Sh, PE_BH), b.ntiles * 2), pb(PE_BW, PE_BH);
        if (dtype == MA_U8)
            fb_polyexp_kernel<uint8_t><<<pg, pb, 0, s>>>((const uint8_t*)mov, (const uint8_t*)ref, pitch, b, cst);
        else
            fb_polyexp_kernel<uint16_t><<<pg, pb, 0, s>>>((const uint16_t*)mov, (const uint16_t*)ref, pitch, b, cst);
        fb_update0_kernel<<<dim3(ceil_div(g.Sw, 64), ceil_div(g.Sh, 4), b.ntiles), 256, 0, s>>>(b);
        // split strips into segments when the grid would not fill the GPU (each segment re-reads a 2m halo)
        int vstrips = ceil_div(g.Sw, kLanes) * b.ntiles * 5, hstrips = ceil_div(g.Sh, kLanes) * b.ntiles;
        int vseg = 1, hseg = 1;
        while (vstrips * vseg < 148 * 4 && ceil_div(g.Sh, vseg * 2) >= 2 * kStep) vseg *= 2;
        while (hstrips * hseg < 148 && ceil_div(g.Sw, hseg * 2) >= 2 * kStep) hseg *= 2;
        int rows_per_seg = ceil_div(ceil_div(g.Sh, vseg), kR) * kR;
        int cols_per_seg = ceil_div(ceil_div(g.Sw, hseg), kR) * kR;
        vseg = ceil_div(g.Sh, rows_per_seg);
        hseg = ceil_div(g.Sw, cols_per_seg);
        for (int it = 0; it < iters; ++it) {
            fb_blur_v_kernel<<<dim3(ceil_div(g.Sw, kLanes), vseg, b.ntiles * 5), 256, 0, s>>>(b, cst, rows_per_seg);
            fb_blur_h_kernel<<<dim3(ceil_div(g.Sh, kLanes), hseg, b.ntiles), 256, kBlurHSmem, s>>>(
                b, cst, cols_per_seg, it == iters - 1, (float2*)flow_out);
        }
        MA_LAUNCH_CHECK("farneback kernels");
    }
    return MA_OK;
}
