// Experimental variants of the Farneback kernels -- included by farneback.cu only, after the default kernels.
//
// None of them is used unless the caller asks for it through the flag bits MA_FB_VARIANT_SHIFT_V / _H / _P of
// ma_farneback_tiles_ex (Python: MA_FB_VARIANT=v,h,p or ops.farneback_tiles(variant=...)).  All compute exactly the
// arithmetic of the default kernels (conv8x2's order of operations, fb_polyexp_kernel's expressions), so results are
// bit-identical; they differ in how work is mapped to warps.  Status and measurements: DESIGN.md section 9,
// profiles/r01_ab_blur_variants.log, scripts/ab_pipeline.py.
#pragma once

namespace ma {

// ------------------------------------------------------------------------------------------------
// K1 experimental variant (flags bits 16..19 = 1): "marching" polynomial expansion without shared memory
// and without block barriers.  A warp owns a strip of 28 output columns (lanes 2..29; lanes 0, 1, 30, 31
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


// ---- epilogues that store straight from the accumulator registers (shared by the experimental variants) ----
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

__device__ __forceinline__ float2 solve_flow(float g11f, float g12f, float g22f, float h1f, float h2f) {
    // FarnebackUpdateFlow_GaussianBlur: 2x2 solve in f64
    const double g11 = g11f, g12 = g12f, g22 = g22f, h1 = h1f, h2 = h2f;
    const double idet = __ddiv_rn(1.0, __dadd_rn(__dsub_rn(__dmul_rn(g11, g22), __dmul_rn(g12, g12)), 1e-3));
    float2 o;
    o.x = (float)__dmul_rn(__dsub_rn(__dmul_rn(g11, h2), __dmul_rn(g12, h1)), idet);
    o.y = (float)__dmul_rn(__dsub_rn(__dmul_rn(g22, h1), __dmul_rn(g12, h2)), idet);
    return o;
}

// H pass: solve and store the thread's 2 rows (y, y + 1) x 8 columns (xb .. xb + 7) -- contiguous along x in the tile
// flow plane (full 64-byte vectors; the pitch Sp is a multiple of 32) and in the stitched flow (checked per pixel)
__device__ __forceinline__ void solve_store_strip(const FbBatch& b, int slot, int y0, int xb, int lane, int last_iter,
                                                  float2* __restrict__ flow_out, const u64 (&a0)[kR], const u64 (&a1)[kR],
                                                  const u64 (&a2)[kR], const u64 (&a3)[kR], const u64 (&a4)[kR]) {
    const TileGeom& g = b.g;
    const int Sh = g.Sh, Sw = g.Sw;
    const int tile = b.tile0 + slot;
    const int ti = tile / g.nx, tj = tile - ti * g.nx;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        const int y = y0 + 2 * lane + half;
        if (y >= Sh) continue;
        float2 f[kR];
#pragma unroll
        for (int j = 0; j < kR; ++j) {
            const float2 G11 = unpack2(a0[j]), G12 = unpack2(a1[j]), G22 = unpack2(a2[j]);
            const float2 H1 = unpack2(a3[j]), H2 = unpack2(a4[j]);
            f[j] = half ? solve_flow(G11.y, G12.y, G22.y, H1.y, H2.y) : solve_flow(G11.x, G12.x, G22.x, H1.x, H2.x);
        }
        if (last_iter) {   // scatter the tile centre into the stitched flow
            const int cy = y - g.ov, gy = ti * g.Th + cy;
            if ((unsigned)cy < (unsigned)g.Th && gy < g.h) {
#pragma unroll
                for (int j = 0; j < kR; ++j) {
                    const int cxx = xb + j - g.ov, gx = tj * g.Tw + cxx;
                    if (xb + j < Sw && (unsigned)cxx < (unsigned)g.Tw && gx < g.w) flow_out[(size_t)gy * g.w + gx] = f[j];
                }
            }
        } else {
            float4* __restrict__ dst = reinterpret_cast<float4*>(reinterpret_cast<float2*>(slot_plane(b, slot, 4, 0)) + (size_t)y * b.Sp + xb);
#pragma unroll
            for (int q = 0; q < kR / 2; ++q) dst[q] = make_float4(f[2 * q].x, f[2 * q].y, f[2 * q + 1].x, f[2 * q + 1].y);
        }
    }
}

// Variant "direct" of fb_blur_v_kernel: same CTA-per-box structure, but no transpose stage and no barrier after the
// convolution -- every thread writes its outputs with store_vt_strip.
template <int OUT, bool FUSED>
__global__ void __launch_bounds__(256) fb_blur_v_direct_kernel(const __grid_constant__ CUtensorMap mapM, FbBatch b,
                                                               const __grid_constant__ FbConsts cst) {
    extern __shared__ __align__(128) float smem[];
    __shared__ __align__(8) uint64_t bar;
    const int Sh = b.g.Sh, Sw = b.g.Sw, m = cst.m;
    const int slot = blockIdx.z / 5, c = blockIdx.z % 5;
    const int x0 = blockIdx.x * kRowF, y0 = blockIdx.y * OUT;
    const int rows = OUT + 2 * m;
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
#pragma unroll
    for (int grp = 0; grp < OUT / 64; ++grp) {
        const int o0 = grp * 64 + warp * kR;
        if (y0 + o0 < Sh) {
            u64 acc[kR];
            conv8x2<FUSED>(reinterpret_cast<const u64*>(smem) + (o0 + m) * kRowU + lane, m, cst.k2, nz, acc);
            store_vt_strip(b, slot, c, x0 + 2 * lane, y0 + o0, Sw, acc);
        }
    }
}

// Variant "direct" of fb_blur_h_kernel: same CTA-per-block structure and plane ring, but the flow is solved and
// stored from registers (no flow stage in shared memory, no barrier after the last plane).
template <bool FUSED>
__global__ void __launch_bounds__(256, 2) fb_blur_h_direct_kernel(const __grid_constant__ CUtensorMap mapVT, FbBatch b,
                                                                   const __grid_constant__ FbConsts cst, int last_iter,
                                                                   float2* __restrict__ flow_out) {
    extern __shared__ __align__(128) float smem[];
    __shared__ __align__(8) uint64_t bars[2];
    const int Sw = b.g.Sw, m = cst.m;
    const int slot = blockIdx.z;
    const int y0 = blockIdx.x * kRowF, x0 = blockIdx.y * kStep;
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
        if (c + 2 < 5) {
            __syncthreads();  // buffer c&1 is free again
            if (threadIdx.x == 0) {
                mbar_expect_tx(&bars[c & 1], box_bytes);
                tma_load_3d(buf[c & 1], &mapVT, y0, x0 - m, slot * kSlotPlanes + 15 + c + 2, &bars[c & 1]);
            }
        }
    }
    if (active) solve_store_strip(b, slot, y0, x0 + warp * kR, lane, last_iter, flow_out, acc[0], acc[1], acc[2], acc[3], acc[4]);
}

// ------------------------------------------------------------------------------------------------
// K3 (pipelined variant, MA_FB_PIPELINED): the same two passes as persistent kernels without CTA barriers.
//
// Grid = 2 CTAs per SM; work items (one 64-output box of one plane for the V pass; one 64 x 64 block of a
// tile = five boxes for the H pass) are dealt round-robin to the CTAs.  A two-stage ring of TMA boxes with
// full / empty mbarriers per stage replaces __syncthreads: thread 0 issues the load of box n+1 when it
// starts box n (the stage it refills was released by all 8 warps after box n-1), every warp convolves its
// own 8-output strip of the box, releases the stage and writes its results straight from registers -- the 8
// outputs a thread holds per column are contiguous along the fast axis of the destination (V^T rows for the
// V pass, flow rows for the H pass), so there is no transpose through shared memory, and the load of the
// next box, the convolution of this one and the stores of the previous one overlap across warps.
// Rows outside the tile are not patched in shared memory; edge boxes clamp the row index instead.
// Arithmetic (order of operations, rounding) is exactly that of conv8x2.
// ------------------------------------------------------------------------------------------------
constexpr int kPipeStages = 2;
constexpr int kPipeWarps = 8;
constexpr int kPipeThreads = kPipeWarps * 32;

// conv8x2 on a box addressed by row number: col -> this lane's pair in box row 0; output j is centred on box row
// r0 + j.  EDGE: rows are clamped to [r_lo, r_hi] (the box rows of the first / last row inside the tile).
template <bool FUSED, bool EDGE>
__device__ __forceinline__ void conv8x2p(const u64* __restrict__ col, int r0, int r_lo, int r_hi, int m,
                                         const float2* __restrict__ k2, u64 nz, u64 (&acc)[kR]) {
    auto row = [&](int r) -> u64 {
        if (EDGE) r = min(max(r, r_lo), r_hi);
        return col[r * kRowU];
    };
    u64 wp[kR], wm[kR];
    const u64 k0 = *reinterpret_cast<const u64*>(&k2[0]);
#pragma unroll
    for (int j = 0; j < kR; ++j) {
        u64 c = row(r0 + j);
        wp[j] = c;
        wm[j] = c;
        acc[j] = mul2(c, k0, nz);
    }
    int i = 1;
#pragma unroll 1
    for (; i + kR - 1 <= m; i += kR) {
#pragma unroll
        for (int s = 0; s < kR; ++s) {
            wp[s] = row(r0 + kR - 1 + i + s);
            wm[(63 - s) & 7] = row(r0 - i - s);
            const u64 kk = *reinterpret_cast<const u64*>(&k2[i + s]);
#pragma unroll
            for (int j = 0; j < kR; ++j) {
                const u64 pr = add2(wp[(j + s + 1) & 7], wm[(j + 63 - s) & 7]);
                acc[j] = FUSED ? fma2(pr, kk, acc[j]) : add2(acc[j], mul2(pr, kk, nz));
            }
        }
    }
#pragma unroll
    for (int s = 0; s < kR - 1; ++s) {
        if (i + s <= m) {
            wp[s] = row(r0 + kR - 1 + i + s);
            wm[(63 - s) & 7] = row(r0 - i - s);
            const u64 kk = *reinterpret_cast<const u64*>(&k2[i + s]);
#pragma unroll
            for (int j = 0; j < kR; ++j) {
                const u64 pr = add2(wp[(j + s + 1) & 7], wm[(j + 63 - s) & 7]);
                acc[j] = FUSED ? fma2(pr, kk, acc[j]) : add2(acc[j], mul2(pr, kk, nz));
            }
        }
    }
}

// one strip of one box: v_first = coordinate (along the convolution axis) of box row 0, n = tile extent
template <bool FUSED>
__device__ __forceinline__ void conv_strip(const float* buf, int lane, int strip, int rows, int v_first, int n, int m,
                                           const float2* __restrict__ k2, u64 nz, u64 (&acc)[kR]) {
    const u64* col = reinterpret_cast<const u64*>(buf) + lane;
    const int r0 = strip * kR + m;
    if (v_first >= 0 && v_first + rows <= n) conv8x2p<FUSED, false>(col, r0, 0, 0, m, k2, nz, acc);
    else conv8x2p<FUSED, true>(col, r0, -v_first, n - 1 - v_first, m, k2, nz, acc);
}

template <bool FUSED>
__global__ void __launch_bounds__(kPipeThreads, 2) fb_blur_v_pipe_kernel(const __grid_constant__ CUtensorMap mapM, FbBatch b,
                                                                         const __grid_constant__ FbConsts cst) {
    extern __shared__ __align__(128) float smem[];
    __shared__ __align__(8) uint64_t full[kPipeStages], empty[kPipeStages];
    const int Sh = b.g.Sh, Sw = b.g.Sw, m = cst.m;
    const int rows = kStep + 2 * m;
    const int nbx = (Sw + kRowF - 1) / kRowF, nby = (Sh + kStep - 1) / kStep;
    const int per_plane = nbx * nby;
    const int total = per_plane * 5 * b.ntiles;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        for (int s = 0; s < kPipeStages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], kPipeWarps);
        }
        fence_mbar_init();
    }
    __syncthreads();
    // box nn of this CTA = item blockIdx.x + nn * gridDim.x, loaded into stage nn % 2 (its use number nn / 2)
    auto issue = [&](int nn) {
        const int item = blockIdx.x + nn * gridDim.x;
        if (item >= total) return;
        const int s = nn % kPipeStages, use = nn / kPipeStages;
        if (use > 0) mbar_wait_guarded(&empty[s], (use - 1) & 1);   // all warps released the box that sat here before
        const int pl = item / per_plane, rem = item - pl * per_plane;
        const int by = rem / nbx, bx = rem - by * nbx;
        const int slot = pl / 5, c = pl - slot * 5;
        mbar_expect_tx(&full[s], rows * kRowF * sizeof(float));
        tma_load_3d(smem + (size_t)s * rows * kRowF, &mapM, bx * kRowF, by * kStep - m, slot * kSlotPlanes + 10 + c, &full[s]);
    };
    if (threadIdx.x == 0) {
        issue(0);
        issue(1);
    }
    const u64 nz = *reinterpret_cast<const u64*>(&cst.negzero2);
    int n = 0;
    for (int item = blockIdx.x; item < total; item += gridDim.x, ++n) {
        if (threadIdx.x == 0 && n > 0) issue(n + 1);
        __syncwarp();
        const int s = n % kPipeStages;
        const int pl = item / per_plane, rem = item - pl * per_plane;
        const int by = rem / nbx, bx = rem - by * nbx;
        const int slot = pl / 5, c = pl - slot * 5;
        const int x0 = bx * kRowF, y0 = by * kStep;
        const bool active = y0 + warp * kR < Sh;
        u64 acc[kR];
        mbar_wait_guarded(&full[s], (n / kPipeStages) & 1);
        if (active) conv_strip<FUSED>(smem + (size_t)s * rows * kRowF, lane, warp, rows, y0 - m, Sh, m, cst.k2, nz, acc);
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);   // this warp no longer reads the stage
        if (active) {
            store_vt_strip(b, slot, c, x0 + 2 * lane, y0 + warp * kR, Sw, acc);
        }
    }
}

template <bool FUSED>
__global__ void __launch_bounds__(kPipeThreads, 2) fb_blur_h_pipe_kernel(const __grid_constant__ CUtensorMap mapVT, FbBatch b,
                                                                         const __grid_constant__ FbConsts cst, int last_iter,
                                                                         float2* __restrict__ flow_out) {
    extern __shared__ __align__(128) float smem[];
    __shared__ __align__(8) uint64_t full[kPipeStages], empty[kPipeStages];
    const TileGeom& g = b.g;
    const int Sh = g.Sh, Sw = g.Sw, m = cst.m;
    const int rows = kStep + 2 * m;
    const int nby = (Sh + kRowF - 1) / kRowF, nbx = (Sw + kStep - 1) / kStep;
    const int per_tile = nby * nbx;
    const int total = per_tile * b.ntiles;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        for (int s = 0; s < kPipeStages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], kPipeWarps);
        }
        fence_mbar_init();
    }
    __syncthreads();
    // box nn of this CTA = V^T plane nn % 5 of its item nn / 5, loaded into stage nn % 2 (use number nn / 2)
    auto issue = [&](int nn) {
        const int item = blockIdx.x + (nn / 5) * gridDim.x, c = nn % 5;
        if (item >= total) return;
        const int s = nn % kPipeStages, use = nn / kPipeStages;
        if (use > 0) mbar_wait_guarded(&empty[s], (use - 1) & 1);
        const int slot = item / per_tile, rem = item - slot * per_tile;
        const int bx = rem / nby, by = rem - bx * nby;
        mbar_expect_tx(&full[s], rows * kRowF * sizeof(float));
        tma_load_3d(smem + (size_t)s * rows * kRowF, &mapVT, by * kRowF, bx * kStep - m, slot * kSlotPlanes + 15 + c, &full[s]);
    };
    if (threadIdx.x == 0) {
        issue(0);
        issue(1);
    }
    const u64 nz = *reinterpret_cast<const u64*>(&cst.negzero2);
    int n = 0;
    for (int item = blockIdx.x; item < total; item += gridDim.x) {
        const int slot = item / per_tile, rem = item - slot * per_tile;
        const int bx = rem / nby, by = rem - bx * nby;
        const int y0 = by * kRowF, x0 = bx * kStep;
        const bool active = x0 + warp * kR < Sw;
        u64 acc[5][kR];
#pragma unroll
        for (int c = 0; c < 5; ++c, ++n) {
            if (threadIdx.x == 0 && n > 0) issue(n + 1);
            __syncwarp();
            const int s = n % kPipeStages;
            mbar_wait_guarded(&full[s], (n / kPipeStages) & 1);
            if (active) conv_strip<FUSED>(smem + (size_t)s * rows * kRowF, lane, warp, rows, x0 - m, Sw, m, cst.k2, nz, acc[c]);
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[s]);
        }
        if (active) solve_store_strip(b, slot, y0, x0 + warp * kR, lane, last_iter, flow_out, acc[0], acc[1], acc[2], acc[3], acc[4]);
    }
}

// Variant of fb_blur_h_pipe_kernel with the plane loop NOT unrolled: one copy of the convolution code (the unrolled
// kernel is ~150 KB of SASS, and its warps -- no longer held together by barriers -- run in different parts of it).
template <bool FUSED>
__global__ void __launch_bounds__(kPipeThreads, 2) fb_blur_h_pipe_rolled_kernel(const __grid_constant__ CUtensorMap mapVT, FbBatch b,
                                                                                const __grid_constant__ FbConsts cst, int last_iter,
                                                                                float2* __restrict__ flow_out) {
    extern __shared__ __align__(128) float smem[];
    __shared__ __align__(8) uint64_t full[kPipeStages], empty[kPipeStages];
    const TileGeom& g = b.g;
    const int Sh = g.Sh, Sw = g.Sw, m = cst.m;
    const int rows = kStep + 2 * m;
    const int nby = (Sh + kRowF - 1) / kRowF, nbx = (Sw + kStep - 1) / kStep;
    const int per_tile = nby * nbx;
    const int total = per_tile * b.ntiles;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        for (int s = 0; s < kPipeStages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], kPipeWarps);
        }
        fence_mbar_init();
    }
    __syncthreads();
    auto issue = [&](int nn) {
        const int item = blockIdx.x + (nn / 5) * gridDim.x, c = nn % 5;
        if (item >= total) return;
        const int s = nn % kPipeStages, use = nn / kPipeStages;
        if (use > 0) mbar_wait_guarded(&empty[s], (use - 1) & 1);
        const int slot = item / per_tile, rem = item - slot * per_tile;
        const int bx = rem / nby, by = rem - bx * nby;
        mbar_expect_tx(&full[s], rows * kRowF * sizeof(float));
        tma_load_3d(smem + (size_t)s * rows * kRowF, &mapVT, by * kRowF, bx * kStep - m, slot * kSlotPlanes + 15 + c, &full[s]);
    };
    if (threadIdx.x == 0) {
        issue(0);
        issue(1);
    }
    const u64 nz = *reinterpret_cast<const u64*>(&cst.negzero2);
    int n = 0;
    for (int item = blockIdx.x; item < total; item += gridDim.x) {
        const int slot = item / per_tile, rem = item - slot * per_tile;
        const int bx = rem / nby, by = rem - bx * nby;
        const int y0 = by * kRowF, x0 = bx * kStep;
        const bool active = x0 + warp * kR < Sw;
        u64 a0[kR], a1[kR], a2[kR], a3[kR], cur[kR];
#pragma unroll 1
        for (int c = 0; c < 5; ++c, ++n) {
            if (threadIdx.x == 0 && n > 0) issue(n + 1);
            __syncwarp();
            const int s = n % kPipeStages;
            mbar_wait_guarded(&full[s], (n / kPipeStages) & 1);
            if (active) conv_strip<FUSED>(smem + (size_t)s * rows * kRowF, lane, warp, rows, x0 - m, Sw, m, cst.k2, nz, cur);
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[s]);
            if (c == 0) {
#pragma unroll
                for (int j = 0; j < kR; ++j) a0[j] = cur[j];
            } else if (c == 1) {
#pragma unroll
                for (int j = 0; j < kR; ++j) a1[j] = cur[j];
            } else if (c == 2) {
#pragma unroll
                for (int j = 0; j < kR; ++j) a2[j] = cur[j];
            } else if (c == 3) {
#pragma unroll
                for (int j = 0; j < kR; ++j) a3[j] = cur[j];
            }
        }
        if (active) solve_store_strip(b, slot, y0, x0 + warp * kR, lane, last_iter, flow_out, a0, a1, a2, a3, cur);
    }
}

// ---- experimental H-pass variant 4: four outputs per thread instead of eight ----
// The 8-output kernel needs 128 registers (five 8-deep packed accumulators + two windows), which caps the H pass at
// two CTAs = 16 warps per SM.  With K = 4 the accumulators and windows take 56 registers, three CTAs of 64 y x 32 x
// fit (boxes of 32 + 2m rows), at the price of twice the shared-memory reads per FP instruction and 1.6x the fill.
template <int K, bool FUSED>
__device__ __forceinline__ void convKx2(const u64* __restrict__ centre, int m, const float2* __restrict__ k2, u64 nz, u64 (&acc)[K]) {
    static_assert(K == 4 || K == 8, "window rotation uses & (K - 1)");
    u64 wp[K], wm[K];
    const u64 k0 = *reinterpret_cast<const u64*>(&k2[0]);
#pragma unroll
    for (int j = 0; j < K; ++j) {
        u64 c = centre[j * kRowU];
        wp[j] = c;
        wm[j] = c;
        acc[j] = mul2(c, k0, nz);
    }
    const u64* pp = centre + K * kRowU;   // row of in[K - 1 + i] for i = 1
    const u64* pm = centre - kRowU;       // row of in[-i] for i = 1
    int i = 1;
    // before step i = K g + 1 + s: in[j + i - 1] lives in wp[(j + i - 1) & (K-1)], in[j - i + 1] in wm[(j - i + 1) & (K-1)]
#pragma unroll 1
    for (; i + K - 1 <= m; i += K) {
#pragma unroll
        for (int s = 0; s < K; ++s) {
            wp[s] = pp[s * kRowU];
            wm[(K - 1 - s) & (K - 1)] = pm[-s * kRowU];
            const u64 kk = *reinterpret_cast<const u64*>(&k2[i + s]);
#pragma unroll
            for (int j = 0; j < K; ++j) {
                const u64 pr = add2(wp[(j + s + 1) & (K - 1)], wm[(j + 8 * K - 1 - s) & (K - 1)]);
                acc[j] = FUSED ? fma2(pr, kk, acc[j]) : add2(acc[j], mul2(pr, kk, nz));
            }
        }
        pp += K * kRowU;
        pm -= K * kRowU;
    }
#pragma unroll
    for (int s = 0; s < K - 1; ++s) {
        if (i + s <= m) {  // warp-uniform tail (m mod K steps)
            wp[s] = pp[s * kRowU];
            wm[(K - 1 - s) & (K - 1)] = pm[-s * kRowU];
            const u64 kk = *reinterpret_cast<const u64*>(&k2[i + s]);
#pragma unroll
            for (int j = 0; j < K; ++j) {
                const u64 pr = add2(wp[(j + s + 1) & (K - 1)], wm[(j + 8 * K - 1 - s) & (K - 1)]);
                acc[j] = FUSED ? fma2(pr, kk, acc[j]) : add2(acc[j], mul2(pr, kk, nz));
            }
        }
    }
}

constexpr int kStep4 = 32;   // x outputs per CTA of the 4-output H kernel (8 warps x 4)

template <bool FUSED>
__global__ void __launch_bounds__(256, 3) fb_blur_h4_kernel(const __grid_constant__ CUtensorMap mapVT, FbBatch b,
                                                             const __grid_constant__ FbConsts cst, int last_iter,
                                                             float2* __restrict__ flow_out) {
    constexpr int K = 4;
    extern __shared__ __align__(128) float smem[];
    __shared__ __align__(8) uint64_t bars[2];
    const TileGeom& g = b.g;
    const int Sh = g.Sh, Sw = g.Sw, m = cst.m;
    const int slot = blockIdx.z;
    const int y0 = blockIdx.x * kRowF, x0 = blockIdx.y * kStep4;
    const int rows = kStep4 + 2 * m;
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
    u64 acc[5][K];
    const int xb = x0 + warp * K;
    const bool active = xb < Sw;
#pragma unroll
    for (int c = 0; c < 5; ++c) {
        mbar_wait(&bars[c & 1], (c >> 1) & 1);
        replicate_edges(buf[c & 1], rows, x0 - m, Sw);
        if (active) convKx2<K, FUSED>(reinterpret_cast<const u64*>(buf[c & 1]) + (warp * K + m) * kRowU + lane, m, cst.k2, nz, acc[c]);
        if (c + 2 < 5) {
            __syncthreads();  // buffer c&1 is free again
            if (threadIdx.x == 0) {
                mbar_expect_tx(&bars[c & 1], box_bytes);
                tma_load_3d(buf[c & 1], &mapVT, y0, x0 - m, slot * kSlotPlanes + 15 + c + 2, &bars[c & 1]);
            }
        }
    }
    if (!active) return;
    // solve + store from registers: rows y0 + 2 lane (+1), columns xb .. xb + 3 (32 contiguous bytes per row)
    const int tile = b.tile0 + slot;
    const int ti = tile / g.nx, tj = tile - ti * g.nx;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        const int y = y0 + 2 * lane + half;
        if (y >= Sh) continue;
        float2 f[K];
#pragma unroll
        for (int j = 0; j < K; ++j) {
            const float2 G11 = unpack2(acc[0][j]), G12 = unpack2(acc[1][j]), G22 = unpack2(acc[2][j]);
            const float2 H1 = unpack2(acc[3][j]), H2 = unpack2(acc[4][j]);
            f[j] = half ? solve_flow(G11.y, G12.y, G22.y, H1.y, H2.y) : solve_flow(G11.x, G12.x, G22.x, H1.x, H2.x);
        }
        if (last_iter) {
            const int cy = y - g.ov, gy = ti * g.Th + cy;
            if ((unsigned)cy < (unsigned)g.Th && gy < g.h) {
#pragma unroll
                for (int j = 0; j < K; ++j) {
                    const int cxx = xb + j - g.ov, gx = tj * g.Tw + cxx;
                    if (xb + j < Sw && (unsigned)cxx < (unsigned)g.Tw && gx < g.w) flow_out[(size_t)gy * g.w + gx] = f[j];
                }
            }
        } else {   // pitch Sp is a multiple of 32: the 32-byte vector stays inside the row
            float4* __restrict__ dst = reinterpret_cast<float4*>(reinterpret_cast<float2*>(slot_plane(b, slot, 4, 0)) + (size_t)y * b.Sp + xb);
            dst[0] = make_float4(f[0].x, f[0].y, f[1].x, f[1].y);
            dst[1] = make_float4(f[2].x, f[2].y, f[3].x, f[3].y);
        }
    }
}

}  // namespace ma
