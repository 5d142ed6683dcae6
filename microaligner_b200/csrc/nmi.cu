// K7: normalized mutual information of two u8 label images over raster-contiguous chunks.
//
//   ma_nmi_chunks   mi_tiled / normalized_mutual_info_score
//                   (reference shared_modules/similarity_scoring.py:27-50; scikit-learn
//                    metrics/cluster/_supervised.py: arithmetic-mean normaliser, natural log)
//
// One thread-block CLUSTER of four CTAs per chunk (nmi_chunk_kernel).  The 256 x 256 joint histogram (256 KB as u32)
// does not fit one SM's shared memory, so it is split into four 64-row slabs, one per CTA of the cluster (64 KB each).
// Every CTA scans the whole chunk -- 16 pixels of both images per thread from two 128-bit loads -- and keeps the pixels
// whose `a` value falls into its slab; equal consecutive (a, b) pairs are merged in registers, and each run is one
// shared-memory atomic.  No global histogram, no L2 atomics, no memset, one launch for both comparisons of the gate.
// After the scan each CTA reduces its slab: row sums, its share of the column sums (exchanged through distributed
// shared memory), and its part of
//   MI = sum_{J>0} J/n (ln J - ln n) + J/n (-ln(a_i b_j) + 2 ln n),  H(a), H(b)
// in f64; CTA 0 adds the four parts in rank order (deterministic) and writes NMI = MI / mean(H_a, H_b) with sklearn's
// special cases (both labelings constant -> 1, MI == 0 -> 0).  One double per chunk leaves the kernel; the mean over
// chunks and the `after > before` decision are taken by the host from two doubles.
//
// Measured on B200, 144 chunks of 10^6 px, ms per call (profiles/r02_ab_variants.log, profiles/r02_nmi_cluster.log):
// histogram in an L2-resident global scratch + separate entropy kernel 1.67 (warp-aggregated atomics) / 1.22 (run-length
// merging) -- bound by the L2 atomic unit serialising the hot bins of smooth images; this kernel with every CTA scanning
// a quarter of the chunk and adding to REMOTE slabs (red.shared::cluster) 1.77 -- remote shared-memory atomics are slow;
// scanning everything and adding locally 0.74.  The scan is latency-bound: 512 threads beat 256 (0.74 vs 1.04), more
// loads in flight per thread or a SIMD test that skips vectors without a pixel of the slab do not pay (0.80 - 0.97).
#include <atomic>
#include <cooperative_groups.h>
#include "common.cuh"

namespace ma {

namespace cg = cooperative_groups;

constexpr int kNmiCluster = 4;                     // CTAs per chunk
constexpr int kSlabRows = 256 / kNmiCluster;       // rows of the joint histogram per CTA
constexpr int kNmiThreads = 512;                   // 3 CTAs x 16 warps per SM: the scan is latency-bound, warps are what helps

__device__ __forceinline__ double block_sum(double v, double* sh) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0;
    for (int i = 0; i < kNmiThreads / 32; ++i) t += sh[i];
    return t;
}



__global__ void __cluster_dims__(kNmiCluster, 1, 1) __launch_bounds__(kNmiThreads)
nmi_chunk_kernel(const uint8_t* __restrict__ a, const uint8_t* __restrict__ b0, const uint8_t* __restrict__ b1, size_t n,
                 size_t chunk, size_t chunk0, double* __restrict__ scores0, double* __restrict__ scores1) {
    // blockIdx.y selects the image compared with `a`: the gate scores (a, b0) and (a, b1) in one launch
    const uint8_t* __restrict__ b = blockIdx.y ? b1 : b0;
    double* __restrict__ scores = blockIdx.y ? scores1 : scores0;
    extern __shared__ unsigned hist[];              // [kSlabRows][256]: rows rank*64 .. rank*64+63 of the joint histogram
    __shared__ unsigned pi[kSlabRows];              // row sums of my slab
    __shared__ unsigned pj_part[256];               // column sums of my slab
    __shared__ unsigned pj[256];                    // column sums of the whole histogram
    __shared__ double red[kNmiThreads / 32];
    __shared__ double part[4];                      // my slab's {MI, H_a, #classes of a, unused}
    cg::cluster_group cl = cg::this_cluster();
    const unsigned rank = cl.block_rank();
    const size_t c = chunk0 + blockIdx.x / kNmiCluster;
    const size_t beg = c * chunk;
    const size_t end = beg + chunk < n ? beg + chunk : n;
    const size_t len = end - beg;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    for (int i = t; i < kSlabRows * 256; i += kNmiThreads) hist[i] = 0;
    if (t < 256) pj_part[t] = 0;
    __syncthreads();
    {
        const uint8_t* pa = a + beg;
        const uint8_t* pb = b + beg;
        auto add = [&](unsigned av, unsigned bv, unsigned cnt) { atomicAdd(&hist[((av % kSlabRows) << 8) | bv], cnt); };
        // [0, head) scalar up to the first 16-byte boundary, [head, tail0) as 16-pixel vectors, [tail0, len) scalar;
        // images whose chunk starts are not equally aligned are processed pixel by pixel
        const bool vec = ((reinterpret_cast<uintptr_t>(pa) ^ reinterpret_cast<uintptr_t>(pb)) & 15) == 0;
        size_t head = vec ? ((16 - (reinterpret_cast<uintptr_t>(pa) & 15)) & 15) : len;
        if (head > len) head = len;
        const size_t nvec = (len - head) / 16;
        const size_t tail0 = head + nvec * 16;
        for (size_t i = t; i < head + (len - tail0); i += kNmiThreads) {
            const size_t p = i < head ? i : tail0 + (i - head);
            const unsigned av = pa[p];
            if (av / kSlabRows == rank) add(av, pb[p], 1u);
        }
        const uint4* va = reinterpret_cast<const uint4*>(pa + head);
        const uint4* vb = reinterpret_cast<const uint4*>(pb + head);
        for (size_t v = t; v < nvec; v += kNmiThreads) {
            const uint4 A = __ldg(va + v), B = __ldg(vb + v);
            const unsigned aw[4] = {A.x, A.y, A.z, A.w}, bw[4] = {B.x, B.y, B.z, B.w};
            unsigned run_a = 0, run_b = 0, cnt = 0;
#pragma unroll
            for (int q = 0; q < 16; ++q) {
                const int sh = 8 * (q & 3);
                const unsigned av = (aw[q >> 2] >> sh) & 255u, bv = (bw[q >> 2] >> sh) & 255u;
                const bool o = av / kSlabRows == rank;
                if (o && cnt != 0 && av == run_a && bv == run_b) {
                    ++cnt;
                } else {
                    if (cnt != 0) add(run_a, run_b, cnt);
                    run_a = av;
                    run_b = bv;
                    cnt = o ? 1u : 0u;
                }
            }
            if (cnt != 0) add(run_a, run_b, cnt);
        }
    }
    __syncthreads();                                // my slab is final
    // marginals of my slab: one warp per row (coalesced, conflict-free), column sums accumulated per lane
    unsigned colsum[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int r = warp; r < kSlabRows; r += kNmiThreads / 32) {
        unsigned rs = 0;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const unsigned v = hist[r * 256 + q * 32 + lane];
            rs += v;
            colsum[q] += v;
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) rs += __shfl_xor_sync(0xffffffffu, rs, o);
        if (lane == 0) pi[r] = rs;
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) atomicAdd(&pj_part[q * 32 + lane], colsum[q]);       // integer: order-independent
    cl.sync();
    if (t < 256) {
        unsigned s = 0;
        for (int k = 0; k < kNmiCluster; ++k) s += cl.map_shared_rank(pj_part, k)[t];
        pj[t] = s;
    }
    __syncthreads();
    const double N = (double)len, logN = log(N);
    // my rows' share of H(a) and of the class count of a; H(b) and its class count are computed redundantly by every CTA
    double ha = 0, hb = 0;
    int ca = 0, cb = 0;
    if (t < kSlabRows && pi[t] > 0) { ca = 1; ha = (pi[t] / N) * (log((double)pi[t]) - logN); }
    if (t < 256 && pj[t] > 0) { cb = 1; hb = (pj[t] / N) * (log((double)pj[t]) - logN); }
    const double Ha_part = -block_sum(ha, red);
    const double Hb = -block_sum(hb, red);
    const double na_part = block_sum((double)ca, red);
    const int nb = (int)block_sum((double)cb, red);
    double mi = 0;
    for (int r = warp; r < kSlabRows; r += kNmiThreads / 32) {
        const unsigned a_r = pi[r];
        if (a_r == 0) continue;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const unsigned v = hist[r * 256 + q * 32 + lane];
            if (v) {
                const double cn = v / N;
                const double outer = (double)((long long)a_r * (long long)pj[q * 32 + lane]);
                const double log_outer = -log(outer) + logN + logN;
                double term = cn * (log((double)v) - logN) + cn * log_outer;
                if (fabs(term) < 2.220446049250313e-16) term = 0.0;
                mi += term;
            }
        }
    }
    mi = block_sum(mi, red);
    if (t == 0) {
        part[0] = mi;
        part[1] = Ha_part;
        part[2] = na_part;
    }
    cl.sync();
    if (rank == 0 && t == 0) {
        double MI = 0, Ha = 0;
        int na = 0;
        for (int k = 0; k < kNmiCluster; ++k) {      // fixed order: the score is reproducible bit for bit
            const double* p = cl.map_shared_rank(part, k);
            MI += p[0];
            Ha += p[1];
            na += (int)p[2];
        }
        double score;
        if (na <= 1 && nb <= 1) score = 1.0;
        else if (na == 1 || nb == 1) score = 0.0;
        else {
            if (MI < 0) MI = 0;
            score = MI == 0 ? 0.0 : MI / (0.5 * (Ha + Hb));
        }
        scores[c] = score;
    }
    cl.sync();                                      // nobody leaves while CTA 0 still reads its shared memory
}

}  // namespace ma

using namespace ma;

extern "C" size_t ma_nmi_workspace_bytes(size_t n, size_t chunk) {
    (void)n; (void)chunk;
    return 0;       // the joint histograms live in (distributed) shared memory; kept in the ABI for callers that size a scratch
}

extern "C" int ma_nmi_chunk_range2(const uint8_t* a, const uint8_t* b0, const uint8_t* b1, size_t n, size_t chunk,
                                   size_t chunk_begin, size_t chunk_end, double* scores0, double* scores1, void* stream) {
    const uint8_t* b = b0;
    double* scores_out = scores0;
    if (b1 && !scores1) return invalid("ma_nmi_chunks: bad argument");
    if (!a || !b || !scores_out || n == 0 || chunk == 0) return invalid("ma_nmi_chunks: bad argument");
    if (chunk > 0xffffffffull) return invalid("ma_nmi_chunks: chunk must fit 32-bit counters");
    cudaStream_t s = (cudaStream_t)stream;
    size_t nchunks = (n + chunk - 1) / chunk;
    if (chunk_begin > chunk_end || chunk_end > nchunks) return invalid("ma_nmi_chunks: bad chunk range");
    if (chunk_begin == chunk_end) return MA_OK;
    const size_t smem = (size_t)kSlabRows * 256 * sizeof(unsigned);
    static std::atomic<bool> attr_set[64];
    int dev_id = 0;
    cudaGetDevice(&dev_id);
    if (dev_id < 0 || dev_id >= 64 || !attr_set[dev_id]) {
        MA_CUDA_CHECK(cudaFuncSetAttribute(nmi_chunk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        if (dev_id >= 0 && dev_id < 64) attr_set[dev_id] = true;
    }
    for (size_t c0 = chunk_begin; c0 < chunk_end; c0 += 16384) {       // grid.x limit is far away; keep launches bounded
        const int g = (int)std::min<size_t>(16384, chunk_end - c0);
        KernelScope ks(K_NMI_HIST, s, (b1 ? 2.0 : 1.0) * (double)std::min<size_t>(n - c0 * chunk, (size_t)g * chunk));
        nmi_chunk_kernel<<<dim3(g * kNmiCluster, b1 ? 2 : 1), kNmiThreads, smem, s>>>(a, b0, b1, n, chunk, c0, scores0, scores1);
    }
    MA_LAUNCH_CHECK("nmi_chunk_kernel");
    return MA_OK;
}

extern "C" int ma_nmi_chunk_range(const uint8_t* a, const uint8_t* b, size_t n, size_t chunk, size_t chunk_begin, size_t chunk_end,
                                  double* scores_out, void* workspace, void* stream) {
    (void)workspace;
    return ma_nmi_chunk_range2(a, b, nullptr, n, chunk, chunk_begin, chunk_end, scores_out, nullptr, stream);
}

extern "C" int ma_nmi_chunks(const uint8_t* a, const uint8_t* b, size_t n, size_t chunk,
                             double* scores_out, void* workspace, void* stream) {
    if (chunk == 0) return invalid("ma_nmi_chunks: bad argument");
    return ma_nmi_chunk_range(a, b, n, chunk, 0, (n + chunk - 1) / chunk, scores_out, workspace, stream);
}
