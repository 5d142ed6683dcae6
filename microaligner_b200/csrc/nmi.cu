// K7: normalized mutual information of two u8 label images over raster-contiguous chunks.
//
//   ma_nmi_chunks   mi_tiled / normalized_mutual_info_score
//                   (reference shared_modules/similarity_scoring.py:27-50; scikit-learn
//                    metrics/cluster/_supervised.py: arithmetic-mean normaliser, natural log)
//
// Per chunk: joint histogram J (256 x 256, u32, L2-resident scratch, warp-aggregated atomics),
// then one CTA reduces it to  MI = sum_{J>0} J/n (ln J - ln n) + J/n (-ln(a_i b_j) + 2 ln n),
// H(a), H(b) in f64 and writes NMI = MI / mean(H_a, H_b) with sklearn's special cases
// (both labelings constant -> 1, MI == 0 -> 0).  One double per chunk leaves the kernel; the mean
// over chunks and the `after > before` decision are taken by the host from two doubles.
#include "common.cuh"

namespace ma {

constexpr int kNmiSlots = 64;                       // chunks processed per group
constexpr size_t kHistBytes = 65536 * sizeof(unsigned);

// 16 pixels per thread from two 128-bit loads, equal consecutive (a, b) pairs merged in registers before the atomic
// -- the DoG images the gate compares are smooth, so runs are long.  (Measured against one byte load per image and a
// warp-wide match_any per pixel: 118 vs 86 Gpx/s, profiles/r02_ab_variants.log.)
__global__ void __launch_bounds__(256) nmi_hist_kernel(const uint8_t* __restrict__ a, const uint8_t* __restrict__ b,
                                                           size_t n, size_t chunk, size_t chunk0, unsigned* __restrict__ hist) {
    const int slot = blockIdx.y;
    const size_t beg = (chunk0 + slot) * chunk;
    const size_t end = beg + chunk < n ? beg + chunk : n;
    const size_t len = end - beg;
    unsigned* H = hist + (size_t)slot * 65536;
    const uint8_t* pa = a + beg;
    const uint8_t* pb = b + beg;
    // [0, head) scalar up to the first 16-byte boundary, [head, tail0) as 16-pixel vectors, [tail0, len) scalar;
    // images whose chunk starts are not equally aligned are processed pixel by pixel
    const bool vec = ((reinterpret_cast<uintptr_t>(pa) ^ reinterpret_cast<uintptr_t>(pb)) & 15) == 0;
    size_t head = vec ? ((16 - (reinterpret_cast<uintptr_t>(pa) & 15)) & 15) : len;
    if (head > len) head = len;
    const size_t nvec = (len - head) / 16;
    const size_t tail0 = head + nvec * 16;
    const size_t gtid = (size_t)blockIdx.x * 256 + threadIdx.x, gsz = (size_t)gridDim.x * 256;
    for (size_t i = gtid; i < head; i += gsz) atomicAdd(&H[((unsigned)pa[i] << 8) | pb[i]], 1u);
    for (size_t i = tail0 + gtid; i < len; i += gsz) atomicAdd(&H[((unsigned)pa[i] << 8) | pb[i]], 1u);
    const uint4* va = reinterpret_cast<const uint4*>(pa + head);
    const uint4* vb = reinterpret_cast<const uint4*>(pb + head);
    for (size_t v = gtid; v < nvec; v += gsz) {
        const uint4 A = __ldg(va + v), B = __ldg(vb + v);
        const unsigned aw[4] = {A.x, A.y, A.z, A.w}, bw[4] = {B.x, B.y, B.z, B.w};
        unsigned run = 0, cnt = 0;
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            const unsigned key = (((aw[q >> 2] >> (8 * (q & 3))) & 255u) << 8) | ((bw[q >> 2] >> (8 * (q & 3))) & 255u);
            if (cnt != 0 && key == run) {
                ++cnt;
            } else {
                if (cnt != 0) atomicAdd(&H[run], cnt);
                run = key;
                cnt = 1;
            }
        }
        atomicAdd(&H[run], cnt);
    }
}

__device__ __forceinline__ double block_sum(double v, double* sh) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += sh[i];
    return t;
}

__global__ void __launch_bounds__(256) nmi_entropy_kernel(const unsigned* __restrict__ hist, size_t n, size_t chunk, size_t chunk0,
                                                          double* __restrict__ scores) {
    __shared__ unsigned pi[256], pj[256];
    __shared__ double red[8];
    int slot = blockIdx.x;
    const unsigned* H = hist + (size_t)slot * 65536;
    size_t beg = (chunk0 + slot) * chunk;
    size_t end = beg + chunk < n ? beg + chunk : n;
    double N = (double)(end - beg);
    int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    pj[t] = 0;
    __syncthreads();
    unsigned colsum[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int r = warp; r < 256; r += 8) {
        unsigned rs = 0;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            unsigned v = H[r * 256 + q * 32 + lane];
            rs += v;
            colsum[q] += v;
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) rs += __shfl_xor_sync(0xffffffffu, rs, o);
        if (lane == 0) pi[r] = rs;
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) atomicAdd(&pj[q * 32 + lane], colsum[q]);
    __syncthreads();
    double logN = log(N);
    // entropies and class counts
    double ha = 0, hb = 0;
    int ca = pi[t] > 0, cb = pj[t] > 0;
    if (ca) ha = (pi[t] / N) * (log((double)pi[t]) - logN);
    if (cb) hb = (pj[t] / N) * (log((double)pj[t]) - logN);
    double Ha = -block_sum(ha, red);
    double Hb = -block_sum(hb, red);
    int na = (int)block_sum((double)ca, red), nb = (int)block_sum((double)cb, red);
    double mi = 0;
    for (int r = warp; r < 256; r += 8) {
        unsigned a_r = pi[r];
        if (a_r == 0) continue;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            unsigned v = H[r * 256 + q * 32 + lane];
            if (v) {
                double cn = v / N;
                double outer = (double)((long long)a_r * (long long)pj[q * 32 + lane]);
                double log_outer = -log(outer) + logN + logN;
                double term = cn * (log((double)v) - logN) + cn * log_outer;
                if (fabs(term) < 2.220446049250313e-16) term = 0.0;
                mi += term;
            }
        }
    }
    mi = block_sum(mi, red);
    if (t == 0) {
        double score;
        if (na <= 1 && nb <= 1) score = 1.0;
        else if (na == 1 || nb == 1) score = 0.0;
        else {
            if (mi < 0) mi = 0;
            if (na == 1) Ha = 0;
            if (nb == 1) Hb = 0;
            score = mi == 0 ? 0.0 : mi / (0.5 * (Ha + Hb));
        }
        scores[chunk0 + slot] = score;
    }
}

}  // namespace ma

using namespace ma;

extern "C" size_t ma_nmi_workspace_bytes(size_t n, size_t chunk) {
    if (n == 0 || chunk == 0) return 0;
    size_t nchunks = (n + chunk - 1) / chunk;
    return std::min<size_t>(nchunks, kNmiSlots) * kHistBytes;
}

extern "C" int ma_nmi_chunk_range(const uint8_t* a, const uint8_t* b, size_t n, size_t chunk, size_t chunk_begin, size_t chunk_end,
                                  double* scores_out, void* workspace, void* stream) {
    if (!a || !b || !scores_out || !workspace || n == 0 || chunk == 0) return invalid("ma_nmi_chunks: bad argument");
    if (chunk > 0xffffffffull) return invalid("ma_nmi_chunks: chunk must fit 32-bit counters");
    cudaStream_t s = (cudaStream_t)stream;
    size_t nchunks = (n + chunk - 1) / chunk;
    if (chunk_begin > chunk_end || chunk_end > nchunks) return invalid("ma_nmi_chunks: bad chunk range");
    unsigned* hist = (unsigned*)workspace;
    for (size_t c0 = chunk_begin; c0 < chunk_end; c0 += kNmiSlots) {
        int g = (int)std::min<size_t>(kNmiSlots, chunk_end - c0);
        MA_CUDA_CHECK(cudaMemsetAsync(hist, 0, (size_t)g * kHistBytes, s));
        int bpc = (int)std::max<size_t>(1, std::min<size_t>((chunk + 256 * 16 - 1) / (256 * 16), (size_t)(148 * 8 + g - 1) / g));
        { KernelScope ks(K_NMI_HIST, s, (double)std::min<size_t>(n - c0 * chunk, (size_t)g * chunk));
        nmi_hist_kernel<<<dim3(bpc, g), 256, 0, s>>>(a, b, n, chunk, c0, hist); }
        { KernelScope ks(K_NMI_ENTROPY, s, (double)g); nmi_entropy_kernel<<<g, 256, 0, s>>>(hist, n, chunk, c0, scores_out); }
        MA_LAUNCH_CHECK("nmi kernels");
    }
    return MA_OK;
}

extern "C" int ma_nmi_chunks(const uint8_t* a, const uint8_t* b, size_t n, size_t chunk,
                             double* scores_out, void* workspace, void* stream) {
    if (chunk == 0) return invalid("ma_nmi_chunks: bad argument");
    return ma_nmi_chunk_range(a, b, n, chunk, 0, (n + chunk - 1) / chunk, scores_out, workspace, stream);
}
