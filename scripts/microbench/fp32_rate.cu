// Microbenchmark: issue rate of scalar FADD/FMUL/FFMA vs packed add/mul/fma.f32x2 on sm_100a.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o fp32_rate fp32_rate.cu && ./fp32_rate
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 4096
#define NACC 16

__global__ void k_scalar_addmul(float* out, float a, float b) {
    float acc[NACC];
#pragma unroll
    for (int j = 0; j < NACC; ++j) acc[j] = threadIdx.x + j;
    for (int i = 0; i < ITERS; ++i) {
#pragma unroll
        for (int j = 0; j < NACC; ++j) {
            float t = __fadd_rn(acc[j], a);
            acc[j] = __fadd_rn(acc[j], __fmul_rn(t, b));   // 3 instr: FADD, FMUL, FADD
        }
    }
    float s = 0;
#pragma unroll
    for (int j = 0; j < NACC; ++j) s += acc[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_scalar_fma(float* out, float a, float b) {
    float acc[NACC];
#pragma unroll
    for (int j = 0; j < NACC; ++j) acc[j] = threadIdx.x + j;
    for (int i = 0; i < ITERS; ++i) {
#pragma unroll
        for (int j = 0; j < NACC; ++j) {
            acc[j] = __fmaf_rn(acc[j], a, b);
            acc[j] = __fmaf_rn(acc[j], b, a);
            acc[j] = __fmaf_rn(acc[j], a, b);
        }
    }
    float s = 0;
#pragma unroll
    for (int j = 0; j < NACC; ++j) s += acc[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__device__ __forceinline__ unsigned long long pk(float x, float y) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(x), "f"(y));
    return r;
}
__device__ __forceinline__ unsigned long long add2(unsigned long long a, unsigned long long b) {
    unsigned long long r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ unsigned long long mul2(unsigned long long a, unsigned long long b) {
    unsigned long long r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}

// NOTE: ptxas contracts mul.rn.f32x2 + add.rn.f32x2 into ONE FFMA2, so this kernel really issues FADD2 + FFMA2
// (2 packed instructions for 3 "operations") -- kept to document that.
__global__ void k_packed_addmul(float* out, float a, float b) {
    unsigned long long acc[NACC / 2], A = pk(a, a), B = pk(b, b);
#pragma unroll
    for (int j = 0; j < NACC / 2; ++j) acc[j] = pk(threadIdx.x + j, threadIdx.x - j);
    for (int i = 0; i < ITERS; ++i) {
#pragma unroll
        for (int j = 0; j < NACC / 2; ++j) {
            unsigned long long t = add2(acc[j], A);
            acc[j] = add2(acc[j], mul2(t, B));
        }
    }
    unsigned long long s = acc[0];
#pragma unroll
    for (int j = 1; j < NACC / 2; ++j) s = add2(s, acc[j]);
    float x, y;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(s));
    out[blockIdx.x * blockDim.x + threadIdx.x] = x + y;
}

// separately rounded add, mul, add with the multiply written as fma(t, B, -0) whose -0 is a runtime value:
// three packed instructions (FADD2, FFMA2, FADD2) -- the sequence the bit-exact blur kernels use
__global__ void k_packed_addmul_unfused(float* out, float a, float nzf) {
    unsigned long long acc[NACC / 2], A = pk(a, a), B = pk(0.9999f, 0.9999f), NZ = pk(nzf, nzf);
#pragma unroll
    for (int j = 0; j < NACC / 2; ++j) acc[j] = pk(threadIdx.x + j, threadIdx.x - j);
    for (int i = 0; i < ITERS; ++i) {
#pragma unroll
        for (int j = 0; j < NACC / 2; ++j) {
            unsigned long long t = add2(acc[j], A);
            acc[j] = add2(acc[j], fma2(t, B, NZ));
        }
    }
    unsigned long long s = acc[0];
#pragma unroll
    for (int j = 1; j < NACC / 2; ++j) s = add2(s, acc[j]);
    float x, y;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(s));
    out[blockIdx.x * blockDim.x + threadIdx.x] = x + y;
}

__global__ void k_packed_fma(float* out, float a, float b) {
    unsigned long long acc[NACC / 2], A = pk(a, a), B = pk(b, b);
#pragma unroll
    for (int j = 0; j < NACC / 2; ++j) acc[j] = pk(threadIdx.x + j, threadIdx.x - j);
    for (int i = 0; i < ITERS; ++i) {
#pragma unroll
        for (int j = 0; j < NACC / 2; ++j) {
            acc[j] = fma2(acc[j], A, B);
            acc[j] = fma2(acc[j], B, A);
            acc[j] = fma2(acc[j], A, B);
        }
    }
    unsigned long long s = acc[0];
#pragma unroll
    for (int j = 1; j < NACC / 2; ++j) s = add2(s, acc[j]);
    float x, y;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(s));
    out[blockIdx.x * blockDim.x + threadIdx.x] = x + y;
}

template <typename K>
void run(const char* name, K kern, float* out, double lane_ops_per_thread, float arg2 = 0.9999f) {
    int blocks = 148 * 8, threads = 256;
    kern<<<blocks, threads>>>(out, 1.0001f, arg2);
    cudaDeviceSynchronize();
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    for (int r = 0; r < 5; ++r) kern<<<blocks, threads>>>(out, 1.0001f, arg2);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    double ops = 5.0 * blocks * threads * lane_ops_per_thread;
    printf("%-58s %8.3f ms  %8.2f T lane-op/s  (%s)\n", name, ms / 5, ops / (ms * 1e-3) / 1e12, cudaGetErrorString(cudaGetLastError()));
}

int main() {
    float* out;
    cudaMalloc(&out, 148 * 8 * 256 * sizeof(float));
    double per_thread = (double)ITERS * NACC * 3;  // float results produced per thread
    run("scalar add,mul,add", k_scalar_addmul, out, per_thread);
    run("scalar fma x3", k_scalar_fma, out, per_thread);
    run("packed add,mul,add (contracted to FADD2+FFMA2 by ptxas)", k_packed_addmul, out, per_thread);
    run("packed add,mul,add unfused (FADD2,FFMA2,FADD2)", k_packed_addmul_unfused, out, per_thread, -0.0f);
    run("packed fma x3 x2", k_packed_fma, out, per_thread);
    int clk;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("peak at %d MHz: %.2f T lane-op/s (148 SM x 128 lanes)\n", clk / 1000, 148.0 * 128 * clk * 1e3 / 1e12);
    return 0;
}
