// Packed f32x2 arithmetic (sm_100 FADD2 / FFMA2) with explicit, separately rounded semantics.
#pragma once
#include <cuda_runtime.h>

namespace ma {

typedef unsigned long long u64;

__device__ __forceinline__ u64 add2(u64 a, u64 b) {
    u64 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ u64 sub2(u64 a, u64 b) {
    u64 r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) {
    u64 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
// a*b rounded once.  ptxas contracts a mul.rn.f32x2 feeding an add.rn.f32x2 into one FFMA2 (observed
// in SASS, also with a literal -0 addend), which would break parity with unfused CPU arithmetic.
// The product is therefore an explicit fma with a -0 addend that arrives as a *runtime* value:
// x*k + (-0) == x*k exactly, and the following add cannot be merged into it.
__device__ __forceinline__ u64 mul2(u64 a, u64 b, u64 negzero) { return fma2(a, b, negzero); }

__device__ __forceinline__ u64 pack2(float x, float y) {
    u64 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(x), "f"(y));
    return r;
}
__device__ __forceinline__ float2 unpack2(u64 v) {
    float2 r;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
    return r;
}

}  // namespace ma
