// Packed fp32 pairs (PTX .f32x2, SASS FFMA2 / FMUL2 / FADD2; sm_100 and later): two fp32 lanes per
// instruction at the scalar FMA-pipe rate but in half the issue slots.
#pragma once

namespace upol {

typedef unsigned long long f32x2;

__device__ __forceinline__ f32x2 pack2(float lo, float hi) {
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ f32x2 bcast2(float v) { return pack2(v, v); }
__device__ __forceinline__ void unpack2(f32x2 v, float &lo, float &hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
    f32x2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
    f32x2 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
    f32x2 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2 sub2(f32x2 a, f32x2 b) {
    f32x2 r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}

// bare MUFU.RSQ / MUFU.RCP (rel. error <= 2^-22.9 / 1 ulp), no denormal fix-up; 0 -> +inf
__device__ __forceinline__ float rsqrt_f32(float x) {
    float y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float rcp_f32(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ f32x2 rsqrt2(f32x2 v) {
    float a, b;
    unpack2(v, a, b);
    return pack2(rsqrt_f32(a), rsqrt_f32(b));
}
__device__ __forceinline__ f32x2 rcp2(f32x2 v) {
    float a, b;
    unpack2(v, a, b);
    return pack2(rcp_f32(a), rcp_f32(b));
}

}  // namespace upol
