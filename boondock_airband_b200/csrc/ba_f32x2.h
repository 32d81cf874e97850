/*
 * ba_f32x2.h — Blackwell's packed FP32 pairs (add / sub / mul / fma .rn.f32x2 -> SASS FADD2 / FMUL2 / FFMA2): two
 * independent, individually rounded IEEE single-precision operations in ONE instruction.  A float2 lives in an aligned register
 * pair; ptxas folds component swaps, sign flips and scalar broadcasts of the operands into the instruction.
 *   K1 (channelize.cu) uses them for the complex arithmetic of the butterflies (half the issue slots);
 *   K2 (demod.cu) uses them in its serial recurrences: a warp that runs alone on its scheduler pays ~4 cycles per instruction
 *   whatever it is, so two independent operations of a loop body in one instruction halve their share of the step.  Each half
 *   rounds exactly as the scalar operation would (-fmad=false concerns contraction only; fma2 is not used there).
 */
#ifndef BA_F32X2_H
#define BA_F32X2_H

#include "ba_port.h"

namespace ba {
#ifdef BA_EMU
__device__ __forceinline__ float2 add2(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 sub2(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 mul2(float2 a, float2 b) { return make_float2(a.x * b.x, a.y * b.y); }
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) { return make_float2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y)); }
#else
__device__ __forceinline__ float2 add2(float2 a, float2 b) {
    float2 r;
    asm("{.reg .b64 ra, rb, rc; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; add.rn.f32x2 rc, ra, rb; mov.b64 {%0,%1}, rc;}"
        : "=f"(r.x), "=f"(r.y)
        : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return r;
}
__device__ __forceinline__ float2 sub2(float2 a, float2 b) {
    float2 r;
    asm("{.reg .b64 ra, rb, rc; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; sub.rn.f32x2 rc, ra, rb; mov.b64 {%0,%1}, rc;}"
        : "=f"(r.x), "=f"(r.y)
        : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return r;
}
__device__ __forceinline__ float2 mul2(float2 a, float2 b) {
    float2 r;
    asm("{.reg .b64 ra, rb, rc; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; mul.rn.f32x2 rc, ra, rb; mov.b64 {%0,%1}, rc;}"
        : "=f"(r.x), "=f"(r.y)
        : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return r;
}
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
    float2 r;
    asm("{.reg .b64 ra, rb, rc, rd; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; mov.b64 rc, {%6,%7}; fma.rn.f32x2 rd, ra, rb, rc; mov.b64 {%0,%1}, rd;}"
        : "=f"(r.x), "=f"(r.y)
        : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
    return r;
}
#endif
}  // namespace ba
#endif
