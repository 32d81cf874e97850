// microbenchmark: packed f32x2 vs scalar FP32 issue/throughput on sm_100a
#include <cuda_runtime.h>
#include <cstdio>
__device__ __forceinline__ float2 add2(float2 a, float2 b) {
    float2 r;
    asm volatile("{.reg .b64 ra, rb, rc; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; add.rn.f32x2 rc, ra, rb; mov.b64 {%0,%1}, rc;}" : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return r;
}
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
    float2 r;
    asm volatile("{.reg .b64 ra, rb, rc, rd; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; mov.b64 rc, {%6,%7}; fma.rn.f32x2 rd, ra, rb, rc; mov.b64 {%0,%1}, rd;}" : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
    return r;
}
template <int MODE>
__global__ void __launch_bounds__(256) k(float2* out, int iters, float2 seed) {
    float2 a[8];
#pragma unroll
    for (int i = 0; i < 8; i++) a[i] = make_float2(seed.x + i + threadIdx.x, seed.y - i);
    const float2 c = make_float2(seed.x * 0.5f, seed.y * 0.25f);
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 4; u++)
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (MODE == 0) { asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[i].x) : "f"(c.x)); asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[i].y) : "f"(c.y)); }
            else if (MODE == 1) a[i] = add2(a[i], c);
            else if (MODE == 2) { asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i].x) : "f"(c.x), "f"(c.y)); asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i].y) : "f"(c.y), "f"(c.x)); }
            else a[i] = fma2(a[i], c, c);
        }
    }
    float2 s = make_float2(0, 0);
#pragma unroll
    for (int i = 0; i < 8; i++) { s.x += a[i].x; s.y += a[i].y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE> float run(float2* d, int blocks, int iters) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<blocks, 256>>>(d, 10, make_float2(1.f, 2.f));
    cudaEventRecord(e0);
    k<MODE><<<blocks, 256>>>(d, iters, make_float2(1.f, 2.f));
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}
int main() {
    float2* d; cudaMalloc(&d, 148 * 8 * 256 * sizeof(float2));
    const int iters = 20000;
    const char* names[] = {"FADD x2 scalar", "FADD2 packed", "FFMA x2 scalar", "FFMA2 packed"};
    for (int wpb = 1; wpb <= 8; wpb *= 2) {
        int blocks = 148 * wpb;
        float t[4] = {run<0>(d, blocks, iters), run<1>(d, blocks, iters), run<2>(d, blocks, iters), run<3>(d, blocks, iters)};
        for (int m = 0; m < 4; m++) {
            double lane_ops = (double)blocks * 256 * iters * 32 * 2; // float results
            printf("ctas/SM %d  %-16s %8.3f ms  %7.2f Tflop-results/s\n", wpb, names[m], t[m], lane_ops / t[m] / 1e9);
        }
    }
    return 0;
}
