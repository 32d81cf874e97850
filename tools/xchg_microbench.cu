// microbenchmark (sm_100a): the exchange between the two passes of the 512-point channelizer (32 x 16 plan, 16 threads per
// FFT, 32 complex values per thread) done through shared memory, as channelize.cu does it, against the same exchange done
// with warp shuffles.  After pass 1 thread t of a half warp holds X[k][t], k = 0..31; pass 2 wants thread u to hold
// X[2u][s] and X[2u+1][s], s = 0..15: a 16 x 16 transpose of items of two complex values between the lanes of a half warp.
//   MODE 0  32 x STS.64 into a skewed buffer, __syncwarp, 32 x LDS.64          (what the kernel does)
//   MODE 1  four butterfly stages (lane ^ 1, 2, 4, 8): per stage 8 items of 4 floats go to the partner lane = 32 SHFL + the
//           selects that pick which half to send and where to put what arrived (registers cannot be indexed by lane)
// Both are followed by a token amount of arithmetic per value (one packed add) so that the loop carries a dependence, and both
// must produce the same checksum.  nvcc -arch=sm_100a -O3 -o /tmp/xchg tools/xchg_microbench.cu && /tmp/xchg
#include <cuda_runtime.h>

#include <cstdio>

constexpr int kThreads = 128;  // the 512-point kernel's CTA: 8 half warps = 8 FFTs in flight
constexpr int kRow = 17;       // float2 per row of the skewed buffer: 16 lanes + 1 (odd stride: no bank conflicts on the transposed read)

template <int MODE>
__global__ void __launch_bounds__(kThreads) xchg(float2* out, int iters, float seed) {
    __shared__ float2 buf[kThreads / 16][32 * kRow];
    const int lane16 = threadIdx.x & 15, grp = threadIdx.x >> 4;
    float2 v[32];
#pragma unroll
    for (int k = 0; k < 32; k++)
        v[k] = make_float2(seed + k + 32 * lane16, seed - k * 0.5f + grp);
    for (int it = 0; it < iters; it++) {
        if (MODE == 0) {
            float2* b = buf[grp];
#pragma unroll
            for (int k = 0; k < 32; k++)
                b[k * kRow + lane16] = v[k];  // X[k][t]
            __syncwarp();
#pragma unroll
            for (int s = 0; s < 16; s++) {
                v[2 * s] = b[(2 * lane16) * kRow + s];  // X[2u][s]
                v[2 * s + 1] = b[(2 * lane16 + 1) * kRow + s];
            }
            __syncwarp();
        } else {
            // item i = (v[2i], v[2i+1]); in-register transpose of the 16 x 16 item matrix over the lanes of a half warp
#pragma unroll
            for (int m = 1; m < 16; m <<= 1) {
                const bool up = (lane16 & m) != 0;
#pragma unroll
                for (int i = 0; i < 16; i++) {
                    if (i & m)
                        continue;
                    const int j = i | m;
                    // the lane with bit m clear keeps item i and trades item j for the partner's item i
                    float2 s0 = up ? v[2 * i] : v[2 * j], s1 = up ? v[2 * i + 1] : v[2 * j + 1];
                    s0.x = __shfl_xor_sync(0xffffffffu, s0.x, m);
                    s0.y = __shfl_xor_sync(0xffffffffu, s0.y, m);
                    s1.x = __shfl_xor_sync(0xffffffffu, s1.x, m);
                    s1.y = __shfl_xor_sync(0xffffffffu, s1.y, m);
                    if (up) {
                        v[2 * i] = s0;
                        v[2 * i + 1] = s1;
                    } else {
                        v[2 * j] = s0;
                        v[2 * j + 1] = s1;
                    }
                }
            }
        }
#pragma unroll
        for (int k = 0; k < 32; k++) {
            v[k].x += 1.0f;
            v[k].y -= 1.0f;
        }
    }
    float2 s = make_float2(0.f, 0.f);
#pragma unroll
    for (int k = 0; k < 32; k++) {
        s.x += v[k].x * (float)(k + 1);
        s.y += v[k].y;
    }
    out[blockIdx.x * kThreads + threadIdx.x] = s;
}

template <int MODE>
static float run(float2* d, int blocks, int iters, double* sum) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    xchg<MODE><<<blocks, kThreads>>>(d, 4, 1.0f);
    cudaEventRecord(e0);
    xchg<MODE><<<blocks, kThreads>>>(d, iters, 1.0f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    static float2 h[1 << 16];
    cudaMemcpy(h, d, sizeof(float2) * 4096, cudaMemcpyDeviceToHost);
    double s = 0.0;
    for (int i = 0; i < 4096; i++)
        s += (double)h[i].x + (double)h[i].y;
    *sum = s;
    return ms;
}

int main() {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    float2* d;
    const int iters = 2000;
    for (int per_sm = 1; per_sm <= 3; per_sm++) {  // the 512-point kernel keeps three CTAs per SM
        const int blocks = sms * per_sm;
        cudaMalloc(&d, sizeof(float2) * blocks * kThreads);
        double s0, s1;
        const float a = run<0>(d, blocks, iters, &s0), b = run<1>(d, blocks, iters, &s1);
        const double ffts = (double)blocks * (kThreads / 16) * iters;
        printf("CTAs/SM %d  shared memory %8.3f ms (%6.2f ps per 512-point exchange)   shuffles %8.3f ms (%6.2f ps)   checksums %s (%.6g %.6g)\n", per_sm, a,
               1e9 * a / ffts, b, 1e9 * b / ffts, s0 == s1 ? "equal" : "DIFFER", s0, s1);
        cudaFree(d);
    }
    return 0;
}
