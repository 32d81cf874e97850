/*
 * cuda_emu.h — a tiny host-side stand-in for the CUDA execution model.  TEST/DEVELOPMENT TOOL ONLY.
 *
 * There is no GPU in the development container, and a round trip to a B200 takes minutes.  To check the index
 * arithmetic, barrier placement and state handling of the kernels in boondock_airband_b200/csrc before spending GPU
 * time, the very same .cu sources are compiled by g++ with -DBA_EMU against this header: every CUDA thread becomes
 * an OS thread, __syncthreads()/named barriers become std::barrier, shared memory becomes a per-CTA heap block and
 * the handful of runtime calls the engine makes become malloc/memcpy.  The result (tests/emu/libba_emu_TESTONLY.so)
 * is loaded ONLY by tests/test_emu_*.py.  It is never built into, linked with, or loaded by the product library
 * libba_cuda.so, which has no CPU path at all and fails with BA_ERR_NO_DEVICE when no CUDA device is present.
 */
#ifndef BA_CUDA_EMU_H
#define BA_CUDA_EMU_H

#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <barrier>
#include <chrono>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __maxnreg__(...)
#define __align__(n) alignas(n)
#define __constant__ static

struct float2 {
    float x, y;
};
struct short2 {
    short x, y;
};
struct alignas(16) uint4 {
    unsigned x, y, z, w;
};
struct alignas(16) float4 {
    float x, y, z, w;
};
struct uint2 {
    unsigned x, y;
};
struct uint3 {
    unsigned x, y, z;
};
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};
static inline float2 make_float2(float x, float y) {
    float2 r = {x, y};
    return r;
}
static inline float4 make_float4(float x, float y, float z, float w) {
    float4 r = {x, y, z, w};
    return r;
}
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) {
    uint4 r = {x, y, z, w};
    return r;
}

namespace emu {

struct Cta {
    unsigned nthreads;
    std::vector<unsigned char> smem_store;
    unsigned char* smem;
    std::barrier<> all;
    std::vector<std::unique_ptr<std::barrier<>>> warp;
    std::mutex named_lock;
    std::unique_ptr<std::barrier<>> named[16];
    std::vector<uint64_t> shfl; /* one 8-byte slot per thread */
    Cta(unsigned n, size_t smem_bytes) : nthreads(n), smem_store(smem_bytes + 64), all(n), shfl(n) {
        smem = (unsigned char*)(((uintptr_t)smem_store.data() + 63) & ~(uintptr_t)63);
        for (unsigned w = 0; w * 32 < n; w++)
            warp.emplace_back(new std::barrier<>(std::min(32u, n - 32 * w)));
    }
};

struct Tls {
    Cta* cta;
    uint3 tid, bid;
    dim3 bdim, gdim;
};
inline Tls& tls() {
    static thread_local Tls t;
    return t;
}

template <class F>
void launch(dim3 grid, dim3 block, size_t smem_bytes, F body) {
    const unsigned max_par = 4;
    std::atomic<unsigned> next(0);
    auto cta_runner = [&]() {
        for (;;) {
            const unsigned lin = next.fetch_add(1);
            if (lin >= grid.x * grid.y * grid.z)
                return;
            const unsigned b = lin % grid.x, by = (lin / grid.x) % grid.y, bz = lin / (grid.x * grid.y);
            Cta cta(block.x, smem_bytes);
            std::vector<std::thread> th;
            th.reserve(block.x);
            for (unsigned t = 0; t < block.x; t++)
                th.emplace_back([&, t, b, by, bz]() {
                    Tls& s = tls();
                    s.cta = &cta;
                    s.tid = {t, 0, 0};
                    s.bid = {b, by, bz};
                    s.bdim = block;
                    s.gdim = grid;
                    body();
                });
            for (auto& x : th)
                x.join();
        }
    };
    std::vector<std::thread> runners;
    const unsigned n = std::min(max_par, grid.x * grid.y * grid.z);
    for (unsigned i = 0; i < n; i++)
        runners.emplace_back(cta_runner);
    for (auto& r : runners)
        r.join();
}

inline void bar_named(int id, int count) {
    Cta* c = tls().cta;
    std::barrier<>* b;
    {
        std::lock_guard<std::mutex> g(c->named_lock);
        if (!c->named[id])
            c->named[id].reset(new std::barrier<>(count));
        b = c->named[id].get();
    }
    b->arrive_and_wait();
}

}  // namespace emu

#define threadIdx (emu::tls().tid)
#define blockIdx (emu::tls().bid)
#define blockDim (emu::tls().bdim)
#define gridDim (emu::tls().gdim)

static inline void __syncthreads() {
    emu::tls().cta->all.arrive_and_wait();
}
static inline void __syncwarp(unsigned = 0xffffffffu) {
    emu::Tls& s = emu::tls();
    s.cta->warp[s.tid.x / 32]->arrive_and_wait();
}

static inline void __threadfence_block() {
    std::atomic_thread_fence(std::memory_order_seq_cst);
}

template <class T>
static inline T __ldg(const T* p) {
    return *p;
}

/* warp collectives: every lane of the warp takes part (full mask) */
template <class T>
static inline T emu_exchange(T v, int src_lane) {
    static_assert(sizeof(T) <= 8, "shuffle payload");
    emu::Tls& s = emu::tls();
    const unsigned base = s.tid.x & ~31u;
    uint64_t raw = 0;
    memcpy(&raw, &v, sizeof(T));
    s.cta->shfl[s.tid.x] = raw;
    __syncwarp();
    const unsigned width = std::min(32u, s.cta->nthreads - base);
    uint64_t got = s.cta->shfl[base + ((unsigned)src_lane % width)];
    __syncwarp();
    T out;
    memcpy(&out, &got, sizeof(T));
    return out;
}
template <class T>
static inline T __shfl_sync(unsigned, T v, int src, int width = 32) {
    const int lane = emu::tls().tid.x & 31;
    return emu_exchange(v, (lane & ~(width - 1)) | (src & (width - 1)));
}
template <class T>
static inline T __shfl_xor_sync(unsigned, T v, int m, int = 32) {
    return emu_exchange(v, (int)(emu::tls().tid.x & 31) ^ m);
}
template <class T>
static inline T __shfl_down_sync(unsigned, T v, unsigned d, int = 32) {
    const int lane = emu::tls().tid.x & 31;
    return emu_exchange(v, lane + (int)d < 32 ? lane + (int)d : lane);
}
template <class T>
static inline T __shfl_up_sync(unsigned, T v, unsigned d, int = 32) {
    const int lane = emu::tls().tid.x & 31;
    return emu_exchange(v, lane >= (int)d ? lane - (int)d : lane);
}
static inline int __popc(unsigned x) {
    return __builtin_popcount(x);
}
static inline int __clz(int x) {
    return x == 0 ? 32 : __builtin_clz((unsigned)x);
}
static inline unsigned __ballot_sync(unsigned, int pred) {
    unsigned bits = 0;
    for (int l = 0; l < 32; l++)
        bits |= (emu_exchange<int>(pred ? 1 : 0, l) ? 1u : 0u) << l;
    return bits;
}
static inline int __any_sync(unsigned m, int pred) {
    return __ballot_sync(m, pred) != 0;
}
static inline int __all_sync(unsigned m, int pred) {
    emu::Tls& s = emu::tls();
    const unsigned width = std::min(32u, s.cta->nthreads - (s.tid.x & ~31u));
    const unsigned full = width == 32 ? 0xffffffffu : ((1u << width) - 1);
    return (__ballot_sync(m, pred) & full) == full;
}

/* arithmetic intrinsics: IEEE single precision, round to nearest (build with -ffp-contract=off) */
static inline float __fmul_rn(float a, float b) {
    return a * b;
}
static inline float __fadd_rn(float a, float b) {
    return a + b;
}
static inline float __fsub_rn(float a, float b) {
    return a - b;
}
static inline float __fdiv_rn(float a, float b) {
    return a / b;
}
static inline float __fsqrt_rn(float a) {
    return sqrtf(a);
}
static inline float __fmaf_rn(float a, float b, float c) {
    return fmaf(a, b, c);
}
static inline double __dmul_rn(double a, double b) {
    return a * b;
}
static inline float __uint_as_float(unsigned u) {
    float f;
    memcpy(&f, &u, 4);
    return f;
}
static inline unsigned __float_as_uint(float f) {
    unsigned u;
    memcpy(&u, &f, 4);
    return u;
}
template <class T>
static inline T min(T a, T b) {
    return b < a ? b : a;
}
template <class T>
static inline T max(T a, T b) {
    return a < b ? b : a;
}

/* ---------------------------------------------------------------- the few runtime calls the engine makes */
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorInvalidValue = 1, cudaErrorMemoryAllocation = 2, cudaErrorNotReady = 600, cudaErrorNoDevice = 100 };
typedef struct emu_stream* cudaStream_t;
struct emu_event {
    std::chrono::steady_clock::time_point t;
};
typedef emu_event* cudaEvent_t;
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
enum { cudaStreamNonBlocking = 1, cudaHostAllocDefault = 0, cudaEventDefault = 0, cudaEventDisableTiming = 2 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
enum cudaDeviceAttr { cudaDevAttrMultiProcessorCount = 16, cudaDevAttrMaxSharedMemoryPerBlockOptin = 97 };

/* Devices: BA_EMU_DEVICES (default 1) emulated GPUs; the current device is per-thread state as in the runtime, and the
 * dynamic shared-memory limit of a kernel is kept per (device, kernel) as the driver keeps it: a launch that asks for more
 * than 48 KB on a device where the attribute was never set fails with cudaErrorInvalidValue, exactly what happens to an
 * engine on a second GPU when the attribute is cached per process. */
namespace emu {
inline int device_count() {
    const char* v = getenv("BA_EMU_DEVICES");
    const int n = v ? atoi(v) : 1;
    return n > 0 ? n : 1;
}
inline int& current_device() {
    static thread_local int d = 0;
    return d;
}
inline int& last_error() {
    static thread_local int e = 0;
    return e;
}
struct FuncAttrs {
    std::mutex lock;
    std::vector<std::pair<std::pair<int, const void*>, size_t>> v;
};
inline FuncAttrs& func_attrs() {
    static FuncAttrs f;
    return f;
}
inline bool launch_allowed(const void* kern, size_t smem) {
    if (smem <= 48 * 1024)
        return true;
    FuncAttrs& f = func_attrs();
    std::lock_guard<std::mutex> g(f.lock);
    for (auto& e : f.v)
        if (e.first.first == current_device() && e.first.second == kern && e.second >= smem)
            return true;
    last_error() = 1; /* cudaErrorInvalidValue */
    return false;
}
}  // namespace emu
static inline cudaError_t cudaGetDeviceCount(int* n) {
    *n = emu::device_count();
    return cudaSuccess;
}
static inline cudaError_t cudaSetDevice(int d) {
    if (d < 0 || d >= emu::device_count())
        return cudaErrorInvalidValue;
    emu::current_device() = d;
    return cudaSuccess;
}
static inline cudaError_t cudaGetDevice(int* d) {
    *d = emu::current_device();
    return cudaSuccess;
}
static inline cudaError_t cudaDeviceGetAttribute(int* v, cudaDeviceAttr a, int) {
    *v = (a == cudaDevAttrMultiProcessorCount) ? 2 : 227 * 1024;
    return cudaSuccess;
}
static inline cudaError_t cudaMalloc(void** p, size_t n) {
    *p = aligned_alloc(256, (n + 255) & ~(size_t)255);
    return *p ? cudaSuccess : cudaErrorMemoryAllocation;
}
static inline cudaError_t cudaFree(void* p) {
    free(p);
    return cudaSuccess;
}
static inline cudaError_t cudaHostAlloc(void** p, size_t n, unsigned) {
    return cudaMalloc(p, n);
}
static inline cudaError_t cudaFreeHost(void* p) {
    free(p);
    return cudaSuccess;
}
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t = 0) {
    memmove(d, s, n);
    return cudaSuccess;
}
static inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) {
    memmove(d, s, n);
    return cudaSuccess;
}
static inline cudaError_t cudaMemcpy2D(void* d, size_t dp, const void* s, size_t sp, size_t w, size_t h, cudaMemcpyKind) {
    for (size_t r = 0; r < h; r++)
        memmove((char*)d + r * dp, (const char*)s + r * sp, w);
    return cudaSuccess;
}
static inline cudaError_t cudaMemcpy2DAsync(void* d, size_t dp, const void* s, size_t sp, size_t w, size_t h, cudaMemcpyKind k, cudaStream_t = 0) {
    return cudaMemcpy2D(d, dp, s, sp, w, h, k);
}
static inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t = 0) {
    memset(d, v, n);
    return cudaSuccess;
}
static inline cudaError_t cudaMemset(void* d, int v, size_t n) {
    memset(d, v, n);
    return cudaSuccess;
}
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) {
    *s = 0;
    return cudaSuccess;
}
static inline cudaError_t cudaStreamDestroy(cudaStream_t) {
    return cudaSuccess;
}
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) {
    return cudaSuccess;
}
static inline cudaError_t cudaDeviceSynchronize() {
    return cudaSuccess;
}
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned = 0) {
    return cudaSuccess;
}
static inline cudaError_t cudaEventCreate(cudaEvent_t* e) {
    *e = new emu_event();
    return cudaSuccess;
}
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) {
    return cudaEventCreate(e);
}
static inline cudaError_t cudaEventDestroy(cudaEvent_t e) {
    delete e;
    return cudaSuccess;
}
static inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t = 0) {
    e->t = std::chrono::steady_clock::now();
    return cudaSuccess;
}
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) {
    return cudaSuccess;
}
static inline cudaError_t cudaEventQuery(cudaEvent_t) {
    return cudaSuccess;
}
static inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b) {
    *ms = std::chrono::duration<float, std::milli>(b->t - a->t).count();
    return cudaSuccess;
}
static inline cudaError_t cudaGetLastError() {
    const int e = emu::last_error();
    emu::last_error() = 0;
    return e;
}
static inline cudaError_t cudaPeekAtLastError() {
    return cudaSuccess;
}
static inline const char* cudaGetErrorString(cudaError_t e) {
    return e == cudaSuccess ? "no error" : "emulated CUDA error";
}
template <class K>
static inline cudaError_t cudaFuncSetAttribute(K kern, cudaFuncAttribute, int bytes) {
    emu::FuncAttrs& f = emu::func_attrs();
    std::lock_guard<std::mutex> g(f.lock);
    const std::pair<int, const void*> key(emu::current_device(), (const void*)kern);
    for (auto& e : f.v)
        if (e.first == key) {
            e.second = (size_t)bytes;
            return cudaSuccess;
        }
    f.v.push_back(std::make_pair(key, (size_t)bytes));
    return cudaSuccess;
}

#endif
