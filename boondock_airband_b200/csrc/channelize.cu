/*
 * channelize.cu — K1: sample expansion + window + batched in-shared-memory FFT + per-channel bin pick (sm_100a).
 *
 * Replaces, for FFT_BATCH-style batches of sliding frames (boondock_airband.h:94, the VideoCore path's shape):
 *   sample expansion tables + window multiply   boondock_airband.cpp:338-346, 426-479   (samplefft on the Pi, rtl_airband_neon.s:28-83)
 *   fftwf_execute / gpu_fft_execute             boondock_airband.cpp:481-485
 *   bin pick                                    boondock_airband.cpp:507-513            (the magnitude is taken in K2)
 *
 * Shape of the kernel
 *   - persistent CTAs walk a list of tiles; a tile = `tile_frames` consecutive frames of one input;
 *   - the byte span of the tile (frames overlap by fft_size - hop samples) is staged once in shared memory with
 *     16-byte coalesced loads, so every input byte is read from HBM/L2 once per tile;
 *   - an FFT is done by a group of G = N/16 threads (one warp for N = 512; a half warp for 256; 2..16 warps above),
 *     16 complex values per thread per pass, mixed radix 2/4/8/16 decimation in frequency, 2..4 passes;
 *     window values and twiddles of a thread do not depend on the frame, they live in registers;
 *   - between passes the data goes through a per-group shared-memory buffer whose layouts are padded/skewed so that
 *     every 8-byte load and store of a half warp hits 16 distinct bank pairs;
 *   - the last pass leaves the spectrum in a digit-reversed order; only the configured bins are read out
 *     (pick table built per input) and written as one contiguous row picks[frame][channel].
 * Sample conversion is exact: u8 levels (i - 127.5f) / 127.5f are reproduced with a Newton-corrected reciprocal
 * (verified for all 256 codes by tests), products use round-to-nearest multiplies that are never contracted.
 * No cuFFT, no tensor cores: the work is FP32 butterflies and shared-memory transposes.
 */
#include <stdint.h>

#include "ba_kernels.h"

namespace ba {
namespace {

template <int N>
struct Plan;
#define BA_PLAN(N_, P_, A, B, C, D)                                                         \
    template <>                                                                             \
    struct Plan<N_> {                                                                       \
        static constexpr int P = P_;                                                        \
        static constexpr int r(int i) { return i == 0 ? A : (i == 1 ? B : (i == 2 ? C : D)); } \
    };
BA_PLAN(256, 2, 16, 16, 1, 1)
BA_PLAN(512, 3, 8, 8, 8, 1)
BA_PLAN(1024, 3, 4, 16, 16, 1)
BA_PLAN(2048, 3, 8, 16, 16, 1)
BA_PLAN(4096, 3, 16, 16, 16, 1)
BA_PLAN(8192, 4, 2, 16, 16, 16)

template <int N>
struct Geo {
    using PL = Plan<N>;
    static constexpr int P = PL::P;
    static constexpr int G = N / 16; /* threads per FFT */
    static constexpr int THREADS = (N >= 8192) ? 512 : 256;
    static constexpr int W = THREADS / G; /* FFT groups per CTA */
    static constexpr int WORK = N + N / 8; /* float2 per group */
    static constexpr int R(int pass) { return PL::r(pass - 1); } /* pass = 1..P */
    static constexpr int Q(int level) { /* blocks after `level` passes */
        int q = 1;
        for (int i = 0; i < level; i++)
            q *= PL::r(i);
        return q;
    }
    static constexpr int M(int level) { return N / Q(level); }
    /* shared-memory position (in float2) of element `pos` of block `q` after `level` passes, 1 <= level <= P-1 */
    static __host__ __device__ constexpr int at(int level, int q, int pos) {
        if (level == P - 1)
            return pos * (Q(level) + 1) + q; /* feeds the last pass: position-major, odd row stride */
        const int next = M(level) / R(level + 1);
        const int stride = M(level) + (next < 16 ? 8 : 0);
        return q * stride + pos;
    }
    /* where the last pass leaves X[bin] */
    static __host__ __device__ constexpr int out_pos(int bin) {
        int q = 0, k = 0;
        for (int pass = 1; pass <= P; pass++) {
            k = bin % R(pass);
            bin /= R(pass);
            if (pass < P)
                q = q * R(pass) + k;
        }
        return k * (N / R(P)) + q;
    }
};

__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ float2 mul_mj(float2 a) { return make_float2(a.y, -a.x); } /* a * (-j) */

template <int R>
struct Dft;
template <>
struct Dft<1> {
    static __device__ __forceinline__ void run(float2*) {}
};
template <>
struct Dft<2> {
    static __device__ __forceinline__ void run(float2* x) {
        float2 a = x[0], b = x[1];
        x[0] = cadd(a, b);
        x[1] = csub(a, b);
    }
};
template <>
struct Dft<4> {
    static __device__ __forceinline__ void run(float2* x) {
        float2 t0 = cadd(x[0], x[2]), t1 = csub(x[0], x[2]);
        float2 t2 = cadd(x[1], x[3]), t3 = mul_mj(csub(x[1], x[3]));
        x[0] = cadd(t0, t2);
        x[1] = cadd(t1, t3);
        x[2] = csub(t0, t2);
        x[3] = csub(t1, t3);
    }
};
template <>
struct Dft<8> {
    static __device__ __forceinline__ void run(float2* x) {
        const float c = 0.70710678118654752440f;
        float2 a[4], b[4];
#pragma unroll
        for (int i = 0; i < 4; i++) {
            a[i] = cadd(x[i], x[i + 4]);
            b[i] = csub(x[i], x[i + 4]);
        }
        b[1] = make_float2(c * (b[1].x + b[1].y), c * (b[1].y - b[1].x));
        b[2] = mul_mj(b[2]);
        b[3] = make_float2(c * (b[3].y - b[3].x), -c * (b[3].x + b[3].y));
        Dft<4>::run(a);
        Dft<4>::run(b);
#pragma unroll
        for (int i = 0; i < 4; i++) {
            x[2 * i] = a[i];
            x[2 * i + 1] = b[i];
        }
    }
};
template <>
struct Dft<16> {
    static __device__ __forceinline__ void run(float2* x) {
        const float c1 = 0.92387953251128675613f, s1 = 0.38268343236508977173f, c = 0.70710678118654752440f;
        float2 a[8], b[8];
#pragma unroll
        for (int i = 0; i < 8; i++) {
            a[i] = cadd(x[i], x[i + 8]);
            b[i] = csub(x[i], x[i + 8]);
        }
        /* b[i] *= exp(-2 pi i * i / 16) */
        b[1] = cmul(b[1], make_float2(c1, -s1));
        b[2] = make_float2(c * (b[2].x + b[2].y), c * (b[2].y - b[2].x));
        b[3] = cmul(b[3], make_float2(s1, -c1));
        b[4] = mul_mj(b[4]);
        b[5] = cmul(b[5], make_float2(-s1, -c1));
        b[6] = make_float2(c * (b[6].y - b[6].x), -c * (b[6].x + b[6].y));
        b[7] = cmul(b[7], make_float2(-c1, -s1));
        Dft<8>::run(a);
        Dft<8>::run(b);
#pragma unroll
        for (int i = 0; i < 8; i++) {
            x[2 * i] = a[i];
            x[2 * i + 1] = b[i];
        }
    }
};

template <int N>
__device__ __forceinline__ void group_sync(int grp) {
    constexpr int G = Geo<N>::G;
    if (G <= 32) {
        __syncwarp();
    } else if (Geo<N>::W == 1) {
        __syncthreads();
    } else {
        BA_BAR_SYNC(1 + grp, G);
    }
}

/* exact (i - 127.5f) / 127.5f for a byte code, boondock_airband.cpp:341-343 */
__device__ __forceinline__ float level_u8(unsigned v) {
    const float a = __fadd_rn(__uint_as_float(0x4A800000u | (v << 1)), -4194431.5f); /* v - 127.5, exact */
    const float r = 1.0f / 127.5f;
    const float q = __fmul_rn(a, r);
    const float e = __fmaf_rn(-q, 127.5f, a);
    return __fmaf_rn(e, r, q);
}
/* i / 128.0f, boondock_airband.cpp:344-346 (code 0x80 is left undefined by the reference; -1.0 here) */
__device__ __forceinline__ float level_s8(unsigned v) {
    return __fmul_rn((float)(int)(signed char)v, 0.0078125f);
}

template <int N, int FMT>
__device__ __forceinline__ float2 load_sample(const unsigned char* frame, int n, float w, float scale) {
    if (FMT == BA_SFMT_U8 || FMT == BA_SFMT_S8) {
        const unsigned v = *reinterpret_cast<const unsigned short*>(frame + 2 * n);
        const float i = (FMT == BA_SFMT_U8) ? level_u8(v & 0xffu) : level_s8(v & 0xffu);
        const float q = (FMT == BA_SFMT_U8) ? level_u8(v >> 8) : level_s8(v >> 8);
        return make_float2(__fmul_rn(i, w), __fmul_rn(q, w));
    } else if (FMT == BA_SFMT_S16) {
        const short2 v = *reinterpret_cast<const short2*>(frame + 4 * n);
        return make_float2(__fmul_rn(__fmul_rn(scale, (float)v.x), w), __fmul_rn(__fmul_rn(scale, (float)v.y), w));
    } else {
        const float2 v = *reinterpret_cast<const float2*>(frame + 8 * n);
        return make_float2(__fmul_rn(__fmul_rn(scale, v.x), w), __fmul_rn(__fmul_rn(scale, v.y), w));
    }
}

/* per-thread constants that do not depend on the frame */
template <int N>
struct ThreadConst {
    float win[16];
    float2 tw[3][15];
};

template <int N, int PASS>
__device__ __forceinline__ void twiddle_setup(ThreadConst<N>& tc, const float2* __restrict__ table, int t) {
    using GE = Geo<N>;
    if constexpr (PASS < GE::P) {
        constexpr int R = GE::R(PASS);
        constexpr int NB = 16 / R;
        constexpr int Mp = GE::M(PASS);       /* positions per block after this pass */
        constexpr int Qprev = GE::Q(PASS - 1);
#pragma unroll
        for (int i = 0; i < NB; i++) {
            const int u = t + GE::G * i;
            const int m = u % Mp;
#pragma unroll
            for (int k = 1; k < R; k++)
                tc.tw[PASS - 1][i * (R - 1) + (k - 1)] = table[k * m * Qprev];
        }
        twiddle_setup<N, PASS + 1>(tc, table, t);
    }
}

template <int N, int FMT, int PASS>
__device__ __forceinline__ void fft_pass(const ThreadConst<N>& tc, float2* __restrict__ work, const unsigned char* frame, float scale, int t,
                                         int grp, float2* dbg_in) {
    using GE = Geo<N>;
    constexpr int R = GE::R(PASS);
    constexpr int NB = 16 / R;
    constexpr int Mp = GE::M(PASS);
    constexpr bool FIRST = PASS == 1, LAST = PASS == GE::P;
    float2 x[NB][R];
#pragma unroll
    for (int i = 0; i < NB; i++) {
        const int u = t + GE::G * i;
        const int q = u / Mp, m = u % Mp;
#pragma unroll
        for (int j = 0; j < R; j++) {
            if (FIRST) {
                const int n = j * Mp + m;
                x[i][j] = load_sample<N, FMT>(frame, n, tc.win[i * R + j], scale);
                if (dbg_in)
                    dbg_in[n] = x[i][j];
            } else {
                x[i][j] = work[GE::at(PASS - 1, q, j * Mp + m)];
            }
        }
    }
    /* every load of this pass (and every read of the previous frame's spectrum) precedes the stores below */
    group_sync<N>(grp);
#pragma unroll
    for (int i = 0; i < NB; i++) {
        const int u = t + GE::G * i;
        const int q = u / Mp, m = u % Mp;
        Dft<R>::run(x[i]);
#pragma unroll
        for (int k = 0; k < R; k++) {
            if (LAST) {
                work[k * (N / R) + q] = x[i][k];
            } else {
                const float2 v = (k == 0) ? x[i][0] : cmul(x[i][k], tc.tw[PASS - 1][i * (R - 1) + (k - 1)]);
                work[GE::at(PASS, q * R + k, m)] = v;
            }
        }
    }
    group_sync<N>(grp);
    if constexpr (!LAST)
        fft_pass<N, FMT, PASS + 1>(tc, work, frame, scale, t, grp, dbg_in);
}

template <int N, int FMT>
__device__ __forceinline__ void run_tile_frames(const ThreadConst<N>& tc, const K1Device& d, const unsigned char* raw0, float2* work,
                                                const uint16_t* picktab, int f0, int nf, int t, int grp) {
    using GE = Geo<N>;
    constexpr int BYTES = (FMT == BA_SFMT_U8 || FMT == BA_SFMT_S8) ? 2 : (FMT == BA_SFMT_S16 ? 4 : 8);
    (void)BYTES;
    for (int base = 0; base < nf; base += GE::W) {
        const bool live = base + grp < nf;
        const int fi = live ? base + grp : nf - 1; /* idle groups redo the last frame so that barriers stay uniform */
        const unsigned char* frame = raw0 + (size_t)fi * d.hop_bytes;
        const int f = f0 + fi;
        float2* dbg_in = (live && d.dbg_in) ? d.dbg_in + (size_t)f * N : nullptr;
        fft_pass<N, FMT, 1>(tc, work, frame, d.scale, t, grp, dbg_in);
        if (live) {
            const uint64_t fs = d.frame0 + (uint64_t)f;
            float2* row = d.picks + (size_t)(fs & d.ring_mask) * d.c_pad;
            float* mrow = d.mags + (size_t)(fs & d.ring_mask) * d.c_pad;
            for (int c = t; c < (int)d.n_channels; c += GE::G) {
                const float2 v = work[picktab[c]];
                row[c] = v;
                /* wavein[] = sqrtf(re*re + im*im), boondock_airband.cpp:507-513: three separately rounded operations and an IEEE square root */
                mrow[c] = __fsqrt_rn(__fadd_rn(__fmul_rn(v.x, v.x), __fmul_rn(v.y, v.y)));
            }
            if (d.dbg_out) {
                float2* o = d.dbg_out + (size_t)f * N;
                for (int k = t; k < N; k += GE::G)
                    o[k] = work[GE::out_pos(k)];
            }
            if (d.spectrum && f == (int)d.n_frames - 1) {
                for (int k = t; k < N; k += GE::G)
                    d.spectrum[k] = work[GE::out_pos(k)];
            }
        }
    }
}

template <int N>
__global__ void __launch_bounds__(Geo<N>::THREADS) channelize_kernel(K1Params p) {
    using GE = Geo<N>;
    BA_SHARED(smem);
    unsigned char* raw = smem;
    float2* work_all = reinterpret_cast<float2*>(smem + p.raw_bytes);
    uint16_t* picktab = reinterpret_cast<uint16_t*>(work_all + GE::W * GE::WORK);
    const int tid = threadIdx.x;
    const int grp = tid / GE::G, t = tid % GE::G;
    float2* work = work_all + grp * GE::WORK;

    ThreadConst<N> tc;
    {
        constexpr int R = GE::R(1);
        constexpr int NB = 16 / R;
        constexpr int M1 = GE::M(1);
#pragma unroll
        for (int i = 0; i < NB; i++) {
            const int m = (t + GE::G * i) % M1;
#pragma unroll
            for (int j = 0; j < R; j++)
                tc.win[i * R + j] = p.window[j * M1 + m];
        }
        twiddle_setup<N, 1>(tc, p.twiddle, t);
    }

    int cur_dev = -1;
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
        /* input that owns this tile: last device with tile0 <= tile */
        int lo = 0, hi = p.n_dev - 1;
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (p.dev[mid].tile0 <= (uint32_t)tile)
                lo = mid;
            else
                hi = mid - 1;
        }
        const K1Device& d = p.dev[lo];
        const int f0 = (tile - (int)d.tile0) * p.tile_frames;
        const int nf = min(p.tile_frames, (int)d.n_frames - f0);
        const int bytes_per = (d.fmt == BA_SFMT_S16) ? 4 : (d.fmt == BA_SFMT_F32 ? 8 : 2);
        const unsigned char* g0 = d.iq + (size_t)f0 * d.hop_bytes;
        const size_t span = (size_t)(nf - 1) * d.hop_bytes + (size_t)N * bytes_per;
        const uintptr_t a0 = reinterpret_cast<uintptr_t>(g0) & ~(uintptr_t)15;
        const int pre = (int)(reinterpret_cast<uintptr_t>(g0) - a0);
        const int chunks = (int)((pre + span + 15) >> 4);

        __syncthreads(); /* the previous tile's frames are done with raw[] and picktab[] */
        for (int i = tid; i < chunks; i += GE::THREADS) {
            const unsigned char* src = reinterpret_cast<const unsigned char*>(a0) + 16 * (size_t)i;
            uint4 v;
            if (src >= d.lo && src + 16 <= d.hi) {
                v = __ldg(reinterpret_cast<const uint4*>(src));
            } else {
                unsigned char tmp[16];
#pragma unroll
                for (int b = 0; b < 16; b++)
                    tmp[b] = (src + b >= d.lo && src + b < d.hi) ? src[b] : 0;
                v = *reinterpret_cast<uint4*>(tmp);
            }
            reinterpret_cast<uint4*>(raw)[i] = v;
        }
        if (lo != cur_dev) {
            for (int c = tid; c < (int)d.n_channels; c += GE::THREADS)
                picktab[c] = (uint16_t)GE::out_pos((int)(d.bins[c] & (N - 1)));
            cur_dev = lo;
        }
        __syncthreads();

        const unsigned char* raw0 = raw + pre;
        switch (d.fmt) {
            case BA_SFMT_U8:
                run_tile_frames<N, BA_SFMT_U8>(tc, d, raw0, work, picktab, f0, nf, t, grp);
                break;
            case BA_SFMT_S8:
                run_tile_frames<N, BA_SFMT_S8>(tc, d, raw0, work, picktab, f0, nf, t, grp);
                break;
            case BA_SFMT_S16:
                run_tile_frames<N, BA_SFMT_S16>(tc, d, raw0, work, picktab, f0, nf, t, grp);
                break;
            default:
                run_tile_frames<N, BA_SFMT_F32>(tc, d, raw0, work, picktab, f0, nf, t, grp);
                break;
        }
    }
}

template <int N>
int launch_n(const K1Params& p, int n_ctas, cudaStream_t s) {
    using GE = Geo<N>;
    const size_t smem = (size_t)p.raw_bytes + sizeof(float2) * GE::W * GE::WORK + sizeof(uint16_t) * ((p.max_channels + 7) & ~7);
    static size_t configured = 0;
    if (smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(channelize_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess)
            return (int)e;
        configured = smem;
    }
    BA_LAUNCH(channelize_kernel<N>, n_ctas, GE::THREADS, smem, s, p);
    return (int)cudaGetLastError();
}

}  // namespace

int k1_threads(int n) {
    return n >= 8192 ? 512 : 256;
}
int k1_groups(int n) {
    return k1_threads(n) / (n / 16);
}
int k1_smem_bytes(int n, int raw_bytes, int max_channels) {
    return raw_bytes + 8 * k1_groups(n) * (n + n / 8) + 2 * ((max_channels + 7) & ~7);
}

int k1_launch(int fft_size, const K1Params& p, int n_ctas, cudaStream_t s) {
    switch (fft_size) {
        case 256:
            return launch_n<256>(p, n_ctas, s);
        case 512:
            return launch_n<512>(p, n_ctas, s);
        case 1024:
            return launch_n<1024>(p, n_ctas, s);
        case 2048:
            return launch_n<2048>(p, n_ctas, s);
        case 4096:
            return launch_n<4096>(p, n_ctas, s);
        case 8192:
            return launch_n<8192>(p, n_ctas, s);
    }
    return (int)cudaErrorInvalidValue;
}

}  // namespace ba
