/*
 * oracle.cpp — CPU restatement of Boondock-Airband's demodulate() loop (the parity oracle).
 *
 * TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may build, load or run this; the shipped engine (boondock_airband_b200/csrc) never does.
 *
 * Two builds of this one file (see oracle/Makefile):
 *   oracle/libba_oracle.so            the per-sample DSP classes come from dsp_restated.h (restatement, travels in git)
 *   oracle/_ref/libba_oracle_ref.so   -DBA_ORACLE_REF: the very same loop drives the reference's OWN squelch.cpp,
 *                                     ctcss.cpp, filters.cpp and logging.cpp, compiled unmodified from /root/reference/src
 * The loop itself cannot be taken from the reference (boondock_airband.cpp needs fftw3, libconfig++, lame, shout —
 * none installed), so it is restated here, each block citing the lines it follows (all under /root/reference/src).
 *
 * FFT: FFTW3f is a third-party dependency of the reference that is neither vendored nor version-pinned
 * (src/CMakeLists.txt:253-260) and is not installed here, and the reference has no test that pins a result at the
 * FFT boundary => FFT PARITY IS UNPINNED.  The oracle's FFT is an in-file float32 Stockham radix-4 transform of
 * the same definition (forward, unnormalised: X[k] = sum x[n] exp(-2 pi i k n / N), fftwf_plan_dft_1d(..., FFTW_FORWARD, ...)
 * boondock_airband.cpp:264); tests bound it against a float64 DFT.
 *
 * Arithmetic: IEEE float, no contraction, no -ffast-math (the reference's own Release build uses -ffast-math
 * -march=native and is therefore not reproducible bit for bit across compilers).
 *
 * Timing builds (-DBA_ORACLE_FAST, the *_fast.so targets of oracle/Makefile; they serve bench.py's CPU legs only and are
 * never compared bit for bit): the reference plans fftwf_plan_dft_1d(..., FFTW_MEASURE) and is built -O3 -ffast-math
 * -march=native (boondock_airband.cpp:262-264, CMakeLists.txt:31-42), so a scalar transform would flatter the GPU.  FFTW is
 * not installed here or on the GPU box; the timing builds therefore run an AVX2 + FMA transform written for this workload:
 * eight consecutive frames at a time, one frame per vector lane (radix-4 Stockham, split real / imaginary arrays), with the
 * sample conversion and window applied by 8 x 8 transposes on the way in.  bench.py reports it as fft_kind and puts a
 * witness beside it (a batched torch.fft on the same cores).
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <atomic>
#include <chrono>
#include <complex>
#include <cstddef>
#include <thread>
#include <memory>
#include <vector>

#include "../include/ba_cuda.h"

#if defined(BA_ORACLE_FAST) && defined(__AVX2__) && defined(__FMA__)
#include <immintrin.h>
#define BA_ORACLE_LANES 8
#endif

#ifdef BA_ORACLE_REF
/* the FSM state is private in the reference and printed only under -DDEBUG_SQUELCH; the decision trace needs it */
#define private public
#include "squelch.h"
#include "filters.h"
#undef private
#include "logging.h"
#define SQ_CURRENT(s) ((int)(s).current_state_)
#else
#include "dsp_restated.h"
using ora::LowpassFilter;
using ora::NotchFilter;
using ora::Squelch;
#define SQ_CURRENT(s) ((s).current_state())
#endif

#include "oracle.h"

namespace {

/* ---------------------------------------------------------------- float32 FFT (stands in for fftwf_execute, .cpp:484) */
struct Fft {
    int n = 0;
    std::vector<float> tw; /* per stage: [p][w1.re w1.im w2.re w2.im w3.re w3.im] */
    std::vector<size_t> stage_off;
    std::vector<float> scratch;
    void plan(int size) {
        n = size;
        tw.clear();
        stage_off.clear();
        for (int len = n; len >= 4; len /= 4) {
            stage_off.push_back(tw.size());
            for (int p = 0; p < len / 4; p++)
                for (int k = 1; k <= 3; k++) {
                    double a = -2.0 * M_PI * (double)(p * k) / (double)len;
                    tw.push_back((float)cos(a));
                    tw.push_back((float)sin(a));
                }
        }
        scratch.assign(2 * (size_t)n, 0.0f);
    }
    /* in: n interleaved complex; out: n interleaved complex (distinct buffers); `in` is clobbered */
    void run(float* in, float* out) {
        int passes = 0, len = n;
        for (; len >= 4; len /= 4)
            passes++;
        if (len == 2)
            passes++;
        float* src = in;
        int s = 1, st = 0;
        len = n;
        for (; len >= 4; len /= 4, s *= 4, st++) {
            float* dst = (st == passes - 1) ? out : ((st & 1) ? in : scratch.data());
            const int q1 = len / 4;
            const float* w = tw.data() + stage_off[st];
            for (int p = 0; p < q1; p++) {
                const float w1r = w[6 * p + 0], w1i = w[6 * p + 1], w2r = w[6 * p + 2], w2i = w[6 * p + 3], w3r = w[6 * p + 4], w3i = w[6 * p + 5];
                const float* xa = src + 2 * (size_t)s * (p);
                const float* xb = src + 2 * (size_t)s * (p + q1);
                const float* xc = src + 2 * (size_t)s * (p + 2 * q1);
                const float* xd = src + 2 * (size_t)s * (p + 3 * q1);
                float* y0 = dst + 2 * (size_t)s * (4 * p + 0);
                float* y1 = dst + 2 * (size_t)s * (4 * p + 1);
                float* y2 = dst + 2 * (size_t)s * (4 * p + 2);
                float* y3 = dst + 2 * (size_t)s * (4 * p + 3);
                for (int q = 0; q < s; q++) {
                    const float ar = xa[2 * q], ai = xa[2 * q + 1], br = xb[2 * q], bi = xb[2 * q + 1];
                    const float cr = xc[2 * q], ci = xc[2 * q + 1], dr = xd[2 * q], di = xd[2 * q + 1];
                    const float apcr = ar + cr, apci = ai + ci, amcr = ar - cr, amci = ai - ci;
                    const float bpdr = br + dr, bpdi = bi + di;
                    const float jr = -(bi - di), ji = (br - dr); /* j*(b-d) */
                    y0[2 * q] = apcr + bpdr;
                    y0[2 * q + 1] = apci + bpdi;
                    const float t1r = amcr - jr, t1i = amci - ji;
                    y1[2 * q] = t1r * w1r - t1i * w1i;
                    y1[2 * q + 1] = t1r * w1i + t1i * w1r;
                    const float t2r = apcr - bpdr, t2i = apci - bpdi;
                    y2[2 * q] = t2r * w2r - t2i * w2i;
                    y2[2 * q + 1] = t2r * w2i + t2i * w2r;
                    const float t3r = amcr + jr, t3i = amci + ji;
                    y3[2 * q] = t3r * w3r - t3i * w3i;
                    y3[2 * q + 1] = t3r * w3i + t3i * w3r;
                }
            }
            src = dst;
        }
        if (len == 2) {
            float* dst = out;
            for (int q = 0; q < s; q++) {
                const float ar = src[2 * q], ai = src[2 * q + 1], br = src[2 * (q + s)], bi = src[2 * (q + s) + 1];
                dst[2 * q] = ar + br;
                dst[2 * q + 1] = ai + bi;
                dst[2 * (q + s)] = ar - br;
                dst[2 * (q + s) + 1] = ai - bi;
            }
        }
    }
};


#ifdef BA_ORACLE_LANES
/* Timing builds only: the same radix-4 Stockham transform on eight frames at once, one frame per AVX2 lane.  Element k of
 * the eight frames is one __m256 in re[] and one in im[]. */
struct Fft8 {
    int n = 0;
    std::vector<float> tw;
    std::vector<size_t> stage_off;
    float* buf = nullptr; /* 4 arrays of n vectors: A.re A.im B.re B.im */
    ~Fft8() { free(buf); }
    Fft8() = default;
    Fft8(const Fft8&) = delete;
    Fft8& operator=(const Fft8&) = delete;
    void plan(int size) {
        n = size;
        tw.clear();
        stage_off.clear();
        for (int len = n; len >= 4; len /= 4) {
            stage_off.push_back(tw.size());
            for (int p = 0; p < len / 4; p++)
                for (int k = 1; k <= 3; k++) {
                    double a = -2.0 * M_PI * (double)(p * k) / (double)len;
                    tw.push_back((float)cos(a));
                    tw.push_back((float)sin(a));
                }
        }
        free(buf);
        buf = (float*)aligned_alloc(64, sizeof(float) * 8 * (size_t)n * 4);
    }
    float* in_re() { return buf; }
    float* in_im() { return buf + 8 * (size_t)n; }
    /* transforms the frames in in_re()/in_im(); *out_re / *out_im point at the result arrays afterwards */
    void run(const float** out_re, const float** out_im) {
        float* a_re = buf;
        float* a_im = buf + 8 * (size_t)n;
        float* b_re = buf + 16 * (size_t)n;
        float* b_im = buf + 24 * (size_t)n;
        int s = 1, st = 0, len = n;
        for (; len >= 4; len /= 4, s *= 4, st++) {
            const int q1 = len / 4;
            const float* w = tw.data() + stage_off[st];
            for (int p = 0; p < q1; p++) {
                const __m256 w1r = _mm256_set1_ps(w[6 * p + 0]), w1i = _mm256_set1_ps(w[6 * p + 1]);
                const __m256 w2r = _mm256_set1_ps(w[6 * p + 2]), w2i = _mm256_set1_ps(w[6 * p + 3]);
                const __m256 w3r = _mm256_set1_ps(w[6 * p + 4]), w3i = _mm256_set1_ps(w[6 * p + 5]);
                const size_t ia = 8 * (size_t)s * p, ib = 8 * (size_t)s * (p + q1), ic = 8 * (size_t)s * (p + 2 * q1), id = 8 * (size_t)s * (p + 3 * q1);
                const size_t o0 = 8 * (size_t)s * (4 * p);
                for (int q = 0; q < s; q++) {
                    const size_t e = 8 * (size_t)q;
                    const __m256 ar = _mm256_load_ps(a_re + ia + e), ai = _mm256_load_ps(a_im + ia + e);
                    const __m256 br = _mm256_load_ps(a_re + ib + e), bi = _mm256_load_ps(a_im + ib + e);
                    const __m256 cr = _mm256_load_ps(a_re + ic + e), ci = _mm256_load_ps(a_im + ic + e);
                    const __m256 dr = _mm256_load_ps(a_re + id + e), di = _mm256_load_ps(a_im + id + e);
                    const __m256 apcr = _mm256_add_ps(ar, cr), apci = _mm256_add_ps(ai, ci), amcr = _mm256_sub_ps(ar, cr), amci = _mm256_sub_ps(ai, ci);
                    const __m256 bpdr = _mm256_add_ps(br, dr), bpdi = _mm256_add_ps(bi, di);
                    const __m256 jr = _mm256_sub_ps(di, bi), ji = _mm256_sub_ps(br, dr); /* j * (b - d) */
                    _mm256_store_ps(b_re + o0 + e, _mm256_add_ps(apcr, bpdr));
                    _mm256_store_ps(b_im + o0 + e, _mm256_add_ps(apci, bpdi));
                    const __m256 t1r = _mm256_sub_ps(amcr, jr), t1i = _mm256_sub_ps(amci, ji);
                    _mm256_store_ps(b_re + o0 + 8 * (size_t)s + e, _mm256_fmsub_ps(t1r, w1r, _mm256_mul_ps(t1i, w1i)));
                    _mm256_store_ps(b_im + o0 + 8 * (size_t)s + e, _mm256_fmadd_ps(t1r, w1i, _mm256_mul_ps(t1i, w1r)));
                    const __m256 t2r = _mm256_sub_ps(apcr, bpdr), t2i = _mm256_sub_ps(apci, bpdi);
                    _mm256_store_ps(b_re + o0 + 16 * (size_t)s + e, _mm256_fmsub_ps(t2r, w2r, _mm256_mul_ps(t2i, w2i)));
                    _mm256_store_ps(b_im + o0 + 16 * (size_t)s + e, _mm256_fmadd_ps(t2r, w2i, _mm256_mul_ps(t2i, w2r)));
                    const __m256 t3r = _mm256_add_ps(amcr, jr), t3i = _mm256_add_ps(amci, ji);
                    _mm256_store_ps(b_re + o0 + 24 * (size_t)s + e, _mm256_fmsub_ps(t3r, w3r, _mm256_mul_ps(t3i, w3i)));
                    _mm256_store_ps(b_im + o0 + 24 * (size_t)s + e, _mm256_fmadd_ps(t3r, w3i, _mm256_mul_ps(t3i, w3r)));
                }
            }
            std::swap(a_re, b_re);
            std::swap(a_im, b_im);
        }
        if (len == 2) {
            for (int q = 0; q < s; q++) {
                const size_t e = 8 * (size_t)q, f = 8 * (size_t)(q + s);
                const __m256 ar = _mm256_load_ps(a_re + e), ai = _mm256_load_ps(a_im + e), br = _mm256_load_ps(a_re + f), bi = _mm256_load_ps(a_im + f);
                _mm256_store_ps(b_re + e, _mm256_add_ps(ar, br));
                _mm256_store_ps(b_im + e, _mm256_add_ps(ai, bi));
                _mm256_store_ps(b_re + f, _mm256_sub_ps(ar, br));
                _mm256_store_ps(b_im + f, _mm256_sub_ps(ai, bi));
            }
            std::swap(a_re, b_re);
            std::swap(a_im, b_im);
        }
        *out_re = a_re;
        *out_im = a_im;
    }
};

/* 8 x 8 transpose: rows r[0..7] (one frame each, eight consecutive samples) -> columns (one sample each, eight frames) */
static inline void transpose8(__m256* r) {
    const __m256 t0 = _mm256_unpacklo_ps(r[0], r[1]), t1 = _mm256_unpackhi_ps(r[0], r[1]);
    const __m256 t2 = _mm256_unpacklo_ps(r[2], r[3]), t3 = _mm256_unpackhi_ps(r[2], r[3]);
    const __m256 t4 = _mm256_unpacklo_ps(r[4], r[5]), t5 = _mm256_unpackhi_ps(r[4], r[5]);
    const __m256 t6 = _mm256_unpacklo_ps(r[6], r[7]), t7 = _mm256_unpackhi_ps(r[6], r[7]);
    const __m256 u0 = _mm256_shuffle_ps(t0, t2, 0x44), u1 = _mm256_shuffle_ps(t0, t2, 0xee);
    const __m256 u2 = _mm256_shuffle_ps(t1, t3, 0x44), u3 = _mm256_shuffle_ps(t1, t3, 0xee);
    const __m256 u4 = _mm256_shuffle_ps(t4, t6, 0x44), u5 = _mm256_shuffle_ps(t4, t6, 0xee);
    const __m256 u6 = _mm256_shuffle_ps(t5, t7, 0x44), u7 = _mm256_shuffle_ps(t5, t7, 0xee);
    r[0] = _mm256_permute2f128_ps(u0, u4, 0x20);
    r[1] = _mm256_permute2f128_ps(u1, u5, 0x20);
    r[2] = _mm256_permute2f128_ps(u2, u6, 0x20);
    r[3] = _mm256_permute2f128_ps(u3, u7, 0x20);
    r[4] = _mm256_permute2f128_ps(u0, u4, 0x31);
    r[5] = _mm256_permute2f128_ps(u1, u5, 0x31);
    r[6] = _mm256_permute2f128_ps(u2, u6, 0x31);
    r[7] = _mm256_permute2f128_ps(u3, u7, 0x31);
}
#endif

/* ---------------------------------------------------------------- helpers restated from the reference */

/* util.cpp:103-127 — 256-entry sine/cosine table, argument formed in double and rounded to float, float sincosf */
struct SinCosLut {
    float s[257], c[257];
    SinCosLut() {
        for (uint32_t i = 0; i < 256; i++)
            sincosf((float)(2.0F * M_PI * (float)i / 256.0f), &s[i], &c[i]);
        s[256] = s[0];
        c[256] = c[0];
    }
    void get(uint32_t phi, float* sine, float* cosine) const {
        uint32_t idx = phi >> 16;
        float fract = (float)(phi & 0xffff) / 65536.0f;
        *sine = s[idx] + (s[idx + 1] - s[idx]) * fract;
        *cosine = c[idx] + (c[idx + 1] - c[idx]) * fract;
    }
};
const SinCosLut g_lut;

/* boondock_airband.cpp:147-176 */
inline float atan2_approx(float y, float x) {
    const float pi4 = (float)M_PI_4, pi34 = (float)(3 * M_PI_4);
    if (x == 0.0f && y == 0.0f)
        return 0;
    float ya = y < 0.0f ? -y : y;
    float ang = (x >= 0.0f) ? pi4 - pi4 * (x - ya) / (x + ya) : pi34 - pi4 * (x + ya) / (ya - x);
    return y < 0.0f ? -ang : ang;
}
inline float disc_polar(float ar, float aj, float br, float bj) {
    /* a * conj(b), boondock_airband.cpp:141-144,168-172 */
    float cr = ar * br - aj * (-bj);
    float cj = aj * br + ar * (-bj);
    return (float)(atan2_approx(cj, cr) * M_1_PI);
}
inline float disc_quadri(float ar, float aj, float br, float bj) {
    return (float)((br * aj - ar * bj) / (ar * ar + aj * aj + 1.0f) * M_1_PI);
}

/* util.cpp:169-176 */
inline float dbfs_offset(int fft_size) {
    return 7.54f + 10.0f * log10f((float)(size_t)(fft_size / 2)) - 2.38f;
}
inline float dbfs_to_level(float dbfs, int fft_size) {
    return (float)(pow(10.0, (dbfs - dbfs_offset(fft_size)) / 20.0f) * (size_t)fft_size);
}

/* freq_t (boondock_airband.h:215-230): what scan mode keeps once per frequency of a channel (config.cpp:364-433) */
struct FreqParms {
    int modulation = 0;
    float agcavgfast = 0.5f, ampfactor = 1.0f;
    uint32_t active_counter = 0;
    Squelch squelch;
    NotchFilter notch;
    LowpassFilter lowpass;
};

struct Channel {
    ba_channel_desc cfg;
    int afc = 0;
    int needs_raw_iq = 0, has_iq_outputs = 0;
    uint32_t dm_dphi = 0, dm_phi = 0;
    float alpha = 0, pr = 0, pj = 0, prev_waveout = 0.5f;
    int axcindicate = BA_NO_SIGNAL;
    std::vector<std::unique_ptr<FreqParms>> freqs; /* channel_t.freqlist; one entry in multichannel mode */
    FreqParms* f = nullptr;                        /* fparms = freqlist + freq_idx, boondock_airband.cpp:522 */
    int freq_idx = 0;
    std::vector<std::pair<uint64_t, int>> freq_plan; /* tests: (batch number, freq_idx) the controller thread would have set by then */
    std::vector<float> wavein, waveout, iq_in, iq_out;
    ba_channel_info info;
    /* recorded streams */
    std::vector<float> rec_wave, rec_iq, rec_picks;
    std::vector<uint8_t> rec_trace;
    std::vector<ba_channel_status> rec_status;
    double checksum = 0.0;
};

struct Device {
    ba_device_desc cfg;
    size_t bps = 0;
    std::vector<size_t> bins, base_bins;
    int waveend = 0;
    std::vector<Channel> ch;
    std::vector<uint8_t> pend;
    uint64_t frames = 0, batches = 0;
    std::vector<float> fftin, fftout;
    Fft fft;
#ifdef BA_ORACLE_LANES
    std::unique_ptr<Fft8> fft8;
#endif
    float scale = 1.0f;
};

}  // namespace

struct ba_oracle {
    int fft_size = 0, wave_rate = 0, wave_batch = 0, fm_demod = 0;
    uint32_t flags = 0;
    bool keep = true;
    std::vector<float> window;
    float levels_u8[256], levels_s8[256];
    std::vector<Device> dev;
};

namespace {

/* boondock_airband.cpp:180-251 (class AFC) as a function; `prev` is the indicator snapshot taken before the batch */
size_t afc_walk(const float* fft, size_t n, size_t base, float base_value, unsigned char afc, int step) {
    float threshold = 0;
    size_t bin;
    auto power = [&](size_t i) { return fft[2 * i] * fft[2 * i] + fft[2 * i + 1] * fft[2 * i + 1]; };
    for (bin = base;; bin += step) {
        if (step < 0) {
            if (bin < (size_t)(-step))
                break;
        } else if ((size_t)(bin + step) >= n)
            break;
        const float value = power((size_t)(bin + step));
        if (value <= base_value)
            break;
        if (base == bin) {
            threshold = (value - base_value) / (float)afc;
        } else {
            if ((value - base_value) < threshold)
                break;
            threshold += threshold / 10.0;
        }
    }
    return bin;
}
void afc_finalize(Device& d, int i, int prev, const float* fft, size_t n) {
    Channel& c = d.ch[i];
    if (c.afc == 0)
        return;
    const int now = c.axcindicate;
    if (now != BA_NO_SIGNAL && prev == BA_NO_SIGNAL) {
        const size_t base = d.base_bins[i];
        const float base_value = fft[2 * base] * fft[2 * base] + fft[2 * base + 1] * fft[2 * base + 1];
        size_t bin = afc_walk(fft, n, base, base_value, (unsigned char)c.afc, -1);
        if (bin == base)
            bin = afc_walk(fft, n, base, base_value, (unsigned char)c.afc, 1);
        if (d.bins[i] != bin) {
            d.bins[i] = bin;
            if (bin > base)
                c.axcindicate = BA_AFC_UP;
            else if (bin < base)
                c.axcindicate = BA_AFC_DOWN;
        }
    } else if (now == BA_NO_SIGNAL && prev != BA_NO_SIGNAL)
        d.bins[i] = d.base_bins[i];
}

/* one batch of the per-channel loop, boondock_airband.cpp:518-679 */
void run_batch(ba_oracle* o, Device& d) {
    const int B = o->wave_batch, E = BA_AGC_EXTRA;
    for (int i = 0; i < (int)d.ch.size(); i++) {
        Channel& c = d.ch[i];
        for (const auto& pl : c.freq_plan) /* channel->freq_idx as the controller thread left it (.cpp:101-139) */
            if (pl.first == d.batches)
                c.freq_idx = pl.second;
        c.f = c.freqs[c.freq_idx].get(); /* .cpp:522 */
        const int prev_axc = c.axcindicate; /* AFC afc(dev, i) */
        c.axcindicate = BA_NO_SIGNAL;
        float* wavein = c.wavein.data();
        float* wout = c.waveout.data();
        for (int j = E; j < B + E; j++) {
            float& real = c.iq_in[2 * (j - E)];
            float& imag = c.iq_in[2 * (j - E) + 1];
            uint8_t tr = 0;

            c.f->squelch.process_raw_sample(wavein[j]);

            if (c.f->squelch.should_filter_sample() && c.needs_raw_iq) {
                float swf, cwf, re_tmp, im_tmp;
                g_lut.get(c.dm_phi, &swf, &cwf);
                /* multiply(real, imag, cwf, -swf, ...) .cpp:141-144,538 */
                re_tmp = real * cwf - imag * (-swf);
                im_tmp = imag * cwf + real * (-swf);
                c.dm_phi += c.dm_dphi;
                c.dm_phi &= 0xffffff;
                c.f->lowpass.apply(re_tmp, im_tmp);
                real = re_tmp;
                imag = im_tmp;
                wavein[j] = sqrtf(real * real + imag * imag); /* sqrt(float) -> float overload, .cpp:548 */
                if (c.f->lowpass.enabled())
                    c.f->squelch.process_filtered_sample(wavein[j]);
                tr |= BA_TRACE_FILTERED;
            }

            if (c.f->modulation == BA_MOD_AM) {
                if (c.f->squelch.first_open_sample()) {
                    for (int k = j - E; k < j; k++)
                        if (wavein[k] >= c.f->squelch.squelch_level())
                            c.f->agcavgfast = c.f->agcavgfast * 0.9f + wavein[k] * 0.1f;
                } else if (c.f->squelch.last_open_sample()) {
                    for (int k = j - E + 1; k < j; k++)
                        wout[k] = wout[k - 1] * 0.94f;
                }
            }

            float& waveout = wout[j];
            if (c.f->squelch.should_process_audio()) {
                if (c.f->modulation == BA_MOD_AM) {
                    if (wavein[j] > c.f->squelch.squelch_level())
                        c.f->agcavgfast = c.f->agcavgfast * 0.995f + wavein[j] * 0.005f;
                    waveout = (wavein[j - E] - c.f->agcavgfast) / (c.f->agcavgfast * 1.5f);
                    if (fabsf(waveout) > 0.8f) {
                        waveout *= 0.85f;
                        c.f->agcavgfast *= 1.15f;
                    }
                } else if (c.f->modulation == BA_MOD_NFM) {
                    if (o->fm_demod == BA_FM_FAST_ATAN2)
                        waveout = disc_polar(real, imag, c.pr, c.pj);
                    else
                        waveout = disc_quadri(real, imag, c.pr, c.pj);
                    c.pr = real;
                    c.pj = imag;
                    c.f->agcavgfast = c.f->agcavgfast * 0.995f + waveout * 0.005f;
                    waveout -= c.f->agcavgfast;
                    waveout = waveout * (1.0f - c.alpha) + c.prev_waveout * c.alpha;
                    c.prev_waveout = waveout;
                }
                c.f->squelch.process_audio_sample(waveout);
                tr |= BA_TRACE_AUDIO;
            }

            if (c.f->squelch.is_open()) {
                c.f->notch.apply(waveout);
                waveout *= c.f->ampfactor;
                if (isnan(waveout))
                    waveout = 0.0;
                else if (waveout > 1.0)
                    waveout = 1.0;
                else if (waveout < -1.0)
                    waveout = -1.0;
                c.axcindicate = BA_SIGNAL;
                if (c.has_iq_outputs) {
                    c.iq_out[2 * (j - E)] = real;
                    c.iq_out[2 * (j - E) + 1] = imag;
                }
                tr |= BA_TRACE_OPEN;
            } else {
                waveout = 0;
                if (c.has_iq_outputs) {
                    c.iq_out[2 * (j - E)] = 0;
                    c.iq_out[2 * (j - E) + 1] = 0;
                }
            }
            if (o->keep && (o->flags & BA_FLAG_TRACE))
                c.rec_trace.push_back((uint8_t)(tr | (SQ_CURRENT(c.f->squelch) & BA_TRACE_STATE_MASK)));
        }
        memmove(wavein, wavein + B, (d.waveend - B) * sizeof(float));
        if (c.needs_raw_iq)
            memmove(c.iq_in.data(), c.iq_in.data() + 2 * B, (d.waveend - B) * sizeof(float) * 2);

        afc_finalize(d, i, prev_axc, d.fftout.data(), (size_t)o->fft_size);

        if (c.axcindicate != BA_NO_SIGNAL)
            c.f->active_counter++;

        /* what output_thread does with the batch (output.cpp:931-951): consume waveout[0..B), keep the tail */
        if (o->keep) {
            c.rec_wave.insert(c.rec_wave.end(), wout, wout + B);
            if (c.has_iq_outputs)
                c.rec_iq.insert(c.rec_iq.end(), c.iq_out.begin(), c.iq_out.begin() + 2 * B);
            ba_channel_status st;
            st.axcindicate = c.axcindicate;
            st.bin = (uint32_t)d.bins[i];
            st.signal_level = c.f->squelch.signal_level();
            st.noise_level = c.f->squelch.noise_level();
            st.squelch_level = c.f->squelch.squelch_level();
            st.open_count = (uint32_t)c.f->squelch.open_count();
            st.flappy_count = (uint32_t)c.f->squelch.flappy_count();
            st.ctcss_count = (uint32_t)c.f->squelch.ctcss_count();
            st.no_ctcss_count = (uint32_t)c.f->squelch.no_ctcss_count();
            st.active_counter = c.f->active_counter;
            c.rec_status.push_back(st);
        } else {
            double s = 0;
            for (int k = 0; k < B; k++)
                s += wout[k];
            c.checksum += s;
        }
        memcpy(wout, wout + B, E * sizeof(float));
    }
    d.waveend -= B;
    d.batches++;
}

/* boondock_airband.cpp:426-516 for one frame at `buf` */
inline void one_frame(ba_oracle* o, Device& d, const unsigned char* buf) {
    const int N = o->fft_size;
    float* fftin = d.fftin.data();
    const float* window = o->window.data();
    if (d.cfg.sample_format == BA_SFMT_S16) {
        const float scale = d.scale;
        const short* b2 = (const short*)buf;
        for (int i = 0; i < N; i++, b2 += 2) {
            fftin[2 * i] = scale * (float)b2[0] * window[i];
            fftin[2 * i + 1] = scale * (float)b2[1] * window[i];
        }
    } else if (d.cfg.sample_format == BA_SFMT_F32) {
        const float scale = d.scale;
        const float* b2 = (const float*)buf;
        for (int i = 0; i < N; i++, b2 += 2) {
            fftin[2 * i] = scale * b2[0] * window[i];
            fftin[2 * i + 1] = scale * b2[1] * window[i];
        }
    } else {
        const float* lv = d.cfg.sample_format == BA_SFMT_U8 ? o->levels_u8 : o->levels_s8;
        for (int i = 0; i < N; i++, buf += 2) {
            fftin[2 * i] = lv[buf[0]] * window[i];
            fftin[2 * i + 1] = lv[buf[1]] * window[i];
        }
    }
    d.fft.run(fftin, d.fftout.data());
    const float* fo = d.fftout.data();
    for (size_t j = 0; j < d.ch.size(); j++) {
        Channel& c = d.ch[j];
        const size_t bin = d.bins[j];
        c.wavein[d.waveend] = sqrtf(fo[2 * bin] * fo[2 * bin] + fo[2 * bin + 1] * fo[2 * bin + 1]);
        if (c.needs_raw_iq) {
            c.iq_in[2 * d.waveend] = fo[2 * bin];
            c.iq_in[2 * d.waveend + 1] = fo[2 * bin + 1];
        }
        if (o->keep && (o->flags & BA_FLAG_TRACE)) {
            c.rec_picks.push_back(fo[2 * bin]);
            c.rec_picks.push_back(fo[2 * bin + 1]);
        }
    }
    d.waveend += 1;
    d.frames++;
    if (d.waveend >= o->wave_batch + BA_AGC_EXTRA)
        run_batch(o, d);
}

#ifdef BA_ORACLE_LANES
/* Timing builds only: boondock_airband.cpp:426-516 for EIGHT consecutive frames starting at `buf` (frame f at buf + f * bps).
 * The caller has made sure that no batch ends inside the group. */
inline void eight_frames(ba_oracle* o, Device& d, const unsigned char* buf) {
    const int N = o->fft_size;
    Fft8& f8 = *d.fft8;
    float* re = f8.in_re();
    float* im = f8.in_im();
    const float* window = o->window.data();
    const __m256i unzip = _mm256_setr_epi32(0, 1, 4, 5, 2, 3, 6, 7);
    const int fmt = d.cfg.sample_format;
    const __m256 scale = _mm256_set1_ps(d.scale);
    for (int n0 = 0; n0 < N; n0 += 8) {
        __m256 r[8], q[8];
        const __m256 w = _mm256_loadu_ps(window + n0);
        for (int f = 0; f < 8; f++) {
            __m256 lo, hi; /* I0 Q0 I1 Q1 I2 Q2 I3 Q3 and I4 Q4 .. Q7 of frame f, samples n0 .. n0 + 7 */
            if (fmt == BA_SFMT_U8) {
                const __m128i v = _mm_loadu_si128((const __m128i*)(buf + (size_t)f * d.bps + 2 * (size_t)n0));
                const __m256 k = _mm256_set1_ps(1.0f / 127.5f), one = _mm256_set1_ps(1.0f);
                lo = _mm256_fmsub_ps(_mm256_cvtepi32_ps(_mm256_cvtepu8_epi32(v)), k, one); /* (i - 127.5) / 127.5, .cpp:341-343 */
                hi = _mm256_fmsub_ps(_mm256_cvtepi32_ps(_mm256_cvtepu8_epi32(_mm_srli_si128(v, 8))), k, one);
            } else if (fmt == BA_SFMT_S8) {
                const __m128i v = _mm_loadu_si128((const __m128i*)(buf + (size_t)f * d.bps + 2 * (size_t)n0));
                const __m256 k = _mm256_set1_ps(1.0f / 128.0f);
                lo = _mm256_mul_ps(_mm256_cvtepi32_ps(_mm256_cvtepi8_epi32(v)), k);
                hi = _mm256_mul_ps(_mm256_cvtepi32_ps(_mm256_cvtepi8_epi32(_mm_srli_si128(v, 8))), k);
            } else if (fmt == BA_SFMT_S16) {
                const __m128i* b = (const __m128i*)(buf + (size_t)f * d.bps + 4 * (size_t)n0);
                lo = _mm256_mul_ps(_mm256_cvtepi32_ps(_mm256_cvtepi16_epi32(_mm_loadu_si128(b))), scale);
                hi = _mm256_mul_ps(_mm256_cvtepi32_ps(_mm256_cvtepi16_epi32(_mm_loadu_si128(b + 1))), scale);
            } else {
                const float* b = (const float*)(buf + (size_t)f * d.bps + 8 * (size_t)n0);
                lo = _mm256_mul_ps(_mm256_loadu_ps(b), scale);
                hi = _mm256_mul_ps(_mm256_loadu_ps(b + 8), scale);
            }
            /* de-interleave: even elements are I, odd are Q */
            const __m256 ev = _mm256_permutevar8x32_ps(_mm256_shuffle_ps(lo, hi, 0x88), unzip);
            const __m256 od = _mm256_permutevar8x32_ps(_mm256_shuffle_ps(lo, hi, 0xdd), unzip);
            r[f] = _mm256_mul_ps(ev, w);
            q[f] = _mm256_mul_ps(od, w);
        }
        transpose8(r);
        transpose8(q);
        for (int k = 0; k < 8; k++) {
            _mm256_store_ps(re + 8 * (size_t)(n0 + k), r[k]);
            _mm256_store_ps(im + 8 * (size_t)(n0 + k), q[k]);
        }
    }
    const float *ore, *oim;
    f8.run(&ore, &oim);
    for (size_t j = 0; j < d.ch.size(); j++) {
        Channel& c = d.ch[j];
        const size_t bin = d.bins[j];
        const __m256 xr = _mm256_load_ps(ore + 8 * bin), xi = _mm256_load_ps(oim + 8 * bin);
        _mm256_storeu_ps(&c.wavein[d.waveend], _mm256_sqrt_ps(_mm256_fmadd_ps(xr, xr, _mm256_mul_ps(xi, xi))));
        if (c.needs_raw_iq) {
            alignas(32) float a[8], b[8];
            _mm256_store_ps(a, xr);
            _mm256_store_ps(b, xi);
            for (int f = 0; f < 8; f++) {
                c.iq_in[2 * (d.waveend + f)] = a[f];
                c.iq_in[2 * (d.waveend + f) + 1] = b[f];
            }
        }
    }
    d.waveend += 8;
    d.frames += 8;
    if (d.waveend >= o->wave_batch + BA_AGC_EXTRA)
        run_batch(o, d);
}
#endif

/* consume as many frames as the reference's availability test allows (.cpp:418-424) from [p, p+len) */
size_t consume(ba_oracle* o, Device& d, const unsigned char* p, size_t len) {
    const size_t need = d.bps + (size_t)o->fft_size * d.cfg.bytes_per_sample * 2;
    size_t off = 0;
    while (len - off >= need) {
#ifdef BA_ORACLE_LANES
        /* timing builds: eight frames at a time where nothing is recorded, no batch ends inside the group and AFC does not need
         * the spectrum of a batch's last frame */
        if (d.fft8 && !o->keep && len - off >= need + 7 * d.bps && d.waveend + 8 <= o->wave_batch + BA_AGC_EXTRA) {
            eight_frames(o, d, p + off);
            off += 8 * d.bps;
            continue;
        }
#endif
        one_frame(o, d, p + off);
        off += d.bps;
    }
    return off;
}

int fill_channel(ba_oracle* o, Device& d, Channel& c, const ba_channel_desc& cd) {
    const int R = o->wave_rate, N = o->fft_size;
    const int B = o->wave_batch, E = BA_AGC_EXTRA;
    c.cfg = cd;
    c.afc = cd.afc & 0xff;
    c.has_iq_outputs = cd.has_iq_outputs ? 1 : 0;
    /* the frequency list: scan mode names it (config.cpp:364-433); otherwise the channel's own fields are freqlist[0] */
    std::vector<ba_freq_desc> fl;
    if (cd.freq_count >= 1 && cd.freqs)
        fl.assign(cd.freqs, cd.freqs + cd.freq_count);
    else
        fl.push_back(ba_freq_desc{cd.frequency, cd.modulation, cd.ampfactor, cd.squelch_threshold_dbfs, cd.squelch_snr_threshold, cd.notch, cd.notch_q, cd.ctcss, cd.bandwidth});
    const int frequency0 = fl[0].frequency; /* bin and dm_dphi come from freqlist[0] only (config.cpp:669,684) */
    c.needs_raw_iq = cd.has_iq_outputs ? 1 : 0;
    for (const ba_freq_desc& fd : fl) /* config.cpp:162,596,674-680 */
        if (fd.modulation == BA_MOD_NFM || fd.bandwidth != 0)
            c.needs_raw_iq = 1;
    c.cfg.freqs = nullptr;
    c.wavein.assign(2 * B + E, 0.0f);
    c.waveout.assign(2 * B + E, 0.0f);
    c.iq_in.assign(2 * (2 * B + E), 0.0f);
    c.iq_out.assign(2 * (2 * B + E), 0.0f);
    for (int k = 0; k < E; k++) { /* config.cpp:319-322 */
        c.wavein[k] = 20;
        c.waveout[k] = 0.5;
    }
    /* de-emphasis constant: global default .cpp:87, device override config.cpp:777-781, channel override config.cpp:650-652 */
    float alpha = (float)exp(-1.0f / (R * 2e-4));
    if (d.cfg.tau_us >= 0)
        alpha = d.cfg.tau_us == 0 ? 0.0f : (float)exp(-1.0f / (R * 1e-6 * d.cfg.tau_us));
    if (cd.tau_us >= 0)
        alpha = cd.tau_us == 0 ? 0.0f : (float)exp(-1.0f / (R * 1e-6 * cd.tau_us));
    c.alpha = alpha;

    /* squelch settings in the order parse_channels applies them (config.cpp:437-515), per frequency */
    for (const ba_freq_desc& fd : fl) {
        std::unique_ptr<FreqParms> fp(new FreqParms());
        fp->modulation = fd.modulation;
        fp->ampfactor = fd.ampfactor;
        if (fd.squelch_threshold_dbfs < 0)
            fp->squelch.set_squelch_level_threshold(dbfs_to_level((float)fd.squelch_threshold_dbfs, N));
        if (fd.squelch_snr_threshold >= 0)
            fp->squelch.set_squelch_snr_threshold(fd.squelch_snr_threshold);
        if (fd.notch > 0) {
            float q = fd.notch_q == 0.0f ? 10.0f : fd.notch_q;
            fp->notch = NotchFilter(fd.notch, (float)R, q);
        }
        if (fd.ctcss > 0)
            fp->squelch.set_ctcss_freq(fd.ctcss, (float)R);
        if (fd.bandwidth > 0)
            fp->lowpass = LowpassFilter((float)fd.bandwidth / 2, (float)R);
        c.freqs.push_back(std::move(fp));
    }
    c.f = c.freqs[0].get();

    /* bin, config.cpp:669-670: Fs / N is an integer division */
    const int fs = d.cfg.sample_rate, cf = d.cfg.centerfreq;
    size_t bin = (size_t)ceil((frequency0 + fs - cf) / (double)((size_t)fs / (size_t)N) - 1.0) % (size_t)N;
    d.bins.push_back(bin);
    d.base_bins.push_back(bin);

    if (c.needs_raw_iq) { /* config.cpp:682-715 */
        double dm = (double)(frequency0 - cf);
        double decim = ((double)fs / (double)R);
        double corr = (double)R / 2.0;
        corr *= (decim - round(decim));
        corr *= (double)(frequency0 - cf) / ((double)fs / 2.0);
        dm -= corr;
        dm /= (double)R;
        dm -= trunc(dm);
        dm *= 256.0 * 65536.0;
        c.dm_dphi = (uint32_t)((int)dm);
        c.dm_phi = 0;
    }

    ba_channel_info& in = c.info;
    memset(&in, 0, sizeof(in));
    in.bin = (uint32_t)bin;
    in.dm_dphi = c.dm_dphi;
    in.needs_raw_iq = c.needs_raw_iq;
    in.alpha = c.alpha;
#ifdef BA_ORACLE_REF
    in.squelch_ratio = c.f->squelch.normal_signal_ratio_;
    in.manual_level = c.f->squelch.using_manual_level_ ? c.f->squelch.manual_signal_level_ : 0.0f;
    in.notch_enabled = c.f->notch.enabled_;
    if (in.notch_enabled)
        memcpy(in.notch_d, c.f->notch.d, sizeof(in.notch_d));
    in.lowpass_enabled = c.f->lowpass.enabled_;
    if (in.lowpass_enabled) {
        in.lowpass_ycoeffs[0] = c.f->lowpass.ycoeffs[0];
        in.lowpass_ycoeffs[1] = c.f->lowpass.ycoeffs[1];
        in.lowpass_gain = c.f->lowpass.gain;
    }
    if (fl[0].ctcss > 0) {
        in.ctcss_fast_tones = (int32_t)c.f->squelch.ctcss_fast_.powers_.tones_.size();
        in.ctcss_slow_tones = (int32_t)c.f->squelch.ctcss_slow_.powers_.tones_.size();
        in.ctcss_fast_window = c.f->squelch.ctcss_fast_.window_size_;
        in.ctcss_slow_window = c.f->squelch.ctcss_slow_.window_size_;
    }
#else
    in.squelch_ratio = c.f->squelch.ratio();
    in.manual_level = c.f->squelch.manual();
    in.notch_enabled = c.f->notch.enabled();
    if (in.notch_enabled)
        memcpy(in.notch_d, c.f->notch.coeffs(), sizeof(in.notch_d));
    in.lowpass_enabled = c.f->lowpass.enabled();
    if (in.lowpass_enabled) {
        in.lowpass_ycoeffs[0] = c.f->lowpass.ycoeffs()[0];
        in.lowpass_ycoeffs[1] = c.f->lowpass.ycoeffs()[1];
        in.lowpass_gain = c.f->lowpass.gain();
    }
    if (fl[0].ctcss > 0) {
        in.ctcss_fast_tones = c.f->squelch.fast_tones();
        in.ctcss_slow_tones = c.f->squelch.slow_tones();
        in.ctcss_fast_window = c.f->squelch.fast_window();
        in.ctcss_slow_window = c.f->squelch.slow_window();
    }
#endif
    return 0;
}

}  // namespace

extern "C" {

int ba_oracle_is_reference_build(void) {
#ifdef BA_ORACLE_REF
    return 1;
#else
    return 0;
#endif
}

int ba_oracle_create(const ba_engine_desc* desc, int keep, ba_oracle** out) {
    if (!desc || !out)
        return BA_ERR_BAD_ARG;
    const int N = desc->fft_size;
    if (N < 256 || N > 8192 || (N & (N - 1)))
        return BA_ERR_BAD_SIZE;
    if (desc->wave_rate <= 0 || desc->wave_rate % 8)
        return BA_ERR_BAD_ARG;
#ifdef BA_ORACLE_REF
    log_destination = NONE;
#endif
    ba_oracle* o = new ba_oracle();
    o->fft_size = N;
    o->wave_rate = desc->wave_rate;
    o->wave_batch = desc->wave_rate / 8;
    o->fm_demod = desc->fm_demod;
    o->flags = desc->flags;
    o->keep = keep != 0;
    /* sample expansion tables, .cpp:338-346 (entry 128 of the s8 table is never written by the reference) */
    for (int i = 0; i < 256; i++)
        o->levels_u8[i] = (i - 127.5f) / 127.5f;
    o->levels_s8[128] = -1.0f;
    for (int i = -127; i < 128; i++)
        o->levels_s8[(uint8_t)i] = i / 128.0f;
    /* window, .cpp:357-373: float literals widened to double, evaluated in double, stored as float */
    const double a0 = 0.27105140069342f, a1 = 0.43329793923448f, a2 = 0.21812299954311f, a3 = 0.06592544638803f;
    const double a4 = 0.01081174209837f, a5 = 0.00077658482522f, a6 = 0.00001388721735f;
    o->window.resize(N);
    for (size_t i = 0; i < (size_t)N; i++) {
        double x = a0 - (a1 * cos((2.0 * M_PI * i) / (N - 1))) + (a2 * cos((4.0 * M_PI * i) / (N - 1))) - (a3 * cos((6.0 * M_PI * i) / (N - 1))) +
                   (a4 * cos((8.0 * M_PI * i) / (N - 1))) - (a5 * cos((10.0 * M_PI * i) / (N - 1))) + (a6 * cos((12.0 * M_PI * i) / (N - 1)));
        o->window[i] = (float)x;
    }
    o->dev.resize(desc->device_count);
    for (int di = 0; di < desc->device_count; di++) {
        Device& d = o->dev[di];
        d.cfg = desc->devices[di];
        d.bps = 2 * (size_t)d.cfg.bytes_per_sample * (size_t)round((double)d.cfg.sample_rate / (double)o->wave_rate); /* .cpp:418 */
        d.scale = 1.0f / d.cfg.fullscale;
        d.fftin.assign(2 * (size_t)N, 0.0f);
        d.fftout.assign(2 * (size_t)N, 0.0f);
        d.fft.plan(N);
        d.ch.resize(d.cfg.channel_count);
        for (int ci = 0; ci < d.cfg.channel_count; ci++)
            fill_channel(o, d, d.ch[ci], d.cfg.channels[ci]);
#ifdef BA_ORACLE_LANES
        {
            bool afc = false;
            for (const Channel& c : d.ch)
                afc = afc || c.afc != 0;
            if (!afc && (N % 8) == 0) {
                d.fft8.reset(new Fft8());
                d.fft8->plan(N);
            }
        }
#endif
        d.cfg.channels = nullptr;
    }
    *out = o;
    return BA_OK;
}

void ba_oracle_destroy(ba_oracle* o) {
    delete o;
}

int ba_oracle_feed(ba_oracle* o, int dev, const void* iq, size_t bytes) {
    if (!o || dev < 0 || dev >= (int)o->dev.size())
        return BA_ERR_BAD_ARG;
    Device& d = o->dev[dev];
    const unsigned char* p = (const unsigned char*)iq;
    if (d.pend.empty()) {
        size_t used = consume(o, d, p, bytes);
        d.pend.assign(p + used, p + bytes);
    } else {
        d.pend.insert(d.pend.end(), p, p + bytes);
        size_t used = consume(o, d, d.pend.data(), d.pend.size());
        d.pend.erase(d.pend.begin(), d.pend.begin() + used);
    }
    return BA_OK;
}

uint64_t ba_oracle_frames(ba_oracle* o, int dev) {
    return o->dev[dev].frames;
}
uint64_t ba_oracle_batches(ba_oracle* o, int dev) {
    return o->dev[dev].batches;
}
double ba_oracle_checksum(ba_oracle* o, int dev, int ch) {
    return o->dev[dev].ch[ch].checksum;
}

static size_t copy_out(const void* src, size_t have, void* dst, size_t want, size_t elem) {
    size_t n = have < want ? have : want;
    if (dst && n)
        memcpy(dst, src, n * elem);
    return have;
}
size_t ba_oracle_waveout(ba_oracle* o, int dev, int ch, float* out, size_t count) {
    Channel& c = o->dev[dev].ch[ch];
    return copy_out(c.rec_wave.data(), c.rec_wave.size(), out, count, sizeof(float));
}
size_t ba_oracle_iq_out(ba_oracle* o, int dev, int ch, float* out, size_t count) {
    Channel& c = o->dev[dev].ch[ch];
    return copy_out(c.rec_iq.data(), c.rec_iq.size(), out, count, sizeof(float));
}
size_t ba_oracle_picks(ba_oracle* o, int dev, int ch, float* out, size_t count) {
    Channel& c = o->dev[dev].ch[ch];
    return copy_out(c.rec_picks.data(), c.rec_picks.size(), out, count, sizeof(float));
}
size_t ba_oracle_trace(ba_oracle* o, int dev, int ch, uint8_t* out, size_t count) {
    Channel& c = o->dev[dev].ch[ch];
    return copy_out(c.rec_trace.data(), c.rec_trace.size(), out, count, 1);
}
size_t ba_oracle_status(ba_oracle* o, int dev, int ch, ba_channel_status* out, size_t count) {
    Channel& c = o->dev[dev].ch[ch];
    return copy_out(c.rec_status.data(), c.rec_status.size(), out, count, sizeof(ba_channel_status));
}
int ba_oracle_channel_info(ba_oracle* o, int dev, int ch, ba_channel_info* out) {
    *out = o->dev[dev].ch[ch].info;
    return BA_OK;
}
/* scan mode: from batch `from_batch` of the device on, the channel runs with freqlist[freq_idx] — what controller_thread's
 * `dev->channels[0].freq_idx = i` (boondock_airband.cpp:115-118) amounts to when demodulate() next reads it (:522) */
int ba_oracle_set_freq_idx(ba_oracle* o, int dev, int ch, uint64_t from_batch, int freq_idx) {
    if (!o || dev < 0 || dev >= (int)o->dev.size() || ch < 0 || ch >= (int)o->dev[dev].ch.size())
        return BA_ERR_BAD_ARG;
    Channel& c = o->dev[dev].ch[ch];
    if (freq_idx < 0 || freq_idx >= (int)c.freqs.size())
        return BA_ERR_BAD_ARG;
    c.freq_plan.emplace_back(from_batch, freq_idx);
    return BA_OK;
}
int ba_oracle_window(ba_oracle* o, float* out, size_t count) {
    if (count != o->window.size())
        return BA_ERR_BAD_ARG;
    memcpy(out, o->window.data(), count * sizeof(float));
    return BA_OK;
}

/* convert+window and FFT of n_frames frames spaced one hop apart; either output may be NULL */
int ba_oracle_debug_frames(ba_oracle* o, int dev, const void* iq, size_t bytes, int n_frames, float* fftin, float* fftout) {
    Device& d = o->dev[dev];
    const size_t N = o->fft_size, fb = N * d.cfg.bytes_per_sample * 2;
    Device scratch;
    scratch.fftin.assign(2 * N, 0.0f);
    scratch.fftout.assign(2 * N, 0.0f);
    scratch.fft.plan((int)N);
    for (int f = 0; f < n_frames; f++) {
        if ((size_t)f * d.bps + fb > bytes)
            return BA_ERR_BAD_ARG;
        /* conversion */
        const unsigned char* buf = (const unsigned char*)iq + (size_t)f * d.bps;
        float* fi = scratch.fftin.data();
        const float* w = o->window.data();
        if (d.cfg.sample_format == BA_SFMT_S16) {
            const short* b2 = (const short*)buf;
            for (size_t i = 0; i < N; i++, b2 += 2) {
                fi[2 * i] = d.scale * (float)b2[0] * w[i];
                fi[2 * i + 1] = d.scale * (float)b2[1] * w[i];
            }
        } else if (d.cfg.sample_format == BA_SFMT_F32) {
            const float* b2 = (const float*)buf;
            for (size_t i = 0; i < N; i++, b2 += 2) {
                fi[2 * i] = d.scale * b2[0] * w[i];
                fi[2 * i + 1] = d.scale * b2[1] * w[i];
            }
        } else {
            const float* lv = d.cfg.sample_format == BA_SFMT_U8 ? o->levels_u8 : o->levels_s8;
            for (size_t i = 0; i < N; i++, buf += 2) {
                fi[2 * i] = lv[buf[0]] * w[i];
                fi[2 * i + 1] = lv[buf[1]] * w[i];
            }
        }
        if (fftin)
            memcpy(fftin + (size_t)f * 2 * N, fi, 2 * N * sizeof(float));
        if (fftout) {
            scratch.fft.run(fi, scratch.fftout.data());
            memcpy(fftout + (size_t)f * 2 * N, scratch.fftout.data(), 2 * N * sizeof(float));
        }
    }
    return BA_OK;
}

/* Timing leg: every device consumes its whole buffer, one thread per device as the reference does with
 * multiple_demod_threads (boondock_airband.cpp:1088-1122), at most `threads` at a time.  Returns wall seconds. */
/* which transform this build runs: 0 scalar radix-4 Stockham (parity builds), 8 the AVX2 eight-frames-per-vector one (timing builds) */
int ba_oracle_fft_lanes(void) {
#ifdef BA_ORACLE_LANES
    return BA_ORACLE_LANES;
#else
    return 0;
#endif
}

double ba_oracle_run_threads(ba_oracle* o, const void* const* iq, const size_t* bytes, int threads) {
    const int nd = (int)o->dev.size();
    if (threads < 1)
        threads = 1;
    if (threads > nd)
        threads = nd;
    std::atomic<int> next(0);
    auto t0 = std::chrono::steady_clock::now();
    auto work = [&]() {
        for (;;) {
            int i = next.fetch_add(1);
            if (i >= nd)
                return;
            Device& d = o->dev[i];
            consume(o, d, (const unsigned char*)iq[i], bytes[i]);
        }
    };
    if (threads == 1) {
        work();
    } else {
        std::vector<std::thread> pool;
        for (int t = 0; t < threads; t++)
            pool.emplace_back(work);
        for (auto& t : pool)
            t.join();
    }
    auto t1 = std::chrono::steady_clock::now();
    return std::chrono::duration<double>(t1 - t0).count();
}

/* ---- bare DSP objects for the behavioural pins of test_squelch.cpp / test_ctcss.cpp / test_filters.cpp ---- */
void* ba_oracle_sq_new(void) {
    return new Squelch();
}
void ba_oracle_sq_free(void* s) {
    delete (Squelch*)s;
}
void ba_oracle_sq_set_ctcss(void* s, float hz, float rate) {
    ((Squelch*)s)->set_ctcss_freq(hz, rate);
}
void ba_oracle_sq_set_level(void* s, float level) {
    ((Squelch*)s)->set_squelch_level_threshold(level);
}
void ba_oracle_sq_set_snr(void* s, float db) {
    ((Squelch*)s)->set_squelch_snr_threshold(db);
}
void ba_oracle_sq_raw(void* s, float v) {
    ((Squelch*)s)->process_raw_sample(v);
}
void ba_oracle_sq_filtered(void* s, float v) {
    ((Squelch*)s)->process_filtered_sample(v);
}
void ba_oracle_sq_audio(void* s, float v) {
    ((Squelch*)s)->process_audio_sample(v);
}
/* out[0..9]: is_open, should_process_audio, should_filter_sample, first_open, last_open, state, open_count, flappy_count, ctcss_count, no_ctcss_count */
void ba_oracle_sq_query(void* sp, int32_t* out, float* levels) {
    Squelch* s = (Squelch*)sp;
    out[0] = s->is_open();
    out[1] = s->should_process_audio();
    out[2] = s->should_filter_sample();
    out[3] = s->first_open_sample();
    out[4] = s->last_open_sample();
    out[5] = SQ_CURRENT(*s);
    out[6] = (int32_t)s->open_count();
    out[7] = (int32_t)s->flappy_count();
    out[8] = (int32_t)s->ctcss_count();
    out[9] = (int32_t)s->no_ctcss_count();
    levels[0] = s->noise_level();
    levels[1] = s->signal_level();
    levels[2] = s->squelch_level();
}
/* run `n` steps: raw[i] always; filtered[i] if not NaN; audio[i] if not NaN; records state byte + is_open per step */
void ba_oracle_sq_run(void* sp, const float* raw, const float* filtered, const float* audio, size_t n, uint8_t* states, float* levels3) {
    Squelch* s = (Squelch*)sp;
    for (size_t i = 0; i < n; i++) {
        s->process_raw_sample(raw[i]);
        if (filtered && !isnan(filtered[i]))
            s->process_filtered_sample(filtered[i]);
        if (audio && !isnan(audio[i]) && s->should_process_audio())
            s->process_audio_sample(audio[i]);
        if (states)
            states[i] = (uint8_t)(SQ_CURRENT(*s) | (s->is_open() ? 8 : 0) | (s->should_process_audio() ? 16 : 0) | (s->should_filter_sample() ? 32 : 0));
        if (levels3) {
            levels3[3 * i] = s->noise_level();
            levels3[3 * i + 1] = s->signal_level();
            levels3[3 * i + 2] = s->squelch_level();
        }
    }
}
/* notch / low-pass responses on a sample stream */
void ba_oracle_notch_run(float hz, float rate, float q, float* x, size_t n, float* coeffs3) {
    NotchFilter f(hz, rate, q);
    for (size_t i = 0; i < n; i++)
        f.apply(x[i]);
#ifndef BA_ORACLE_REF
    if (coeffs3 && f.enabled())
        memcpy(coeffs3, f.coeffs(), 3 * sizeof(float));
#else
    (void)coeffs3;
#endif
}
void ba_oracle_lowpass_run(float hz, float rate, float* re, float* im, size_t n) {
    LowpassFilter f(hz, rate);
    for (size_t i = 0; i < n; i++)
        f.apply(re[i], im[i]);
}
/* a single CTCSS bank as test_ctcss.cpp drives it: returns has_tone after `n` samples; enough[0] = enough_samples */
int ba_oracle_ctcss_run(float hz, float rate, int window, const float* x, size_t n, int32_t* enough);
}

#ifdef BA_ORACLE_REF
extern "C" int ba_oracle_ctcss_run(float hz, float rate, int window, const float* x, size_t n, int32_t* enough) {
    CTCSS c(hz, rate, window);
    for (size_t i = 0; i < n; i++)
        c.process_audio_sample(x[i]);
    if (enough)
        *enough = c.enough_samples();
    return c.has_tone();
}
#else
extern "C" int ba_oracle_ctcss_run(float hz, float rate, int window, const float* x, size_t n, int32_t* enough) {
    ora::ToneBank c;
    c.configure(hz, rate, window);
    for (size_t i = 0; i < n; i++)
        c.feed(x[i]);
    if (enough)
        *enough = c.full;
    return c.has_tone();
}
#endif

#if defined(BA_ORACLE_REF) && defined(BA_ORACLE_UNITY)
/* timing build of the reference-objects oracle as ONE translation unit: the reference's own sources, compiled where they lie */
#include "squelch.cpp"
#include "ctcss.cpp"
#include "filters.cpp"
#include "logging.cpp"
#endif
