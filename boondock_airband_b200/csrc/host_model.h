/*
 * host_model.h — configuration-time constants of the channel model, computed on the host exactly as the
 * reference's config parser and constructors compute them (double where it uses double, float where it uses float).
 * These feed the kernels; nothing here runs per sample.
 */
#ifndef BA_HOST_MODEL_H
#define BA_HOST_MODEL_H

#include <math.h>
#include <stddef.h>
#include <stdint.h>

#include <complex>
#include <vector>

#include "ba_kernels.h"

namespace ba {
namespace model {

/* the analysis window, boondock_airband.cpp:357-373: seven cosine terms whose coefficients are float literals
 * widened to double; evaluated in double, stored as float */
inline void window7(int n, float* out) {
    const double a0 = 0.27105140069342f, a1 = 0.43329793923448f, a2 = 0.21812299954311f, a3 = 0.06592544638803f;
    const double a4 = 0.01081174209837f, a5 = 0.00077658482522f, a6 = 0.00001388721735f;
    for (size_t i = 0; i < (size_t)n; i++) {
        const double d = (double)(n - 1);
        double x = a0 - (a1 * cos((2.0 * M_PI * i) / d)) + (a2 * cos((4.0 * M_PI * i) / d)) - (a3 * cos((6.0 * M_PI * i) / d)) + (a4 * cos((8.0 * M_PI * i) / d)) -
                   (a5 * cos((10.0 * M_PI * i) / d)) + (a6 * cos((12.0 * M_PI * i) / d));
        out[i] = (float)x;
    }
}

/* FFT twiddles exp(-2 pi i k / n), rounded once from double */
inline void twiddles(int n, float2* out) {
    for (int k = 0; k < n; k++) {
        const double a = -2.0 * M_PI * (double)k / (double)n;
        out[k] = make_float2((float)cos(a), (float)sin(a));
    }
}

/* sincosf_lut_init, util.cpp:105-110: out[0..256] sine, out[257..513] cosine */
inline void sincos_table(float* out) {
    for (uint32_t i = 0; i < 256; i++)
        sincosf((float)(2.0F * M_PI * (float)i / 256.0f), &out[i], &out[257 + i]);
    out[256] = out[0];
    out[257 + 256] = out[257];
}

/* dev->bins[], config.cpp:669-670 — sample_rate / fft_size is an integer division there */
inline uint32_t bin_index(int freq, int sample_rate, int centerfreq, int fft_size) {
    const size_t hz_per_bin = (size_t)sample_rate / (size_t)fft_size;
    return (uint32_t)((size_t)ceil((freq + sample_rate - centerfreq) / (double)hz_per_bin - 1.0) % (size_t)fft_size);
}

/* channel_t.dm_dphi, config.cpp:682-715 */
inline uint32_t derotation_step(int freq, int centerfreq, int sample_rate, int wave_rate) {
    double dm = (double)(freq - centerfreq);
    const double decim = (double)sample_rate / (double)wave_rate;
    double corr = (double)wave_rate / 2.0;
    corr *= (decim - round(decim));
    corr *= (double)(freq - centerfreq) / ((double)sample_rate / 2.0);
    dm -= corr;
    dm /= (double)wave_rate;
    dm -= trunc(dm);
    dm *= 256.0 * 65536.0;
    return (uint32_t)((int)dm);
}

/* de-emphasis constant alpha = exp(-1 / (WAVE_RATE * tau)); boondock_airband.cpp:87, config.cpp:650-652, 777-781 */
inline float alpha_default(int wave_rate) {
    return (float)exp(-1.0f / (wave_rate * 2e-4));
}
inline float alpha_from_tau_us(int wave_rate, int tau_us) {
    return tau_us == 0 ? 0.0f : (float)exp(-1.0f / (wave_rate * 1e-6 * tau_us));
}

/* dBFS_to_level, util.cpp:169-176 */
inline float dbfs_to_level(float dbfs, int fft_size) {
    const float offset = 7.54f + 10.0f * log10f((float)(size_t)(fft_size / 2)) - 2.38f;
    return (float)(pow(10.0, (dbfs - offset) / 20.0f) * (size_t)fft_size);
}

/* Squelch::set_squelch_snr_threshold, squelch.cpp:93-104 */
inline float snr_ratio(float db) {
    return (float)pow(10.0, db / 20.0);
}

/* NotchFilter constructor, filters.cpp:30-50 */
inline bool notch_design(float hz, float rate, float q, float d[3]) {
    if (hz <= 0.0)
        return false;
    const float w0 = (float)(2 * M_PI * (hz / rate));
    const float e = 1 / (1 + tanf(w0 / (q * 2)));
    const float p = cosf(w0);
    d[0] = e;
    d[1] = 2 * e * p;
    d[2] = (2 * e - 1);
    return true;
}

/* LowpassFilter constructor, filters.cpp:70-144: 2nd-order Bessel through the bilinear transform, in double */
inline bool lowpass_design(float hz, float rate, float ycoeffs[2], float* gain) {
    typedef std::complex<double> cd;
    if (hz <= 0.0)
        return false;
    const double raw = (double)hz / rate;
    const double warped = tan(M_PI * raw) / M_PI;
    const cd s0(-1.10160133059e+00, 6.36009824757e-01);
    auto blt = [](cd s) { return (2.0 + s) / (2.0 - s); };
    const cd poles[2] = {blt(M_PI * 2 * warped * s0), blt(M_PI * 2 * warped * std::conj(s0))};
    const cd zeros[2] = {-1.0, -1.0};
    auto expand = [](const cd r[2], cd c[3]) {
        c[0] = 1.0;
        c[1] = c[2] = 0.0;
        for (int k = 0; k < 2; k++) {
            const cd m = -r[k];
            for (int i = 2; i >= 1; i--)
                c[i] = (m * c[i]) + c[i - 1];
            c[0] = m * c[0];
        }
    };
    cd top[3], bot[3];
    expand(zeros, top);
    expand(poles, bot);
    auto eval = [](const cd c[3], cd z) {
        cd sum = 0.0;
        for (int i = 2; i >= 0; i--)
            sum = (sum * z) + c[i];
        return sum;
    };
    const cd dc = eval(top, cd(1.0, 0.0)) / eval(bot, cd(1.0, 0.0));
    *gain = (float)hypot(dc.imag(), dc.real());
    ycoeffs[0] = (float)(-(bot[0].real() / bot[2].real()));
    ycoeffs[1] = (float)(-(bot[1].real() / bot[2].real()));
    return true;
}

/* ToneDetector constructor, ctcss.cpp:31-43 */
inline float goertzel_coeff(float tone_hz, float rate, int window) {
    const int k = (int)(0.5 + window * tone_hz / rate);
    const float omega = (float)((2.0 * M_PI * k) / window);
    return (float)(2.0 * cosf(omega));
}

/* CTCSS constructor + ToneDetectorSet::add, ctcss.cpp:61-72, 105-122: target first, then the standard tones that
 * are at least 5 Hz away; a tone whose coefficient equals an earlier one is dropped */
inline int tone_bank(float target_hz, float rate, int window, float* coeff) {
    static const float standard[51] = {67.0f,  69.3f,  71.9f,  74.4f,  77.0f,  79.7f,  82.5f,  85.4f,  88.5f,  91.5f,  94.8f,  97.4f,  100.0f,
                                       103.5f, 107.2f, 110.9f, 114.8f, 118.8f, 123.0f, 127.3f, 131.8f, 136.5f, 141.3f, 146.2f, 150.0f, 151.4f,
                                       156.7f, 159.8f, 162.2f, 165.5f, 167.9f, 171.3f, 173.8f, 177.3f, 179.9f, 183.5f, 186.2f, 189.9f, 192.8f,
                                       196.6f, 199.5f, 203.5f, 206.5f, 210.7f, 218.1f, 225.7f, 229.1f, 233.6f, 241.8f, 250.3f, 254.1f};
    int n = 0;
    auto add = [&](float hz) {
        const float c = goertzel_coeff(hz, rate, window);
        for (int i = 0; i < n; i++)
            if (coeff[i] == c)
                return;
        coeff[n++] = c;
    };
    add(target_hz);
    for (int i = 0; i < 51; i++) {
        if (fabsf(target_hz - standard[i]) < 5)
            continue;
        add(standard[i]);
    }
    return n;
}

/* input ring length, config.cpp:796-805 with FFT_BATCH = 1 */
inline size_t ring_bytes(size_t min_bytes, int bytes_per_sample, int sample_rate, int wave_rate) {
    const size_t block = 2 * (size_t)bytes_per_sample * (size_t)ceil((double)sample_rate / (double)wave_rate);
    size_t n = min_bytes;
    if (n % block != 0)
        n += block - n % block;
    return n;
}

}  // namespace model
}  // namespace ba
#endif
