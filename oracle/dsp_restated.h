/*
 * dsp_restated.h — CPU restatement of the per-sample DSP blocks on Boondock-Airband's demodulate() path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the shipped engine: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may build, load or run it.
 *
 * What is restated here (reference file:line, all under /root/reference/src):
 *   sq_*      Squelch           squelch.cpp:36-70 (defaults), :195-246 (raw sample), :248-276 (filtered sample),
 *                               :278-295 (audio sample), :297-361 (transition coercion), :363-460 (FSM step),
 *                               :462-514 (signal tests, noise floor, capped moving averages), :118-177 (queries)
 *   ToneBank  CTCSS             ctcss.cpp:31-59 (Goertzel), :61-99 (set, de-duplication), :101-172 (decision)
 *   Notch     NotchFilter       filters.cpp:30-64
 *   Bessel2   LowpassFilter     filters.cpp:70-99 (design), :146-163 (apply)
 *
 * Pinning: tests/test_oracle_pins.py replays the scenarios of test_squelch.cpp:56-281 and
 * test_ctcss.cpp:122-155 against this file, and (when oracle/_ref exists) compares it bit for bit with the
 * reference's own squelch.cpp / ctcss.cpp / filters.cpp compiled unmodified from /root/reference.
 *
 * Arithmetic is IEEE float with no contraction: build with -ffp-contract=off and without -ffast-math.
 */
#ifndef BA_ORACLE_DSP_RESTATED_H
#define BA_ORACLE_DSP_RESTATED_H

#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <complex>
#include <vector>

namespace ora {

/* ---------------------------------------------------------------- CTCSS tone bank (ctcss.cpp) */

static const float k_standard_tones[51] = {67.0f,  69.3f,  71.9f,  74.4f,  77.0f,  79.7f,  82.5f,  85.4f,  88.5f,  91.5f,  94.8f,  97.4f,  100.0f,
                                           103.5f, 107.2f, 110.9f, 114.8f, 118.8f, 123.0f, 127.3f, 131.8f, 136.5f, 141.3f, 146.2f, 150.0f, 151.4f,
                                           156.7f, 159.8f, 162.2f, 165.5f, 167.9f, 171.3f, 173.8f, 177.3f, 179.9f, 183.5f, 186.2f, 189.9f, 192.8f,
                                           196.6f, 199.5f, 203.5f, 206.5f, 210.7f, 218.1f, 225.7f, 229.1f, 233.6f, 241.8f, 250.3f, 254.1f};

/* Goertzel coefficient exactly as ToneDetector's constructor forms it (ctcss.cpp:31-43):
 * k from a float product plus a double 0.5, omega rounded to float, cosine taken in float, doubled in double. */
static inline float goertzel_coeff(float tone_hz, float rate, int window) {
    int k = (int)(0.5 + window * tone_hz / rate);
    float omega = (float)((2.0 * M_PI * k) / window);
    return (float)(2.0 * cosf(omega));
}

struct Goertzel {
    float hz, coeff, s1, s2, power;
    int n;
};

struct ToneBank {
    bool on = false;
    float target = 0.0f;
    int window = 0;
    size_t hits = 0, misses = 0;
    bool full = false; /* enough_samples_ */
    int fed = 0;
    bool tone = false; /* has_tone_ */
    std::vector<Goertzel> det;

    void try_add(float hz, float rate) {
        float c = goertzel_coeff(hz, rate, window);
        for (size_t i = 0; i < det.size(); i++)
            if (det[i].coeff == c)
                return; /* same Goertzel bin as an earlier detector: dropped (ctcss.cpp:61-72) */
        Goertzel g = {hz, c, 0.0f, 0.0f, 0.0f, 0};
        det.push_back(g);
    }
    void configure(float target_hz, float rate, int window_len) {
        on = true;
        target = target_hz;
        window = window_len;
        hits = misses = 0;
        det.clear();
        try_add(target_hz, rate);
        for (int i = 0; i < 51; i++) {
            if (fabsf(target_hz - k_standard_tones[i]) < 5)
                continue;
            try_add(k_standard_tones[i], rate);
        }
        clear();
    }
    void clear_detectors() {
        for (size_t i = 0; i < det.size(); i++) {
            det[i].n = 0;
            det[i].s1 = det[i].s2 = 0.0f;
        }
    }
    void clear() { /* CTCSS::reset */
        if (!on)
            return;
        clear_detectors();
        full = false;
        fed = 0;
        tone = false;
    }
    bool has_tone() const { return !on || tone; }
    void feed(float s) {
        if (!on)
            return;
        for (size_t i = 0; i < det.size(); i++) {
            Goertzel& g = det[i];
            float s0 = g.coeff * g.s1 - g.s2 + s;
            g.s2 = g.s1;
            g.s1 = s0;
            if (++g.n == window) {
                g.power = g.s1 * g.s1 + g.s2 * g.s2 - g.s1 * g.s2 * g.coeff;
                g.n = 0;
            }
        }
        if (++fed < window)
            return;
        full = true;
        /* window complete: target must hold the largest power and exceed the mean (ctcss.cpp:139-159);
         * the mean is accumulated in detector order, in float */
        float total = 0.0f, best = det[0].power, mine = 0.0f;
        bool have_mine = false;
        for (size_t i = 0; i < det.size(); i++) {
            total += det[i].power;
            if (det[i].power > best)
                best = det[i].power;
        }
        /* the reference sorts by power (descending) and takes the first entry whose freq equals the target;
         * only one detector carries the target frequency, so that is its power */
        for (size_t i = 0; i < det.size() && !have_mine; i++)
            if (det[i].hz == target) {
                mine = det[i].power;
                have_mine = true;
            }
        float mean = total / det.size();
        if (mine == best && mine > mean) {
            tone = true;
            hits++;
        } else {
            tone = false;
            misses++;
        }
        clear_detectors();
        fed = 0;
    }
};

/* ---------------------------------------------------------------- Squelch (squelch.cpp) */

enum { ST_CLOSED = 0, ST_OPENING = 1, ST_CLOSING = 2, ST_LOW_SIGNAL_ABORT = 3, ST_OPEN = 4 };

struct Ema {
    float full, capped;
};

class Squelch {
   public:
    Squelch() {
        noise_ = 5.0f;
        set_squelch_snr_threshold(9.54f);
        manual_level_ = -1.0f;
        pre_ = {0.001f, 0.001f};
        post_ = {0.001f, 0.001f};
        level_cache_ = 0.0f;
        post_active_ = false;
        pre_post_factor_ = 0.9f;
        open_delay_ = 197;
        close_delay_ = 197;
        abort_after_ = 88;
        next_ = cur_ = ST_CLOSED;
        delay_ = 0;
        opens_ = 0;
        samples_ = (size_t)-1;
        flappy_ = 0;
        low_run_ = 0;
        recent_span_ = 1000;
        flap_opens_ = 3;
        recent_opens_ = 0;
        closed_run_ = 0;
        for (int i = 0; i < kRing; i++)
            ring_[i] = 0.0f;
        head_ = 0;
        tail_ = 1;
    }

    void set_squelch_level_threshold(const float& level) {
        if (level > 0) {
            manual_ = true;
            manual_level_ = level;
        } else {
            manual_ = false;
        }
        refresh_cap();
    }
    void set_squelch_snr_threshold(const float& db) {
        manual_ = false;
        ratio_ = (float)pow(10.0, db / 20.0);
        flappy_ratio_ = ratio_ * 0.9f;
        refresh_cap();
    }
    void set_ctcss_freq(const float& hz, const float& rate) {
        fast_.configure(hz, rate, (int)(rate * 0.05));
        slow_.configure(hz, rate, (int)(rate * 0.4));
    }

    bool is_open() const {
        if (cur_ == ST_OPEN || cur_ == ST_CLOSING) {
            if (slow_.on)
                return slow_.full ? slow_.has_tone() : fast_.has_tone();
            return true;
        }
        return false;
    }
    bool should_filter_sample() { return (pre_has_signal() || cur_ != ST_CLOSED) && cur_ != ST_LOW_SIGNAL_ABORT; }
    bool should_process_audio() { return cur_ == ST_OPEN || cur_ == ST_CLOSING; }
    bool first_open_sample() const { return cur_ != ST_OPEN && next_ == ST_OPEN; }
    bool last_open_sample() const { return (cur_ == ST_CLOSING && next_ == ST_CLOSED) || (cur_ != ST_LOW_SIGNAL_ABORT && next_ == ST_LOW_SIGNAL_ABORT); }
    bool signal_outside_filter() { return post_active_ && pre_has_signal() && !post_has_signal(); }

    const float& noise_level() const { return noise_; }
    const float& signal_level() const { return pre_.full; }
    const float& squelch_level() {
        if (manual_)
            return manual_level_;
        if (level_cache_ == 0.0f) {
            if (flapping() && flappy_ratio_ < ratio_)
                level_cache_ = flappy_ratio_ * noise_;
            else
                level_cache_ = ratio_ * noise_;
        }
        return level_cache_;
    }
    const size_t& open_count() const { return opens_; }
    const size_t& flappy_count() const { return flappy_; }
    const size_t& ctcss_count() const { return slow_.hits; }
    const size_t& no_ctcss_count() const { return slow_.misses; }

    void process_raw_sample(const float& s) {
        step_fsm();
        samples_++;
        if (samples_ % 16 == 0)
            track_noise();
        ema_update(pre_, s);
        ring_[head_] = pre_.capped * pre_post_factor_;

        if (cur_ == ST_OPEN && !has_signal())
            request(ST_CLOSING);
        if (cur_ == ST_CLOSED && has_signal())
            request(ST_OPENING);

        if (cur_ != ST_CLOSED && cur_ != ST_LOW_SIGNAL_ABORT) {
            if (s >= squelch_level()) {
                low_run_ = 0;
            } else {
                low_run_++;
                if (low_run_ >= abort_after_)
                    request(ST_LOW_SIGNAL_ABORT);
            }
        }
    }
    void process_filtered_sample(const float& s) {
        if (!should_filter_sample())
            return;
        if (cur_ == ST_OPENING) {
            if (delay_ < kRing)
                return;
            if (delay_ == kRing)
                post_ = {ring_[tail_], ring_[tail_]};
        }
        post_active_ = true;
        ema_update(post_, s);
        if (post_.capped < ring_[tail_])
            request(ST_CLOSED);
    }
    void process_audio_sample(const float& s) {
        if (!slow_.on)
            return;
        if (cur_ != ST_CLOSED) {
            slow_.feed(s);
            if (!slow_.full)
                fast_.feed(s);
        }
    }

    /* test-only view of the FSM (the reference exposes it only under -DDEBUG_SQUELCH, squelch.cpp:520-633) */
    int current_state() const { return cur_; }
    int next_state() const { return next_; }
    float ratio() const { return ratio_; }
    float manual() const { return manual_ ? manual_level_ : 0.0f; }
    int fast_tones() const { return (int)fast_.det.size(); }
    int slow_tones() const { return (int)slow_.det.size(); }
    int fast_window() const { return fast_.window; }
    int slow_window() const { return slow_.window; }

   private:
    enum { kRing = 102 };
    float noise_;
    bool manual_;
    float manual_level_, ratio_, flappy_ratio_;
    float cap_;
    Ema pre_, post_;
    float level_cache_;
    bool post_active_;
    float pre_post_factor_;
    int open_delay_, close_delay_, abort_after_;
    int next_, cur_;
    int delay_;
    size_t opens_, samples_, flappy_;
    int low_run_;
    size_t recent_span_, flap_opens_, recent_opens_, closed_run_;
    float ring_[kRing];
    int head_, tail_;
    ToneBank fast_, slow_;

    bool flapping() const { return recent_opens_ >= flap_opens_; }
    bool pre_has_signal() { return pre_.capped >= squelch_level(); }
    bool post_has_signal() { return post_active_ && post_.capped >= ring_[tail_]; }
    bool has_signal() {
        if (post_active_)
            return pre_has_signal() && post_has_signal();
        return pre_has_signal();
    }
    void refresh_cap() { cap_ = manual_ ? 1.5f * manual_level_ : 1.5f * ratio_ * noise_; }
    void track_noise() {
        static const float keep = 0.97f;
        static const float take = 1.0 - keep; /* double subtraction, rounded to float (squelch.cpp:478-479) */
        noise_ = noise_ * keep + (pre_.capped < noise_ ? pre_.capped : noise_) * take + 1e-6f;
        refresh_cap();
        level_cache_ = 0.0f;
    }
    void ema_update(Ema& a, const float& s) {
        static const float keep = 0.99f;
        static const float take = 1.0 - keep;
        a.full = a.full * keep + s * take;
        if (a.capped >= cap_ && s >= cap_) {
            a.capped = cap_;
        } else {
            float v = a.capped * keep + s * take;
            a.capped = cap_ < v ? cap_ : v;
        }
    }
    /* squelch.cpp:297-361: illegal requests are redirected */
    void request(int want) {
        if (cur_ == ST_CLOSED && (want == ST_CLOSING || want == ST_LOW_SIGNAL_ABORT))
            want = ST_CLOSED;
        else if (cur_ == ST_CLOSED && want == ST_OPEN)
            want = ST_OPENING;
        else if (cur_ == ST_OPENING && want == ST_LOW_SIGNAL_ABORT)
            want = ST_CLOSED;
        else if (cur_ == ST_LOW_SIGNAL_ABORT && want != ST_LOW_SIGNAL_ABORT && want != ST_CLOSED)
            want = ST_CLOSED;
        else if (cur_ == ST_OPEN && want == ST_CLOSED)
            want = ST_CLOSING;
        else if (cur_ == ST_OPEN && want == ST_OPENING)
            want = ST_OPEN;
        next_ = want;
    }
    /* squelch.cpp:363-460 */
    void step_fsm() {
        switch (next_) {
            case ST_OPENING:
                if (cur_ != ST_OPENING) {
                    delay_ = 0;
                    low_run_ = 0;
                    post_active_ = false;
                    cur_ = ST_OPENING;
                } else if (++delay_ >= open_delay_) {
                    if (closed_run_ < recent_span_) {
                        recent_opens_++;
                        if (flapping())
                            flappy_++;
                        level_cache_ = 0.0f;
                    }
                    next_ = has_signal() ? ST_OPEN : ST_CLOSED;
                }
                break;
            case ST_CLOSING:
                if (cur_ != ST_CLOSING) {
                    delay_ = 0;
                    cur_ = ST_CLOSING;
                } else if (++delay_ >= close_delay_) {
                    if (!has_signal()) {
                        next_ = ST_CLOSED;
                    } else {
                        cur_ = ST_OPEN;
                        next_ = ST_OPEN;
                    }
                }
                break;
            case ST_LOW_SIGNAL_ABORT:
                if (cur_ != ST_LOW_SIGNAL_ABORT) {
                    if (cur_ != ST_CLOSING)
                        delay_ = 0;
                    cur_ = ST_LOW_SIGNAL_ABORT;
                } else if (++delay_ >= close_delay_) {
                    next_ = ST_CLOSED;
                }
                break;
            case ST_OPEN:
                if (cur_ != ST_OPEN) {
                    opens_++;
                    cur_ = ST_OPEN;
                }
                break;
            default: /* ST_CLOSED */
                if (cur_ != ST_CLOSED) {
                    post_active_ = false;
                    closed_run_ = 0;
                    cur_ = ST_CLOSED;
                    fast_.clear();
                    slow_.clear();
                } else if (closed_run_ < recent_span_) {
                    closed_run_++;
                } else if (closed_run_ == recent_span_) {
                    recent_opens_ = 0;
                    level_cache_ = 0.0f;
                }
                break;
        }
        tail_ = (tail_ + 1) % kRing;
        head_ = (head_ + 1) % kRing;
    }
};

/* ---------------------------------------------------------------- filters (filters.cpp) */

class NotchFilter {
   public:
    NotchFilter() : on_(false) {}
    NotchFilter(float hz, float rate, float q) : on_(true) {
        for (int i = 0; i < 3; i++)
            in_[i] = out_[i] = 0.0f;
        if (hz <= 0.0) {
            on_ = false;
            return;
        }
        float w0 = (float)(2 * M_PI * (hz / rate));
        float e = 1 / (1 + tanf(w0 / (q * 2)));
        float p = cosf(w0);
        k_[0] = e;
        k_[1] = 2 * e * p;
        k_[2] = (2 * e - 1);
    }
    bool enabled() const { return on_; }
    void apply(float& v) {
        if (!on_)
            return;
        in_[0] = in_[1];
        in_[1] = in_[2];
        in_[2] = v;
        out_[0] = out_[1];
        out_[1] = out_[2];
        out_[2] = k_[0] * in_[2] - k_[1] * in_[1] + k_[0] * in_[0] + k_[1] * out_[1] - k_[2] * out_[0];
        v = out_[2];
    }
    const float* coeffs() const { return k_; }

   private:
    bool on_;
    float k_[3], in_[3], out_[3];
};

class LowpassFilter {
   public:
    LowpassFilter() : on_(false) {}
    LowpassFilter(float hz, float rate) : on_(true) {
        typedef std::complex<double> cd;
        if (hz <= 0.0) {
            on_ = false;
            return;
        }
        for (int i = 0; i < 3; i++)
            xr_[i] = xi_[i] = yr_[i] = yi_[i] = 0.0f;
        double raw = (double)hz / rate;
        double warped = tan(M_PI * raw) / M_PI;
        const cd bessel(-1.10160133059e+00, 6.36009824757e-01);
        cd pole[2] = {bilinear(M_PI * 2 * warped * bessel), bilinear(M_PI * 2 * warped * std::conj(bessel))};
        cd zero[2] = {-1.0, -1.0};
        cd top[3], bot[3];
        poly_from_roots(zero, top);
        poly_from_roots(pole, bot);
        /* response at z = 1 (DC) */
        cd num = (top[2] * 1.0 + top[1]) * 1.0 + top[0];
        cd den = (bot[2] * 1.0 + bot[1]) * 1.0 + bot[0];
        /* the reference accumulates from sum = 0: ((0*z + c2)*z + c1)*z + c0 — identical in value */
        cd g = num / den;
        gain_ = (float)hypot(g.imag(), g.real());
        for (int i = 0; i <= 2; i++)
            yc_[i] = (float)(-(bot[i].real() / bot[2].real()));
    }
    bool enabled() const { return on_; }
    void apply(float& r, float& j) {
        if (!on_)
            return;
        xr_[0] = xr_[1], xi_[0] = xi_[1];
        xr_[1] = xr_[2], xi_[1] = xi_[2];
        xr_[2] = r / gain_, xi_[2] = j / gain_;
        yr_[0] = yr_[1], yi_[0] = yi_[1];
        yr_[1] = yr_[2], yi_[1] = yi_[2];
        yr_[2] = (xr_[0] + xr_[2]) + (2.0f * xr_[1]) + (yc_[0] * yr_[0]) + (yc_[1] * yr_[1]);
        yi_[2] = (xi_[0] + xi_[2]) + (2.0f * xi_[1]) + (yc_[0] * yi_[0]) + (yc_[1] * yi_[1]);
        r = yr_[2];
        j = yi_[2];
    }
    float gain() const { return gain_; }
    const float* ycoeffs() const { return yc_; }

   private:
    typedef std::complex<double> cdbl;
    static cdbl bilinear(cdbl s) { return (2.0 + s) / (2.0 - s); }
    /* (z - r0)(z - r1) as c[0] + c[1] z + c[2] z^2, built factor by factor like filters.cpp:121-144 */
    static void poly_from_roots(const cdbl r[2], cdbl c[3]) {
        c[0] = 1.0;
        c[1] = c[2] = 0.0;
        for (int k = 0; k < 2; k++) {
            cdbl m = -r[k];
            for (int i = 2; i >= 1; i--)
                c[i] = (m * c[i]) + c[i - 1];
            c[0] = m * c[0];
        }
    }
    bool on_;
    float yc_[3], gain_;
    float xr_[3], xi_[3], yr_[3], yi_[3];
};

}  // namespace ora
#endif
