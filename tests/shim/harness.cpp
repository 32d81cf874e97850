/*
 * harness.cpp — TEST INFRASTRUCTURE: the pieces of Boondock-Airband that surround demodulate(), played by the test so that
 * boondock_airband_b200/csrc/demodulate_cuda.cpp can run exactly as it would inside the reference:
 *   - the globals it reads (boondock_airband.cpp:71-90);
 *   - one rx thread per input appending with circbuffer_append()'s arithmetic under buffer_lock (input-helpers.cpp:37-63;
 *     restated here: the reference's function lives in a file that needs libconfig++);
 *   - the output thread's share of the hand-off (output.cpp:931-951): take waveout[0..WAVE_BATCH) of every channel while
 *     waveavail is set, move the tail down, clear waveavail;
 *   - main()'s pthread_create of the demodulator (boondock_airband.cpp:1146-1148).
 * Built twice by tests/test_shim.py, without and with -DNFM (WAVE_RATE 8000 / 16000), together with demodulate_cuda.cpp and
 * include/ba_ref_layout.h, and linked against libba_cuda.so.  Never part of the product.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <vector>

#include "../../include/ba_ref_layout.h"
#include "../../include/ba_cuda.h"

/* ---- globals of boondock_airband.cpp ---- */
size_t fft_size = 512;
int device_count = 0;
volatile int do_exit = 0;
int devices_running = 0;
device_t* devices = NULL;
int fm_demod = 0;

void* demodulate_cuda(void* params);

void disable_device_outputs(device_t*) {}

namespace {
std::vector<std::vector<ba_channel_desc> > g_cfg;
std::vector<ba_device_desc> g_devcfg;
struct Feed {
    input_t* in;
    const unsigned char* data;
    size_t len, chunk;
    std::atomic<int> done;
};
struct Taken {
    std::vector<float> wave;
    std::vector<int> axc;
};
std::vector<std::vector<Taken> > g_taken;
std::atomic<long> g_batches(0);
long g_expected = 0; /* hand-offs the streams of this run amount to (all devices); 0 = unknown */
volatile int g_stop_output = 0;

/* circbuffer_append(), input-helpers.cpp:37-63 */
void append(input_t* const input, const unsigned char* buf, size_t len) {
    if (len == 0)
        return;
    pthread_mutex_lock(&input->buffer_lock);
    const size_t space_left = input->buf_size - input->bufe;
    const size_t tail = 2 * (size_t)input->bytes_per_sample * fft_size;
    if (space_left >= len) {
        memcpy(input->buffer + input->bufe, buf, len);
        if (input->bufe == 0)
            memcpy(input->buffer + input->buf_size, input->buffer, std::min(len, tail));
    } else {
        memcpy(input->buffer + input->bufe, buf, space_left);
        memcpy(input->buffer, buf + space_left, len - space_left);
        memcpy(input->buffer + input->buf_size, input->buffer, std::min(len - space_left, tail));
    }
    const size_t old_end = input->bufe;
    input->bufe = (input->bufe + len) % input->buf_size;
    if (old_end < input->bufs && input->bufe >= input->bufs)
        input->overflow_count++;
    pthread_mutex_unlock(&input->buffer_lock);
}

/* an rx thread that is handed its samples faster than real time: like file_rx_thread (input-file.cpp:120-160) it looks at the
 * room between bufe and bufs before it appends */
void* rx_thread(void* arg) {
    Feed* f = (Feed*)arg;
    input_t* in = f->in;
    size_t off = 0;
    while (off < f->len && !do_exit) {
        const size_t n = std::min(f->chunk, f->len - off);
        pthread_mutex_lock(&in->buffer_lock);
        const size_t used = in->bufe >= in->bufs ? in->bufe - in->bufs : in->buf_size - in->bufs + in->bufe;
        pthread_mutex_unlock(&in->buffer_lock);
        if (in->buf_size - used <= n + 1) {
            usleep(200);
            continue;
        }
        append(in, f->data + off, n);
        off += n;
    }
    f->done = 1;
    return NULL;
}

/* the output thread's side of the hand-off, output.cpp:931-951 */
void* output_thread(void*) {
    while (!g_stop_output) {
        bool any = false;
        for (int i = 0; i < device_count; i++) {
            device_t* dev = devices + i;
            if (dev->input->state == INPUT_RUNNING && dev->waveavail) {
                for (int j = 0; j < dev->channel_count; j++) {
                    channel_t* channel = dev->channels + j;
                    Taken& t = g_taken[i][j];
                    t.wave.insert(t.wave.end(), channel->waveout, channel->waveout + WAVE_BATCH); /* process_outputs() reads these */
                    t.axc.push_back((int)channel->axcindicate);
                    memcpy(channel->waveout, channel->waveout + WAVE_BATCH, AGC_EXTRA * 4);
                }
                dev->waveavail = 0;
                g_batches++;
                any = true;
            }
        }
        if (!any)
            usleep(20);
    }
    return NULL;
}
}  // namespace

extern "C" const ba_channel_desc* ba_ref_channel_cfg(int device, int channel) {
    if (device < 0 || device >= (int)g_cfg.size() || channel < 0 || channel >= (int)g_cfg[device].size())
        return NULL;
    return &g_cfg[device][channel];
}

extern "C" const ba_device_desc* ba_ref_device_cfg(int device) {
    if (device < 0 || device >= (int)g_devcfg.size())
        return NULL;
    return &g_devcfg[device];
}

extern "C" {
__attribute__((visibility("default"))) int ba_shim_wave_rate(void) { return WAVE_RATE; }

/* Builds devices[] from `desc` the way parse_devices()/parse_channels() leave them for demodulate() (only the members that
 * thread reads), starts one rx thread per input on iq[i] (appended `chunk` bytes at a time), the output thread and the
 * demodulator thread, lets them run until every byte has been appended and the hand-offs have dried up, and stops them. */
__attribute__((visibility("default"))) int ba_shim_run(const ba_engine_desc* desc, const unsigned char* const* iq, const size_t* bytes, size_t chunk) {
    if (desc->wave_rate != WAVE_RATE)
        return -100;
    fft_size = (size_t)desc->fft_size;
    fm_demod = desc->fm_demod;
    device_count = desc->device_count;
    do_exit = 0;
    g_stop_output = 0;
    g_batches = 0;
    std::vector<device_t> devs(device_count);
    std::vector<input_t> inputs(device_count);
    std::vector<std::vector<channel_t> > chans(device_count);
    std::vector<std::vector<freq_t> > freqs(device_count);
    std::vector<std::vector<size_t> > bins(device_count);
    g_cfg.assign(device_count, std::vector<ba_channel_desc>());
    g_devcfg.assign(desc->devices, desc->devices + device_count);
    g_taken.assign(device_count, std::vector<Taken>());
    for (int i = 0; i < device_count; i++) {
        const ba_device_desc& dd = desc->devices[i];
        input_t& in = inputs[i];
        memset(&in, 0, sizeof(in));
        in.sfmt = (sample_format_t)dd.sample_format;
        in.bytes_per_sample = dd.bytes_per_sample;
        in.fullscale = dd.fullscale;
        in.sample_rate = dd.sample_rate;
        in.centerfreq = dd.centerfreq;
        in.state = INPUT_RUNNING;
        pthread_mutex_init(&in.buffer_lock, NULL);
        device_t& dev = devs[i];
        memset(&dev, 0, sizeof(dev));
        dev.input = &in;
        dev.channel_count = dd.channel_count;
        dev.mode = R_MULTICHANNEL;
        chans[i].resize(dd.channel_count);
        freqs[i].resize(dd.channel_count);
        bins[i].assign(2 * (size_t)dd.channel_count, 0);
        dev.channels = chans[i].data();
        dev.bins = bins[i].data();
        dev.base_bins = bins[i].data() + dd.channel_count;
        g_taken[i].resize(dd.channel_count);
        for (int c = 0; c < dd.channel_count; c++) {
            const ba_channel_desc& cd = dd.channels[c];
            g_cfg[i].push_back(cd);
            channel_t& ch = chans[i][c];
            memset(&ch, 0, sizeof(ch));
            freq_t& f = freqs[i][c];
            memset(&f, 0, sizeof(f));
            f.frequency = cd.frequency;
            f.modulation = (enum modulations)cd.modulation;
            f.ampfactor = cd.ampfactor;
            f.agcavgfast = 0.5f;
            ch.freqlist = &f;
            ch.freq_count = 1;
            ch.freq_idx = 0;
            ch.afc = (unsigned char)cd.afc;
            ch.has_iq_outputs = cd.has_iq_outputs;
            ch.axcindicate = NO_SIGNAL;
        }
    }
    devices = devs.data();
    devices_running = device_count;

    Signal sig;
    demod_params_t dp;
    dp.mp3_signal = &sig;
    dp.device_start = 0;
    dp.device_end = device_count;
    pthread_t demod, outp;
    pthread_create(&outp, NULL, output_thread, NULL);
    pthread_create(&demod, NULL, demodulate_cuda, &dp); /* boondock_airband.cpp:1146-1148 */
    /* the rx threads start once the demodulator has pointed input_t::buffer at its ring (in the reference that happens before
     * input_start()) */
    for (int spin = 0; spin < 200000 && !do_exit; spin++) {
        bool ready = true;
        for (int i = 0; i < device_count; i++)
            ready = ready && inputs[i].buffer != NULL;
        if (ready)
            break;
        usleep(100);
    }
    std::vector<Feed*> feeds;
    for (int i = 0; i < device_count && !do_exit; i++) {
        Feed* f = new Feed();
        f->in = &inputs[i];
        f->data = iq[i];
        f->len = bytes[i];
        f->chunk = chunk;
        f->done = 0;
        feeds.push_back(f);
        pthread_create(&inputs[i].rx_thread, NULL, rx_thread, f);
    }
    int rc = do_exit ? -101 : 0;
    for (Feed* f : feeds)
        pthread_join(f->in->rx_thread, NULL);
    /* everything is in the rings: wait until the hand-offs the streams amount to have been taken (bounded: 120 s without a new
     * one), or, if the caller did not say how many that is, until they have stopped coming for two seconds */
    long seen = -1;
    int quiet = 0;
    while (!do_exit && (g_expected > 0 ? (g_batches < g_expected && quiet < 12000) : quiet < 200)) {
        usleep(10000);
        const long now = g_batches;
        quiet = (now == seen) ? quiet + 1 : 0;
        seen = now;
    }
    do_exit = 1;
    pthread_join(demod, NULL);
    g_stop_output = 1;
    pthread_join(outp, NULL);
    for (Feed* f : feeds)
        delete f;
    for (int i = 0; i < device_count; i++)
        if (inputs[i].overflow_count)
            rc = -102;
    devices = NULL;
    return rc;
}

__attribute__((visibility("default"))) long ba_shim_batches(void) { return g_batches; }
__attribute__((visibility("default"))) void ba_shim_expect_batches(long n) { g_expected = n; }
__attribute__((visibility("default"))) size_t ba_shim_wave(int dev, int ch, const float** data) {
    *data = g_taken[dev][ch].wave.data();
    return g_taken[dev][ch].wave.size();
}
__attribute__((visibility("default"))) size_t ba_shim_axc(int dev, int ch, const int** data) {
    *data = g_taken[dev][ch].axc.data();
    return g_taken[dev][ch].axc.size();
}
/* sizes and offsets of the restated structures, for the layout probe */
__attribute__((visibility("default"))) size_t ba_shim_sizeof(int what) {
    switch (what) {
        case 0: return sizeof(input_t);
        case 1: return sizeof(freq_t);
        case 2: return sizeof(channel_t);
        case 3: return sizeof(device_t);
        case 4: return offsetof(freq_t, squelch);
        case 5: return offsetof(freq_t, notch_filter);
        case 6: return offsetof(freq_t, lowpass_filter);
        case 7: return offsetof(channel_t, waveout);
        case 8: return offsetof(channel_t, axcindicate);
        case 9: return offsetof(device_t, waveavail);
        case 10: return offsetof(input_t, buffer_lock);
    }
    return 0;
}
}
