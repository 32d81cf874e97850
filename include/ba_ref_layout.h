/*
 * ba_ref_layout.h — the structures Boondock-Airband's demodulate() thread shares with the input drivers above it and
 * the output / mixer threads below it, restated with the reference's names, member order and types so that
 * boondock_airband_b200/csrc/demodulate_cuda.cpp compiles — unchanged — either inside the reference tree (against
 * boondock_airband.h, -DBA_WITH_REFERENCE_HEADERS) or stand-alone against this header (this repository's tests, which
 * play the rx thread and the output thread themselves).
 *
 * Follows (all under /root/reference/src):
 *   input_t, sample_format_t, input_state_t          input-common.h:32-57
 *   WAVE_RATE / WAVE_BATCH / AGC_EXTRA / WAVE_LEN     boondock_airband.h:66-75
 *   status, ch_states, mix_modes, modulations         boondock_airband.h:99-101, 202-208
 *   Signal                                            boondock_airband.h:210-230
 *   freq_t, channel_t, device_t, demod_params_t       boondock_airband.h:232-326 (demod_params_t without the FFTW members,
 *                                                     as under WITH_BCM_VC, :321-325)
 * Members the hot path never touches keep their size and position but not their type: Squelch, NotchFilter,
 * LowpassFilter (squelch.h, filters.h), output_t and the libconfig++ reference in input_t::parse_config are opaque here.
 * Their sizes are the x86-64 / GCC ones of the reference's own headers; tests/test_shim.py compiles a probe against
 * /root/reference/src (where it is mounted) and compares sizeof / alignof.  Nothing here is copied code: these are
 * declarations an ABI-compatible peer has to repeat.
 */
#ifndef BA_REF_LAYOUT_H
#define BA_REF_LAYOUT_H

#include <pthread.h>
#include <stddef.h>
#include <stdint.h>
#include <sys/time.h>

/* ---- input-common.h:32-57 ---- */
typedef enum { SFMT_UNDEF = 0, SFMT_U8, SFMT_S8, SFMT_S16, SFMT_F32 } sample_format_t;
typedef enum { INPUT_UNKNOWN = 0, INPUT_INITIALIZED, INPUT_RUNNING, INPUT_FAILED, INPUT_STOPPED, INPUT_DISABLED } input_state_t;

typedef struct input_t input_t;
struct input_t {
    unsigned char* buffer;
    void* dev_data;
    size_t buf_size, bufs, bufe;
    size_t overflow_count;
    input_state_t state;
    sample_format_t sfmt;
    float fullscale;
    int bytes_per_sample;
    int sample_rate;
    int centerfreq;
    int (*parse_config)(input_t* const input, void* /* libconfig::Setting& */ cfg);
    int (*init)(input_t* const input);
    void* (*run_rx_thread)(void* input_ptr);
    int (*set_centerfreq)(input_t* const input, int const centerfreq);
    int (*stop)(input_t* const input);
    pthread_t rx_thread;
    pthread_mutex_t buffer_lock;
};

/* ---- boondock_airband.h:66-75 ---- */
#ifdef NFM
#define WAVE_RATE 16000
#else
#define WAVE_RATE 8000
#endif
#define WAVE_BATCH WAVE_RATE / 8
#define AGC_EXTRA 100
#define WAVE_LEN 2 * WAVE_BATCH + AGC_EXTRA
#define TAG_QUEUE_LEN 16

enum status { NO_SIGNAL = ' ', SIGNAL = '*', AFC_UP = '<', AFC_DOWN = '>' };
enum ch_states { CH_DIRTY, CH_WORKING, CH_READY };
enum mix_modes { MM_MONO, MM_STEREO };
enum modulations {
    MOD_AM
#ifdef NFM
    ,
    MOD_NFM
#endif
};
enum rec_modes { R_MULTICHANNEL, R_SCAN };

/* ---- boondock_airband.h:210-230 ---- */
class Signal {
   public:
    Signal(void) {
        pthread_cond_init(&cond_, NULL);
        pthread_mutex_init(&mutex_, NULL);
    }
    void send(void) {
        pthread_mutex_lock(&mutex_);
        pthread_cond_signal(&cond_);
        pthread_mutex_unlock(&mutex_);
    }
    void wait(void) {
        pthread_mutex_lock(&mutex_);
        pthread_cond_wait(&cond_, &mutex_);
        pthread_mutex_unlock(&mutex_);
    }

   private:
    pthread_cond_t cond_;
    pthread_mutex_t mutex_;
};

/* opaque stand-ins, sizes of squelch.h / filters.h on x86-64 (checked by tests/test_shim.py against the reference's headers) */
struct alignas(8) ba_ref_opaque_squelch {
    unsigned char bytes[312];
};
struct alignas(4) ba_ref_opaque_notch {
    unsigned char bytes[48];
};
struct alignas(4) ba_ref_opaque_lowpass {
    unsigned char bytes[68];
};
struct output_t; /* boondock_airband.h:181-196: only ever a pointer here */

struct freq_tag {
    int freq;
    struct timeval tv;
};

/* ---- boondock_airband.h:232-242 ---- */
struct freq_t {
    int frequency;
    char* label;
    float agcavgfast;
    float ampfactor;
    ba_ref_opaque_squelch squelch;
    size_t active_counter;
    ba_ref_opaque_notch notch_filter;
    ba_ref_opaque_lowpass lowpass_filter;
    enum modulations modulation;
};

/* ---- boondock_airband.h:243-269 ---- */
struct channel_t {
    float wavein[WAVE_LEN];
    float waveout[WAVE_LEN];
    float waveout_r[WAVE_LEN];
    float iq_in[2 * WAVE_LEN];
    float iq_out[2 * WAVE_LEN];
#ifdef NFM
    float pr;
    float pj;
    float prev_waveout;
    float alpha;
#endif
    uint32_t dm_dphi, dm_phi;
    enum mix_modes mode;
    status axcindicate;
    unsigned char afc;
    struct freq_t* freqlist;
    int freq_count;
    int freq_idx;
    int needs_raw_iq;
    int has_iq_outputs;
    enum ch_states state;
    int output_count;
    output_t* outputs;
    int highpass;
    int lowpass;
};

/* ---- boondock_airband.h:272-292 ---- */
struct device_t {
    input_t* input;
#ifdef NFM
    float alpha;
#endif
    int channel_count;
    size_t *base_bins, *bins;
    channel_t* channels;
    int waveend;
    int waveavail;
    pthread_t controller_thread;
    struct freq_tag tag_queue[TAG_QUEUE_LEN];
    int tq_head, tq_tail;
    int last_frequency;
    pthread_mutex_t tag_queue_lock;
    int row;
    int failed;
    enum rec_modes mode;
    size_t output_overrun_count;
};

/* ---- boondock_airband.h:316-326, the members left when the FFTW ones are compiled out ---- */
struct demod_params_t {
    Signal* mp3_signal;
    int device_start;
    int device_end;
};

/* ---- globals of boondock_airband.cpp the thread body reads (:71-90; boondock_airband.h:347-358) ---- */
extern size_t fft_size;
extern int device_count;
extern volatile int do_exit;
extern int devices_running;
extern device_t* devices;
extern int fm_demod; /* enum fm_demod_algo, boondock_airband.cpp:88; FM_FAST_ATAN2 = 0, FM_QUADRI_DEMOD = 1 */

#endif
