/*
 * ba_host.h — host-side pieces either side of the CUDA hot path, plain C ABI.
 *
 *   ba_conf_*        a reader for the subset of the libconfig grammar the reference's configuration files use and the
 *                    translation of `devices` / `channels` into the descriptors of include/ba_cuda.h, with the rules of
 *                    parse_devices() / parse_channels() (src/config.cpp:298-836): number forms (int Hz, float MHz, "118.5M"
 *                    strings), defaults, validation errors, the disabled-entry skipping and the two silent channel drops.
 *                    libconfig++ is not needed.
 *   ba_file_input_*  the file input driver (src/input-file.cpp:35-181) with a selectable sample format and unpaced replay:
 *                    a reader thread that appends to the input's ring with the arithmetic of circbuffer_append
 *                    (src/input-helpers.cpp:37-63) and waits for a full ring in 20 us .. 1 ms naps instead of 10 ms per poll.
 *
 *   ba_handoff_*     the demodulator -> output thread hand-off (waveavail + Signal, boondock_airband.cpp:673-679,728;
 *                    output.cpp:899-961) with N slots and back-pressure instead of one slot and overruns (row f-4).
 *
 * Nothing here touches CUDA; libba_host.so loads on a machine without a GPU.  The engine is reached only through the
 * function pointers of ba_ring_sink, which ba_file_input_sink_for_engine() fills from libba_cuda.so's entry points.
 */
#ifndef BA_HOST_H
#define BA_HOST_H

#include <stddef.h>
#include <stdint.h>

#include "ba_cuda.h"

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define BA_HOST_API
#else
#define BA_HOST_API __attribute__((visibility("default")))
#endif

/* ---------------------------------------------------------------- configuration front-end (row f-2) */

typedef struct ba_conf ba_conf;

/* error codes of this library (negative), besides BA_OK */
#define BA_HOST_ERR_SYNTAX (-20) /* the text is not in the grammar; ba_host_last_error() names line and column */
#define BA_HOST_ERR_CONFIG (-21) /* a "Configuration error" of config.cpp (the reference prints it and calls error() = _Exit(1)) */
#define BA_HOST_ERR_IO (-22)
#define BA_HOST_ERR_UNSUPPORTED (-23) /* valid for the reference, outside this engine */

/* wave_rate: 8000 = the reference built without -DNFM, 16000 = with it (boondock_airband.h:67-71); 0 = 16000 when any
 * channel says modulation = "nfm", else 8000.  With 8000 a "nfm" channel is the reference's "unknown modulation" error. */
BA_HOST_API int ba_conf_parse_file(const char* path, int wave_rate, ba_conf** out);
BA_HOST_API int ba_conf_parse_text(const char* text, int wave_rate, ba_conf** out);
BA_HOST_API void ba_conf_free(ba_conf* c);
BA_HOST_API const char* ba_host_last_error(void);

/* The engine descriptor built from `devices` (pointers stay valid until ba_conf_free).  cuda_device, flags and
 * max_batches_per_step are left 0 for the caller to fill. */
BA_HOST_API const ba_engine_desc* ba_conf_engine_desc(const ba_conf* c);
BA_HOST_API int ba_conf_device_count(const ba_conf* c);
/* Settings of the i-th enabled device the input driver reads itself: "type", "filepath", "speedup_factor", "sample_format",
 * "index", "serial", "gain", "device_string", ...  Strings come back as written, numbers formatted with %.17g. NULL if absent. */
BA_HOST_API const char* ba_conf_device_setting(const ba_conf* c, int device, const char* key);
/* 1 if the i-th enabled device says mode = "scan" (R_SCAN): one channel whose ba_channel_desc carries the frequency list */
BA_HOST_API int ba_conf_device_is_scan(const ba_conf* c, int device);
/* controller_thread() of scan mode (src/boondock_airband.cpp:101-139) as a step function, one call per 200 ms poll of
 * channels[0].axcindicate.  state = { i, consecutive_squelch_off, last_frequency (start at -1), freq_count }.  Returns the
 * freq_idx to run next (-> ba_cuda_set_freq_idx + input_set_centerfreq when it changed); *tag_freq (optional) receives i
 * when the reference queues a metadata tag (tag_queue_put, :131-134), else -1. */
BA_HOST_API int ba_scan_controller_poll(int32_t state[4], int has_signal, int* tag_freq);
/* name of the m-th enabled entry of the `mixers` section = ba_engine_desc.mixers[m] (parse_mixers, config.cpp:838-889);
 * its inputs are the channel outputs of type "mixer" naming it, in the order the reference connects them */
BA_HOST_API const char* ba_conf_mixer_name(const ba_conf* c, int mixer);
/* root-level switches demodulate()'s callers read (boondock_airband.cpp:852-893) */
BA_HOST_API int ba_conf_multiple_demod_threads(const ba_conf* c);
/* Warnings the reference prints to stderr while parsing (obsolete 'squelch', conflicting thresholds, frequency outside the
 * device's bandwidth) plus a note for every silently dropped channel; '\n'-separated, "" if none. */
BA_HOST_API const char* ba_conf_warnings(const ba_conf* c);
/* index of the configuration's channel entry (position in the `channels` list, disabled ones counted) for channel `ch` of
 * enabled device `device`; -1 if out of range */
BA_HOST_API int ba_conf_channel_source_index(const ba_conf* c, int device, int ch);

/* ---------------------------------------------------------------- file input (row f-1) */

/* Where the reader thread puts bytes: the input_t ring (input-common.h:39-57).  `space` returns the free bytes of the
 * ring, `append` copies n bytes in (it is only called with n <= the space last reported), both are called from the
 * reader thread only. */
typedef struct ba_ring_sink {
    void* ctx;
    size_t (*space)(void* ctx);
    int (*append)(void* ctx, const void* data, size_t n); /* 0 or a negative BA_ERR_* */
} ba_ring_sink;

typedef struct ba_file_input_desc {
    const char* filepath;     /* "filepath" (input-file.cpp:40-45) */
    int32_t sample_format;    /* BA_SFMT_*; the reference's driver is BA_SFMT_U8 only (input-file.cpp:170-172) */
    int32_t sample_rate;      /* Hz, for pacing */
    double speedup_factor;    /* "speedup_factor": replay speed relative to real time, the reference's default is 4; 0 = unpaced */
    size_t chunk_bytes;       /* bytes per read; 0 = half the ring minus one, as input-file.cpp:96 */
    size_t ring_bytes;        /* size of the ring behind `sink` (for the default chunk size) */
    int32_t loop;             /* != 0: rewind at end of file instead of stopping (benchmarks) */
} ba_file_input_desc;

typedef struct ba_file_input ba_file_input;

/* input state, as input_t.state (input-common.h:33-37) */
#define BA_INPUT_UNKNOWN 0
#define BA_INPUT_INITIALIZED 1
#define BA_INPUT_RUNNING 2
#define BA_INPUT_FAILED 3
#define BA_INPUT_STOPPED 4

BA_HOST_API int ba_file_input_open(const ba_file_input_desc* desc, const ba_ring_sink* sink, ba_file_input** out); /* file_init */
BA_HOST_API int ba_file_input_start(ba_file_input* f);                                                              /* run_rx_thread */
BA_HOST_API int ba_file_input_state(const ba_file_input* f); /* BA_INPUT_FAILED after end of file or a read error, like the reference */
BA_HOST_API uint64_t ba_file_input_bytes(const ba_file_input* f); /* bytes appended so far */
BA_HOST_API int ba_file_input_stop(ba_file_input* f);              /* file_stop: joins the thread, closes the file, frees f */

/* A sink that feeds input `dev` of an engine through ba_cuda_submit(); `submit` and `space` are libba_cuda.so's
 * ba_cuda_submit and ba_cuda_input_space (passed as pointers so that this library does not link against CUDA). */
typedef int (*ba_submit_fn)(ba_engine* e, int dev, const void* iq, size_t bytes);
typedef int (*ba_space_fn)(ba_engine* e, int dev, size_t* free_bytes);
BA_HOST_API int ba_file_input_sink_for_engine(ba_engine* e, int dev, ba_submit_fn submit, ba_space_fn space, ba_ring_sink* out);
/* frees what ba_file_input_sink_for_engine() put behind out->ctx (after ba_file_input_stop) */
BA_HOST_API void ba_file_input_sink_release(ba_ring_sink* sink);

/* ---------------------------------------------------------------- output hand-off (row f-4) */

/* The reference hands a batch to the output thread through ONE slot per device: demodulate() sets dev->waveavail = 1 and
 * Signal::send(); if the slot is still full it counts output_overrun_count and overwrites (boondock_airband.cpp:673-679,
 * 728; consumer output.cpp:899-961).  That is right at 1x real time and loses audio at many times real time.  This is the
 * same hand-off with N slots: the producer either waits for a free slot (back-pressure, file replay) or, with timeout 0,
 * counts an overrun like the reference and carries on (live input).  Slots are `slot_bytes` each, 64-byte aligned, and
 * reach the consumer in publish order. */
typedef struct ba_handoff ba_handoff;

#define BA_HANDOFF_TIMEOUT (-30) /* nothing became available within timeout_ms */
#define BA_HANDOFF_CLOSED (-31)  /* ba_handoff_close() was called (do_exit) and, for the consumer, the queue is drained */

BA_HOST_API int ba_handoff_create(int slots, size_t slot_bytes, ba_handoff** out);
/* producer: a free slot to fill.  timeout_ms < 0 waits, 0 polls (BA_HANDOFF_TIMEOUT counts as one overrun). */
BA_HOST_API int ba_handoff_acquire(ba_handoff* h, int timeout_ms, void** slot);
/* producer: hand the filled slot over with a caller-defined tag (device index, batch number ...); wakes a consumer
 * (the Signal::send() of boondock_airband.cpp:728) */
BA_HOST_API int ba_handoff_publish(ba_handoff* h, void* slot, uint64_t tag);
/* consumer: the oldest published slot (the Signal::wait() + waveavail test of output.cpp:908-931) */
BA_HOST_API int ba_handoff_take(ba_handoff* h, int timeout_ms, void** slot, uint64_t* tag);
/* consumer: done with it (waveavail = 0, output.cpp:951) */
BA_HOST_API int ba_handoff_release(ba_handoff* h, void* slot);
/* do_exit: wakes every waiter; producers get BA_HANDOFF_CLOSED at once, consumers after the queue has drained */
BA_HOST_API void ba_handoff_close(ba_handoff* h);
BA_HOST_API uint64_t ba_handoff_overruns(const ba_handoff* h); /* output_overrun_count */
BA_HOST_API void ba_handoff_destroy(ba_handoff* h);

#ifdef __cplusplus
}
#endif
#endif
