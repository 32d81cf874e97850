/*
 * oracle.h — C entry points of the CPU parity oracle (TEST INFRASTRUCTURE ONLY; see oracle.cpp).
 * The descriptors are the ones of include/ba_cuda.h so that one configuration drives both sides of a parity test.
 */
#ifndef BA_ORACLE_H
#define BA_ORACLE_H
#include <stddef.h>
#include <stdint.h>
#include "../include/ba_cuda.h"
#ifdef __cplusplus
extern "C" {
#endif
typedef struct ba_oracle ba_oracle;

int ba_oracle_is_reference_build(void); /* 1: DSP classes are the reference's own objects (oracle/_ref) */
/* keep != 0: record every output stream (tests); keep == 0: checksum only (timing) */
int ba_oracle_create(const ba_engine_desc* desc, int keep, ba_oracle** out);
void ba_oracle_destroy(ba_oracle* o);
/* append interleaved IQ bytes of device `dev` and run every frame the reference's availability test admits */
int ba_oracle_feed(ba_oracle* o, int dev, const void* iq, size_t bytes);
uint64_t ba_oracle_frames(ba_oracle* o, int dev);
uint64_t ba_oracle_batches(ba_oracle* o, int dev);
double ba_oracle_checksum(ba_oracle* o, int dev, int ch);
/* recorded streams; each returns the number of elements available and copies min(count, available) */
size_t ba_oracle_waveout(ba_oracle* o, int dev, int ch, float* out, size_t count);
size_t ba_oracle_iq_out(ba_oracle* o, int dev, int ch, float* out, size_t count);
size_t ba_oracle_picks(ba_oracle* o, int dev, int ch, float* out, size_t count);
size_t ba_oracle_trace(ba_oracle* o, int dev, int ch, uint8_t* out, size_t count);
size_t ba_oracle_status(ba_oracle* o, int dev, int ch, ba_channel_status* out, size_t count);
int ba_oracle_channel_info(ba_oracle* o, int dev, int ch, ba_channel_info* out);
int ba_oracle_set_freq_idx(ba_oracle* o, int dev, int ch, uint64_t from_batch, int freq_idx);
int ba_oracle_window(ba_oracle* o, float* out, size_t count);
int ba_oracle_debug_frames(ba_oracle* o, int dev, const void* iq, size_t bytes, int n_frames, float* fftin, float* fftout);
double ba_oracle_run_threads(ba_oracle* o, const void* const* iq, const size_t* bytes, int threads);
/* 0: scalar transform (parity builds); 8: AVX2 transform, eight frames per vector (timing builds, -DBA_ORACLE_FAST) */
int ba_oracle_fft_lanes(void);

void* ba_oracle_sq_new(void);
void ba_oracle_sq_free(void* s);
void ba_oracle_sq_set_ctcss(void* s, float hz, float rate);
void ba_oracle_sq_set_level(void* s, float level);
void ba_oracle_sq_set_snr(void* s, float db);
void ba_oracle_sq_raw(void* s, float v);
void ba_oracle_sq_filtered(void* s, float v);
void ba_oracle_sq_audio(void* s, float v);
void ba_oracle_sq_query(void* s, int32_t* out10, float* levels3);
void ba_oracle_sq_run(void* s, const float* raw, const float* filtered, const float* audio, size_t n, uint8_t* states, float* levels3);
void ba_oracle_notch_run(float hz, float rate, float q, float* x, size_t n, float* coeffs3);
void ba_oracle_lowpass_run(float hz, float rate, float* re, float* im, size_t n);
int ba_oracle_ctcss_run(float hz, float rate, int window, const float* x, size_t n, int32_t* enough);
#ifdef __cplusplus
}
#endif
#endif
