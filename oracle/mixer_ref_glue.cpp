/*
 * mixer_ref_glue.cpp — TEST INFRASTRUCTURE: drives the reference's OWN mixer (src/mixer.cpp, compiled unmodified where it lies
 * with oracle/ref_stubs standing in for the missing third-party headers) so that the restated mixing arithmetic
 * (oracle/ba_oracle.py: mix_reference, which the device-side mixer K3 is compared with) is pinned against a run of the
 * reference itself.  Provides the globals and helpers mixer.cpp links against (boondock_airband.cpp:71-90, util.cpp) and plays
 * the two neighbours of mixer_thread: the output thread's mixer_put_samples() calls (output.cpp:562-566) and its consumption of
 * the mixed channel (CH_READY -> CH_DIRTY, output.cpp:957-961).
 */
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include "boondock_airband.h"

mixer_t* mixers = NULL;
int mixer_count = 0;
volatile int do_exit = 0;

void disable_channel_outputs(channel_t*) {}
void* xcalloc(size_t nmemb, size_t size, const char*, const int, const char*) { return calloc(nmemb, size); }
void* xrealloc(void* ptr, size_t size, const char*, const int, const char*) { return realloc(ptr, size); }

extern "C" {
__attribute__((visibility("default"))) int ba_mixref_wave_batch(void) { return WAVE_BATCH; }

/* One mixer with n inputs (mixer_connect_input, mixer.cpp:57-94), fed n_batches batches: in[batch][input][WAVE_BATCH],
 * has_signal[batch][input]; mixer_thread (mixer.cpp:160-261) mixes them at its own pace.  out_l / out_r: [n_batches][WAVE_BATCH],
 * axc[n_batches].  Returns MM_STEREO ? 1 : 0, or -1. */
__attribute__((visibility("default"))) int ba_mixref_run(int n_inputs, const float* ampfactor, const float* balance, int n_batches, const float* in, const unsigned char* has_signal,
                                                            float* out_l, float* out_r, int* axc) {
    static mixer_t the_mixer;
    memset(&the_mixer, 0, sizeof(the_mixer));
    the_mixer.name = "pin";
    the_mixer.interval = MIX_DIVISOR;
    the_mixer.channel.state = CH_DIRTY;
    mixers = &the_mixer;
    mixer_count = 1;
    do_exit = 0;
    for (int i = 0; i < n_inputs; i++)
        if (mixer_connect_input(&the_mixer, ampfactor[i], balance[i]) < 0)
            return -1;
    Signal sig;
    pthread_t th;
    pthread_create(&th, NULL, mixer_thread, &sig);
    const int B = WAVE_BATCH;
    for (int b = 0; b < n_batches; b++) {
        for (int i = 0; i < n_inputs; i++)
            mixer_put_samples(&the_mixer, i, in + ((size_t)b * n_inputs + i) * B, has_signal[(size_t)b * n_inputs + i] != 0, B);
        while (the_mixer.channel.state != CH_READY)
            usleep(1000);
        memcpy(out_l + (size_t)b * B, the_mixer.channel.waveout, B * sizeof(float));
        if (the_mixer.channel.mode == MM_STEREO)
            memcpy(out_r + (size_t)b * B, the_mixer.channel.waveout_r, B * sizeof(float));
        axc[b] = (int)the_mixer.channel.axcindicate;
        the_mixer.channel.state = CH_DIRTY; /* output.cpp:960 */
    }
    do_exit = 1;
    pthread_join(th, NULL);
    return the_mixer.channel.mode == MM_STEREO ? 1 : 0;
}
}
