#!/usr/bin/env python
"""Golden vector for the mixer: a run of the reference's OWN mixer.cpp (oracle/_ref/libba_mixer_ref.so = src/mixer.cpp +
src/logging.cpp compiled unmodified, driven by oracle/mixer_ref_glue.cpp through mixer_connect_input / mixer_put_samples /
mixer_thread).  Needs /root/reference (make -C oracle ref).  Writes tests/golden/golden_mixer.npz.

    python tests/golden/make_golden_mixer.py
"""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
LIB = os.path.join(ROOT, "oracle", "_ref", "libba_mixer_ref.so")


def case():
    rng = np.random.default_rng(0xB00D)
    ampfactor = np.array([1.0, 0.5, 2.5, 0.0, 0.8], np.float32)     # one input with a zero multiplier (mix_waveforms returns early)
    balance = np.array([0.0, -0.4, 0.4, 0.2, 1.0], np.float32)      # hard right included: ampl = 0
    n_in, n_b, B = len(ampfactor), 5, 1000
    x = (rng.standard_normal((n_b, n_in, B)) * 0.3).astype(np.float32)
    sig = np.array([[1, 1, 1, 1, 1], [1, 0, 1, 0, 0], [0, 0, 0, 0, 0], [0, 1, 0, 1, 1], [1, 1, 0, 0, 1]], np.uint8)
    return ampfactor, balance, x, sig


def run_reference(ampfactor, balance, x, sig):
    L = C.CDLL(LIB)
    L.ba_mixref_wave_batch.restype = C.c_int
    n_b, n_in, B = x.shape
    assert L.ba_mixref_wave_batch() == B
    fp = C.POINTER(C.c_float)
    L.ba_mixref_run.argtypes = [C.c_int, fp, fp, C.c_int, fp, C.POINTER(C.c_ubyte), fp, fp, C.POINTER(C.c_int)]
    L.ba_mixref_run.restype = C.c_int
    left = np.zeros((n_b, B), np.float32)
    right = np.zeros((n_b, B), np.float32)
    axc = np.zeros(n_b, np.int32)
    x = np.ascontiguousarray(x)
    sig = np.ascontiguousarray(sig)
    rc = L.ba_mixref_run(n_in, ampfactor.ctypes.data_as(fp), balance.ctypes.data_as(fp), n_b, x.ctypes.data_as(fp), sig.ctypes.data_as(C.POINTER(C.c_ubyte)),
                         left.ctypes.data_as(fp), right.ctypes.data_as(fp), axc.ctypes.data_as(C.POINTER(C.c_int)))
    assert rc in (0, 1), rc
    return left, right, axc, rc


if __name__ == "__main__":
    if not os.path.exists(LIB):
        sys.exit("build it first: make -C oracle ref (needs /root/reference)")
    a, b, x, s = case()
    left, right, axc, stereo = run_reference(a, b, x, s)
    np.savez_compressed(os.path.join(HERE, "golden_mixer.npz"), ampfactor=a, balance=b, x=x.astype(np.float16).astype(np.float32) if False else x, has_signal=s,
                        left=left, right=right, axcindicate=axc, stereo=np.int32(stereo))
    print("golden_mixer.npz: %d batches, %d inputs, stereo=%d" % (x.shape[0], x.shape[1], stereo))
