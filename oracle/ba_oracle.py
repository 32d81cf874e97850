"""ctypes front-end of the CPU parity oracle.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / ``--impl reference`` legs may import this.
``Oracle(cfg)`` loads oracle/libba_oracle.so (restated DSP); ``Oracle(cfg, ref=True)`` loads
oracle/_ref/libba_oracle_ref.so, where the same loop drives the reference's own squelch/ctcss/filters objects.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Optional

import numpy as np

from boondock_airband_b200 import abi

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE_SRC = "/root/reference/src"

_LIBS = {}


def lib_path(ref: bool = False, fast: bool = False) -> str:
    name = "libba_oracle" + ("_ref" if ref else "") + ("_fast" if fast else "") + ".so"
    return os.path.join(HERE, "_ref", name) if ref else os.path.join(HERE, name)


def build(ref: Optional[bool] = None, quiet: bool = True) -> None:
    """Compile the oracle (and oracle/_ref when the reference sources are mounted)."""
    out = subprocess.DEVNULL if quiet else None
    subprocess.check_call(["make", "-C", HERE, "all"], stdout=out)
    if ref is None:
        ref = os.path.isdir(REFERENCE_SRC)
    if ref:
        subprocess.check_call(["make", "-C", HERE, "ref"], stdout=out)


def have_ref() -> bool:
    return os.path.exists(lib_path(ref=True))


def load(ref: bool = False, fast: bool = False):
    key = (ref, fast)
    if key in _LIBS:
        return _LIBS[key]
    path = lib_path(ref, fast)
    if not os.path.exists(path):
        if ref:
            raise FileNotFoundError(path + " (built only where /root/reference is mounted: make -C oracle ref)")
        build(ref=False)
    L = C.CDLL(path)
    vp, f32p, u8p = C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_uint8)
    L.ba_oracle_create.argtypes = [C.POINTER(abi.EngineDesc), C.c_int, C.POINTER(vp)]
    L.ba_oracle_destroy.argtypes = [vp]
    L.ba_oracle_destroy.restype = None
    L.ba_oracle_feed.argtypes = [vp, C.c_int, vp, C.c_size_t]
    L.ba_oracle_set_freq_idx.argtypes = [vp, C.c_int, C.c_int, C.c_uint64, C.c_int]
    for n in ("frames", "batches"):
        f = getattr(L, "ba_oracle_" + n)
        f.argtypes = [vp, C.c_int]
        f.restype = C.c_uint64
    L.ba_oracle_checksum.argtypes = [vp, C.c_int, C.c_int]
    L.ba_oracle_checksum.restype = C.c_double
    for n in ("waveout", "iq_out", "picks", "trace", "status"):
        f = getattr(L, "ba_oracle_" + n)
        f.argtypes = [vp, C.c_int, C.c_int, vp, C.c_size_t]
        f.restype = C.c_size_t
    L.ba_oracle_channel_info.argtypes = [vp, C.c_int, C.c_int, C.POINTER(abi.ChannelInfo)]
    L.ba_oracle_window.argtypes = [vp, f32p, C.c_size_t]
    L.ba_oracle_debug_frames.argtypes = [vp, C.c_int, vp, C.c_size_t, C.c_int, vp, vp]
    L.ba_oracle_run_threads.argtypes = [vp, C.POINTER(vp), C.POINTER(C.c_size_t), C.c_int]
    L.ba_oracle_run_threads.restype = C.c_double
    L.ba_oracle_fft_lanes.argtypes = []
    L.ba_oracle_fft_lanes.restype = C.c_int
    L.ba_oracle_sq_new.restype = vp
    L.ba_oracle_sq_free.argtypes = [vp]
    L.ba_oracle_sq_free.restype = None
    for n in ("set_level", "set_snr", "raw", "filtered", "audio"):
        f = getattr(L, "ba_oracle_sq_" + n)
        f.argtypes = [vp, C.c_float]
        f.restype = None
    L.ba_oracle_sq_set_ctcss.argtypes = [vp, C.c_float, C.c_float]
    L.ba_oracle_sq_set_ctcss.restype = None
    L.ba_oracle_sq_query.argtypes = [vp, C.POINTER(C.c_int32), f32p]
    L.ba_oracle_sq_query.restype = None
    L.ba_oracle_sq_run.argtypes = [vp, vp, vp, vp, C.c_size_t, vp, vp]
    L.ba_oracle_sq_run.restype = None
    L.ba_oracle_notch_run.argtypes = [C.c_float, C.c_float, C.c_float, vp, C.c_size_t, vp]
    L.ba_oracle_notch_run.restype = None
    L.ba_oracle_lowpass_run.argtypes = [C.c_float, C.c_float, vp, vp, C.c_size_t]
    L.ba_oracle_lowpass_run.restype = None
    L.ba_oracle_ctcss_run.argtypes = [C.c_float, C.c_float, C.c_int, vp, C.c_size_t, C.POINTER(C.c_int32)]
    _LIBS[key] = L
    return L


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Oracle:
    """The reference's demodulate() loop on the CPU, fed from memory."""

    def __init__(self, cfg: abi.EngineCfg, ref: bool = False, keep: bool = True, fast: bool = False):
        self.cfg = cfg
        self.L = load(ref, fast)
        desc, self._keep = abi.build_desc(cfg)
        h = C.c_void_p()
        rc = self.L.ba_oracle_create(C.byref(desc), 1 if keep else 0, C.byref(h))
        if rc != 0:
            raise RuntimeError("ba_oracle_create: %s" % abi.ERRORS.get(rc, rc))
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.L.ba_oracle_destroy(self.h)
            self.h = None

    __del__ = close

    def feed(self, dev: int, iq: np.ndarray):
        iq = np.ascontiguousarray(iq)
        rc = self.L.ba_oracle_feed(self.h, dev, _ptr(iq), iq.nbytes)
        if rc != 0:
            raise RuntimeError("ba_oracle_feed: %s" % abi.ERRORS.get(rc, rc))

    def frames(self, dev=0) -> int:
        return int(self.L.ba_oracle_frames(self.h, dev))

    def batches(self, dev=0) -> int:
        return int(self.L.ba_oracle_batches(self.h, dev))

    def checksum(self, dev, ch) -> float:
        return float(self.L.ba_oracle_checksum(self.h, dev, ch))

    def _stream(self, name, dev, ch, dtype):
        f = getattr(self.L, "ba_oracle_" + name)
        n = f(self.h, dev, ch, None, 0)
        out = np.empty(n, dtype=dtype)
        f(self.h, dev, ch, _ptr(out), n)
        return out

    def waveout(self, dev, ch) -> np.ndarray:
        return self._stream("waveout", dev, ch, np.float32)

    def iq_out(self, dev, ch) -> np.ndarray:
        return self._stream("iq_out", dev, ch, np.float32).reshape(-1, 2)

    def picks(self, dev, ch) -> np.ndarray:
        """Raw picked-bin IQ of every frame (needs abi.FLAG_TRACE)."""
        return self._stream("picks", dev, ch, np.float32).reshape(-1, 2)

    def trace(self, dev, ch) -> np.ndarray:
        return self._stream("trace", dev, ch, np.uint8)

    def status(self, dev, ch):
        n = self.L.ba_oracle_status(self.h, dev, ch, None, 0)
        arr = (abi.ChannelStatus * n)()
        self.L.ba_oracle_status(self.h, dev, ch, C.cast(arr, C.c_void_p), n)
        return list(arr)

    def set_freq_idx(self, dev, ch, from_batch, freq_idx):
        """Scan mode: batches >= from_batch of the device run with freqlist[freq_idx] (boondock_airband.cpp:101-139,522)."""
        rc = self.L.ba_oracle_set_freq_idx(self.h, dev, ch, from_batch, freq_idx)
        if rc != 0:
            raise ValueError("ba_oracle_set_freq_idx: %d" % rc)

    def channel_info(self, dev, ch) -> abi.ChannelInfo:
        info = abi.ChannelInfo()
        self.L.ba_oracle_channel_info(self.h, dev, ch, C.byref(info))
        return info

    def window(self) -> np.ndarray:
        w = np.empty(self.cfg.fft_size, np.float32)
        self.L.ba_oracle_window(self.h, w.ctypes.data_as(C.POINTER(C.c_float)), w.size)
        return w

    def debug_frames(self, dev, iq: np.ndarray, n_frames: int, want_in=True, want_out=True):
        iq = np.ascontiguousarray(iq)
        n = self.cfg.fft_size
        fi = np.empty((n_frames, n, 2), np.float32) if want_in else None
        fo = np.empty((n_frames, n, 2), np.float32) if want_out else None
        rc = self.L.ba_oracle_debug_frames(self.h, dev, _ptr(iq), iq.nbytes, n_frames, _ptr(fi), _ptr(fo))
        if rc != 0:
            raise RuntimeError("ba_oracle_debug_frames: %s" % abi.ERRORS.get(rc, rc))
        return fi, fo

    def fft_lanes(self) -> int:
        """0: scalar transform (parity builds); 8: the AVX2 eight-frames-per-vector transform of the timing builds."""
        return int(self.L.ba_oracle_fft_lanes())

    def run_threads(self, iqs, threads: int) -> float:
        """Timing leg: device i consumes iqs[i] entirely; one thread per device, `threads` at a time. Returns seconds."""
        iqs = [np.ascontiguousarray(a) for a in iqs]
        ptrs = (C.c_void_p * len(iqs))(*[a.ctypes.data for a in iqs])
        sizes = (C.c_size_t * len(iqs))(*[a.nbytes for a in iqs])
        return float(self.L.ba_oracle_run_threads(self.h, ptrs, sizes, threads))


class SquelchProbe:
    """A bare Squelch object, driven the way test_squelch.cpp drives it."""

    def __init__(self, ref=False):
        self.L = load(ref)
        self.h = C.c_void_p(self.L.ba_oracle_sq_new())

    def __del__(self):
        if getattr(self, "h", None):
            self.L.ba_oracle_sq_free(self.h)
            self.h = None

    def set_ctcss(self, hz, rate):
        self.L.ba_oracle_sq_set_ctcss(self.h, hz, rate)

    def set_level(self, level):
        self.L.ba_oracle_sq_set_level(self.h, level)

    def set_snr(self, db):
        self.L.ba_oracle_sq_set_snr(self.h, db)

    def raw(self, v):
        self.L.ba_oracle_sq_raw(self.h, v)

    def filtered(self, v):
        self.L.ba_oracle_sq_filtered(self.h, v)

    def audio(self, v):
        self.L.ba_oracle_sq_audio(self.h, v)

    def query(self):
        q = (C.c_int32 * 10)()
        lv = (C.c_float * 3)()
        self.L.ba_oracle_sq_query(self.h, q, lv)
        names = ("is_open", "should_process_audio", "should_filter_sample", "first_open", "last_open", "state",
                 "open_count", "flappy_count", "ctcss_count", "no_ctcss_count")
        d = dict(zip(names, list(q)))
        d.update(noise_level=lv[0], signal_level=lv[1], squelch_level=lv[2])
        return d

    def run(self, raw, filtered=None, audio=None, want_levels=False):
        raw = np.ascontiguousarray(raw, np.float32)
        n = raw.size
        fl = None if filtered is None else np.ascontiguousarray(filtered, np.float32)
        au = None if audio is None else np.ascontiguousarray(audio, np.float32)
        st = np.empty(n, np.uint8)
        lv = np.empty((n, 3), np.float32) if want_levels else None
        self.L.ba_oracle_sq_run(self.h, _ptr(raw), _ptr(fl), _ptr(au), n, _ptr(st), _ptr(lv))
        return (st, lv) if want_levels else st


def notch_run(hz, rate, q, x, ref=False):
    L = load(ref)
    y = np.array(x, np.float32)
    co = np.zeros(3, np.float32)
    L.ba_oracle_notch_run(hz, rate, q, _ptr(y), y.size, _ptr(co))
    return y, co


def lowpass_run(hz, rate, z, ref=False):
    L = load(ref)
    re = np.ascontiguousarray(z.real, np.float32).copy()
    im = np.ascontiguousarray(z.imag, np.float32).copy()
    L.ba_oracle_lowpass_run(hz, rate, _ptr(re), _ptr(im), re.size)
    return re + 1j * im


def ctcss_run(hz, rate, window, x, ref=False):
    L = load(ref)
    x = np.ascontiguousarray(x, np.float32)
    enough = C.c_int32(0)
    tone = L.ba_oracle_ctcss_run(hz, rate, window, _ptr(x), x.size, C.byref(enough))
    return bool(tone), bool(enough.value)


def mix_reference(cfg: abi.EngineCfg, oracles_waveout, oracles_status, mixer: abi.MixerCfg, masked=()):
    """Restates the summing of the reference's mixer (src/mixer.cpp) for batches paired by number.

    mixer_put_samples (mixer.cpp:114-131): an input's batch is waveout[0..WAVE_BATCH) of its channel plus
    has_signal = (axcindicate != NO_SIGNAL) (output.cpp:562-564).  mixer_thread (mixer.cpp:183-206): waveout (and
    waveout_r for MM_STEREO) start at zero, then for every unmasked input with signal, in input order,
    mix_waveforms adds in[s] * (ampfactor * ampl) — product and sum rounded to float separately — unless the factor
    is zero (mixer.cpp:133-141); axcindicate = SIGNAL if any input had signal.  ampl = min(1, 1 - balance),
    ampr = min(1, 1 + balance) (mixer.cpp:79-81).

    oracles_waveout(dev, ch) -> float32 [n], oracles_status(dev, ch) -> list of per-batch status with .axcindicate.
    Returns (left, right or None, axcindicate per batch) for the batches every unmasked input has delivered.
    """
    B = cfg.wave_batch
    live = [(j, i) for j, i in enumerate(mixer.inputs) if j not in masked]
    if not live:
        return np.zeros(0, np.float32), (np.zeros(0, np.float32) if mixer.stereo else None), np.zeros(0, np.int32)
    waves = {j: oracles_waveout(i.device, i.channel) for j, i in live}
    stats = {j: oracles_status(i.device, i.channel) for j, i in live}
    n_batches = min(len(s) for s in stats.values())
    left = np.zeros(n_batches * B, np.float32)
    right = np.zeros(n_batches * B, np.float32) if mixer.stereo else None
    sig = np.full(n_batches, abi.NO_SIGNAL, np.int32)
    for k in range(n_batches):
        sl = slice(k * B, (k + 1) * B)
        for j, i in live:
            if stats[j][k].axcindicate == abi.NO_SIGNAL:
                continue
            sig[k] = abi.SIGNAL
            amp = np.float32(i.ampfactor)
            ml = amp * np.minimum(np.float32(1.0), np.float32(1.0) - np.float32(i.balance))
            mr = amp * np.minimum(np.float32(1.0), np.float32(1.0) + np.float32(i.balance))
            x = waves[j][sl].astype(np.float32)
            if ml != 0:
                left[sl] = left[sl] + x * np.float32(ml)
            if right is not None and mr != 0:
                right[sl] = right[sl] + x * np.float32(mr)
    return left, right, sig
