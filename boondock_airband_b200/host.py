"""ctypes binding of libba_host.so (include/ba_host.h): configuration front-end and file input.

Rows f-2 and f-1 of SURVEY.md section 8.  The C++ library does the work (grammar, parse_channels() rules of
src/config.cpp:298-836, reader thread of src/input-file.cpp:82-147); this module only turns its descriptors into the
``EngineCfg`` dataclasses of ``abi.py`` so that a configuration file can be handed to ``Engine`` (and, in the tests,
to the oracle).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import List, Optional

from . import abi

HERE = os.path.dirname(os.path.abspath(__file__))
DEFAULT_LIB = os.path.join(HERE, "libba_host.so")
_LIB = {}

ERR_SYNTAX, ERR_CONFIG, ERR_IO, ERR_UNSUPPORTED = -20, -21, -22, -23
INPUT_UNKNOWN, INPUT_INITIALIZED, INPUT_RUNNING, INPUT_FAILED, INPUT_STOPPED = range(5)

# every symbol include/ba_host.h declares
SYMBOLS = (
    "ba_conf_parse_file", "ba_conf_parse_text", "ba_conf_free", "ba_host_last_error", "ba_conf_engine_desc",
    "ba_conf_device_count", "ba_conf_device_setting", "ba_conf_device_is_scan", "ba_conf_mixer_name", "ba_conf_multiple_demod_threads", "ba_conf_warnings",
    "ba_conf_channel_source_index", "ba_file_input_open", "ba_file_input_start", "ba_file_input_state",
    "ba_file_input_bytes", "ba_file_input_stop", "ba_file_input_sink_for_engine", "ba_file_input_sink_release",
    "ba_handoff_create", "ba_handoff_acquire", "ba_handoff_publish", "ba_handoff_take", "ba_handoff_release", "ba_handoff_close",
    "ba_handoff_overruns", "ba_handoff_destroy", "ba_scan_controller_poll",
)
HANDOFF_TIMEOUT, HANDOFF_CLOSED = -30, -31


class RingSink(C.Structure):
    _fields_ = [("ctx", C.c_void_p), ("space", C.c_void_p), ("append", C.c_void_p)]


class FileInputDesc(C.Structure):
    _fields_ = [
        ("filepath", C.c_char_p),
        ("sample_format", C.c_int32),
        ("sample_rate", C.c_int32),
        ("speedup_factor", C.c_double),
        ("chunk_bytes", C.c_size_t),
        ("ring_bytes", C.c_size_t),
        ("loop", C.c_int32),
    ]


class ConfigError(RuntimeError):
    def __init__(self, code: int, text: str):
        super().__init__(text)
        self.code = code


def load_library(path: Optional[str] = None):
    path = path or DEFAULT_LIB
    if path in _LIB:
        return _LIB[path]
    if not os.path.exists(path):
        raise RuntimeError("%s is missing: run `make -C boondock_airband_b200/csrc`" % path)
    L = C.CDLL(path)
    vp = C.c_void_p
    L.ba_conf_parse_file.argtypes = [C.c_char_p, C.c_int, C.POINTER(vp)]
    L.ba_conf_parse_text.argtypes = [C.c_char_p, C.c_int, C.POINTER(vp)]
    L.ba_conf_free.argtypes = [vp]
    L.ba_conf_free.restype = None
    L.ba_host_last_error.restype = C.c_char_p
    L.ba_conf_engine_desc.argtypes = [vp]
    L.ba_conf_engine_desc.restype = C.POINTER(abi.EngineDesc)
    L.ba_conf_device_count.argtypes = [vp]
    L.ba_conf_device_setting.argtypes = [vp, C.c_int, C.c_char_p]
    L.ba_conf_device_setting.restype = C.c_char_p
    L.ba_conf_device_is_scan.argtypes = [vp, C.c_int]
    L.ba_conf_mixer_name.argtypes = [vp, C.c_int]
    L.ba_conf_mixer_name.restype = C.c_char_p
    L.ba_conf_multiple_demod_threads.argtypes = [vp]
    L.ba_conf_warnings.argtypes = [vp]
    L.ba_conf_warnings.restype = C.c_char_p
    L.ba_conf_channel_source_index.argtypes = [vp, C.c_int, C.c_int]
    L.ba_file_input_open.argtypes = [C.POINTER(FileInputDesc), C.POINTER(RingSink), C.POINTER(vp)]
    L.ba_file_input_start.argtypes = [vp]
    L.ba_file_input_state.argtypes = [vp]
    L.ba_file_input_bytes.argtypes = [vp]
    L.ba_file_input_bytes.restype = C.c_uint64
    L.ba_file_input_stop.argtypes = [vp]
    L.ba_file_input_sink_for_engine.argtypes = [vp, C.c_int, vp, vp, C.POINTER(RingSink)]
    L.ba_file_input_sink_release.argtypes = [C.POINTER(RingSink)]
    L.ba_file_input_sink_release.restype = None
    L.ba_handoff_create.argtypes = [C.c_int, C.c_size_t, C.POINTER(vp)]
    L.ba_handoff_acquire.argtypes = [vp, C.c_int, C.POINTER(vp)]
    L.ba_handoff_publish.argtypes = [vp, vp, C.c_uint64]
    L.ba_handoff_take.argtypes = [vp, C.c_int, C.POINTER(vp), C.POINTER(C.c_uint64)]
    L.ba_handoff_release.argtypes = [vp, vp]
    L.ba_handoff_close.argtypes = [vp]
    L.ba_handoff_close.restype = None
    L.ba_handoff_overruns.argtypes = [vp]
    L.ba_handoff_overruns.restype = C.c_uint64
    L.ba_handoff_destroy.argtypes = [vp]
    L.ba_handoff_destroy.restype = None
    _LIB[path] = L
    return L


_SFMT_NAME = {v: k for k, v in abi.SFMT.items()}
_MOD_NAME = {v: k for k, v in abi.MOD.items()}


class Config:
    """A parsed configuration file: ``cfg`` is the EngineCfg for Engine(); ``settings(dev)`` the driver keys."""

    def __init__(self, handle, L):
        self._h, self._L = handle, L
        d = L.ba_conf_engine_desc(handle).contents
        cfg = abi.EngineCfg(fft_size=d.fft_size, wave_rate=d.wave_rate, fm_demod=d.fm_demod)
        for i in range(d.device_count):
            dd = d.devices[i]
            dev = abi.DeviceCfg(sample_rate=dd.sample_rate, centerfreq=dd.centerfreq, sample_format=_SFMT_NAME[dd.sample_format],
                                fullscale=dd.fullscale, tau=dd.tau_us)
            for j in range(dd.channel_count):
                c = dd.channels[j]
                dev.channels.append(abi.ChannelCfg(
                    freq=c.frequency, modulation=_MOD_NAME[c.modulation], afc=c.afc, ampfactor=c.ampfactor,
                    squelch_threshold=c.squelch_threshold_dbfs, squelch_snr_threshold=c.squelch_snr_threshold, notch=c.notch,
                    notch_q=c.notch_q, ctcss=c.ctcss, bandwidth=c.bandwidth, tau=c.tau_us, has_iq_outputs=bool(c.has_iq_outputs),
                    freqs=[abi.FreqCfg(f.frequency, _MOD_NAME[f.modulation], f.ampfactor, f.squelch_threshold_dbfs, f.squelch_snr_threshold, f.notch,
                                       f.notch_q, f.ctcss, f.bandwidth) for f in (c.freqs[k] for k in range(c.freq_count))]))
            cfg.devices.append(dev)
        for m in range(d.mixer_count):
            md = d.mixers[m]
            cfg.mixers.append(abi.MixerCfg(L.ba_conf_mixer_name(handle, m).decode(), [
                abi.MixerInputCfg(md.inputs[j].device, md.inputs[j].channel, md.inputs[j].ampfactor, md.inputs[j].balance) for j in range(md.input_count)]))
        self.cfg = cfg
        self.multiple_demod_threads = bool(L.ba_conf_multiple_demod_threads(handle))
        w = L.ba_conf_warnings(handle).decode()
        self.warnings: List[str] = w.split("\n") if w else []

    def setting(self, dev: int, key: str) -> Optional[str]:
        v = self._L.ba_conf_device_setting(self._h, dev, key.encode())
        return None if v is None else v.decode()

    def is_scan(self, dev: int) -> bool:
        return bool(self._L.ba_conf_device_is_scan(self._h, dev))

    def source_index(self, dev: int, ch: int) -> int:
        return self._L.ba_conf_channel_source_index(self._h, dev, ch)

    def close(self):
        if self._h:
            self._L.ba_conf_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _finish(L, rc, h) -> Config:
    if rc != 0:
        raise ConfigError(rc, L.ba_host_last_error().decode())
    return Config(h, L)


def parse_text(text: str, wave_rate: int = 0) -> Config:
    L = load_library()
    h = C.c_void_p()
    return _finish(L, L.ba_conf_parse_text(text.encode(), wave_rate, C.byref(h)), h)


def parse_file(path: str, wave_rate: int = 0) -> Config:
    L = load_library()
    h = C.c_void_p()
    return _finish(L, L.ba_conf_parse_file(path.encode(), wave_rate, C.byref(h)), h)


class FileInput:
    """The file input driver feeding input ``dev`` of an engine (file_init / run_rx_thread / file_stop)."""

    def __init__(self, engine, dev: int, path: str, sample_format: str = "u8", sample_rate: int = 0,
                 speedup_factor: float = 0.0, chunk_bytes: int = 0, loop: bool = False):
        self.L = load_library()
        self.sink = RingSink()
        _, ring_bytes, _ = engine.input_ring(dev)
        submit = C.cast(engine.L.ba_cuda_submit, C.c_void_p)
        space = C.cast(engine.L.ba_cuda_input_space, C.c_void_p)
        rc = self.L.ba_file_input_sink_for_engine(engine.h, dev, submit, space, C.byref(self.sink))
        if rc != 0:
            raise ConfigError(rc, self.L.ba_host_last_error().decode())
        desc = FileInputDesc(path.encode(), abi.SFMT[sample_format], sample_rate, speedup_factor, chunk_bytes, ring_bytes, 1 if loop else 0)
        self.h = C.c_void_p()
        rc = self.L.ba_file_input_open(C.byref(desc), C.byref(self.sink), C.byref(self.h))
        if rc != 0:
            self.L.ba_file_input_sink_release(C.byref(self.sink))
            raise ConfigError(rc, self.L.ba_host_last_error().decode())

    def start(self):
        rc = self.L.ba_file_input_start(self.h)
        if rc != 0:
            raise ConfigError(rc, self.L.ba_host_last_error().decode())

    @property
    def state(self) -> int:
        return self.L.ba_file_input_state(self.h)

    @property
    def bytes(self) -> int:
        return int(self.L.ba_file_input_bytes(self.h))

    def stop(self):
        if self.h:
            self.L.ba_file_input_stop(self.h)
            self.h = None
            self.L.ba_file_input_sink_release(C.byref(self.sink))
