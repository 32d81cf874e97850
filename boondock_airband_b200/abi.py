"""ctypes mirror of include/ba_cuda.h plus the host-side channel model.

The dataclasses follow the reference's libconfig channel model (src/config.cpp:312-836): one
``DeviceCfg`` per ``devices[]`` entry (its ``input_t`` fields: sample format, rate, centre frequency)
and one ``ChannelCfg`` per ``channels[]`` entry in multichannel mode, with the same option names
(``freq``, ``modulation``, ``afc``, ``ampfactor``, ``squelch_threshold``, ``squelch_snr_threshold``,
``notch``, ``notch_q``, ``ctcss``, ``bandwidth``, ``tau``).  Derived constants (bin, dm_dphi,
filter coefficients) are computed by the engine, never here.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

ABI_VERSION = 3

SFMT = {"u8": 1, "s8": 2, "s16": 3, "f32": 4}
SFMT_BYTES = {"u8": 1, "s8": 1, "s16": 2, "f32": 4}
MOD = {"am": 0, "nfm": 1}
FM_FAST_ATAN2, FM_QUADRI_DEMOD = 0, 1
AGC_EXTRA = 100
NO_SIGNAL, SIGNAL, AFC_UP, AFC_DOWN = ord(" "), ord("*"), ord("<"), ord(">")
SQ_CLOSED, SQ_OPENING, SQ_CLOSING, SQ_LOW_SIGNAL_ABORT, SQ_OPEN = range(5)
FLAG_TRACE = 0x1
FLAG_KEEP_PICKS = 0x2
FLAG_RESULTS_ON_DEVICE = 0x4
FLAG_SKIP_SILENT_ROWS = 0x8
TRACE_STATE_MASK, TRACE_OPEN, TRACE_AUDIO, TRACE_FILTERED = 0x07, 0x08, 0x10, 0x20

OK = 0
ERRORS = {-1: "NO_DEVICE", -2: "BAD_SIZE", -3: "NOMEM", -4: "BAD_ARG", -5: "CUDA", -6: "OVERRUN", -7: "STATE"}


class FreqDesc(C.Structure):
    _fields_ = [
        ("frequency", C.c_int32),
        ("modulation", C.c_int32),
        ("ampfactor", C.c_float),
        ("squelch_threshold_dbfs", C.c_int32),
        ("squelch_snr_threshold", C.c_float),
        ("notch", C.c_float),
        ("notch_q", C.c_float),
        ("ctcss", C.c_float),
        ("bandwidth", C.c_int32),
    ]


class ChannelDesc(C.Structure):
    _fields_ = [
        ("frequency", C.c_int32),
        ("modulation", C.c_int32),
        ("afc", C.c_int32),
        ("ampfactor", C.c_float),
        ("squelch_threshold_dbfs", C.c_int32),
        ("squelch_snr_threshold", C.c_float),
        ("notch", C.c_float),
        ("notch_q", C.c_float),
        ("ctcss", C.c_float),
        ("bandwidth", C.c_int32),
        ("tau_us", C.c_int32),
        ("has_iq_outputs", C.c_int32),
        ("freq_count", C.c_int32),
        ("freqs", C.POINTER(FreqDesc)),
    ]


class DeviceDesc(C.Structure):
    _fields_ = [
        ("sample_format", C.c_int32),
        ("bytes_per_sample", C.c_int32),
        ("fullscale", C.c_float),
        ("sample_rate", C.c_int32),
        ("centerfreq", C.c_int32),
        ("tau_us", C.c_int32),
        ("channel_count", C.c_int32),
        ("channels", C.POINTER(ChannelDesc)),
    ]


class MixerInputDesc(C.Structure):
    _fields_ = [("device", C.c_int32), ("channel", C.c_int32), ("ampfactor", C.c_float), ("balance", C.c_float)]


class MixerDesc(C.Structure):
    _fields_ = [("input_count", C.c_int32), ("inputs", C.POINTER(MixerInputDesc))]


class MixerOut(C.Structure):
    _fields_ = [
        ("n_batches", C.c_int32),
        ("wave_batch", C.c_int32),
        ("stereo", C.c_int32),
        ("pad0", C.c_int32),
        ("first_batch", C.c_uint64),
        ("waveout", C.POINTER(C.c_float)),
        ("waveout_r", C.POINTER(C.c_float)),
        ("axcindicate", C.POINTER(C.c_int32)),
    ]


class EngineDesc(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32),
        ("fft_size", C.c_int32),
        ("wave_rate", C.c_int32),
        ("fm_demod", C.c_int32),
        ("cuda_device", C.c_int32),
        ("device_count", C.c_int32),
        ("devices", C.POINTER(DeviceDesc)),
        ("max_batches_per_step", C.c_int32),
        ("flags", C.c_uint32),
        ("ring_bytes", C.c_uint64),
        ("mixer_count", C.c_int32),
        ("mixers", C.POINTER(MixerDesc)),
    ]


class ChannelStatus(C.Structure):
    _fields_ = [
        ("axcindicate", C.c_int32),
        ("bin", C.c_uint32),
        ("signal_level", C.c_float),
        ("noise_level", C.c_float),
        ("squelch_level", C.c_float),
        ("open_count", C.c_uint32),
        ("flappy_count", C.c_uint32),
        ("ctcss_count", C.c_uint32),
        ("no_ctcss_count", C.c_uint32),
        ("active_counter", C.c_uint32),
    ]


class StepOut(C.Structure):
    _fields_ = [
        ("n_batches", C.c_int32),
        ("wave_batch", C.c_int32),
        ("channel_count", C.c_int32),
        ("wave_stride", C.c_int32),
        ("waveout", C.POINTER(C.c_float)),
        ("iq_out", C.POINTER(C.c_float)),
        ("trace", C.POINTER(C.c_uint8)),
        ("status", C.POINTER(ChannelStatus)),
        ("frames_done", C.c_uint64),
        ("rows", C.POINTER(C.c_float)),
        ("row_of", C.POINTER(C.c_int32)),
        ("n_rows", C.c_uint32),
        ("row_of_stride", C.c_int32),
    ]


class ChannelInfo(C.Structure):
    _fields_ = [
        ("bin", C.c_uint32),
        ("dm_dphi", C.c_uint32),
        ("needs_raw_iq", C.c_int32),
        ("alpha", C.c_float),
        ("squelch_ratio", C.c_float),
        ("manual_level", C.c_float),
        ("notch_enabled", C.c_int32),
        ("notch_d", C.c_float * 3),
        ("lowpass_enabled", C.c_int32),
        ("lowpass_ycoeffs", C.c_float * 2),
        ("lowpass_gain", C.c_float),
        ("ctcss_fast_tones", C.c_int32),
        ("ctcss_slow_tones", C.c_int32),
        ("ctcss_fast_window", C.c_int32),
        ("ctcss_slow_window", C.c_int32),
    ]

    def as_dict(self):
        d = {}
        for name, _ in self._fields_:
            v = getattr(self, name)
            d[name] = list(v) if hasattr(v, "__len__") else v
        return d


@dataclass
class FreqCfg:
    """One frequency of a scan-mode channel: the freq_t fields (config.cpp:364-433)."""

    freq: int
    modulation: str = "am"
    ampfactor: float = 1.0
    squelch_threshold: int = 0
    squelch_snr_threshold: float = -1.0
    notch: float = 0.0
    notch_q: float = 0.0
    ctcss: float = 0.0
    bandwidth: int = 0


@dataclass
class ChannelCfg:
    """One ``channels[]`` entry of a multichannel device (config.cpp:312-729)."""

    freq: int  # Hz
    modulation: str = "am"
    afc: int = 0
    ampfactor: float = 1.0
    squelch_threshold: int = 0  # dBFS, 0 = automatic
    squelch_snr_threshold: float = -1.0  # dB, <0 = keep the default 9.54 dB
    notch: float = 0.0
    notch_q: float = 0.0  # 0 = default 10.0
    ctcss: float = 0.0
    bandwidth: int = 0
    tau: int = -1  # microseconds, <0 = inherit
    has_iq_outputs: bool = False
    freqs: List[FreqCfg] = field(default_factory=list)  # scan mode: the channel's frequency list; empty = multichannel


@dataclass
class DeviceCfg:
    """One ``devices[]`` entry and its ``input_t`` (config.cpp:731-836, input-common.h:39-57)."""

    sample_rate: int
    centerfreq: int
    sample_format: str = "u8"
    fullscale: Optional[float] = None
    tau: int = -1
    channels: List[ChannelCfg] = field(default_factory=list)

    @property
    def bytes_per_sample(self) -> int:
        return SFMT_BYTES[self.sample_format]

    def default_fullscale(self) -> float:
        if self.fullscale is not None:
            return float(self.fullscale)
        # input-file.cpp:170-172 (u8), input-mirisdr.cpp:229-232 (s8), input-soapysdr.cpp:56-66 (s16, f32)
        return {"u8": 126.5, "s8": 126.5, "s16": 32766.5, "f32": 1.0}[self.sample_format]


@dataclass
class MixerInputCfg:
    """One channel output of type "mixer" (config.cpp:173-194): which channel, its ampfactor and balance."""

    device: int
    channel: int
    ampfactor: float = 1.0
    balance: float = 0.0


@dataclass
class MixerCfg:
    """One entry of the ``mixers`` section; inputs in connection order (mixer_connect_input, mixer.cpp:55-93)."""

    name: str = ""
    inputs: List[MixerInputCfg] = field(default_factory=list)

    @property
    def stereo(self) -> bool:
        return any(i.balance != 0.0 for i in self.inputs)


@dataclass
class EngineCfg:
    """The globals ``demodulate()`` reads (boondock_airband.cpp:71-90) plus the device list."""

    fft_size: int = 512
    wave_rate: int = 8000  # 8000 = AM-only build, 16000 = NFM build (boondock_airband.h:67-71)
    fm_demod: int = FM_FAST_ATAN2
    devices: List[DeviceCfg] = field(default_factory=list)
    cuda_device: int = 0
    max_batches_per_step: int = 0
    flags: int = 0
    ring_bytes: int = 0  # 0 = the reference's MIN_BUF_SIZE
    mixers: List[MixerCfg] = field(default_factory=list)

    @property
    def wave_batch(self) -> int:
        return self.wave_rate // 8

    def hop(self, dev: int) -> int:
        """Complex samples between consecutive frames, round(Fs / WAVE_RATE) (boondock_airband.cpp:418)."""
        return int(round(self.devices[dev].sample_rate / self.wave_rate))

    def hop_bytes(self, dev: int) -> int:
        return 2 * self.devices[dev].bytes_per_sample * self.hop(dev)

    def frame_bytes(self, dev: int) -> int:
        return 2 * self.devices[dev].bytes_per_sample * self.fft_size

    def frames_for(self, dev: int, nbytes: int) -> int:
        """Frames the reference's availability test admits for a stream of nbytes (boondock_airband.cpp:419)."""
        need = self.hop_bytes(dev) + self.frame_bytes(dev)
        if nbytes < need:
            return 0
        return (nbytes - need) // self.hop_bytes(dev) + 1

    def batches_for(self, dev: int, nbytes: int) -> int:
        f = self.frames_for(dev, nbytes)
        return max(0, (f - AGC_EXTRA) // self.wave_batch)


def build_desc(cfg: EngineCfg):
    """EngineCfg -> (EngineDesc, keepalive list). The arrays must outlive the C call."""
    keep = []
    devs = (DeviceDesc * len(cfg.devices))()
    for i, d in enumerate(cfg.devices):
        chans = (ChannelDesc * len(d.channels))()
        for j, c in enumerate(d.channels):
            fl = (FreqDesc * max(1, len(c.freqs)))()
            for k, f in enumerate(c.freqs):
                fl[k] = FreqDesc(int(f.freq), MOD[f.modulation], float(f.ampfactor), int(f.squelch_threshold), float(f.squelch_snr_threshold),
                                 float(f.notch), float(f.notch_q), float(f.ctcss), int(f.bandwidth))
            keep.append(fl)
            chans[j] = ChannelDesc(
                int(c.freq), MOD[c.modulation], int(c.afc), float(c.ampfactor), int(c.squelch_threshold),
                float(c.squelch_snr_threshold), float(c.notch), float(c.notch_q), float(c.ctcss), int(c.bandwidth),
                int(c.tau), 1 if c.has_iq_outputs else 0, len(c.freqs), fl if c.freqs else None)
        keep.append(chans)
        devs[i] = DeviceDesc(SFMT[d.sample_format], d.bytes_per_sample, d.default_fullscale(), int(d.sample_rate),
                             int(d.centerfreq), int(d.tau), len(d.channels), chans)
    keep.append(devs)
    mixers = (MixerDesc * max(1, len(cfg.mixers)))()
    for m, mx in enumerate(cfg.mixers):
        ins = (MixerInputDesc * max(1, len(mx.inputs)))()
        for j, i in enumerate(mx.inputs):
            ins[j] = MixerInputDesc(int(i.device), int(i.channel), float(i.ampfactor), float(i.balance))
        keep.append(ins)
        mixers[m] = MixerDesc(len(mx.inputs), ins)
    keep.append(mixers)
    desc = EngineDesc(ABI_VERSION, cfg.fft_size, cfg.wave_rate, cfg.fm_demod, cfg.cuda_device, len(cfg.devices), devs,
                      cfg.max_batches_per_step, cfg.flags, cfg.ring_bytes, len(cfg.mixers), mixers)
    return desc, keep
