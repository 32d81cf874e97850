"""The configurations behind the committed golden fixtures (tests/golden/*.npz).  Sample rates are kept low so that the
input IQ itself fits in the fixture; the code path is the same as for the BASELINE workloads."""
from __future__ import annotations

import numpy as np

from boondock_airband_b200 import abi, synth
from boondock_airband_b200.abi import ChannelCfg, DeviceCfg, EngineCfg

NAMES = ("golden_am_u8", "golden_nfm_s16")


def build(name: str, with_iq: bool = True):
    if name == "golden_am_u8":
        fs, cf = 256_000, 120_000_000
        ch = [ChannelCfg(freq=cf - 90_000), ChannelCfg(freq=cf - 31_000, bandwidth=5000), ChannelCfg(freq=cf + 27_500, squelch_threshold=-50, ampfactor=1.7),
              ChannelCfg(freq=cf + 77_000, squelch_snr_threshold=6.0, has_iq_outputs=True), ChannelCfg(freq=cf + 110_000, afc=6)]
        dev = DeviceCfg(sample_rate=fs, centerfreq=cf, sample_format="u8", channels=ch)
        cfg = EngineCfg(fft_size=256, wave_rate=8000, devices=[dev], flags=abi.FLAG_TRACE, max_batches_per_step=2)
        iq = synth.synth(dev, 0.66, 101, gate_on=0.22, gate_off=0.09) if with_iq else None
        return cfg, iq
    if name == "golden_nfm_s16":
        fs, cf = 240_000, 162_500_000
        ch = [ChannelCfg(freq=cf - 75_000, modulation="nfm", bandwidth=12_500, ctcss=100.0, notch=100.0),
              ChannelCfg(freq=cf - 25_000, modulation="nfm", bandwidth=12_500, ctcss=100.0, notch=100.0),
              ChannelCfg(freq=cf + 25_000, modulation="nfm", bandwidth=12_500, ctcss=100.0, notch=100.0),
              ChannelCfg(freq=cf + 75_000, modulation="nfm", tau=75, has_iq_outputs=True)]
        dev = DeviceCfg(sample_rate=fs, centerfreq=cf, sample_format="s16", channels=ch)
        cfg = EngineCfg(fft_size=256, wave_rate=16000, devices=[dev], flags=abi.FLAG_TRACE, max_batches_per_step=2)
        iq = synth.synth(dev, 0.78, 102, gate_on=0.62, gate_off=0.1) if with_iq else None
        return cfg, iq
    raise KeyError(name)
