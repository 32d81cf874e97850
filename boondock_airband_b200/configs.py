"""The five BASELINE.json workloads as channel models.

Frequencies and options are the ones a libconfig file for the reference would carry
(config/basic_multichannel.conf, config/noaa.conf for the two that exist upstream).
"""
from __future__ import annotations

from .abi import ChannelCfg, DeviceCfg, EngineCfg


def cfg1() -> EngineCfg:
    """basic_multichannel.conf-style AM: 1 input, 2.56 Msps u8, fft_size 512, 8 AM channels, WAVE_RATE 8000."""
    freqs = [119_000_000, 119_300_000, 119_500_000, 119_800_000, 120_225_000, 120_500_000, 120_800_000, 121_000_000]
    dev = DeviceCfg(sample_rate=2_560_000, centerfreq=120_000_000, sample_format="u8",
                    channels=[ChannelCfg(freq=f) for f in freqs])
    return EngineCfg(fft_size=512, wave_rate=8000, devices=[dev])


def cfg2(n_channels: int = 32) -> EngineCfg:
    """NFM with CTCSS + de-emphasis: 1 input, 2.4 Msps cs16, fft_size 1024, 32 channels, squelch + notch + low-pass on."""
    chans = []
    for i in range(n_channels):
        f = 162_087_500 + 25_000 * i
        chans.append(ChannelCfg(freq=f, modulation="nfm", bandwidth=12_500, ctcss=100.0, notch=100.0))
    dev = DeviceCfg(sample_rate=2_400_000, centerfreq=162_482_000, sample_format="s16", channels=chans)
    return EngineCfg(fft_size=1024, wave_rate=16000, devices=[dev])


def _am16(center: int, index: int):
    offs = []
    for k in range(1, 9):
        offs += [-75_000 * k, 75_000 * k]
    shift = 8_330 * (index % 8)
    return [ChannelCfg(freq=center + o + shift) for o in sorted(offs)]


def cfg3(n_inputs: int = 64, first_index: int = 0) -> EngineCfg:
    """Synthetic dongles: 2.4 Msps u8, fft_size 512, 16 AM channels each (8 inputs per GPU on 8 GPUs)."""
    devs = []
    for i in range(first_index, first_index + n_inputs):
        center = 118_000_000 + 2_000_000 * (i % 10)
        devs.append(DeviceCfg(sample_rate=2_400_000, centerfreq=center, sample_format="u8", channels=_am16(center, i)))
    return EngineCfg(fft_size=512, wave_rate=8000, devices=devs)


def cfg4(n_channels: int = 2000, sample_rate: int = 61_440_000) -> EngineCfg:
    """Wideband SoapySDR-style stream: 61.44 Msps cf32, fft_size 8192, 2000 mixed AM/NFM channels on a 25 kHz raster."""
    center = 140_000_000
    chans = []
    first = center - 25_000 * (n_channels // 2)
    nfm_seen = 0
    for i in range(n_channels):
        f = first + 25_000 * i
        if i % 2 == 0:
            chans.append(ChannelCfg(freq=f, modulation="am"))
        else:
            if nfm_seen % 4 == 0:
                chans.append(ChannelCfg(freq=f, modulation="nfm", bandwidth=12_500, ctcss=100.0, notch=100.0))
            else:
                chans.append(ChannelCfg(freq=f, modulation="nfm", bandwidth=12_500))
            nfm_seen += 1
    dev = DeviceCfg(sample_rate=sample_rate, centerfreq=center, sample_format="f32", channels=chans)
    return EngineCfg(fft_size=8192, wave_rate=16000, devices=[dev])


def cfg5(n_inputs: int = 512, fft_size: int = 512, first_index: int = 0) -> EngineCfg:
    """Box-scale sweep: synthetic inputs of 2.56 Msps u8, 16 AM channels each, fft_size 512..4096."""
    devs = []
    for i in range(first_index, first_index + n_inputs):
        center = 118_000_000 + 2_000_000 * (i % 10)
        devs.append(DeviceCfg(sample_rate=2_560_000, centerfreq=center, sample_format="u8", channels=_am16(center, i)))
    return EngineCfg(fft_size=fft_size, wave_rate=8000, devices=devs)


WORKLOADS = {"cfg1": cfg1, "cfg2": cfg2, "cfg3": cfg3, "cfg4": cfg4, "cfg5": cfg5}
