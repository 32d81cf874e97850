"""Parity scenarios: small versions of the BASELINE.json workloads plus option coverage the reference has no test for."""
from __future__ import annotations

import numpy as np

from boondock_airband_b200 import abi, configs, synth
from boondock_airband_b200.abi import ChannelCfg, DeviceCfg, EngineCfg


def cfg1_short(seconds=0.9):
    cfg = configs.cfg1()
    cfg.flags = abi.FLAG_TRACE
    cfg.max_batches_per_step = 3
    iq = synth.synth(cfg.devices[0], seconds, 0, gate_on=0.25, gate_off=0.1)
    return cfg, [iq]


def am_stress(seconds=1.5, index=11):
    """Plain AM channels driven hard: carriers from strong to below the squelch threshold (decisions flip on noise),
    a threefold amplitude step halfway through each transmission (the AGC clip branch, .cpp:583-586) and 130 ms transmissions 40 ms apart (flap detection,
    low-signal aborts, every delay running out).  For the bit-exact checks of the demodulator."""
    cfg = configs.cfg1()
    cfg.flags = abi.FLAG_TRACE
    cfg.max_batches_per_step = 3
    levels = [-25.0, -35.0, -44.0, -46.0, -47.0, -48.0, -49.0, -50.0]
    iq = synth.synth(cfg.devices[0], seconds, index, gate_on=0.13, gate_off=0.04, am_depth=0.9, carrier_dbfs=levels, burst=3.0)
    return cfg, [iq]


def cfg2_small(n_channels=8, seconds=1.3):
    cfg = configs.cfg2(n_channels)
    cfg.flags = abi.FLAG_TRACE
    cfg.max_batches_per_step = 2
    iq = synth.synth(cfg.devices[0], seconds, 0, gate_on=0.8, gate_off=0.15)
    return cfg, [iq]


def mixed_options(seconds=0.8, fmt="s16", fft_size=1024, fm_demod=abi.FM_FAST_ATAN2, afc=True):
    """One input carrying every per-channel option of parse_channels (config.cpp:312-729)."""
    fs, cf = 2_400_000, 145_000_000
    ch = [
        ChannelCfg(freq=cf - 600_000),                                              # plain AM
        ChannelCfg(freq=cf - 450_000, bandwidth=8000),                              # AM through the low-pass (needs_raw_iq)
        ChannelCfg(freq=cf - 300_000, modulation="nfm"),                            # NFM, no filters
        ChannelCfg(freq=cf - 150_000, modulation="nfm", bandwidth=12_500, tau=0),   # NFM, no de-emphasis
        ChannelCfg(freq=cf + 150_000, modulation="nfm", bandwidth=12_500, ctcss=100.0, notch=100.0, tau=530),
        ChannelCfg(freq=cf + 300_000, squelch_threshold=-45, ampfactor=2.5),        # manual squelch level, amplified
        ChannelCfg(freq=cf + 450_000, squelch_snr_threshold=6.0, has_iq_outputs=True),
        ChannelCfg(freq=cf + 600_000, modulation="nfm", notch=1000.0, notch_q=5.0, has_iq_outputs=True),
        ChannelCfg(freq=cf + 750_000, afc=8 if afc else 0),                                       # AFC walk
        ChannelCfg(freq=cf + 900_000 + 4_000, afc=20 if afc else 0, modulation="nfm", bandwidth=12_500),
    ]
    dev = DeviceCfg(sample_rate=fs, centerfreq=cf, sample_format=fmt, channels=ch, tau=300)
    cfg = EngineCfg(fft_size=fft_size, wave_rate=16000, fm_demod=fm_demod, devices=[dev], flags=abi.FLAG_TRACE, max_batches_per_step=2)
    iq = synth.synth(dev, seconds, 7, gate_on=0.3, gate_off=0.12)
    return cfg, [iq]


def afc_walk(seconds=1.6):
    """AFC (boondock_airband.cpp:180-251) with carriers that do NOT sit in the channel's bin: the transmitters are placed by
    one configuration, the receiver listens with another whose channels are tuned a few bins off.  Channel 0 is tuned two
    bins below its carrier (the walk goes up: AFC_UP), channel 1 two bins above (AFC_DOWN), channel 2 one bin below with a
    coarse afc step, channel 3 is on its carrier (no move), channel 4 is NFM through the low-pass, tuned one bin above.  The carriers gate on and off, so every open->closed edge
    restores base_bins (.cpp:246-249) and every closed->open edge walks again."""
    fs, cf, n = 2_560_000, 120_000_000, 512
    bw = fs // n
    tx_freqs = [cf - 800_000, cf - 400_000, cf + 100_000, cf + 500_000, cf + 900_000]
    tx = DeviceCfg(sample_rate=fs, centerfreq=cf, sample_format="u8",
                   channels=[ChannelCfg(freq=f) for f in tx_freqs[:4]] + [ChannelCfg(freq=tx_freqs[4], modulation="nfm")])
    rx_ch = [ChannelCfg(freq=tx_freqs[0] - 2 * bw, afc=10), ChannelCfg(freq=tx_freqs[1] + 2 * bw, afc=10), ChannelCfg(freq=tx_freqs[2] - bw, afc=2),
             ChannelCfg(freq=tx_freqs[3], afc=10), ChannelCfg(freq=tx_freqs[4] + bw, afc=10, modulation="nfm", bandwidth=12_500)]
    rx = DeviceCfg(sample_rate=fs, centerfreq=cf, sample_format="u8", channels=rx_ch)
    cfg = EngineCfg(fft_size=n, wave_rate=16000, devices=[rx], flags=abi.FLAG_TRACE, max_batches_per_step=3)
    iq = synth.synth(tx, seconds, 77, gate_on=0.33, gate_off=0.2)
    return cfg, [iq]


def held_by_post_filter(seconds=1.2, fm_demod=abi.FM_FAST_ATAN2):
    """NFM channels through the low-pass, tuned a little off their carriers (still inside the carrier's bin): the raw magnitude
    says "signal", the derotated carrier sits at 0.6-2 kHz where the narrow low-pass takes more than a tenth of it away, and the
    filtered average never reaches 0.9 x the raw one (Squelch::buffer_[tail]).  The reference then either keeps the squelch
    CLOSED for the whole transmission while filtering every sample (the open request is taken back within the sample,
    squelch.cpp:223-232 with :268-274), or flaps CLOSED -> OPENING -> CLOSED, or opens late - one channel of each kind, plus one
    on its carrier.  Transmitters by one configuration, receiver by another (as afc_walk)."""
    fs, cf = 2_400_000, 145_000_000
    ks = range(-3, 4)
    offs = [0, 600, 900, 1200, 900, 1500, 2000]
    bws = [12_500, 4000, 4000, 4000, 2000, 6000, 6000]
    tx = DeviceCfg(sample_rate=fs, centerfreq=cf, sample_format="s16", channels=[ChannelCfg(freq=cf + 100_000 * k, modulation="nfm") for k in ks])
    rx_ch = [ChannelCfg(freq=cf + 100_000 * k + offs[i], modulation="nfm", bandwidth=bws[i]) for i, k in enumerate(ks)]
    rx_ch[2].ctcss = 100.0
    rx_ch[3].notch = 100.0
    rx = DeviceCfg(sample_rate=fs, centerfreq=cf, sample_format="s16", channels=rx_ch)
    cfg = EngineCfg(fft_size=1024, wave_rate=16000, fm_demod=fm_demod, devices=[rx], flags=abi.FLAG_TRACE, max_batches_per_step=2)
    iq = synth.synth(tx, seconds, 5, gate_on=0.45, gate_off=0.15)
    return cfg, [iq]


def multi_device(seconds=0.5):
    """Three inputs of different sample formats and rates behind one engine (device_start..device_end of one demod thread)."""
    devs = []
    streams = []
    specs = (("u8", 2_560_000, 120_000_000), ("s8", 2_048_000, 131_000_000), ("f32", 1_920_000, 460_000_000))
    for i, (fmt, fs, cf) in enumerate(specs):
        ch = [ChannelCfg(freq=cf + k * 100_000 + 12_500 * i) for k in (-3, -1, 2)]
        ch.append(ChannelCfg(freq=cf + 350_000, bandwidth=6000))
        d = DeviceCfg(sample_rate=fs, centerfreq=cf, sample_format=fmt, channels=ch)
        devs.append(d)
        streams.append(synth.synth(d, seconds * (1.0 + 0.3 * i), 20 + i, gate_on=0.2, gate_off=0.07))
    cfg = EngineCfg(fft_size=512, wave_rate=8000, devices=devs, flags=abi.FLAG_TRACE, max_batches_per_step=2)
    return cfg, streams


def mixers_on_multi_device(seconds=0.5):
    """multi_device() plus three mixers (row f-4): a mono one across all inputs, a stereo one with balances (one input
    panned fully right, so its left factor is zero), a single-input one.  The inputs run at different rates and lengths,
    so their batches arrive unevenly and the mixer FIFO is exercised."""
    from boondock_airband_b200.abi import MixerCfg, MixerInputCfg
    cfg, streams = multi_device(seconds)
    cfg.mixers = [
        MixerCfg("mono", [MixerInputCfg(0, 0), MixerInputCfg(1, 1, ampfactor=0.5), MixerInputCfg(2, 2, ampfactor=2.0)]),
        MixerCfg("stereo", [MixerInputCfg(0, 1, balance=-0.6), MixerInputCfg(1, 0, ampfactor=0.7, balance=0.3), MixerInputCfg(2, 3, balance=1.0),
                            MixerInputCfg(0, 0, ampfactor=0.0)]),
        MixerCfg("single", [MixerInputCfg(0, 3, ampfactor=1.25)]),
    ]
    return cfg, streams


def scan_mode(seconds=1.2):
    """Row f-3: a scan-mode input (one channel, three frequencies with different freq_t settings: plain AM, NFM through
    the low-pass with CTCSS and notch, AM with a manual squelch level and gain) next to an ordinary multichannel input.
    The file does not retune, so every frequency sees the carrier that sits in the channel's bin; what the test
    exercises is the freq_t switch: each frequency keeps its own Squelch, filters, AGC level and counters."""
    from boondock_airband_b200.abi import FreqCfg
    fs, n = 2_560_000, 512
    f0 = 121_500_000
    cf = f0 + 20 * (fs // n)  # config.cpp:431: tuned 20 bins above the first frequency
    scan = ChannelCfg(freq=f0, modulation="nfm", tau=100, freqs=[
        FreqCfg(f0),
        FreqCfg(f0 + 25_000, modulation="nfm", bandwidth=12_500, ctcss=100.0, notch=100.0, ampfactor=0.8),
        FreqCfg(f0 + 50_000, squelch_threshold=-48, ampfactor=1.6, notch=1000.0, notch_q=4.0)])
    d0 = DeviceCfg(sample_rate=fs, centerfreq=cf, sample_format="u8", channels=[scan])
    d1 = DeviceCfg(sample_rate=2_400_000, centerfreq=145_000_000, sample_format="s16",
                   channels=[ChannelCfg(freq=145_000_000 - 300_000), ChannelCfg(freq=145_000_000 + 250_000, modulation="nfm", bandwidth=12_500)])
    cfg = EngineCfg(fft_size=n, wave_rate=16000, devices=[d0, d1], flags=abi.FLAG_TRACE, max_batches_per_step=1)
    streams = [synth.synth(d0, seconds, 41, gate_on=0.22, gate_off=0.09), synth.synth(d1, seconds, 42, gate_on=0.3, gate_off=0.1)]
    return cfg, streams


def squelch_steps(levels, fft_size=512):
    """Picked-bin series with constant magnitudes, the stimulus of test_squelch.cpp (0.05 = noise, 0.75 = signal)."""
    z = np.zeros((len(levels), 1, 2), np.float32)
    z[:, 0, 0] = levels
    return z


FILE_REPLAY_CONF = """
# two file inputs in one configuration file, as an unmodified airband .conf would name them (rows f-1 / f-2)
fft_size = 512;
multiple_demod_threads = false;
mixers: { both: { %(o)s } };
devices: (
  { type = "file"; filepath = "%(a)s"; sample_rate = 2.4; centerfreq = 145.0; sample_format = "S16"; speedup_factor = 100;
    channels: (
      { freq = 144.7; modulation = "nfm"; bandwidth = 12500; ctcss = 100.0; notch = 100.0; %(o)s },
      { freq = 144.85; outputs: ( { type = "mixer"; name = "both"; balance = -0.4; } ); },
      { freq = 145.15; disable = true; %(o)s },
      { freq = 145.3; modulation = "nfm"; outputs: ( { type = "rawfile"; directory = "/tmp"; filename_template = "iq"; } ); },
      { freq = 145.45; bandwidth = "8k"; ampfactor = 1.5; squelch_snr_threshold = 6; %(o)s }
    ); },
  { type = "file"; filepath = "%(b)s"; sample_rate = 2560000; centerfreq = 120000000;
    channels: ( { freq = 119.5; outputs: ( { type = "mixer"; name = "both"; ampfactor = 0.8; balance = 0.4; } ); }, { freq = 120.225; afc = 6; %(o)s }, { freq = 120800000; squelch_threshold = -42; %(o)s } ); }
);
"""


def file_replay(tmp_path, seconds=0.6):
    """Writes the IQ files and returns (configuration text, parsed EngineCfg, the streams as written)."""
    from boondock_airband_b200 import host
    out = 'outputs: ( { type = "file"; directory = "/tmp"; filename_template = "x"; } );'
    a, b = str(tmp_path / "dev0.cs16"), str(tmp_path / "dev1.cu8")
    text = FILE_REPLAY_CONF % {"a": a, "b": b, "o": out}
    conf = host.parse_text(text)
    cfg = conf.cfg
    cfg.flags = abi.FLAG_TRACE
    cfg.max_batches_per_step = 2
    streams = [synth.synth(cfg.devices[0], seconds, 31, gate_on=0.25, gate_off=0.1), synth.synth(cfg.devices[1], seconds * 1.2, 32, gate_on=0.2, gate_off=0.08)]
    for path, s in zip((a, b), streams):
        s.tofile(path)
    return conf, cfg, streams
