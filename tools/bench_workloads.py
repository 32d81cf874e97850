#!/usr/bin/env python
"""Device-resident throughput of the BASELINE.json workloads other than the headline one (bench.py measures cfg5).

    python tools/bench_workloads.py [--only cfg2,cfg4] [--steps 4] [--warmup 2] > gpurun_out/workloads.jsonl

One JSON line per workload: aggregate Msps, x real-time, and the per-step device time of the channelizer (K1) and the
demodulator (K2).  A step is one second of signal for every input of the workload (8 WAVE_BATCH batches); IQ is
synthesised on the GPU (boondock_airband_b200.synth.synth_torch) and attached as a device-resident stream, results
return to pinned host memory inside the timed region.  Parity of these workloads is covered by tests/ at reduced size;
this script only measures.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def workloads():
    from boondock_airband_b200 import configs
    return {
        "cfg1": ("basic_multichannel AM: 1 input, 2.56 Msps u8, fft 512, 8 AM channels, R 8000", configs.cfg1),
        "cfg2": ("NFM + CTCSS + de-emphasis: 1 input, 2.4 Msps cs16, fft 1024, 32 channels (low-pass, notch, CTCSS)", configs.cfg2),
        "cfg2x64": ("64 inputs of cfg2's shape on one GPU (2048 NFM/CTCSS channels)", lambda: _replicate(configs.cfg2(), 64)),
        "cfg3": ("8 of the 64 synthetic dongles (one GPU's share): 2.4 Msps u8, fft 512, 16 AM channels each", lambda: configs.cfg3(8)),
        "cfg4": ("wideband: 1 input, 61.44 Msps cf32, fft 8192, 2000 mixed AM/NFM channels (250 with CTCSS + notch)", configs.cfg4),
        "cfg4_k1": ("cfg4's input with only 64 of its channels: the wideband channelizer (fft 8192, cf32) on its own", _cfg4_few),
        "cfg3_mixers": ("cfg3's 8 inputs with 17 mixers summed on the GPU (K3): mixer k = channel k of every input, plus one stereo mixer over input 0", _cfg3_mixers),
        "cfg5_n1024": ("cfg5 at fft 1024: 512 inputs x 2.56 Msps u8, 16 AM channels", lambda: configs.cfg5(512, 1024)),
        "cfg5_n2048": ("cfg5 at fft 2048", lambda: configs.cfg5(512, 2048)),
        "cfg5_n4096": ("cfg5 at fft 4096", lambda: configs.cfg5(512, 4096)),
    }


def _cfg4_few():
    from boondock_airband_b200 import configs
    cfg = configs.cfg4()
    cfg.devices[0].channels = cfg.devices[0].channels[::32][:64]
    return cfg


def _cfg3_mixers():
    from boondock_airband_b200 import configs
    from boondock_airband_b200.abi import MixerCfg, MixerInputCfg
    cfg = configs.cfg3(8)
    for k in range(16):
        cfg.mixers.append(MixerCfg("m%d" % k, [MixerInputCfg(d, k, ampfactor=0.5) for d in range(8)]))
    cfg.mixers.append(MixerCfg("stereo", [MixerInputCfg(0, k, balance=(k - 7.5) / 8.0) for k in range(16)]))
    return cfg


def run_file_replay(seconds, ring_bytes):
    """Row f-1 measured: cfg1's IQ replayed from a file (tmpfs) by the file input's reader thread through the pinned input
    ring, unpaced, with the demodulator loop on this thread - the path an unmodified input driver takes."""
    import tempfile

    import numpy as np

    from boondock_airband_b200 import configs, host, synth
    from boondock_airband_b200.engine import Engine

    cfg = configs.cfg1()
    cfg.flags = 0
    cfg.max_batches_per_step = 8
    cfg.ring_bytes = ring_bytes
    one = synth.synth(cfg.devices[0], 2.0, 0)
    reps = int(seconds / 2.0)
    d = "/dev/shm" if os.path.isdir("/dev/shm") else None
    with tempfile.NamedTemporaryFile(dir=d, suffix=".cu8") as f:
        for _ in range(reps):
            f.write(one.tobytes())
        f.flush()
        nbytes = reps * one.nbytes
        eng = Engine(cfg)
        inp = host.FileInput(eng, 0, f.name, sample_format="u8", sample_rate=cfg.devices[0].sample_rate, speedup_factor=0.0)
        t0 = time.perf_counter()
        inp.start()
        batches = passes = 0
        while True:
            ended = inp.state in (host.INPUT_FAILED, host.INPUT_STOPPED)
            t = eng.process()
            r = eng.collect_raw(t, 0)
            passes += 1
            batches += r.n_batches
            if ended and r.n_batches == 0:
                break
        wall = time.perf_counter() - t0
        launches = eng.launch_count()
        inp.stop()
        eng.close()
    samples = nbytes / 2
    fs = cfg.devices[0].sample_rate
    return {"workload": "cfg1_file_ring%d" % ring_bytes, "what": "cfg1 replayed from a tmpfs file through the file input's reader thread and the pinned input ring (ring base %d bytes), unpaced" % (ring_bytes or 2560000),
            "inputs": 1, "channels": 8, "fft_size": 512, "wave_rate": 8000, "signal_seconds": samples / fs, "msps": samples / wall / 1e6, "x_realtime": samples / fs / wall,
            "process_calls": passes, "batches": batches, "launches_total": launches, "data": "synthetic IQ file; host ring -> H2D -> kernels -> D2H inside the timed region"}


def _replicate(cfg, n):
    import copy
    d0 = cfg.devices[0]
    cfg.devices = [copy.deepcopy(d0) for _ in range(n)]
    return cfg


def run(name, text, make, steps, warmup, device):
    import torch

    from boondock_airband_b200 import synth
    from boondock_airband_b200.engine import Engine

    cfg = make()
    cfg.flags = 0
    cfg.max_batches_per_step = 8
    R = cfg.wave_rate
    B = R // 8
    secs = 8 * B / R  # one step
    total = steps + warmup
    fmt_bytes = {"u8": 1, "s8": 1, "s16": 2, "f32": 4}
    streams, step_bytes, lead_bytes = [], [], []
    templates = {}
    for i, dev in enumerate(cfg.devices):
        hop = int(round(dev.sample_rate / R))
        bps = 2 * fmt_bytes[dev.sample_format] * hop
        sb = bps * 8 * B
        lead = bps * 128 + 2 * fmt_bytes[dev.sample_format] * cfg.fft_size
        n_samples = (sb * total + lead) // (2 * fmt_bytes[dev.sample_format]) + 16
        key = (dev.sample_rate, dev.sample_format, len(dev.channels), i % 8)
        if key not in templates:
            templates[key] = synth.synth_torch(dev, n_samples, i % 8, device)
        t = templates[key].clone() if len(cfg.devices) > 1 else templates[key]
        streams.append(t)
        step_bytes.append(sb)
        lead_bytes.append(lead)
    eng = Engine(cfg)
    for i, t in enumerate(streams):
        eng.attach_device_stream(i, t.data_ptr(), t.numel() * t.element_size())
    torch.cuda.synchronize()
    nd = len(cfg.devices)
    depth = 3

    def go(n, first):
        k1 = k2 = 0.0
        nb = 0
        pend = []

        def fin(old):
            nonlocal k1, k2, nb
            r = eng.collect_raw(old, 0)
            a, b = eng.kernel_ms(old)
            k1 += a
            k2 += b
            nb += r.n_batches

        for s in range(n):
            for i in range(nd):
                eng.advance_device_stream(i, step_bytes[i] + (lead_bytes[i] if first and s == 0 else 0))
            pend.append(eng.process())
            if len(pend) >= depth:
                fin(pend.pop(0))
        while pend:
            fin(pend.pop(0))
        return k1, k2, nb

    go(warmup, True)
    torch.cuda.synchronize()
    eng.mark(0)
    t0 = time.perf_counter()
    k1, k2, nb = go(steps, False)
    eng.mark(1)
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    dev_ms = eng.mark_ms(0, 1)
    assert nb == steps * 8, (name, nb)
    launches = eng.launch_count()
    eng.close()
    elapsed = max(wall, dev_ms / 1e3)
    samples = sum(d.sample_rate for d in cfg.devices) * secs * steps
    n_ch = sum(len(d.channels) for d in cfg.devices)
    return {"workload": name, "what": text, "inputs": nd, "channels": n_ch, "fft_size": cfg.fft_size, "wave_rate": R, "steps": steps, "warmup": warmup,
            "msps": samples / elapsed / 1e6, "x_realtime": secs * steps / elapsed, "x_realtime_aggregate": nd * secs * steps / elapsed,
            "ms_per_step": 1e3 * elapsed / steps, "channelize_ms_per_step": k1 / steps, "demod_ms_per_step": k2 / steps, "launches_total": launches,
            "data": "synthetic, device-resident; audio returns to pinned host memory"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="")
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--warmup", type=int, default=2)
    a = ap.parse_args()
    import torch
    device = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    wl = workloads()
    names = [n for n in a.only.split(",") if n] or (list(wl) + ["cfg1_file"])
    for n in names:
        if n.startswith("cfg1_file"):
            for ring in (0, 64 << 20):
                print(json.dumps(run_file_replay(40.0, ring)), flush=True)
            continue
        text, make = wl[n]
        line = run(n, text, make, a.steps, a.warmup, device)
        print(json.dumps(line), flush=True)
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
