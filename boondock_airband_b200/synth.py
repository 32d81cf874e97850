"""Deterministic synthetic IQ for the BASELINE.json workloads.

Per input:  x[n] = sum_c g_c(n) A m_c(n) exp(j 2 pi f_c n / Fs) + w[n]
  w    complex AWGN at -45 dBFS rms
  A    -20 dBFS per carrier (lowered to 0.7/sqrt(C) when many carriers would clip)
  g_c  on `gate_on` s / off `gate_off` s, staggered by 0.1 c seconds
  AM   m = 1 + 0.5 sin(2 pi 1000 t)
  NFM  phase modulation: +-2.5 kHz deviation 1 kHz tone, plus a CTCSS tone at +-500 Hz deviation
       (half of the CTCSS channels carry the configured tone, a quarter 107.2 Hz, a quarter none)
Quantisation: u8 clip(round(127.5 + 127.5 x)), s8 clip(round(128 x)), s16 round(32766.5 x), f32 as is.
Seed: 0xB00D0000 + input index (numpy Philox).  `synth_torch` evaluates the same formulas on a CUDA
device for the benchmark-sized inputs (plumbing only; it is not bit-identical to the numpy path).
"""
from __future__ import annotations

import math
from typing import Optional

import numpy as np

from .abi import DeviceCfg

NOISE_DBFS = -45.0
CARRIER_DBFS = -20.0
SEED_BASE = 0xB00D0000


def _plan(dev: DeviceCfg):
    """Per-carrier parameters: offset Hz, kind (0 AM / 1 NFM), ctcss Hz carried (0 = none)."""
    out = []
    k = 0
    for ch in dev.channels:
        off = float(ch.freq - dev.centerfreq)
        if ch.modulation == "nfm":
            tone = 0.0
            if ch.ctcss > 0:
                sel = k % 4
                tone = ch.ctcss if sel in (0, 1) else (107.2 if sel == 2 else 0.0)
                k += 1
            out.append((off, 1, tone))
        else:
            out.append((off, 0, 0.0))
    return out


def amplitude(n_carriers: int) -> float:
    return min(10.0 ** (CARRIER_DBFS / 20.0), 0.7 / math.sqrt(max(1, n_carriers)))


def quantise(x: np.ndarray, fmt: str) -> np.ndarray:
    """complex64 -> interleaved samples in the input driver's format."""
    iq = np.empty((x.size, 2), np.float32)
    iq[:, 0] = x.real
    iq[:, 1] = x.imag
    if fmt == "u8":
        return np.clip(np.rint(127.5 + 127.5 * iq), 0, 255).astype(np.uint8).reshape(-1)
    if fmt == "s8":
        return np.clip(np.rint(128.0 * iq), -127, 127).astype(np.int8).reshape(-1)
    if fmt == "s16":
        return np.clip(np.rint(32766.5 * iq), -32767, 32767).astype(np.int16).reshape(-1)
    if fmt == "f32":
        return iq.reshape(-1)
    raise ValueError(fmt)


def synth(dev: DeviceCfg, seconds: float, index: int = 0, gate_on: float = 1.5, gate_off: float = 0.5,
          n_samples: Optional[int] = None, chunk: int = 1 << 18, am_depth: float = 0.5, carrier_dbfs=None, burst: float = 1.0) -> np.ndarray:
    """Interleaved IQ (dtype per dev.sample_format) for one input.  `am_depth`, `carrier_dbfs` (one level per
    carrier) and `burst` (amplitude factor over the second half of every transmission) are for the stress scenarios of the parity tests (marginal carriers, AGC clipping)."""
    fs = float(dev.sample_rate)
    n = int(round(seconds * fs)) if n_samples is None else int(n_samples)
    plan = _plan(dev)
    amp = amplitude(len(plan))
    rng = np.random.Generator(np.random.Philox(SEED_BASE + index))
    sigma = 10.0 ** (NOISE_DBFS / 20.0) / math.sqrt(2.0)
    period = gate_on + gate_off
    pieces = []
    for start in range(0, n, chunk):
        m = min(chunk, n - start)
        t = (start + np.arange(m, dtype=np.float64)) / fs
        x = (rng.standard_normal(m, dtype=np.float32) + 1j * rng.standard_normal(m, dtype=np.float32)) * np.float32(sigma)
        x = x.astype(np.complex64)
        am_env = (1.0 + am_depth * np.sin(2 * np.pi * 1000.0 * t)).astype(np.float32)
        fm_voice = 2.5 * np.sin(2 * np.pi * 1000.0 * t)
        for c, (off, kind, tone) in enumerate(plan):
            if carrier_dbfs is not None:
                amp = 10.0 ** (carrier_dbfs[c] / 20.0)
            gate_phase = np.mod(t + 0.1 * c, period)
            gate = (gate_phase < gate_on)
            if not gate.any():
                continue
            if burst != 1.0:
                gate = gate * np.where(gate_phase >= 0.5 * gate_on, np.float32(burst), np.float32(1.0))
            ph = 2 * np.pi * off * t
            if kind == 1:
                ph = ph + fm_voice
                if tone > 0:
                    ph = ph + (500.0 / tone) * np.sin(2 * np.pi * tone * t)
                car = np.exp(1j * ph).astype(np.complex64) * np.float32(amp)
            else:
                car = np.exp(1j * ph).astype(np.complex64) * (am_env * np.float32(amp))
            x += car * gate
        pieces.append(quantise(x, dev.sample_format))
    return np.concatenate(pieces) if pieces else np.empty(0, np.uint8)


def synth_torch(dev: DeviceCfg, n_samples: int, index: int, device, gate_on: float = 1.5, gate_off: float = 0.5,
                chunk: int = 1 << 20, carrier_block: int = 64):
    """Same signal model evaluated with torch on `device`; returns a 1-D tensor of interleaved samples."""
    import torch

    fs = float(dev.sample_rate)
    plan = _plan(dev)
    amp = amplitude(len(plan))
    g = torch.Generator(device=device)
    g.manual_seed(SEED_BASE + index)
    sigma = 10.0 ** (NOISE_DBFS / 20.0) / math.sqrt(2.0)
    period = gate_on + gate_off
    offs = torch.tensor([p[0] for p in plan], dtype=torch.float64, device=device)
    kinds = torch.tensor([p[1] for p in plan], dtype=torch.float32, device=device)
    tones = torch.tensor([p[2] for p in plan], dtype=torch.float64, device=device)
    stag = 0.1 * torch.arange(len(plan), dtype=torch.float64, device=device)
    fmt = dev.sample_format
    tdt = {"u8": torch.uint8, "s8": torch.int8, "s16": torch.int16, "f32": torch.float32}[fmt]
    out = torch.empty(2 * n_samples, dtype=tdt, device=device)
    for start in range(0, n_samples, chunk):
        m = min(chunk, n_samples - start)
        t = (start + torch.arange(m, dtype=torch.float64, device=device)) / fs
        re = torch.randn(m, generator=g, device=device, dtype=torch.float32) * sigma
        im = torch.randn(m, generator=g, device=device, dtype=torch.float32) * sigma
        am_env = (1.0 + 0.5 * torch.sin(2 * math.pi * 1000.0 * t)).float()
        fm_voice = 2.5 * torch.sin(2 * math.pi * 1000.0 * t)
        for c0 in range(0, len(plan), carrier_block):
            c1 = min(len(plan), c0 + carrier_block)
            o = offs[c0:c1, None]
            # phase reduced modulo 1 turn in float64 before the float32 trig
            turns = torch.remainder(o * t[None, :], 1.0)
            isfm = kinds[c0:c1, None]
            tn = tones[c0:c1, None]
            tone_ph = torch.where(tn > 0, (500.0 / torch.clamp(tn, min=1.0)) * torch.sin(2 * math.pi * tn * t[None, :]), torch.zeros_like(turns))
            ph = (2 * math.pi * turns + (fm_voice[None, :] + tone_ph) * isfm.double()).float()
            gate = (torch.remainder(t[None, :] + stag[c0:c1, None], period) < gate_on).float()
            env = (isfm + (1.0 - isfm) * am_env[None, :]) * gate * amp
            re += (torch.cos(ph) * env).sum(0)
            im += (torch.sin(ph) * env).sum(0)
        iq = torch.stack((re, im), dim=1).reshape(-1)
        sl = slice(2 * start, 2 * (start + m))
        if fmt == "u8":
            out[sl] = torch.clamp(torch.round(127.5 + 127.5 * iq), 0, 255).to(torch.uint8)
        elif fmt == "s8":
            out[sl] = torch.clamp(torch.round(128.0 * iq), -127, 127).to(torch.int8)
        elif fmt == "s16":
            out[sl] = torch.clamp(torch.round(32766.5 * iq), -32767, 32767).to(torch.int16)
        else:
            out[sl] = iq
    return out
