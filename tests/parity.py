"""Shared parity checks: the same assertions run against the product library on a B200 (-m gpu) and, for the host
logic and kernel index arithmetic, against the thread-emulated build of the same sources on the CPU (tests/emu)."""
from __future__ import annotations

import os
import subprocess

import copy

import numpy as np

from boondock_airband_b200 import abi
from boondock_airband_b200.engine import Engine, LIB_PATH
from oracle.ba_oracle import Oracle

HERE = os.path.dirname(os.path.abspath(__file__))
EMU_DIR = os.path.join(HERE, "emu")
EMU_LIB = os.path.join(EMU_DIR, "libba_emu_TESTONLY.so")

# stated tolerances (BASELINE.json north_star)
TOL_BASEBAND = 1e-4  # relative, picked-bin IQ
TOL_AUDIO = 1e-3     # demodulated audio samples (full scale is 1.0)


def build_emu():
    subprocess.check_call(["make", "-C", EMU_DIR], stdout=subprocess.DEVNULL)
    return EMU_LIB


def lib_for(kind: str) -> str:
    if kind == "emu":
        return build_emu()
    assert kind == "cuda"
    return LIB_PATH


def random_iq(rng, fmt: str, n_complex: int) -> np.ndarray:
    if fmt == "u8":
        return rng.integers(0, 256, 2 * n_complex, dtype=np.uint8)
    if fmt == "s8":
        return rng.integers(-127, 128, 2 * n_complex, dtype=np.int8)
    if fmt == "s16":
        return rng.integers(-32768, 32768, 2 * n_complex, dtype=np.int16)
    return (rng.standard_normal(2 * n_complex) * 0.3).astype(np.float32)


def check_frames(cfg: abi.EngineCfg, iq: np.ndarray, n_frames: int, lib: str):
    """Sample conversion x window must be bit-exact; spectra within TOL_BASEBAND of the float64 DFT."""
    o = Oracle(cfg)
    e = Engine(cfg, lib)
    try:
        assert np.array_equal(o.window(), e.window())
        fi_o, fo_o = o.debug_frames(0, iq, n_frames)
        fi_e, fo_e = e.debug_frames(0, iq, n_frames)
    finally:
        e.close()
    assert np.array_equal(fi_o.view(np.uint32), fi_e.view(np.uint32)), "converted+windowed frames differ"
    ref = np.fft.fft(fi_o[..., 0].astype(np.float64) + 1j * fi_o[..., 1].astype(np.float64), axis=1)
    got = fo_e[..., 0].astype(np.float64) + 1j * fo_e[..., 1].astype(np.float64)
    orc = fo_o[..., 0].astype(np.float64) + 1j * fo_o[..., 1].astype(np.float64)
    scale = np.abs(ref).max(axis=1, keepdims=True)
    err_ref = float((np.abs(got - ref) / scale).max())
    err_orc = float((np.abs(got - orc) / scale).max())
    assert err_ref < TOL_BASEBAND and err_orc < TOL_BASEBAND, (err_ref, err_orc)
    return err_ref, err_orc


def check_channel_info(cfg: abi.EngineCfg, lib: str, ref: bool = False):
    """Bins, dm_dphi and every derived filter/squelch/CTCSS constant must equal the oracle's bit for bit."""
    o = Oracle(cfg, ref=ref)
    e = Engine(cfg, lib)
    try:
        for d, dev in enumerate(cfg.devices):
            for c in range(len(dev.channels)):
                a, b = o.channel_info(d, c), e.channel_info(d, c)
                assert bytes(a) == bytes(b), (d, c, a.as_dict(), b.as_dict())
    finally:
        e.close()


def run_both(cfg: abi.EngineCfg, streams, lib: str, chunk_bytes: int = 1 << 20, ref: bool = False):
    o = Oracle(cfg, ref=ref)
    for d, s in enumerate(streams):
        o.feed(d, s)
    e = Engine(cfg, lib)
    try:
        res = e.run_stream(streams, chunk_bytes=chunk_bytes)
        picks = None
        launches = e.launch_count()
    finally:
        e.close()
    return o, res, launches


def run_scan(cfg: abi.EngineCfg, streams, lib: str, every: int, order, chunk_bytes: int = 200_000):
    """Engine: after every `every` batches of device 0 its scan channel moves on through `order` (the controller thread's
    freq_idx = i, boondock_airband.cpp:115-118).  The oracle gets the same switches at the same batch numbers."""
    e = Engine(cfg, lib)
    plan = []
    try:
        nd = len(cfg.devices)
        views = [np.ascontiguousarray(s).view(np.uint8).reshape(-1) for s in streams]
        pos = [0] * nd
        acc = e._new_acc()
        done0, step = 0, 0
        while True:
            fed = False
            for d in range(nd):
                n = min(chunk_bytes, views[d].size - pos[d], e.input_space(d))
                if n > 0:
                    e.submit(d, views[d][pos[d]:pos[d] + n])
                    pos[d] += n
                    fed = True
            produced, advanced = e._step_into(acc)
            n0 = sum(w.shape[1] for w in acc[0]["waveout"]) // cfg.wave_batch
            if n0 // every > done0 // every:
                step += 1
                idx = order[step % len(order)]
                plan.append((e.set_freq_idx(0, 0, idx), idx))
            done0 = n0
            if not fed and not produced and not advanced:
                break
        res = e._finish_acc(acc)
        launches = e.launch_count()
    finally:
        e.close()
    o = Oracle(cfg)
    for b, idx in plan:
        o.set_freq_idx(0, 0, b, idx)
    for d, s in enumerate(streams):
        o.feed(d, s)
    return o, res, plan, launches


def check_mixers(cfg: abi.EngineCfg, streams, lib: str, chunk_bytes: int, masked=None):
    """K3 against the restated mixer: bit-exact on the engine's own channel audio (the mixer's arithmetic alone), and within
    the audio tolerance on the oracle's (the whole path)."""
    from oracle.ba_oracle import mix_reference
    masked = masked or {}
    o = Oracle(cfg)
    for d, s in enumerate(streams):
        o.feed(d, s)
    e = Engine(cfg, lib)
    try:
        for m, js in masked.items():
            for j in js:
                e.mixer_input_mask(m, j, False)
        res = e.run_stream(streams, chunk_bytes=chunk_bytes)
        mixed = e.mixer_results()
        launches = e.launch_count()
    finally:
        e.close()

    class S:  # the engine's per-batch status rows as objects with .axcindicate
        def __init__(self, v):
            self.axcindicate = v

    for m, mx in enumerate(cfg.mixers):
        own = mix_reference(cfg, lambda d, c: res[d]["waveout"][c], lambda d, c: [S(row[c]["axcindicate"]) for row in res[d]["status"]], mx, masked.get(m, ()))
        ref = mix_reference(cfg, o.waveout, o.status, mx, masked.get(m, ()))
        got = mixed[m]
        n = len(got["left"])
        assert n > 0 and n == len(own[0]) == len(ref[0]), (m, n, len(own[0]), len(ref[0]))
        assert np.array_equal(got["left"].view(np.uint32), own[0].view(np.uint32)), "mixer %d: left plane not bit-exact" % m
        assert np.array_equal(got["axcindicate"], own[2]) and np.array_equal(got["axcindicate"], ref[2])
        gain = sum(abs(i.ampfactor) for i in mx.inputs)
        assert float(np.abs(got["left"] - ref[0]).max()) <= TOL_AUDIO * max(1.0, gain)
        if mx.stereo:
            assert np.array_equal(got["right"].view(np.uint32), own[1].view(np.uint32)), "mixer %d: right plane not bit-exact" % m
            assert float(np.abs(got["right"] - ref[1]).max()) <= TOL_AUDIO * max(1.0, gain)
        else:
            assert got["right"] is None
        assert (got["axcindicate"] == abi.SIGNAL).any(), "mixer %d never had signal" % m
    return mixed, res, launches


def run_files(conf, cfg: abi.EngineCfg, streams, lib: str, ref: bool = False, chunk_bytes: int = 0):
    """The file replay: one host.FileInput reader thread per device (as the reference's input threads), the demodulator
    loop on this thread.  The oracle gets the same bytes in one piece."""
    from boondock_airband_b200 import host
    o = Oracle(cfg, ref=ref)
    for d, s in enumerate(streams):
        o.feed(d, s)
    e = Engine(cfg, lib)
    inputs = []
    try:
        for d, dev in enumerate(cfg.devices):
            assert conf.setting(d, "type") == "file"
            inputs.append(host.FileInput(e, d, conf.setting(d, "filepath"), sample_format=dev.sample_format, sample_rate=dev.sample_rate,
                                         speedup_factor=0.0, chunk_bytes=chunk_bytes))  # unpaced, whatever speedup_factor the file says
        for i in inputs:
            i.start()
        res = e.run_file_inputs(inputs)
        mixed = e.mixer_results()
        for d, i in enumerate(inputs):
            assert i.state == host.INPUT_FAILED  # end of file disables the input, as input-file.cpp:107-111
            assert i.bytes == np.ascontiguousarray(streams[d]).nbytes
    finally:
        for i in inputs:
            i.stop()
        e.close()
    from oracle.ba_oracle import mix_reference
    for m, mx in enumerate(cfg.mixers):  # mixers named by the configuration file: whole-path check against the oracle
        left, right, sig = mix_reference(cfg, o.waveout, o.status, mx)
        got = mixed[m]
        assert len(left) > 0 and len(got["left"]) == len(left) and np.array_equal(got["axcindicate"], sig)
        gain = max(1.0, sum(abs(i.ampfactor) for i in mx.inputs))
        assert float(np.abs(got["left"] - left).max()) <= TOL_AUDIO * gain
        if mx.stereo:
            assert float(np.abs(got["right"] - right).max()) <= TOL_AUDIO * gain
    return o, res


def compare_streams(cfg: abi.EngineCfg, o: Oracle, res, exact: bool = False, min_open: int = 0):
    """Squelch decisions (trace) identical; audio within TOL_AUDIO (or bit-exact); status scalars consistent."""
    report = []
    for d, dev in enumerate(cfg.devices):
        r = res[d]
        assert r["frames_done"] == o.frames(d), (r["frames_done"], o.frames(d))
        for c in range(len(dev.channels)):
            wo = o.waveout(d, c)
            we = r["waveout"][c]
            assert len(wo) == len(we) and len(wo) == o.batches(d) * cfg.wave_batch, (len(wo), len(we))
            if r["trace"] is not None:
                to = o.trace(d, c)
                te = r["trace"][c]
                flips = int((to != te).sum())
                assert flips == 0, "device %d channel %d: %d squelch decision/state differences" % (d, c, flips)
            if exact:
                assert np.array_equal(wo.view(np.uint32), we.view(np.uint32)), "device %d channel %d: audio not bit-exact" % (d, c)
                err = 0.0
            else:
                err = float(np.abs(wo - we).max()) if len(wo) else 0.0
                assert err <= TOL_AUDIO, "device %d channel %d: audio differs by %g" % (d, c, err)
            if dev.channels[c].has_iq_outputs:
                io = o.iq_out(d, c)
                ie = r["iq_out"][c]
                if exact:
                    assert np.array_equal(io.view(np.uint32), ie.view(np.uint32))
                else:
                    sc = max(1e-30, float(np.abs(io).max()))
                    assert float(np.abs(io - ie).max()) / sc <= TOL_BASEBAND
            so = o.status(d, c)
            assert len(so) == len(r["status"])
            for b, s in enumerate(so):
                g = r["status"][b][c]
                assert g["axcindicate"] == s.axcindicate and g["bin"] == s.bin, (d, c, b, g, s.axcindicate, s.bin)
                assert g["open_count"] == s.open_count and g["flappy_count"] == s.flappy_count
                assert g["ctcss_count"] == s.ctcss_count and g["no_ctcss_count"] == s.no_ctcss_count
                assert g["active_counter"] == s.active_counter
                for key, want in (("signal_level", s.signal_level), ("noise_level", s.noise_level), ("squelch_level", s.squelch_level)):
                    if exact:
                        assert np.float32(g[key]) == np.float32(want), (key, g[key], want)
                    else:
                        assert abs(g[key] - want) <= 1e-4 * max(1.0, abs(want)), (key, g[key], want)
            report.append((d, c, err))
    if min_open:
        opened = sum(int(((o.trace(d, c) & abi.TRACE_OPEN) != 0).sum()) for d, dev in enumerate(cfg.devices) for c in range(len(dev.channels)))
        assert opened >= min_open, "test signal never opened the squelch (%d open samples)" % opened
    return report


def check_picks(cfg: abi.EngineCfg, iq: np.ndarray, lib: str):
    """Picked-bin IQ (the channel baseband) within TOL_BASEBAND of the oracle's, relative to the channel's peak."""
    ocfg = cfg
    o = Oracle(ocfg)
    o.feed(0, iq)
    cfg.flags |= abi.FLAG_KEEP_PICKS  # plain AM inputs keep only |X| on the device otherwise
    e = Engine(cfg, lib)
    try:
        e.submit(0, iq)
        t = e.process()
        r = e.collect(t, 0)
        n = int(r.frames_done)
        worst = 0.0
        for c in range(len(cfg.devices[0].channels)):
            po = o.picks(0, c)[:n]
            pe = e.debug_picks(0, c, 0, n)
            sc = max(1e-30, float(np.abs(po).max()))
            worst = max(worst, float(np.abs(po - pe).max()) / sc)
    finally:
        e.close()
    assert n == o.frames(0)
    assert worst <= TOL_BASEBAND, worst
    return worst


def check_demod_exact(cfg: abi.EngineCfg, streams, lib: str, frames_per_call: int = 1500, ref: bool = False):
    """Feed the oracle's own picked-bin IQ to the demodulator: everything downstream must be bit-exact.
    With FLAG_TRACE every channel runs the general kernel; without it plain AM channels run demod_plain_kernel."""
    ocfg = copy.copy(cfg)
    ocfg.flags |= abi.FLAG_TRACE  # the oracle records picks and decisions only when asked to trace
    o = Oracle(ocfg, ref=ref)
    for d, s in enumerate(streams):
        o.feed(d, s)
    e = Engine(cfg, lib)
    try:
        nd = len(cfg.devices)
        picks = []
        for d, dev in enumerate(cfg.devices):
            per = [o.picks(d, c) for c in range(len(dev.channels))]
            picks.append(np.stack(per, axis=1))  # [frames][C][2]
        pos = [0] * nd
        acc = [dict(waveout=[], iq_out=[], trace=[], status=[], frames_done=0) for _ in range(nd)]
        while True:
            fed = False
            for d in range(nd):
                if pos[d] < picks[d].shape[0]:
                    n = min(frames_per_call, picks[d].shape[0] - pos[d])
                    e.inject_picks(d, picks[d][pos[d]:pos[d] + n])
                    pos[d] += n
                    fed = True
            t = e.process()
            got = False
            for d in range(nd):
                r = e.collect(t, d)
                acc[d]["frames_done"] = r.frames_done
                if r.n_batches:
                    got = True
                    acc[d]["waveout"].append(r.waveout)
                    if r.iq_out is not None:
                        acc[d]["iq_out"].append(r.iq_out)
                    if r.trace is not None:
                        acc[d]["trace"].append(r.trace)
                    for b in range(r.n_batches):
                        acc[d]["status"].append([r.status(b, c) for c in range(r.channel_count)])
            if not fed and not got:
                break
        res = []
        for d in range(nd):
            a = acc[d]
            res.append(dict(waveout=np.concatenate(a["waveout"], axis=1), iq_out=np.concatenate(a["iq_out"], axis=1) if a["iq_out"] else None,
                            trace=np.concatenate(a["trace"], axis=1) if a["trace"] else None, status=a["status"], frames_done=a["frames_done"]))
    finally:
        e.close()
    return compare_streams(cfg, o, res, exact=True)


def squelch_regimes(o: Oracle, cfg: abi.EngineCfg, dev: int = 0):
    """What the oracle's decision trace says the scenario exercised, per channel: the longest stretch of samples that were
    filtered while the squelch stayed CLOSED ("held" by the filtered average), the samples filtered in OPENING, the longest
    LOW_SIGNAL_ABORT stretch, and the number of OPENING episodes."""
    out = []
    for c in range(len(cfg.devices[dev].channels)):
        tr = o.trace(dev, c)
        st, fil = tr & abi.TRACE_STATE_MASK, (tr & abi.TRACE_FILTERED) != 0

        def longest(mask):
            edges = np.flatnonzero(np.diff(np.concatenate(([0], mask.astype(np.int8), [0]))))
            return int((edges[1::2] - edges[::2]).max()) if edges.size else 0

        opening = st == abi.SQ_OPENING
        out.append(dict(held=longest((st == abi.SQ_CLOSED) & fil), opening_filtered=int((opening & fil).sum()),
                        aborted=longest(st == abi.SQ_LOW_SIGNAL_ABORT), openings=int(np.count_nonzero(opening[1:] & ~opening[:-1]))))
    return out


def require_squelch_regimes(cfg: abi.EngineCfg, streams):
    """The scenario does what it is for (checked on the oracle, so that a green parity test means the paths were walked)."""
    o = Oracle(cfg)
    o.feed(0, streams[0])
    reg = squelch_regimes(o, cfg)
    assert max(r["held"] for r in reg) >= 2000, reg             # a whole transmission held CLOSED, every sample filtered
    assert sum(r["held"] >= 500 for r in reg) >= 3, reg
    assert max(r["openings"] for r in reg) >= 8, reg             # one channel flaps CLOSED -> OPENING -> CLOSED
    assert max(r["opening_filtered"] for r in reg) >= 1500, reg
    assert sum(r["aborted"] >= 100 for r in reg) >= 4, reg       # most channels sit out a LOW_SIGNAL_ABORT
    o.close()
    return reg
