"""CPU-side checks of the engine's host logic and kernel index arithmetic through the thread-emulated build of the
product sources (tests/emu).  These do not replace the GPU parity tests (tests/test_gpu_parity.py, -m gpu); they catch
logic errors before a B200 is involved.  The emulation library is test infrastructure and is never loaded by the package."""
import numpy as np
import pytest

from boondock_airband_b200 import abi
from boondock_airband_b200.abi import ChannelCfg, DeviceCfg, EngineCfg

import parity
import scenarios


@pytest.fixture(scope="module")
def emu(oracle_built):
    return parity.lib_for("emu")


@pytest.mark.parametrize("n", [256, 512, 1024, 2048, 4096, 8192])
@pytest.mark.parametrize("fmt", ["u8", "s8", "s16", "f32"])
def test_emu_frames(emu, n, fmt):
    rng = np.random.default_rng(n + len(fmt))
    dev = DeviceCfg(sample_rate=2_400_000, centerfreq=100_000_000, sample_format=fmt,
                    channels=[ChannelCfg(freq=100_000_000 + 12_500 * k) for k in range(-5, 6)])
    cfg = EngineCfg(fft_size=n, wave_rate=16000, devices=[dev])
    nfr = 9
    iq = parity.random_iq(rng, fmt, 150 * (nfr - 1) + n + 3)
    parity.check_frames(cfg, iq, nfr, emu)


def test_emu_u8_all_codes(emu):
    """Every u8 code converts to exactly (i - 127.5f) / 127.5f times the window (boondock_airband.cpp:341-343)."""
    dev = DeviceCfg(sample_rate=2_560_000, centerfreq=120_000_000, channels=[ChannelCfg(freq=120_100_000)])
    cfg = EngineCfg(fft_size=256, wave_rate=8000, devices=[dev])
    iq = np.repeat(np.arange(256, dtype=np.uint8), 2)[:512]
    iq = np.concatenate([iq, iq[::-1]]).astype(np.uint8)
    iq = np.resize(iq, 2 * (320 * 3 + 256))
    parity.check_frames(cfg, iq, 4, emu)


def test_emu_cfg1(emu):
    cfg, streams = scenarios.cfg1_short(0.9)
    o, res, _ = parity.run_both(cfg, streams, emu, chunk_bytes=700_001)
    parity.compare_streams(cfg, o, res, min_open=1000)


def test_emu_cfg1_fast_path(emu):
    """Without the trace flag plain AM channels take the specialised (PLAIN) instantiation of the demodulator."""
    cfg, streams = scenarios.cfg1_short(0.9)
    cfg.flags = 0
    o, res, _ = parity.run_both(cfg, streams, emu, chunk_bytes=700_001)
    assert res[0]["trace"] is None
    parity.compare_streams(cfg, o, res)
    assert float(np.abs(res[0]["waveout"]).max()) > 0.05


def test_emu_plain_kernel_exact(emu):
    """demod_plain_kernel fed with the oracle's picks: audio, levels and counters bit-exact (ragged frame counts per call)."""
    cfg, streams = scenarios.cfg1_short(0.9)
    cfg.flags = 0
    parity.check_demod_exact(cfg, streams, emu, frames_per_call=1777)


def test_emu_plain_kernel_stress_exact(emu):
    """The speculative quad path of demod_plain_kernel against the sequential loop on marginal, flapping, clipping
    carriers (every squelch state, flap detection, low-signal aborts, AGC clip): bit-exact."""
    cfg, streams = scenarios.am_stress(1.4)
    cfg.flags = 0
    parity.check_demod_exact(cfg, streams, emu, frames_per_call=2333)


def test_emu_cfg2(emu):
    cfg, streams = scenarios.cfg2_small(6, 1.1)
    o, res, _ = parity.run_both(cfg, streams, emu, chunk_bytes=1_000_003)
    parity.compare_streams(cfg, o, res, min_open=1000)


@pytest.mark.parametrize("fm_demod", [abi.FM_FAST_ATAN2, abi.FM_QUADRI_DEMOD])
def test_emu_mixed_options_exact(emu, fm_demod):
    cfg, streams = scenarios.mixed_options(0.7, fm_demod=fm_demod, afc=False)  # AFC needs K1's spectrum: end-to-end test below
    parity.check_channel_info(cfg, emu)
    parity.check_demod_exact(cfg, streams, emu, frames_per_call=1777)


def test_emu_squelch_held_by_the_filtered_average_exact(emu):
    """The chunk-at-a-time paths of the general demodulator for a squelch that only counts - held CLOSED while every sample is
    filtered, OPENING (with and without the filtered average running), LOW_SIGNAL_ABORT - against the sequential loop, bit for
    bit (audio, per-sample decisions, counters), on a scenario the oracle's trace shows to spend thousands of samples in each."""
    cfg, streams = scenarios.held_by_post_filter(1.0)
    parity.require_squelch_regimes(cfg, streams)
    parity.check_demod_exact(cfg, streams, emu, frames_per_call=2111)


def test_emu_mixed_options_end_to_end(emu):
    cfg, streams = scenarios.mixed_options(0.7)
    o, res, _ = parity.run_both(cfg, streams, emu, chunk_bytes=555_555)
    parity.compare_streams(cfg, o, res, min_open=1000)


def test_emu_multi_device(emu):
    cfg, streams = scenarios.multi_device(0.5)
    o, res, _ = parity.run_both(cfg, streams, emu, chunk_bytes=300_000)
    parity.compare_streams(cfg, o, res, min_open=500)


def test_emu_file_replay_from_a_configuration_file(emu, tmp_path):
    """Rows f-1/f-2 end to end: configuration text -> ba_conf -> engine; IQ files -> reader threads -> input rings ->
    ba_cuda_process on the demodulator thread, against the oracle fed the same bytes."""
    conf, cfg, streams = scenarios.file_replay(tmp_path, 0.45)
    assert [len(d.channels) for d in cfg.devices] == [4, 3] and cfg.wave_rate == 16000
    assert [(i.device, i.channel) for i in cfg.mixers[0].inputs] == [(0, 1), (1, 0)] and cfg.mixers[0].stereo
    parity.check_channel_info(cfg, emu)
    o, res = parity.run_files(conf, cfg, streams, emu)
    parity.compare_streams(cfg, o, res, min_open=500)


@pytest.mark.parametrize("chunk", [300_000, 1_100_000])
def test_emu_mixers(emu, chunk):
    """Row f-4: mixers summed on the device (K3), inputs arriving unevenly from three devices."""
    cfg, streams = scenarios.mixers_on_multi_device(0.5)
    mixed, res, _ = parity.check_mixers(cfg, streams, emu, chunk)
    assert mixed[1]["right"] is not None and mixed[0]["right"] is None


def test_emu_mixer_masked_input(emu):
    """mixer_disable_input (mixer.cpp:96-112): a masked input is neither waited for nor summed."""
    cfg, streams = scenarios.mixers_on_multi_device(0.5)
    streams[1] = streams[1][: len(streams[1]) // 3]  # device 1 dies early; without the mask mixers 0 and 1 would stop with it
    mixed, res, _ = parity.check_mixers(cfg, streams, emu, 400_000, masked={0: [1], 1: [1]})
    assert len(mixed[0]["left"]) > len(res[1]["waveout"][0])


def test_emu_scan_mode(emu):
    """Row f-3: freq_idx switches between batches; every frequency resumes its own Squelch/filters/AGC/counters."""
    cfg, streams = scenarios.scan_mode(1.3)
    o, res, plan, _ = parity.run_scan(cfg, streams, emu, every=2, order=[0, 1, 2, 1, 0, 2])
    assert len(plan) >= 4 and len({i for _, i in plan}) == 3
    assert [b for b, _ in plan] == sorted(b for b, _ in plan) and plan[0][0] >= 2
    parity.compare_streams(cfg, o, res, min_open=1000)
    # without the switches the oracle hears something else: the comparison above is not vacuous
    from oracle.ba_oracle import Oracle
    plain = Oracle(cfg)
    plain.feed(0, streams[0])
    assert not np.array_equal(plain.waveout(0, 0), o.waveout(0, 0))


def test_emu_results_do_not_depend_on_how_the_stream_is_cut(emu):
    """The same bytes fed in ragged pieces, in large pieces, and after an odd number of leading frames give bit-identical
    audio: a frame meets the same arithmetic (incl. the rotation of the odd FFT group, channelize.cu) wherever the step
    and tile boundaries fall."""
    from boondock_airband_b200.engine import Engine
    cfg, streams = scenarios.cfg1_short(0.5)
    cfg.flags = 0
    runs = []
    for chunk in (99_991, 640 * 3 + 1024, 2_000_000):  # the middle one: a first step of exactly three frames, then the rest
        e = Engine(cfg, emu)
        try:
            if chunk < 10_000:
                e.submit(0, streams[0][:chunk])
                t = e.process()
                assert e.collect(t, 0).frames_done == 3  # the next launch starts on an odd frame of the stream
                e.submit(0, streams[0][chunk:chunk + 1_000_000])
                rest = streams[0][chunk + 1_000_000:]
                res = e.run_stream([rest], chunk_bytes=777_777)
            else:
                res = e.run_stream(streams, chunk_bytes=chunk)
        finally:
            e.close()
        runs.append(res[0]["waveout"])
    assert runs[0].shape == runs[1].shape == runs[2].shape and runs[0].shape[1] >= 2 * cfg.wave_batch
    assert np.array_equal(runs[0].view(np.uint32), runs[2].view(np.uint32))
    assert np.array_equal(runs[1].view(np.uint32), runs[2].view(np.uint32))


def test_emu_two_engines_on_two_devices_in_one_process(emu, monkeypatch):
    """One engine per GPU inside one process (one demod thread per device, boondock_airband.cpp:1088-1122): the kernels'
    dynamic shared-memory limit is an attribute of the (device, function) pair, so an engine created on a second device must set
    it there itself.  The emulation refuses a launch whose limit was not set on the current device; the two engines run
    interleaved from one thread, so every entry point has to select its own device."""
    import threading
    from boondock_airband_b200.engine import Engine
    monkeypatch.setenv("BA_EMU_DEVICES", "2")
    cfg, streams = scenarios.cfg1_short(0.5)
    cfgs = []
    for dev in (0, 1):
        c, _ = scenarios.cfg1_short(0.5)
        c.cuda_device = dev
        cfgs.append(c)
    engines = [Engine(c, emu) for c in cfgs]
    try:
        res = [None, None]
        # interleaved from ONE thread first (the current device must follow the engine that is called) ...
        views = np.ascontiguousarray(streams[0]).view(np.uint8).reshape(-1)
        half = views.size // 2
        accs = [e._new_acc() for e in engines]
        for part in (views[:half], views[half:]):
            for e, acc in zip(engines, accs):
                e.submit(0, part)
                while True:
                    produced, advanced = e._step_into(acc)
                    if not produced and not advanced:
                        break
        res = [e._finish_acc(acc) for e, acc in zip(engines, accs)]
        # ... then one engine per thread, as the reference's demod threads would
        out = [None, None]

        def work(k):
            out[k] = engines[k].launch_count()
        th = [threading.Thread(target=work, args=(k,)) for k in range(2)]
        for t in th:
            t.start()
        for t in th:
            t.join()
        assert out[0] > 0 and out[1] > 0
    finally:
        for e in engines:
            e.close()
    a, b = res[0][0]["waveout"], res[1][0]["waveout"]
    assert len(a) == len(b) and all(np.array_equal(x, y) for x, y in zip(a, b))
    assert res[0][0]["frames_done"] == res[1][0]["frames_done"] > 0


def test_emu_skip_silent_rows(emu):
    """BA_FLAG_SKIP_SILENT_ROWS: only the (channel, batch) rows that hold anything but +0.0f come back, with a map; put back
    together they are the audio of a run without the flag, bit for bit, and silence was in fact skipped."""
    from boondock_airband_b200.engine import Engine
    cfg, streams = scenarios.cfg1_short(0.9)
    cfg.flags = 0
    e = Engine(cfg, emu)
    want = e.run_stream(streams, chunk_bytes=600_000)
    e.close()
    cfg2, _ = scenarios.cfg1_short(0.9)
    cfg2.flags = abi.FLAG_SKIP_SILENT_ROWS
    e = Engine(cfg2, emu)
    got = e.run_stream(streams, chunk_bytes=450_001)
    e.close()
    assert got[0]["rows_skipped"] > 0
    for c in range(len(cfg.devices[0].channels)):
        assert np.array_equal(want[0]["waveout"][c].view(np.uint32), got[0]["waveout"][c].view(np.uint32))
    assert got[0]["status"] == want[0]["status"]
