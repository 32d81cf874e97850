"""The C-ABI boundary without a GPU: the library loads, exports every function include/ba_cuda.h declares, the ctypes
mirrors have the C layout, and the engine refuses to run without a CUDA device (there is no CPU path)."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from boondock_airband_b200 import abi, configs, engine

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "ba_cuda.h")


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(engine.LIB_PATH):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "boondock_airband_b200", "csrc")], stdout=subprocess.DEVNULL)
    return engine.load_library()


def declared_functions():
    text = open(HEADER).read()
    return sorted(set(re.findall(r"^BA_API\s+[\w\s\*]+?\b(ba_cuda_\w+)\s*\(", text, re.M)))


def test_header_and_binding_agree():
    names = declared_functions()
    assert len(names) >= 20
    assert sorted(engine.SYMBOLS) == names


def test_library_exports_every_declared_symbol(lib):
    out = subprocess.check_output(["nm", "-D", "--defined-only", engine.LIB_PATH], text=True)
    exported = set(re.findall(r"\bT (ba_cuda_\w+)", out))
    assert exported == set(declared_functions())
    for name in declared_functions():
        assert hasattr(lib, name)


def test_struct_layouts_match_c(tmp_path):
    src = tmp_path / "sizes.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "ba_cuda.h"\nint main(void){printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\\n",'
                   "sizeof(ba_channel_desc),sizeof(ba_device_desc),sizeof(ba_engine_desc),sizeof(ba_channel_status),sizeof(ba_step_out),sizeof(ba_channel_info),"
                   "offsetof(ba_engine_desc,ring_bytes),offsetof(ba_step_out,frames_done),offsetof(ba_device_desc,channels),"
                   "sizeof(ba_mixer_input_desc),sizeof(ba_mixer_desc),sizeof(ba_mixer_out),offsetof(ba_engine_desc,mixers),offsetof(ba_mixer_out,axcindicate));return 0;}\n")
    exe = tmp_path / "sizes"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src)])
    got = [int(x) for x in subprocess.check_output([str(exe)], text=True).split()]
    want = [C.sizeof(abi.ChannelDesc), C.sizeof(abi.DeviceDesc), C.sizeof(abi.EngineDesc), C.sizeof(abi.ChannelStatus), C.sizeof(abi.StepOut), C.sizeof(abi.ChannelInfo),
            abi.EngineDesc.ring_bytes.offset, abi.StepOut.frames_done.offset, abi.DeviceDesc.channels.offset,
            C.sizeof(abi.MixerInputDesc), C.sizeof(abi.MixerDesc), C.sizeof(abi.MixerOut), abi.EngineDesc.mixers.offset, abi.MixerOut.axcindicate.offset]
    assert got == want


def test_header_is_plain_c(tmp_path):
    src = tmp_path / "c89.c"
    src.write_text('#include "ba_cuda.h"\nint main(void){return BA_OK;}\n')
    subprocess.check_call(["gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-c", "-o", str(tmp_path / "c.o"), str(src)])


def test_no_cuda_device_means_no_engine(lib):
    """On a box without a GPU the engine must fail loudly, never compute on the CPU."""
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("a CUDA device is present")
    assert lib.ba_cuda_visible_devices() == -1
    with pytest.raises(engine.EngineError) as ei:
        engine.Engine(configs.cfg1())
    assert ei.value.code == -1 and "no CPU path" in str(ei.value)


def test_bad_arguments_are_rejected_before_cuda(lib):
    desc, keep = abi.build_desc(configs.cfg1())
    desc.abi_version = 99
    h = C.c_void_p()
    assert lib.ba_cuda_create(C.byref(desc), C.byref(h)) == -4
    desc, keep = abi.build_desc(configs.cfg1())
    desc.fft_size = 300
    assert lib.ba_cuda_create(C.byref(desc), C.byref(h)) == -2  # same meaning as gpu_fft_prepare's -2
    assert lib.ba_cuda_process(None) == -4
    assert b"null" in lib.ba_cuda_last_error()


def test_product_never_touches_the_oracle():
    """Nothing under boondock_airband_b200/ may import, include, link or load the CPU oracle or the emulation build."""
    pkg = os.path.join(ROOT, "boondock_airband_b200")
    for dirpath, _, files in os.walk(pkg):
        if "build" in dirpath:
            continue
        for f in files:
            path = os.path.join(dirpath, f)
            if f.endswith(".py"):
                text = open(path).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, re.M), path
                assert "libba_oracle" not in text and "TESTONLY" not in text, path
            elif f.endswith((".cu", ".cpp", ".h", "Makefile")):
                text = open(path).read()
                assert not re.search(r'#include\s+"[^"]*oracle', text), path
                assert "libba_oracle" not in text and "TESTONLY" not in text, path
                if "cuda_emu.h" in text:  # the only mention allowed: the BA_EMU branch of ba_port.h
                    assert f == "ba_port.h" and re.search(r'#ifdef BA_EMU\s*\n#include "cuda_emu.h"', text), path
    assert "oracle" not in subprocess.check_output(["ldd", engine.LIB_PATH], text=True)
    syms = subprocess.check_output(["nm", "-C", engine.LIB_PATH], text=True)
    assert "emu::" not in syms and "ba_oracle" not in syms


def test_frame_arithmetic_matches_the_reference_rule():
    """boondock_airband.cpp:418-424: a frame runs while available >= bps + fft_size * bytes_per_sample * 2."""
    cfg = configs.cfg1()
    hop, frame = cfg.hop_bytes(0), cfg.frame_bytes(0)
    assert hop == 640 and frame == 1024
    assert cfg.frames_for(0, hop + frame - 1) == 0
    assert cfg.frames_for(0, hop + frame) == 1
    assert cfg.frames_for(0, 2 * hop + frame) == 2
    assert cfg.batches_for(0, 1099 * hop + hop + frame) == 1  # the first batch needs WAVE_BATCH + AGC_EXTRA frames
    assert cfg.batches_for(0, 1098 * hop + hop + frame) == 0
    c2 = configs.cfg2()
    assert c2.hop(0) == 150 and c2.hop_bytes(0) == 600 and c2.wave_batch == 2000
