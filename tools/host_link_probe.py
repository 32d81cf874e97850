#!/usr/bin/env python
"""Host link probe: what one pinned host<->device copy reaches on this box, per GPU alone and with all GPUs copying at once,
for ordinary and write-combined pinned host memory.  Used to name what bounds the end-to-end leg of bench.py (PCIe) and how the
GPUs of a box share the host side (profiles/).  cudart through ctypes; no engine involved.

    python tools/host_link_probe.py [--mb 1024] [--reps 4] > gpurun_out/host_link.json
"""
import argparse
import ctypes as C
import json
import subprocess
import threading
import time


def cudart():
    for name in ("libcudart.so", "libcudart.so.12", "/usr/local/cuda/lib64/libcudart.so"):
        try:
            return C.CDLL(name)
        except OSError:
            continue
    raise SystemExit("no libcudart")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mb", type=int, default=1024)
    ap.add_argument("--reps", type=int, default=4)
    a = ap.parse_args()
    rt = cudart()
    n = C.c_int(0)
    assert rt.cudaGetDeviceCount(C.byref(n)) == 0
    ng = n.value
    nbytes = a.mb << 20
    out = {"gpus": ng, "bytes_per_copy": nbytes, "reps": a.reps, "rows": []}
    try:
        out["topo"] = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True, timeout=30).stdout
    except Exception as ex:  # noqa: BLE001
        out["topo"] = repr(ex)

    def setup(dev, wc):
        assert rt.cudaSetDevice(dev) == 0
        h = C.c_void_p()
        d = C.c_void_p()
        assert rt.cudaHostAlloc(C.byref(h), C.c_size_t(nbytes), C.c_uint(4 if wc else 0)) == 0  # cudaHostAllocWriteCombined = 4
        assert rt.cudaMalloc(C.byref(d), C.c_size_t(nbytes)) == 0
        C.memset(h, 1, nbytes)
        s = C.c_void_p()
        assert rt.cudaStreamCreate(C.byref(s)) == 0
        return h, d, s

    def run(dev, bufs, kind, res, key, barrier):
        h, d, s = bufs
        rt.cudaSetDevice(dev)
        dst, src = (d, h) if kind == 1 else (h, d)
        rt.cudaMemcpyAsync(dst, src, C.c_size_t(nbytes), C.c_int(kind), s)
        rt.cudaStreamSynchronize(s)
        barrier.wait()
        t0 = time.perf_counter()
        for _ in range(a.reps):
            rt.cudaMemcpyAsync(dst, src, C.c_size_t(nbytes), C.c_int(kind), s)
        rt.cudaStreamSynchronize(s)
        res[key] = a.reps * nbytes / (time.perf_counter() - t0) / 1e9

    for wc in (False, True):
        bufs = [setup(g, wc) for g in range(ng)]
        for kind, name in ((1, "h2d"), (2, "d2h")):
            if wc and kind == 2:
                continue  # the CPU reads write-combined memory uncached: not what a result buffer should be
            alone = {}
            for g in range(ng):
                run(g, bufs[g], kind, alone, g, threading.Barrier(1))
            together = {}
            if ng > 1:
                bar = threading.Barrier(ng)
                th = [threading.Thread(target=run, args=(g, bufs[g], kind, together, g, bar)) for g in range(ng)]
                for t in th:
                    t.start()
                for t in th:
                    t.join()
            out["rows"].append({"memory": "write-combined pinned" if wc else "pinned", "direction": name, "alone_gbs": [round(alone[g], 1) for g in range(ng)],
                                "all_at_once_gbs": [round(together[g], 1) for g in range(ng)] if together else None,
                                "all_at_once_sum_gbs": round(sum(together.values()), 1) if together else None})
        for h, d, s in bufs:
            rt.cudaFreeHost(h)
            rt.cudaFree(d)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
