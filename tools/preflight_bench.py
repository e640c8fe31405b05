#!/usr/bin/env python
"""Pre-flight of bench.py on a machine without a GPU.

Runs bench.py's own main() — every leg, unchanged — with the C ABI answered by tests/hostmodel/libplsvo_hostmodel.so (the
product's host code on a model CUDA runtime, model kernels backed by the CPU oracle; lean end-to-end batches, which the
oracle cannot take, get the digest kernel) and with the handful of torch.cuda entry points bench.py touches replaced by
host stand-ins (events that read the wall clock, a stream whose handle is a model-runtime stream, no-op synchronise,
generation and pinned memory on the CPU).  What it is for: catching a Python error, a broken leg or a malformed JSON line
before GPU time is spent.  What it is not: a measurement — every number in the line it prints is meaningless.

usage: python tools/preflight_bench.py [bench.py arguments]      (defaults: a reduced batch so that it ends in minutes)"""
from __future__ import annotations

import ctypes as C
import importlib.util
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main(argv):
    spec = importlib.util.spec_from_file_location("hm_build", os.path.join(ROOT, "tests", "hostmodel", "build.py"))
    hm = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(hm)
    os.environ["PLSVO_LIB"] = hm.build()
    sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
    import oracle_lib

    oracle_lib.build()
    os.environ["PLSVO_FAKE_ORACLE"] = os.path.join(ROOT, "oracle", "libplsvo_oracle.so")
    os.environ["PLSVO_FAKE_LEAN_DIGEST"] = "1"

    import torch

    import plsvo_b200
    from plsvo_b200 import abi, synth

    lib = abi.load_library()
    lib.fake_cuda_errors.restype = C.c_char_p

    class Event:
        def __init__(self, enable_timing=True):
            self.t = None

        def record(self, stream=None):
            self.t = time.perf_counter()

        def elapsed_time(self, other):
            return max(1e3 * (other.t - self.t), 1e-3)

    class Stream:
        def __init__(self, device=None):
            h = C.c_void_p()
            assert lib.cudaStreamCreateWithFlags(C.byref(h), 1) == 0
            self.cuda_stream = h.value

    torch.cuda.is_available = lambda: True
    torch.cuda.set_device = lambda i: None
    torch.cuda.synchronize = lambda dev=None: None
    torch.cuda.Event, torch.cuda.Stream = Event, Stream
    torch.cuda.set_stream = lambda s: None
    torch.Tensor.pin_memory = lambda self, *a, **k: self
    real_device = torch.device
    for name in ("empty", "tensor", "zeros"):
        def on_cpu(*a, _f=getattr(torch, name), **k):
            k.pop("device", None)
            return _f(*a, **k)
        setattr(torch, name, on_cpu)
    import inspect

    for name, fn in list(vars(synth).items()):  # every generator that takes a device generates on the CPU
        if inspect.isfunction(fn) and "device" in inspect.signature(fn).parameters:
            def gen_on_cpu(*a, _f=fn, **k):
                k["device"] = "cpu"
                return _f(*a, **k)
            setattr(synth, name, gen_on_cpu)
    torch.cuda.get_device_name = lambda i=0: "host model (no GPU)"
    _to = torch.Tensor.to

    def to_cpu(self, *a, **k):  # tensor.to(cuda device) stays where it is
        a = tuple("cpu" if isinstance(x, torch.device) and x.type == "cuda" else x for x in a)
        if isinstance(k.get("device"), torch.device) and k["device"].type == "cuda":
            k["device"] = "cpu"
        return _to(self, *a, **k)

    torch.Tensor.to = to_cpu

    spec = importlib.util.spec_from_file_location("bench", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    if not any(a.startswith("--batch") for a in argv):
        argv = ["--batch", "260", "--n-pts", "60", "--n-segs", "16", "--steps", "3", "--warmup", "3", "--cpu-sample", "64"] + argv
    sys.argv = ["bench.py"] + argv
    import contextlib
    import io

    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        rc = bench.main()
    out = buf.getvalue().strip().splitlines()
    assert rc == 0 and out, "bench.main() printed nothing"
    line = json.loads(out[-1])  # one JSON line, parseable
    errs = lib.fake_cuda_errors()
    report = {"model_runtime_errors": errs.decode() if errs else "", "keys": sorted(line)}
    for k in ("e2e", "e2e_chain", "poseopt", "next_rows", "cpu_baseline", "roofline", "clocks"):
        v = line.get(k)
        report[k] = ("error: " + str(v["error"])) if isinstance(v, dict) and "error" in v else ("present" if v is not None else None)
    print(json.dumps(report, indent=1))
    for k in ("e2e", "e2e_chain"):
        print(k, "=", json.dumps(line.get(k)))
    print("poseopt.e2e =", json.dumps((line.get("poseopt") or {}).get("e2e")))
    print(out[-1][:600] + " ...")
    bad = [k for k, v in report.items() if isinstance(v, str) and v.startswith("error")]
    assert real_device is torch.device
    return 1 if bad or report["model_runtime_errors"] else 0


if __name__ == "__main__":
    sys.exit(main(sys.argv[1:]))
