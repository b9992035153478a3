#!/usr/bin/env python
"""Key-gated recorder of the REFERENCE engine's outputs and CPU timings (SURVEY.md section 8c "Plan").

The reference engine is a closed binary whose `pv_koala_init` needs a Picovoice AccessKey validated online, so nothing here can
run in the build container (no key, no network).  On a machine that has both, with the reference checkout at
$PV_KOALA_REFERENCE_DIR (default /root/reference):

    PV_ACCESS_KEY=... python tools/record_reference.py [--iterations 20]

drives the UNMODIFIED reference Python binding (binding/python/_koala.py:122-254) over the fixture WAVs, their sum and a seeded
synthetic signal, exactly like the reference's own tests do (binding/python/test_koala.py:71-114), and stores
  tests/golden/ref_<signal>.npy        enhanced int16 of the reference engine (golden vectors that would pin parity)
  tests/golden/ref_record.json         delay_sample, version, per-signal checksums, timings R-ref-1 / R-ref-N / R-ref-py
Timing follows binding/python/test_koala_perf.py:31-58 (one discarded warm-up pass + N timed passes over test.wav):
`cpu:1` and `cpu`, through the binding (R-ref-py) and through a bare ctypes loop with a preallocated frame (R-ref-1 / R-ref-N:
no Python list boxing inside the loop).  If a koala_b200 engine can be created (a B200 is visible) the per-sample deltas against it
are printed as INFORMATION: the two engines implement different, independently specified networks.

Without a key, a reference checkout or a reachable licence server the script says why, writes nothing and exits 0.
`bench.py --impl reference` calls `time_reference()` and falls back to the CPU oracle port when it returns None."""
from __future__ import annotations

import argparse
import hashlib
import importlib.util
import json
import os
import sys
import time
import wave

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
FRAME = 256


def reference_dir() -> str:
    return os.environ.get("PV_KOALA_REFERENCE_DIR", "/root/reference")


def load_reference_binding():
    """Imports the reference package from its checkout without modifying or copying it.  Returns (module, why_not)."""
    pkg_dir = os.path.join(reference_dir(), "binding", "python")
    init = os.path.join(pkg_dir, "__init__.py")
    if not os.path.exists(init):
        return None, "no reference checkout at %s" % reference_dir()
    try:
        spec = importlib.util.spec_from_file_location("pvkoala_reference", init, submodule_search_locations=[pkg_dir])
        mod = importlib.util.module_from_spec(spec)
        sys.modules["pvkoala_reference"] = mod
        spec.loader.exec_module(mod)
        return mod, None
    except Exception as e:      # wrong platform, missing shared object, ...
        return None, "reference binding failed to import: %r" % (e,)


def reference_paths():
    ref = reference_dir()
    return (os.path.join(ref, "lib", "linux", "x86_64", "libpv_koala.so"), os.path.join(ref, "lib", "common", "koala_params.pv"))


def open_reference(device: str):
    """(Koala instance, None) or (None, reason)."""
    key = os.environ.get("PV_ACCESS_KEY", "")
    if not key:
        return None, "PV_ACCESS_KEY is not set"
    mod, why = load_reference_binding()
    if mod is None:
        return None, why
    lib, model = reference_paths()
    try:
        return mod.Koala(access_key=key, model_path=model, device=device, library_path=lib), None
    except Exception as e:      # KoalaActivationError etc.: no network / bad key
        return None, "reference pv_koala_init failed: %s" % (str(e).replace("\n", " | "),)


def read_wav(path):
    with wave.open(path, "rb") as f:
        return np.frombuffer(f.readframes(f.getnframes()), dtype="<i2").copy()


def signals():
    test, noise = read_wav(os.path.join(GOLDEN, "test.wav")), read_wav(os.path.join(GOLDEN, "noise.wav"))
    mixed = np.clip(test.astype(np.int32) + noise.astype(np.int32), -32768, 32767).astype(np.int16)
    rng = np.random.default_rng(0x4B4F414C)
    t = np.arange(4 * 16000) / 16000.0
    synth = 2030.0 * np.sin(2 * np.pi * 180.0 * t) * 0.5 * (1 + np.sin(2 * np.pi * 4.0 * t)) + 760.0 * rng.standard_normal(t.size)
    return {"test": test, "noise": noise, "mixed": mixed, "synthetic": np.clip(np.rint(synth), -32768, 32767).astype(np.int16)}


def run_stream(koala, pcm):
    fl = koala.frame_length
    out = [koala.process(pcm[i:i + fl].tolist()) for i in range(0, len(pcm) - fl + 1, fl)]
    return np.asarray(out, np.int16).reshape(-1)


def time_binding(koala, pcm, iterations):
    """test_koala_perf.py:45-52: wall time of one pass through Koala.process, first pass discarded."""
    fl, n = koala.frame_length, len(pcm) // koala.frame_length
    frames = [tuple(pcm[j * fl:(j + 1) * fl].tolist()) for j in range(n)]
    res = []
    for i in range(iterations + 1):
        t0 = time.perf_counter()
        for f in frames:
            koala.process(f)
        if i > 0:
            res.append(time.perf_counter() - t0)
    return sum(res) / len(res), n


def time_bare(koala, pcm, iterations):
    """The same pass through pv_koala_process with preallocated ctypes frames (no per-frame boxing): the engine's own time."""
    from ctypes import c_short
    fl, n = koala.frame_length, len(pcm) // koala.frame_length
    frames = [(c_short * fl)(*pcm[j * fl:(j + 1) * fl].tolist()) for j in range(n)]
    out = (c_short * fl)()
    fn, handle = koala.process_func, koala._handle
    res = []
    for i in range(iterations + 1):
        t0 = time.perf_counter()
        for f in frames:
            if fn(handle, f, out) != 0:
                raise RuntimeError("pv_koala_process failed")
        if i > 0:
            res.append(time.perf_counter() - t0)
    return sum(res) / len(res), n


def time_reference(device: str = "cpu", iterations: int = 5):
    """frames/s of the reference engine on this machine, or None with a reason.  Used by bench.py --impl reference."""
    koala, why = open_reference(device)
    if koala is None:
        return None, why
    try:
        pcm = read_wav(os.path.join(GOLDEN, "test.wav"))
        sec, n = time_bare(koala, pcm, iterations)
        return {"frames_per_second": n / sec, "seconds_per_pass": sec, "frames_per_pass": n, "device": device,
                "version": koala.version}, None
    finally:
        koala.delete()


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(description=__doc__.split("\n")[0])
    ap.add_argument("--iterations", type=int, default=20)
    ap.add_argument("--no-write", action="store_true")
    args = ap.parse_args(argv)
    koala, why = open_reference("cpu:1")
    if koala is None:
        print("reference recorder skipped: %s" % why)
        print("reference CPU baseline: not measurable (no AccessKey/network); see BASELINE.md section 3")
        return 0
    record = {"reference_version": koala.version, "delay_sample": koala.delay_sample, "frame_length": koala.frame_length,
              "signals": {}, "timings": {}}
    outs = {}
    for name, pcm in signals().items():
        koala.reset()
        outs[name] = run_stream(koala, pcm)
        record["signals"][name] = {"samples": int(outs[name].size), "sha1": hashlib.sha1(outs[name].tobytes()).hexdigest()}
        if not args.no_write:
            np.save(os.path.join(GOLDEN, "ref_%s.npy" % name), outs[name])
    test = signals()["test"]
    sec, n = time_binding(koala, test, args.iterations)
    record["timings"]["R-ref-py cpu:1"] = {"seconds_per_pass": sec, "frames_per_second": n / sec}
    sec, n = time_bare(koala, test, args.iterations)
    record["timings"]["R-ref-1 cpu:1"] = {"seconds_per_pass": sec, "frames_per_second": n / sec}
    koala.delete()
    many, why = open_reference("cpu")
    if many is not None:
        sec, n = time_bare(many, test, args.iterations)
        record["timings"]["R-ref-N cpu"] = {"seconds_per_pass": sec, "frames_per_second": n / sec, "host_threads": os.cpu_count()}
        many.delete()
    if not args.no_write:
        with open(os.path.join(GOLDEN, "ref_record.json"), "w") as f:
            json.dump(record, f, indent=1)
    print(json.dumps(record["timings"], indent=1))
    try:                                   # information only: deltas against this repository's engine
        sys.path.insert(0, ROOT)
        import koala_b200 as kb
        ours = kb.create(kb.ANY_ACCESS_KEY)
        for name, pcm in signals().items():
            ours.reset()
            mine = run_stream(ours, pcm)
            d, shift = record["delay_sample"], ours.delay_sample
            a, b = outs[name][d:], mine[shift:]
            m = min(len(a), len(b))
            diff = np.abs(a[:m].astype(np.int32) - b[:m].astype(np.int32))
            print("delta vs koala_b200 on %-9s (delay-aligned, information only): max %d LSB, rms %.1f LSB" % (name, diff.max(), float(np.sqrt(np.mean(diff.astype(np.float64) ** 2)))))
        ours.delete()
    except Exception as e:
        print("no koala_b200 engine here for the informational comparison: %s" % (str(e).split("\n")[0],))
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
