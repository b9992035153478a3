#!/usr/bin/env python
"""bench.py -- throughput of the koala_b200 hot path (enhanced frames/s, BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

A "step" is one pass of the hot path over one batch: one 256-sample frame for every stream resident on the GPU
(analysis/STFT -> mask estimator -> synthesis/iSTFT).  Default workload = the per-GPU partition of BASELINE.json
configs[3] ("65 536 streams sharded 8xB200"): 8192 streams per GPU, bf16 tensor-core mask estimator, weak scaling, no
data-path collective -- at N = 8 it is exactly that config.  The steps are fed `frames_per_process_call` frames per call
(the library walks a call's frames in chunks: analysis of the chunk, ONE persistent mask-estimator launch stepping through
its frames, synthesis of the chunk); the one-frame-per-call rate is reported beside it (`one_frame_per_call`).  `value` times
the steps with PCM already resident in HBM;
`e2e` times the same metric through the public API with pinned HOST buffers (H2D + D2H inside the timed region).
At N = 1 the other BASELINE workloads are measured too (shorter runs) and reported under `others`.
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: streams per GPU, precision, frames per process() call, description
    "cfg4_8192_per_gpu_bf16": dict(streams=8192, precision="bf16", frames_per_call=32, time_major=True,
                                   desc="BASELINE configs[3] per-GPU partition: 8192 concurrent 16 kHz streams/GPU, bf16 tcgen05 mask estimator"),
    "cfg3_4096_bf16": dict(streams=4096, precision="bf16", frames_per_call=32, time_major=True,
                           desc="BASELINE configs[2]: 4096 concurrent streams, 1xB200, bf16 tensor-core mask-estimator GEMMs"),
    "cfg2_256_fp32": dict(streams=256, precision="fp32", frames_per_call=64,
                          desc="BASELINE configs[1]: 256 concurrent streams, 1xB200, fp32 mask path (tcgen05 with every activation split into "
                               "three bf16 planes, fp32 accumulation), fed 64 frames per process() call"),
    "cfg5_128_per_gpu_bf16": dict(streams=128, precision="bf16", frames_per_call=64, steps=37504,
                                  desc="BASELINE configs[4] per-GPU partition: 128 streams/GPU (1024 over 8 GPUs), 10-minute clips (37 504 frames) "
                                       "fed in 64-frame process() calls with the state carried across all 586 calls"),
    "fixed_point_4096_int8": dict(streams=4096, precision="int8", frames_per_call=16,
                                  desc="SURVEY section 8f row 3: fixed-point variant (int8 weights x int16 activations on tcgen05 kind::i8, integer "
                                       "gates; SPEC.md section 6), 4096 concurrent streams, 1xB200"),
}
DEFAULT_WORKLOAD = "cfg4_8192_per_gpu_bf16"
FRAME = 256
HIDDEN, LAYERS, BINS = 512, 2, 256
MACS_PER_FRAME = BINS * HIDDEN + LAYERS * 2 * 3 * HIDDEN * HIDDEN + HIDDEN * BINS      # 3 407 872
FLOPS_PER_FRAME = 2 * MACS_PER_FRAME                                                   # mask-estimator GEMMs only
GRU_FLOPS_PER_STREAM = 2 * 2 * 3 * HIDDEN * HIDDEN                                     # one GRU layer launch, per stream
STATE_BYTES = 512 + 1024 + LAYERS * HIDDEN * 4                                         # tail + OLA + fp32 h
BYTES_PER_FRAME = 1024 + 2 * STATE_BYTES
BURST_MAX_SECONDS = 1.0      # a timed region shorter than this is a burst: compare with the burst peak, else with the sustained one


def source_hash() -> str:
    """Identifies the kernel sources a profile belongs to (profiles/traffic.json is keyed by it)."""
    h = hashlib.sha1()
    d = os.path.join(ROOT, "koala_b200", "csrc")
    for name in sorted(os.listdir(d)):
        if name.endswith((".cu", ".cuh", ".h")):
            with open(os.path.join(d, name), "rb") as f:
                h.update(name.encode() + b"\0" + f.read())
    return h.hexdigest()[:16]


def ncu_traffic(workload: str):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed `ncu --set full` capture
    (profiles/traffic.json, written by tools/ncu_traffic.py from the .ncu-rep).  Returns (bytes | None, provenance)."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            t = json.load(f)
    except (OSError, ValueError):
        return None, "no profiles/traffic.json"
    e = t.get(workload)
    if not e:
        return None, "no ncu capture of this workload in profiles/traffic.json"
    same = e.get("source_hash") == source_hash()
    return e["dram_bytes_per_launch"], f"{e.get('source', 'profiles/traffic.json')} ({'same kernel sources' if same else 'captured on OLDER kernel sources'})"


def synth_pcm(n_streams: int, n_frames: int, seed: int) -> np.ndarray:
    """SURVEY.md section 8d synthetic input: half band-limited noise (rms 760 LSB), half speech-like harmonic stack
    (rms 2030 LSB, 4 Hz syllabic AM) + noise.  Built from a small pool and tiled so set-up stays fast."""
    rng = np.random.default_rng(seed)
    pool = min(n_streams, 256)
    n = n_frames * FRAME
    t = np.arange(n) / 16000.0
    out = np.empty((pool, n), np.float32)
    for s in range(pool):
        noise = rng.standard_normal(n).astype(np.float32) * 760.0
        if s % 2 == 0:
            out[s] = noise
        else:
            f0 = rng.uniform(100, 250)
            harm = sum(np.sin(2 * np.pi * f0 * k * t + rng.uniform(0, 6.28)) / k for k in range(1, 10))
            sp = harm * 0.5 * (1 + np.sin(2 * np.pi * 4.0 * t + rng.uniform(0, 6.28)))
            out[s] = sp / (np.sqrt(np.mean(sp ** 2)) + 1e-9) * 2030.0 + noise
    pcm = np.clip(np.rint(out), -32768, 32767).astype(np.int16).reshape(pool, n_frames, FRAME)
    reps = (n_streams + pool - 1) // pool
    return np.ascontiguousarray(np.tile(pcm, (reps, 1, 1))[:n_streams])


def bench_model_path() -> str:
    """Random-init weights of the spec's architecture (seeded), as the contract asks for synthetic benchmarks."""
    from koala_b200 import spec
    d = os.path.join(ROOT, "gpurun_out")
    os.makedirs(d, exist_ok=True)
    p = os.path.join(d, f"bench_random_{os.getpid()}.kpv")
    spec.save_model(p, spec.random_model())
    return p


class ClockSampler:
    """nvidia-smi clocks line of B200_PROFILING.md, sampled every 20 ms; started before warm-up so that samples exist
    for short timed regions, and reduced over the samples that fall inside the timed window."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows, self.proc, self.idx = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20",
                                          "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def wait_first(self, timeout=3.0):
        t0 = time.time()
        while self.proc and not self.rows and time.time() - t0 < timeout:
            time.sleep(0.01)

    def stop(self, t_begin=None, t_end=None):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                pass
        rows = self.rows
        if t_begin is not None and rows:
            inside = [r for r in rows if t_begin - 0.03 <= r[0] <= t_end + 0.03]
            rows = inside if inside else [min(rows, key=lambda r: abs(r[0] - 0.5 * (t_begin + t_end)))]
        sm, mx, pw, reasons = [], [], [], set()
        for _, r in rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); pw.append(float(r[3]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except (ValueError, IndexError):
                pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm),
                "power_w_max": float(max(pw))}


def cpu_baseline(model_path: str, precision: str, seconds: float, threads: int, streams: int):
    """Times the CPU oracle (a port: the reference engine is closed and licence-gated) on a bounded sample of the workload."""
    from oracle import OracleBatch, OracleModel
    frames = 4
    pcm = synth_pcm(streams, frames, seed=0x4B4F414C)
    ob = OracleBatch(OracleModel(model_path), streams, precision)
    ob.process(pcm[:, :1], threads=threads)           # warm-up
    done, t0 = 0, time.perf_counter()
    while time.perf_counter() - t0 < seconds:
        ob.process(pcm, threads=threads)
        done += streams * frames
    dt = time.perf_counter() - t0
    return done / dt, f"{streams} streams x {done // streams} frames, {dt:.1f} s, C oracle (oracle/koala_oracle.c), {precision} mode"


def run_reference(args, rank: int):
    """--impl reference: times the reference's own CPU engine when it can run here (PV_ACCESS_KEY set, reference checkout present,
    licence server reachable: tools/record_reference.py), `kind: "reference"`.  Otherwise -- always, in the build container and on
    the GPU box: the engine is a closed binary that validates an AccessKey online -- the CPU oracle port on all
    host threads, `kind: "port"`, per the tier contract."""
    if rank != 0:
        return
    w = WORKLOADS[args.workload]
    streams_gpu, precision, desc = w["streams"], w["precision"], w["desc"]
    threads = os.cpu_count() or 1
    base = {"impl": "reference", "metric": "enhanced_frames_per_second", "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": precision, "data": "synthetic"}
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    try:
        import record_reference
        ref, why = record_reference.time_reference("cpu", max(1, args.steps))
    except Exception as e:          # the recorder must never take the arm down
        ref, why = None, repr(e)
    if ref is not None:
        value = ref["frames_per_second"]
        sample = (f"reference engine {ref['version']} (lib/linux/x86_64/libpv_koala.so), device=cpu ({threads} host threads), "
                  f"{args.steps} passes over test.wav ({ref['frames_per_pass']} frames), bare pv_koala_process loop")
        line = dict(base, value=value, ms_per_step=1e3 * ref["seconds_per_pass"],
                    config={"workload": args.workload, "description": desc, "streams_per_gpu": streams_gpu, "frame_length": FRAME,
                            "note": "single-stream reference engine; a step = one pass over the fixture WAV"},
                    rtf_x=value * 0.016, cpu_baseline={"value": value, "unit": "frames/s", "cores": threads, "kind": "reference", "sample": sample},
                    e2e={"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0})
        print(json.dumps(line), flush=True)
        return
    sample_streams = min(streams_gpu, max(threads * 8, 64))
    from oracle import OracleBatch, OracleModel, build_oracle
    build_oracle()
    model = bench_model_path()
    pcm = synth_pcm(sample_streams, 1, seed=0x4B4F414C)
    ob = OracleBatch(OracleModel(model), sample_streams, precision)
    # keep the whole run within a few minutes: calibrate frames per step to ~1 s
    t0 = time.perf_counter(); ob.process(pcm, threads=threads); one = time.perf_counter() - t0
    frames_per_step = int(min(64, max(1, 1.0 / max(one, 1e-4))))
    pcm = synth_pcm(sample_streams, frames_per_step, seed=0x4B4F414C)
    for _ in range(args.warmup):
        ob.process(pcm, threads=threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ob.process(pcm, threads=threads)
    dt = time.perf_counter() - t0
    value = args.steps * sample_streams * frames_per_step / dt
    sample = f"each step = {sample_streams} streams x {frames_per_step} frames of the workload on {threads} host threads (C oracle port)"
    line = dict(base, value=value, ms_per_step=1e3 * dt / args.steps,
                config={"workload": args.workload, "description": desc, "streams_per_gpu": streams_gpu, "frame_length": FRAME,
                        "note": "reference engine unrunnable here (%s); CPU oracle port timed instead" % why},
                rtf_x=value * 0.016, cpu_baseline={"value": value, "unit": "frames/s", "cores": threads, "kind": "port", "sample": sample},
                e2e={"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0})
    print(json.dumps(line), flush=True)
    try:
        os.remove(model)
    except OSError:
        pass


class Runner:
    """One workload on this rank's GPU: device-resident timing, per-kernel timing, end-to-end timing through host buffers."""

    def __init__(self, name, streams, model, local_rank, world, total_streams=None):
        import torch
        import koala_b200 as kb
        w = WORKLOADS[name]
        self.name, self.w, self.torch, self.world = name, w, torch, world
        self.streams, self.precision, self.fpc = streams, w["precision"], w["frames_per_call"]
        self.time_major = bool(w.get("time_major")) or self.fpc == 1
        self.dev = torch.device("cuda", local_rank)
        if total_streams is None:
            self.eng = kb.BatchKoala(streams, model_path=model, device=f"gpu:{local_rank}", precision=self.precision)
        else:       # under torchrun: the product's multi-GPU front door owns the partition
            self.sharded = kb.ShardedKoala(total_streams, model_path=model, precision=self.precision)
            assert self.sharded.num_streams == streams
            self.eng = self.sharded.engine
        self.stream = torch.cuda.Stream(self.dev)        # the launching stream: kernels AND timing events go here

    def barrier(self):
        if self.world > 1:
            import torch.distributed as dist
            dist.barrier()
        self.torch.cuda.synchronize(self.dev)

    def setup_ring(self, ring, seed):
        torch = self.torch
        self.ring = ring
        self.host_pcm = synth_pcm(self.streams, ring, seed=seed)           # [B][ring][256]
        if self.time_major:
            # time-major [ring][B][256]: each step's frames are one contiguous [B][256] block, what a caller that collects one
            # frame of every stream per 16 ms tick has (stream stride 256, frame stride B * 256)
            self.d_in = torch.from_numpy(np.ascontiguousarray(self.host_pcm.transpose(1, 0, 2))).to(self.dev)
        else:
            self.d_in = torch.from_numpy(self.host_pcm).to(self.dev)       # stream-major [B][ring][256]: multi-frame calls
        self.d_out = torch.empty_like(self.d_in)

    def run_steps(self, first, count):
        """Enqueues steps [first, first + count) on the launching stream; returns the number of process() calls made."""
        from ctypes import c_void_p
        lib, handle, st = self.eng._library, self.eng._handle, c_void_p(self.stream.cuda_stream)
        calls = 0
        i, end = first, first + count
        while i < end:
            t0 = i % self.ring
            n = min(self.fpc, end - i, self.ring - t0)                     # frames of this call (state carries to the next one)
            if self.time_major:                                            # ring slot = a [B][256] block: frame of stream s at + s*256
                off = t0 * self.streams * FRAME * 2
                rc = lib.pv_koala_batch_process_async_strided(handle, self.d_in.data_ptr() + off, self.d_out.data_ptr() + off, n, FRAME,
                                                              self.streams * FRAME, st)
            else:
                off = t0 * FRAME * 2
                rc = lib.pv_koala_batch_process_async(handle, self.d_in.data_ptr() + off, self.d_out.data_ptr() + off, n, self.ring * FRAME, st)
            if rc != 0:
                raise RuntimeError(f"pv_koala_batch_process_async failed with status {rc}")
            i += n
            calls += 1
        return calls

    def timed(self, steps, warmup, sampler=None):
        torch = self.torch
        torch.cuda.set_stream(self.stream)
        self.run_steps(0, warmup)
        self.barrier()
        if sampler:
            sampler.wait_first()
        launches0 = self.eng.kernel_launches
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self.barrier()
        t_begin = time.time()
        ev0.record(self.stream)
        calls = self.run_steps(warmup, steps)
        ev1.record(self.stream)
        self.barrier()
        clocks = sampler.stop(t_begin, time.time()) if sampler else None
        return ev0.elapsed_time(ev1), self.eng.kernel_launches - launches0, calls, clocks

    def profile(self, first, steps):
        self.eng.profile(True)
        self.run_steps(first, steps)
        prof = self.eng.profile_read()
        self.eng.profile(False)
        return prof

    def e2e(self, e2e_steps, extras=True):
        """Through the public API with pinned HOST buffers: one call carries e2e_steps frames of every stream in the time-major layout
        [steps][B][256] (a frame of every stream per 16 ms tick); inside it every step's frames go host -> device and its enhanced
        frames device -> host, chunk by chunk, overlapped with compute by the library's ingest path.  Wall clock around the call."""
        torch, eng, ring = self.torch, self.eng, self.ring
        tm = np.ascontiguousarray(self.host_pcm.transpose(1, 0, 2))        # [ring][B][256]
        h_in = torch.from_numpy(np.concatenate([tm] * ((e2e_steps + ring - 1) // ring), axis=0)[:e2e_steps]).pin_memory()
        h_out = torch.empty_like(h_in).pin_memory()
        eng.process(h_in, out=h_out, time_major=True)                 # warm-up: the same call once (sizes the staging buffers, touches the pinned pages)
        each = []
        for _ in range(3):                                            # the same call three times; the MEDIAN is reported (a single 10 ms call
            self.barrier()                                            # through the host link scattered by +-10 % from run to run)
            t0 = time.perf_counter()
            eng.process(h_in, out=h_out, time_major=True)             # synchronous: returns when h_out is valid
            torch.cuda.synchronize(self.dev)
            each.append(time.perf_counter() - t0)
        res = {"seconds": sorted(each)[1], "seconds_each": each}
        if extras:
            # the same thing one step per call (the latency-bound way to drive the API), for reference
            h1_in = h_in[0].contiguous().pin_memory()
            h1_out = torch.empty_like(h1_in).pin_memory()
            eng.process(h1_in, out=h1_out)
            t0 = time.perf_counter()
            for _ in range(20):
                eng.process(h1_in, out=h1_out)
            res["one_step_per_call_fps"] = self.streams * 20 / (time.perf_counter() - t0)
            # ... and one call in the stream-major layout [B][steps][256] (pitched copies, wide output blocks)
            s_in = torch.from_numpy(np.ascontiguousarray(h_in.numpy().transpose(1, 0, 2))).pin_memory()
            s_out = torch.empty_like(s_in).pin_memory()
            eng.process(s_in, out=s_out)                                  # sizes the staging buffers for this layout
            t0 = time.perf_counter()
            eng.process(s_in, out=s_out)
            torch.cuda.synchronize(self.dev)
            res["stream_major_call_fps"] = self.streams * e2e_steps / (time.perf_counter() - t0)
        return res

    def host_link(self, mib=64, reps=12):
        """What the host side of the end-to-end path can carry, measured in the same job with every rank copying at once: pinned
        host <-> device copies of `mib` MiB in both directions simultaneously (the shape of the ingest path's traffic: one contiguous
        chunk each way) and a plain host memcpy.  GB/s of THIS rank; the caller sums over ranks."""
        torch = self.torch
        n = mib << 20
        h_in, h_out = torch.empty(n, dtype=torch.uint8).pin_memory(), torch.empty(n, dtype=torch.uint8).pin_memory()
        d_in, d_out = torch.empty(n, dtype=torch.uint8, device=self.dev), torch.zeros(n, dtype=torch.uint8, device=self.dev)
        s_in, s_out = torch.cuda.Stream(self.dev), torch.cuda.Stream(self.dev)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        def both(k):
            with torch.cuda.stream(s_in):
                for _ in range(k):
                    d_in.copy_(h_in, non_blocking=True)
            with torch.cuda.stream(s_out):
                for _ in range(k):
                    h_out.copy_(d_out, non_blocking=True)
        both(2)
        torch.cuda.synchronize(self.dev)
        self.barrier()
        ev[0].record(s_in); ev[2].record(s_out)
        both(reps)
        ev[1].record(s_in); ev[3].record(s_out)
        torch.cuda.synchronize(self.dev)
        h2d = n * reps / (ev[0].elapsed_time(ev[1]) * 1e-3) / 1e9
        d2h = n * reps / (ev[2].elapsed_time(ev[3]) * 1e-3) / 1e9
        a, b = np.empty(n * 2, np.uint8), np.ones(n * 2, np.uint8)
        np.copyto(a, b)
        self.barrier()
        t0 = time.perf_counter()
        for _ in range(4):
            np.copyto(a, b)
        host = 4 * a.nbytes / (time.perf_counter() - t0) / 1e9
        return {"h2d_gbs": h2d, "d2h_gbs": d2h, "host_memcpy_gbs": host}

    def close(self):
        self.eng.delete()


def roofline_of(name, streams, precision, prof, prof_steps, timed_seconds, peaks):
    """The roofline object of the dominant kernel: algorithmic flops per launch / its CUDA-event duration against the measured peak."""
    fused = prof["masknet"][1] > 0                               # encoder -> GRU layers -> decoder in ONE kernel
    dom_ms, dom_n = prof["masknet"] if fused else prof["gru"]
    dom_flops = (FLOPS_PER_FRAME if fused else GRU_FLOPS_PER_STREAM) * streams * (prof_steps / max(dom_n, 1) if fused else 1.0)
    tot = max(sum(v[0] for v in prof.values()), 1e-12)
    shares = {k: v[0] / tot for k, v in prof.items()}
    burst = timed_seconds < BURST_MAX_SECONDS
    key = "bf16_tflops" if burst else "bf16_tflops_sustained"
    if peaks.get(key):
        peak, src = peaks[key], f"MEASURED_PEAKS.json {key} (timed region {timed_seconds * 1e3:.1f} ms: a {'burst' if burst else 'seconds-long run'}), of measured"
    else:
        peak, src = (1590.0, "fallback 1.59 PFLOP/s burst (B200_PROFILING.md), of fallback") if burst else (1400.0, "fallback 1.4 PFLOP/s sustained (B200_PROFILING.md), of fallback")
    achieved = dom_flops / (dom_ms / max(dom_n, 1) * 1e-3) / 1e12 if dom_n else None
    if precision == "int8":      # integer tensor pipe: no measured peak in MEASURED_PEAKS.json to divide by
        peak, src = None, "no measured int8 tensor peak available (MEASURED_PEAKS.json has bf16 only); achieved is in Tera integer op/s"
    traffic, traffic_src = ncu_traffic(name) if streams == WORKLOADS[name]["streams"] else (None, "stream count overridden")
    parts = -(-streams // 4096) if (fused and streams > 4096) else 1      # engine.cu kPartStreams: big batches run as partitions that fit the L2
    per_launch = prof_steps * parts // max(dom_n, 1)
    kernel = ("tc_fused_kernel (encoder + GRU layers + decoder GEMMs, " + (f"one of {parts} partitions of {streams // parts} streams" if parts > 1 else "all streams") +
              ("" if per_launch == 1 else f", {per_launch} steps per launch") +
              (", fp32 operands as three bf16 planes: 3x the algorithmic flops are executed" if precision == "fp32" else "") + ")") if fused \
        else ("i8_layer_kernel<GRU> (one GRU layer, tcgen05 kind::i8, hi and lo byte planes: 2x the algorithmic multiply-adds are executed, "
              "plus a quarter of zero rows)" if precision == "int8" else "gru_fp32_kernel (CUDA-core FMA GRU layer)")
    r = {"bound": "tensor", "kernel": kernel, "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
         "frac": (achieved / peak) if achieved and peak else None, "traffic": traffic, "traffic_source": traffic_src, "peak_source": src,
         "avg_launch_ms": dom_ms / max(dom_n, 1), "algorithmic_flops_per_launch": dom_flops, "kernel_share_of_step": shares,
         "note": "launch duration from CUDA events bracketing the kernel; bracketing disables the dependent-launch overlap with the "
                 "neighbouring kernels, so the bracketed durations sum to more than ms_per_step"}
    return r


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=32)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--streams", type=int, default=0, help="override streams per GPU")
    ap.add_argument("--ring-frames", type=int, default=64, help="distinct input frames per stream kept in HBM")
    ap.add_argument("--e2e-steps", type=int, default=128)
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-others", action="store_true", help="skip the other BASELINE workloads (measured at N = 1 only)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist
    from koala_b200 import _build

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: no CUDA device visible (there is no CPU fallback)")
    _build.build()
    w = WORKLOADS[args.workload]
    streams, precision, desc = (args.streams or w["streams"]), w["precision"], w["desc"]
    model = bench_model_path() if rank == 0 or world == 1 else None

    # CPU baseline first, at N = 1 only, before any GPU work or process group exists: the oracle port on all host threads,
    # on a bounded sample of the same workload (with other ranks spinning in NCCL barriers the host cores are not free)
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        from oracle import build_oracle
        build_oracle()
        threads = os.cpu_count() or 1
        v, sample = cpu_baseline(model, precision, args.cpu_seconds, threads, min(streams, max(64, threads * 8)))
        cpu = {"value": v, "unit": "frames/s", "cores": threads, "kind": "port", "sample": sample,
               "note": "reference CPU engine not measurable (closed binary, needs AccessKey + licence server); "
                       "its CI ceilings: >456 frames/s cpu:1 on GitHub runners (BASELINE.md section 1)"}

    # one rank = one GPU + its own slice of the host cores: the copy-issuing thread and the pinned buffers it first touches
    # stay on cores no other rank uses (a no-op at N = 1)
    cores = None
    if world > 1 and hasattr(os, "sched_setaffinity"):
        try:
            avail = sorted(os.sched_getaffinity(0))
            per = max(1, len(avail) // world)
            cores = avail[(local_rank * per) % len(avail):][:per]
            os.sched_setaffinity(0, cores)
        except OSError:
            cores = None
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)   # NCCL only for barriers / reductions of timings: no data-path collective
        box = [model]
        dist.broadcast_object_list(box, src=0)           # every rank reads the same model file (one node)
        model = box[0]

    def reduce(v, op):
        if world == 1:
            return v
        t = torch.tensor([float(v)], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=op)
        return t.item()

    SUM, MAX = (dist.ReduceOp.SUM, dist.ReduceOp.MAX) if world > 1 else (None, None)
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except OSError:
        pass

    run = Runner(args.workload, streams, model, local_rank, world, total_streams=streams * world if world > 1 else None)
    run.setup_ring(args.ring_frames, seed=0x4B4F414C + rank)
    sampler = ClockSampler(local_rank)
    sampler.start()
    ms_local, launches_local, calls_local, clocks = run.timed(args.steps, args.warmup, sampler)
    prof_steps = min(args.steps, 128) // w["frames_per_call"] * w["frames_per_call"] or min(args.steps, 128)
    prof = run.profile(-(-(args.warmup + args.steps) // args.ring_frames) * args.ring_frames, prof_steps)     # starts at a ring boundary: whole calls only
    one_frame = None
    if w["frames_per_call"] > 1:        # the same steps driven one frame per call (three launches per frame)
        fpc, run.fpc = run.fpc, 1
        o_ms, o_launches, _, _ = run.timed(min(args.steps, 256), args.warmup)
        run.fpc = fpc
        one_frame = {"value": reduce(streams * min(args.steps, 256), SUM) / (reduce(o_ms, MAX) * 1e-3), "unit": "frames/s",
                     "ms_per_step": reduce(o_ms, MAX) / min(args.steps, 256), "gpu_launches_rank0": o_launches}
    e2e_steps = max(8, args.e2e_steps)
    e2e = run.e2e(e2e_steps)
    link = run.host_link()

    ms = reduce(ms_local, MAX)
    total_frames = reduce(streams * args.steps, SUM)
    e2e_each = [reduce(s, MAX) for s in e2e["seconds_each"]]          # per repetition: the slowest rank
    e2e_s = sorted(e2e_each)[1]
    e2e_frames = reduce(streams * e2e_steps, SUM)
    launches = int(reduce(launches_local, SUM))
    value = total_frames / (ms * 1e-3)
    e2e_value = e2e_frames / e2e_s
    # the host's ceiling for the end-to-end metric: every frame crosses the link once each way (512 B in, 512 B out)
    link_sum = {k: reduce(v, SUM) for k, v in link.items()}
    link_min = {k: -reduce(-v, MAX) for k, v in link.items()}
    ceiling = min(link_sum["h2d_gbs"], link_sum["d2h_gbs"]) * 1e9 / (FRAME * 2)
    # the ranks do not get equal shares of the host link, and the job's time is the MAX over ranks: the slowest rank's share sets it
    ceiling_slowest = world * min(link_min["h2d_gbs"], link_min["d2h_gbs"]) * 1e9 / (FRAME * 2)
    host_link = {"h2d_gbs": link_sum["h2d_gbs"], "d2h_gbs": link_sum["d2h_gbs"], "host_memcpy_gbs": link_sum["host_memcpy_gbs"],
                 "per_rank_min": link_min, "ranks_copying_at_once": world, "cores_per_rank": len(cores) if cores else (os.cpu_count() or 1),
                 "e2e_ceiling_frames_per_s": ceiling, "e2e_fraction_of_ceiling": e2e_value / ceiling,
                 "e2e_ceiling_at_slowest_rank_frames_per_s": ceiling_slowest, "e2e_fraction_of_slowest_rank_ceiling": e2e_value / ceiling_slowest,
                 "how": "64 MiB pinned copies, both directions at once on two streams, all ranks at once (sum over ranks); ceiling = min(h2d, d2h) / 512 B per frame; at_slowest_rank = ranks x the smallest per-rank share / 512 B (time is max over ranks)"}
    run.close()

    others = None
    if world == 1 and not args.no_others and not args.streams and args.workload == DEFAULT_WORKLOAD:
        # the other BASELINE workloads, same build, shorter runs: one compact line each (device-resident value, end-to-end value,
        # the dominant kernel's roofline fraction)
        others = {}
        for name, ow in WORKLOADS.items():
            if name == args.workload:
                continue
            try:
                r = Runner(name, ow["streams"], model, local_rank, 1)
                r.setup_ring(args.ring_frames, seed=0x4B4F414C)
                o_steps = max(ow["frames_per_call"] * 4, min(args.steps, 512))
                if args.steps >= 512 and ow.get("steps"):
                    o_steps = ow["steps"]        # the config's own length (configs[4]: a 10-minute clip per stream)
                o_ms, o_launches, o_calls, _ = r.timed(o_steps, max(args.warmup, ow["frames_per_call"]))
                o_prof_steps = max(ow["frames_per_call"], 64)
                o_prof = r.profile(-(-o_steps // args.ring_frames) * args.ring_frames, o_prof_steps)
                o_e2e = r.e2e(max(8, min(e2e_steps, 64)), extras=False)
                o_value = ow["streams"] * o_steps / (o_ms * 1e-3)
                roof = roofline_of(name, ow["streams"], ow["precision"], o_prof, o_prof_steps, o_ms * 1e-3, peaks)
                others[name] = {"description": ow["desc"], "value": o_value, "unit": "frames/s", "steps": o_steps, "ms_per_step": o_ms / o_steps,
                                "frames_per_call": ow["frames_per_call"], "process_calls": o_calls, "gpu_launches": o_launches,
                                "dtype": ow["precision"], "rtf_x": o_value * 0.016,
                                "rtf_x_at_8_gpus_if_linear": o_value * 0.016 * 8,
                                "e2e": {"value": ow["streams"] * max(8, min(e2e_steps, 64)) / o_e2e["seconds"], "unit": "frames/s",
                                        "h2d_bytes_per_step": ow["streams"] * FRAME * 2, "d2h_bytes_per_step": ow["streams"] * FRAME * 2},
                                "roofline": {k: roof[k] for k in ("kernel", "achieved", "peak", "frac", "avg_launch_ms", "peak_source")}}
                r.close()
            except Exception as e:      # a side workload must never take the headline line down
                others[name] = {"error": repr(e)[:300]}

    if others is not None:
        # BASELINE configs[0]: ONE stream through the reference-shaped per-frame API (create / process / delete), the fixture WAV's length
        try:
            import koala_b200 as kb
            k = kb.create(access_key=kb.ANY_ACCESS_KEY, device=f"gpu:{local_rank}")
            frame = [int(v) for v in synth_pcm(1, 1, 7)[0, 0]]
            for _ in range(20):
                k.process(frame)
            t0 = time.perf_counter()
            for _ in range(365):
                k.process(frame)
            dt = time.perf_counter() - t0
            k.delete()
            others["cfg1_single_stream_api"] = {
                "description": "BASELINE configs[0] shape: one stream, koala_b200.create(...).process(frame) per 256-sample frame (Python list in, list out; "
                               "H2D, three launches, D2H and a synchronisation per call), 365 frames = the fixture WAV's length, shipped weights",
                "value": 365 / dt, "unit": "frames/s", "us_per_call": dt / 365 * 1e6, "rtf_x": 365 / dt * 0.016, "dtype": "bf16",
                "note": "latency of one call, not throughput; the reference's CI floor for its CPU engine is 456 frames/s (BASELINE.md section 1)"}
        except Exception as e:
            others["cfg1_single_stream_api"] = {"error": repr(e)[:300]}

    if rank == 0:
        roofline = roofline_of(args.workload, streams, precision, prof, prof_steps, ms * 1e-3, peaks)
        step_prof_ms = sum(v[0] for v in prof.values()) / max(prof_steps, 1)
        step_peak = roofline["peak"]
        roofline["step_tensor_frac"] = (value / world) * FLOPS_PER_FRAME / 1e12 / step_peak if step_peak else None
        roofline["step_hbm_frac"] = (value / world) * BYTES_PER_FRAME / 1e9 / (peaks.get("hbm_gbs", 6650.0))
        line = {
            "metric": "enhanced_frames_per_second", "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": precision, "data": "synthetic",
            "config": {"workload": args.workload, "description": desc, "streams_per_gpu": streams, "total_streams": streams * world,
                       "frame_length": FRAME, "sample_rate": 16000, "hidden": HIDDEN, "gru_layers": LAYERS,
                       "frames_per_process_call": w["frames_per_call"],
                       "parallelism": f"stream-partition x{world} (koala_b200.ShardedKoala, no data-path collective)" if world > 1 else "one GPU",
                       "l2": f"input/output rings of {args.ring_frames} frames/stream = {2 * streams * args.ring_frames * FRAME * 2 / 2**20:.0f} MiB (> 126 MB L2 at the "
                             f"default size); per-stream recurrent state is re-read every step by construction",
                       "weights": "random-init, seeded (koala_b200.spec.random_model)"},
            "rtf_x": value * 0.016, "rtf_reference_convention": 1.0 / (value * 0.016),
            "flops_per_frame": FLOPS_PER_FRAME, "bytes_per_frame": BYTES_PER_FRAME,
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": streams * FRAME * 2,
                    "d2h_bytes_per_step": streams * FRAME * 2, "steps": e2e_steps,
                    "repetitions": {"seconds_each_max_over_ranks": e2e_each, "reported": "median"},
                    "api": "koala_b200.BatchKoala.process(pinned host tensor [steps][B][256], time_major=True) -> "
                           "pv_koala_batch_process_time_major, one call",
                    "one_step_per_call_value_rank0": e2e.get("one_step_per_call_fps"), "stream_major_call_value_rank0": e2e.get("stream_major_call_fps")},
            "host_link": host_link,
            "gpu_launches": launches,
            "process_calls": calls_local,
            "one_frame_per_call": one_frame,
            "roofline": roofline,
            "cpu_baseline": cpu,
            "kernel_ms_per_step": {k: v[0] / max(prof_steps, 1) for k, v in prof.items()},
            "profiled_step_ms": step_prof_ms,
            "others": others,
        }
        print(json.dumps(line), flush=True)
        try:
            os.remove(model)
        except OSError:
            pass
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
