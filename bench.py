#!/usr/bin/env python
"""bench.py -- throughput of the koala_b200 hot path (enhanced frames/s, BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

A "step" is one pass of the hot path over one batch: one 256-sample frame for every stream resident on the GPU
(analysis/STFT -> mask estimator -> synthesis/iSTFT).  Default workload = the per-GPU partition of BASELINE.json
configs[3] ("65 536 streams sharded 8xB200"): 8192 streams per GPU, bf16 tensor-core mask estimator, weak scaling, no
data-path collective -- at N = 8 it is exactly that config.  `value` times the steps with PCM already resident in HBM;
`e2e` times the same metric through the public API with pinned HOST buffers (H2D + D2H inside the timed region).
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
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
    # name: (streams per GPU, precision, description)
    "cfg4_8192_per_gpu_bf16": (8192, "bf16", "BASELINE configs[3] per-GPU partition: 8192 concurrent 16 kHz streams/GPU, bf16 tcgen05 mask estimator"),
    "cfg3_4096_bf16": (4096, "bf16", "BASELINE configs[2]: 4096 concurrent streams, 1xB200, bf16 tensor-core mask-estimator GEMMs"),
    "cfg2_256_fp32": (256, "fp32", "BASELINE configs[1]: 256 concurrent streams, 1xB200, fp32 mask path"),
    "cfg5_128_per_gpu_bf16": (128, "bf16", "BASELINE configs[4] per-GPU partition: 128 streams/GPU, state carried across calls"),
}
FRAME = 256
HIDDEN, LAYERS, BINS = 512, 2, 256
MACS_PER_FRAME = BINS * HIDDEN + LAYERS * 2 * 3 * HIDDEN * HIDDEN + HIDDEN * BINS      # 3 407 872
FLOPS_PER_FRAME = 2 * MACS_PER_FRAME                                                   # mask-estimator GEMMs only
GRU_FLOPS_PER_STREAM = 2 * 2 * 3 * HIDDEN * HIDDEN                                     # one GRU layer launch, per stream
STATE_BYTES = 512 + 1024 + LAYERS * HIDDEN * 4                                         # tail + OLA + fp32 h
BYTES_PER_FRAME = 1024 + 2 * STATE_BYTES


# dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed `ncu --set full` capture
# of this workload (fused kernel, profiles/r01d_step_ncu_full_selected_metrics.csv); other workloads have no capture -> null
NCU_DRAM_TRAFFIC_BYTES = {"cfg4_8192_per_gpu_bf16": 61.4e6 + 16.9e6}


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
        sm, mx, reasons = [], [], set()
        for _, r in rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except (ValueError, IndexError):
                pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


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
    """--impl reference: the reference's own CPU implementation cannot run (closed binary, AccessKey + licence server
    needed), so this arm times the oracle port on all host threads, per the tier contract."""
    if rank != 0:
        return
    streams_gpu, precision, desc = WORKLOADS[args.workload]
    threads = os.cpu_count() or 1
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
    line = {
        "impl": "reference", "metric": "enhanced_frames_per_second", "value": value, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": precision, "data": "synthetic",
        "config": {"workload": args.workload, "description": desc, "streams_per_gpu": streams_gpu, "frame_length": FRAME,
                   "note": "reference engine unrunnable here (closed binary + licence key); CPU oracle port timed instead"},
        "rtf_x": value * 0.016,
        "cpu_baseline": {"value": value, "unit": "frames/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)
    try:
        os.remove(model)
    except OSError:
        pass


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=32)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg4_8192_per_gpu_bf16", choices=sorted(WORKLOADS))
    ap.add_argument("--streams", type=int, default=0, help="override streams per GPU")
    ap.add_argument("--ring-frames", type=int, default=64, help="distinct input frames per stream kept in HBM")
    ap.add_argument("--e2e-steps", type=int, default=128)
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
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
    import koala_b200 as kb
    from koala_b200 import _build

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: no CUDA device visible (there is no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)   # NCCL only for barriers / reductions of timings: no data-path collective
    _build.build()

    streams, precision, desc = WORKLOADS[args.workload]
    if args.streams:
        streams = args.streams
    model = bench_model_path()
    eng = kb.BatchKoala(streams, model_path=model, device=f"gpu:{local_rank}", precision=precision)

    ring = args.ring_frames
    host_pcm = synth_pcm(streams, ring, seed=0x4B4F414C + rank)
    # resident in HBM, time-major [ring][B][256]: each step's frames are one contiguous [B][256] block, exactly what a
    # caller of the one-frame-per-call API hands over (stream stride 256)
    d_in = torch.from_numpy(np.ascontiguousarray(host_pcm.transpose(1, 0, 2))).to(dev)
    d_out = torch.empty_like(d_in)
    stream = torch.cuda.Stream(dev)                                # the launching stream: kernels AND timing events go here
    torch.cuda.set_stream(stream)
    lib, handle = eng._library, eng._handle
    from ctypes import c_void_p

    def step(i):
        off = (i % ring) * streams * FRAME * 2                     # ring slot = a [B][256] block: frame of stream s at + s*256
        rc = lib.pv_koala_batch_process_async(handle, d_in.data_ptr() + off, d_out.data_ptr() + off, 1, FRAME,
                                              c_void_p(stream.cuda_stream))
        if rc != 0:
            raise RuntimeError(f"pv_koala_batch_process_async failed with status {rc}")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    sampler = ClockSampler(local_rank)
    sampler.start()
    for i in range(args.warmup):
        step(i)
    barrier()
    sampler.wait_first()
    launches0 = eng.kernel_launches
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_begin = time.time()
    ev0.record(stream)
    for i in range(args.steps):
        step(args.warmup + i)
    ev1.record(stream)
    barrier()
    clocks = sampler.stop(t_begin, time.time())
    ms_local = ev0.elapsed_time(ev1)
    launches_local = eng.kernel_launches - launches0

    # ---- per-kernel-class timing with CUDA events on the launching stream (dominant kernel -> roofline)
    eng.profile(True)
    prof_steps = min(args.steps, 100)
    for i in range(prof_steps):
        step(args.warmup + args.steps + i)
    prof = eng.profile_read()
    eng.profile(False)

    # ---- end to end through the public API with pinned HOST buffers: one call carries e2e_steps frames of every stream in the
    # time-major layout [steps][B][256] (a frame of every stream per 16 ms tick); inside it every step's 256-sample frames go
    # host -> device and its enhanced frames device -> host, chunk by chunk, overlapped with compute by the library's ingest
    # path.  Timed by wall clock around the synchronous call.
    e2e_steps = max(8, args.e2e_steps)
    tm = np.ascontiguousarray(host_pcm.transpose(1, 0, 2))        # [ring][B][256]
    h_in = torch.from_numpy(np.concatenate([tm] * ((e2e_steps + ring - 1) // ring), axis=0)[:e2e_steps]).pin_memory()
    h_out = torch.empty_like(h_in).pin_memory()
    eng.process(h_in[:8].contiguous().pin_memory(), out=torch.empty_like(h_in[:8]).pin_memory(), time_major=True)   # warm-up (allocates staging)
    barrier()
    t0 = time.perf_counter()
    eng.process(h_in, out=h_out, time_major=True)                 # synchronous: returns when h_out is valid
    torch.cuda.synchronize(dev)
    e2e_s_local = time.perf_counter() - t0
    # the same thing one step per call (the latency-bound way to drive the API), for reference
    h1_in = h_in[0].contiguous().pin_memory()
    h1_out = torch.empty_like(h1_in).pin_memory()
    eng.process(h1_in, out=h1_out)
    t0 = time.perf_counter()
    for i in range(20):
        eng.process(h1_in, out=h1_out)
    e2e_single_call_fps = streams * 20 / (time.perf_counter() - t0)
    # ... and one call in the stream-major layout [B][steps][256] (pitched copies, wide output blocks)
    s_in = torch.from_numpy(np.ascontiguousarray(h_in.numpy().transpose(1, 0, 2))).pin_memory()
    s_out = torch.empty_like(s_in).pin_memory()
    eng.process(s_in, out=s_out)                                  # sizes the staging buffers for this layout
    t0 = time.perf_counter()
    eng.process(s_in, out=s_out)
    torch.cuda.synchronize(dev)
    e2e_stream_major_fps = streams * e2e_steps / (time.perf_counter() - t0)
    del s_in, s_out

    # ---- reduce over ranks: SUM of units, MAX of time
    def reduce(v, op):
        if world == 1:
            return v
        t = torch.tensor([float(v)], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=op)
        return t.item()

    SUM, MAX = (dist.ReduceOp.SUM, dist.ReduceOp.MAX) if world > 1 else (None, None)
    ms = reduce(ms_local, MAX)
    total_frames = reduce(streams * args.steps, SUM)
    e2e_s = reduce(e2e_s_local, MAX)
    e2e_frames = reduce(streams * e2e_steps, SUM)
    launches = int(reduce(launches_local, SUM))
    value = total_frames / (ms * 1e-3)
    e2e_value = e2e_frames / e2e_s

    if rank == 0:
        peaks = {}
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                peaks = json.load(f)
        except OSError:
            pass
        fused = precision == "bf16"                            # bf16 path: encoder -> GRU layers -> decoder in ONE kernel
        gru_ms, gru_n = prof["masknet"] if fused else prof["gru"]   # fp32 path: the GRU layer kernel dominates
        dom_flops = (FLOPS_PER_FRAME if fused else GRU_FLOPS_PER_STREAM) * streams
        step_prof_ms = sum(v[0] for v in prof.values()) / max(prof_steps, 1)
        shares = {k: (v[0] / max(sum(x[0] for x in prof.values()), 1e-12)) for k, v in prof.items()}
        if precision == "bf16":
            peak = peaks.get("bf16_tflops_sustained", 1400.0)
            peak_src = "MEASURED_PEAKS.json bf16_tflops_sustained (kernel timed inside a long step)" if peaks else "fallback 1.4 PFLOP/s sustained (B200_PROFILING.md)"
            achieved = dom_flops / (gru_ms / max(gru_n, 1) * 1e-3) / 1e12 if gru_n else None
            roofline = {"bound": "tensor", "kernel": "tc_fused_kernel (encoder + GRU layers + decoder GEMMs of one step, all streams)", "achieved": achieved,
                        "peak": peak, "unit": "TFLOP/s", "frac": (achieved / peak) if achieved else None,
                        "traffic": NCU_DRAM_TRAFFIC_BYTES.get(args.workload) if not args.streams else None,
                        "peak_source": peak_src, "avg_launch_ms": gru_ms / max(gru_n, 1),
                        "algorithmic_flops_per_launch": dom_flops}
        else:
            # fp32 CUDA-core path: no measured fp32 peak in MEASURED_PEAKS.json; nominal 148 SM x 128 FMA x 2 x sm_max_mhz
            peak = 148 * 128 * 2 * (peaks.get("sm_max_mhz", 1965.0) * 1e6) / 1e12
            achieved = GRU_FLOPS_PER_STREAM * streams / (gru_ms / max(gru_n, 1) * 1e-3) / 1e12 if gru_n else None
            roofline = {"bound": "tensor", "kernel": "gru_fp32_kernel (CUDA-core FMA; bound is the fp32 FMA pipe, not the tensor pipe)",
                        "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": (achieved / peak) if achieved else None,
                        "traffic": None, "peak_source": "nominal fp32 FMA peak at clocks.max.sm (no measured fp32 figure)",
                        "avg_launch_ms": gru_ms / max(gru_n, 1), "algorithmic_flops_per_launch": GRU_FLOPS_PER_STREAM * streams}
        roofline["kernel_share_of_step"] = shares
        roofline["step_tensor_frac"] = (value / world) * FLOPS_PER_FRAME / 1e12 / (peaks.get("bf16_tflops_sustained", 1400.0))
        roofline["step_hbm_frac"] = (value / world) * BYTES_PER_FRAME / 1e9 / (peaks.get("hbm_gbs", 6650.0))
        cpu = None
        if not args.no_cpu_baseline:
            from oracle import build_oracle
            build_oracle()
            threads = os.cpu_count() or 1
            v, sample = cpu_baseline(model, precision, args.cpu_seconds, threads, min(streams, max(64, threads * 8)))
            cpu = {"value": v, "unit": "frames/s", "cores": threads, "kind": "port", "sample": sample,
                   "note": "reference CPU engine not measurable (closed binary, needs AccessKey + licence server); "
                           "its CI ceilings: >456 frames/s cpu:1 on GitHub runners (BASELINE.md section 1)"}
        line = {
            "metric": "enhanced_frames_per_second", "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": precision, "data": "synthetic",
            "config": {"workload": args.workload, "description": desc, "streams_per_gpu": streams, "total_streams": streams * world,
                       "frame_length": FRAME, "sample_rate": 16000, "hidden": HIDDEN, "gru_layers": LAYERS,
                       "parallelism": f"stream-partition x{world} (no data-path collective)",
                       "l2": f"input/output rings of {ring} frames/stream = {2 * streams * ring * FRAME * 2 / 2**20:.0f} MiB (> 126 MB L2 at the "
                             f"default size); per-stream recurrent state is re-read every step by construction",
                       "weights": "random-init, seeded (koala_b200.spec.random_model)"},
            "rtf_x": value * 0.016, "rtf_reference_convention": 1.0 / (value * 0.016),
            "flops_per_frame": FLOPS_PER_FRAME, "bytes_per_frame": BYTES_PER_FRAME,
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": streams * FRAME * 2,
                    "d2h_bytes_per_step": streams * FRAME * 2, "steps": e2e_steps,
                    "api": "koala_b200.BatchKoala.process(pinned host tensor [steps][B][256], time_major=True) -> "
                           "pv_koala_batch_process_time_major, one call",
                    "one_step_per_call_value_rank0": e2e_single_call_fps, "stream_major_call_value_rank0": e2e_stream_major_fps},
            "gpu_launches": launches,
            "roofline": roofline,
            "cpu_baseline": cpu,
            "kernel_ms_per_step": {k: v[0] / max(prof_steps, 1) for k, v in prof.items()},
            "profiled_step_ms": step_prof_ms,
        }
        print(json.dumps(line), flush=True)
    try:
        os.remove(model)
    except OSError:
        pass
    eng.delete()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
