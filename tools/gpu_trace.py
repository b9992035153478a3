"""Dump the clock64() timeline of one fused mask-estimator launch -- development aid for the tcgen05 pipeline.
Needs a trace build of the library:  python -m koala_b200._build -DKOALA_FU_TRACE=1 -otrace.so
    python tools/gpu_trace.py <streams> <trace.so> [frames per launch] [first tile recorded] [tile for per-k-block detail]
Per tile of cluster 0 (CTA 0 = pair leader, CTA 1 = its peer): when the producers got it, when its dependencies were met, when
the MMA issuer started / finished, when the accumulator was full and when the epilogue handed the buffer back."""
import os, sys
import numpy as np
os.environ["KOALA_FU_TRACE_BUF"] = "1"
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
lib = os.path.abspath(sys.argv[2]) if len(sys.argv) > 2 else None
frames = int(sys.argv[3]) if len(sys.argv) > 3 else 16
skip = int(sys.argv[4]) if len(sys.argv) > 4 else 0
os.environ["KOALA_FU_TRACE_SKIP"] = str(skip)
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import koala_b200 as kb
from koala_b200 import spec
m = "gpurun_out/r.kpv"; os.makedirs("gpurun_out", exist_ok=True); spec.save_model(m, spec.random_model())
eng = kb.BatchKoala(n, model_path=m, precision="bf16", library_path=lib)
frames = min(frames, eng.chunk_frames)
pcm = torch.from_numpy((np.random.default_rng(0).standard_normal((frames, n, 256)) * 2000).astype(np.int16)).cuda()
for _ in range(3):
    eng.process(pcm, time_major=True)       # the trace buffer keeps the last launch
torch.cuda.synchronize()
tr = eng.debug_read("trace", (2, 1024), np.int64)
print(f"{n} streams, {frames} steps per launch, tiles {skip}.. of pair 0")
for cta in range(2):
    t = tr[cta]; t0 = t[1020]
    r = lambda x: int(x - t0) if x else -1
    ns = int(t[1015] - t[1012])
    print(f"== CTA {cta}: start 0, after setup {r(t[1021])}, roles done {r(t[1022])}, end {r(t[1023])}; {ns} ns by globaltimer = {r(t[1023]) / max(ns, 1):.3f} GHz SM clock")
    for it in range(21):
        b = it * 48
        if t[b+4] == 0 and t[b+0] == 0: continue
        kbs = [r(x) for x in t[b+32:b+48] if x]
        print(f" tile {skip + it:3d}: prod got {r(t[b+0])} dep ok {r(t[b+12])}/{r(t[b+13])} last load {r(t[b+1])} | mma start {r(t[b+2])} done {r(t[b+3])} ({len(kbs)} kb, {(kbs[-1]-kbs[0])//max(len(kbs)-1,1) if kbs else 0}/kb) | epi ready {r(t[b+4])} acc_full {r(t[b+5])} handback {r(t[b+6])} | stores {r(t[b+9])} {r(t[b+11])} | pass0 box wait {r(t[b+14])}->{r(t[b+8])} pass1 {r(t[b+15])}->{r(t[b+10])}")
if len(sys.argv) > 5:       # per-k-block detail of one tile of CTA 0: when the A producers saw the slot free, when the MMA issuer saw it full
    it = int(sys.argv[5]) - skip; t = tr[0]; t0 = t[1020]; b = it * 48
    free = [int(x - t0) for x in t[b+16:b+32]]; full = [int(x - t0) for x in t[b+32:b+48]]
    print(f"tile {it + skip}: slot free (producer)  ", free)
    print(f"tile {it + skip}: data full (MMA issuer)", full)
    print("   full - free (load latency)      ", [f - e for f, e in zip(full, free)])
    print("   d(full) per k-block             ", [full[i + 1] - full[i] for i in range(15)])
