"""Dump the clock64() timeline of one GRU launch (KOALA_TC_TRACE=1) -- development aid for the tcgen05 pipeline."""
import os, sys
import numpy as np
os.environ["KOALA_TC_TRACE"] = sys.argv[2] if len(sys.argv) > 2 else "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import koala_b200 as kb
from koala_b200 import spec
m = "gpurun_out/r.kpv"; os.makedirs("gpurun_out", exist_ok=True); spec.save_model(m, spec.random_model())
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
eng = kb.BatchKoala(n, model_path=m, precision="bf16")
pcm = (np.random.default_rng(0).standard_normal((n, 4, 256)) * 2000).astype(np.int16)
eng.process(pcm)
tr = eng.debug_read("trace", (2, 512), np.int64)
for cta in range(2):
    t = tr[cta]; t0 = t[500]
    print(f"== CTA {cta}: start 0, after setup {t[501]-t0}, roles done {t[502]-t0}, end {t[503]-t0}")
    for it in range(5):
        b = it * 48
        if t[b+4] == 0 and t[b+0] == 0: continue
        r = lambda x: (x - t0) if x else -1
        print(f" tile {it}: prod first {r(t[b+0])} last {r(t[b+1])} | mma buf_free {r(t[b+2])} done_issue {r(t[b+3])} | epi ready {r(t[b+4])} acc_full {r(t[b+5])} pre_arrive {r(t[b+6])} arrived {r(t[b+7])}")
        print("    epi detail: chunk0 loaded", r(t[b+8]), "chunk0 done", r(t[b+9]), "chunk1 loaded", r(t[b+10]), "chunk1 done", r(t[b+11]))
        print("    prod slot-free per kb:", [int(r(x)) for x in t[b+16:b+32]])
        print("    mma  data-full per kb:", [int(r(x)) for x in t[b+32:b+48]])

t = tr[0]; t0 = t[500]
print("tile 1 MMA thread detail per kb: [before wait, after wait, after 4 mma issued, after commit]")
for kb in range(16):
    print("   kb", kb, [int(x - t0) for x in t[256 + kb * 4: 256 + kb * 4 + 4]])
