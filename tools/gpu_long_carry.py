"""Development aid: GPU vs oracle differences after a long state carry (shipped weights, fixtures)."""
import os, sys, wave
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import koala_b200 as kb
from oracle import OracleBatch, OracleModel
def load(n):
    f = wave.open(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", n + ".wav"))
    return np.frombuffer(f.readframes(f.getnframes()), dtype="<i2").copy()
t, z = load("test"), load("noise")
mixed = np.clip(t.astype(np.int32) + z, -32768, 32767).astype(np.int16)
n, frames = 6, 1400
src = [t, z, mixed, t[5000:], z[::-1].copy(), mixed[20000:]]
pcm = np.stack([np.resize(s, frames * 256) for s in src]).reshape(n, frames, 256)
mp = kb.default_model_path()
for prec in ("bf16", "fp32"):
    eng = kb.BatchKoala(n, model_path=mp, precision=prec)
    out = eng.process(pcm)
    ob = OracleBatch(OracleModel(mp), n, prec); ref = ob.process(pcm, threads=8)
    d = np.abs(out.astype(int) - ref.astype(int))
    mask = eng.debug_read("mask", (n, 256), np.float32); rm = np.stack([ob.stream(s).last_mask for s in range(n)])
    print(prec, "max LSB diff", d.max(), "count>1", int((d > 1).sum()), "| mask max abs", np.abs(mask - rm).max(), "max rel", (np.abs(mask - rm) / np.maximum(rm, 1e-6)).max())
    for l in range(2):
        h = eng.debug_read(f"h{l}", (n, 512), np.float32); rh = np.stack([ob.stream(s).h[l] for s in range(n)])
        print("   h", l, "max abs diff", np.abs(h - rh).max())
