"""Time the steps of the bf16 path with a given build of the library (tuning variants built by
`python -m koala_b200._build -DNAME=VALUE -o<path>`): python tools/variant_bench.py <lib.so> [streams] [steps]."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import koala_b200 as kb
from koala_b200 import spec
from ctypes import c_void_p
lib_path = os.path.abspath(sys.argv[1])
n = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 300
m = "gpurun_out/r.kpv"; os.makedirs("gpurun_out", exist_ok=True); spec.save_model(m, spec.random_model())
eng = kb.BatchKoala(n, model_path=m, precision="bf16", library_path=lib_path)
ring = int(os.environ.get("RING", "64"))      # frames per stream resident in HBM: 64 = bench.py (512 MiB in + out, larger than L2)
pcm = torch.from_numpy((np.random.default_rng(0).standard_normal((ring, n, 256)) * 2000).astype(np.int16)).cuda()
out = torch.empty_like(pcm)
st = torch.cuda.Stream(); torch.cuda.set_stream(st)
lib, h = eng._library, eng._handle
fpc = int(os.environ.get("FPC", "16"))        # frames per call (bench.py's headline: 16); the timed unit below is one such call
def step(i):
    off = ((i * fpc) % ring) * n * 512
    rc = lib.pv_koala_batch_process_async_strided(h, pcm.data_ptr() + off, out.data_ptr() + off, fpc, 256, n * 256, c_void_p(st.cuda_stream))
    assert rc == 0
for i in range(20): step(i)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(st)
for i in range(steps): step(i)
e1.record(st); torch.cuda.synchronize()
us = e0.elapsed_time(e1) / (steps * fpc) * 1e3
eng.profile(True)
for i in range(100): step(i)
prof = eng.profile_read(); eng.profile(False)
print(f"{os.path.basename(lib_path):40s} {fpc} frames/call: step {us:7.2f} us | per step: " + " ".join(f"{k} {v[0] / (100 * fpc) * 1e3:6.2f}" for k, v in prof.items() if v[1]))
