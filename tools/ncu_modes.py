"""Drives a few chunks of one numeric mode so that ncu can capture its kernels:
    ncu --set full --clock-control none -k regex:i8_layer -s 8 -c 4 -o gpurun_out/i8 python tools/ncu_modes.py int8 4096 16
    ncu --set full --clock-control none -k regex:tc_fused -s 2 -c 1 -o gpurun_out/fp32 python tools/ncu_modes.py fp32 256 64"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import koala_b200 as kb
from koala_b200 import spec
precision, n, frames = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
m = "gpurun_out/r.kpv"; os.makedirs("gpurun_out", exist_ok=True); spec.save_model(m, spec.random_model())
eng = kb.BatchKoala(n, model_path=m, precision=precision)
pcm = torch.from_numpy((np.random.default_rng(0).standard_normal((frames, n, 256)) * 2000).astype(np.int16)).cuda()
for _ in range(4):
    eng.process(pcm, time_major=True)
torch.cuda.synchronize()
