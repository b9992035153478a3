"""Does a device->host copy make progress while the step's kernels run?  Times contiguous and 2-D (pitched) D2H copies of
32 MiB alone and concurrently with a loop of device-resident steps -- development aid for Engine::process_host."""
import os, sys, time, ctypes, glob
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import koala_b200 as kb
from koala_b200 import spec
from ctypes import c_void_p, c_size_t, c_int
rt = None
for pat in ("/usr/local/cuda/lib64/libcudart.so*", os.path.join(os.path.dirname(torch.__file__), "lib", "libcudart*.so*")):
    for f in sorted(glob.glob(pat)):
        try: rt = ctypes.CDLL(f); break
        except OSError: pass
    if rt: break
rt.cudaMemcpy2DAsync.argtypes = [c_void_p, c_size_t, c_void_p, c_size_t, c_size_t, c_size_t, c_int, c_void_p]
m = "gpurun_out/r.kpv"; os.makedirs("gpurun_out", exist_ok=True); spec.save_model(m, spec.random_model())
n = 8192
eng = kb.BatchKoala(n, model_path=m, precision="bf16")
pcm = torch.from_numpy((np.random.default_rng(0).standard_normal((16, n, 256)) * 2000).astype(np.int16)).cuda()
out = torch.empty_like(pcm)
comp = torch.cuda.Stream(); copy = torch.cuda.Stream()
lib, h = eng._library, eng._handle
def steps(k):
    for i in range(k):
        off = (i % 16) * n * 512
        lib.pv_koala_batch_process_async(h, pcm.data_ptr() + off, out.data_ptr() + off, 1, 256, c_void_p(comp.cuda_stream))
d = torch.empty(32 << 20, dtype=torch.uint8, device="cuda")
hbuf = torch.empty(520 << 20, dtype=torch.uint8).pin_memory()
rt.cudaMemcpy2DAsync.restype = c_int
def chk(rc):
    assert rc == 0, f'cudaMemcpy2DAsync failed: {rc}'
def d2h_1d():
    with torch.cuda.stream(copy): hbuf[: 32 << 20].copy_(d, non_blocking=True)
def d2h_2d(width, hpitch):
    chk(rt.cudaMemcpy2DAsync(hbuf.data_ptr(), hpitch, d.data_ptr(), width, width, (32 << 20) // width, 2, c_void_p(copy.cuda_stream)))
def h2d_2d(width, hpitch):
    chk(rt.cudaMemcpy2DAsync(d.data_ptr(), width, hbuf.data_ptr(), hpitch, width, (32 << 20) // width, 1, c_void_p(copy.cuda_stream)))
def timed(fn, busy):
    torch.cuda.synchronize()
    if busy: steps(40)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(copy)
    for _ in range(4): fn()
    e1.record(copy)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 4
steps(20); torch.cuda.synchronize()
for name, fn in (("D2H contiguous", d2h_1d), ("D2H 2-D width 4 KB pitch 64 KB", lambda: d2h_2d(4096, 65536)),
                 ("D2H 2-D width 8 KB pitch 64 KB", lambda: d2h_2d(8192, 65536)), ("D2H 2-D width 16 KB pitch 64 KB", lambda: d2h_2d(16384, 65536)), ("D2H 2-D width 32 KB pitch 64 KB", lambda: d2h_2d(32768, 65536)), ("H2D 2-D width 4 KB pitch 64 KB", lambda: h2d_2d(4096, 65536))):
    a, b = timed(fn, False), timed(fn, True)
    print(f"{name:34s}: alone {a:6.3f} ms ({32 / 1024 / a * 1e3:5.1f} GiB/s)   with steps running {b:6.3f} ms ({32 / 1024 / b * 1e3:5.1f} GiB/s)")
