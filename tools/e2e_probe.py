"""PCIe copy rates (pinned host <-> device, contiguous and 2-D with short rows) and the end-to-end rate of one multi-frame
host call for a few chunk sizes (KOALA_HOST_CHUNK) -- development aid for the host ingest path (Engine::process_host)."""
import os, sys, time, subprocess
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
if len(sys.argv) > 1 and sys.argv[1] == "copy":
    h = torch.empty(256 << 20, dtype=torch.uint8).pin_memory(); d = torch.empty_like(h, device="cuda")
    for name, fn in (("H2D", lambda: d.copy_(h, non_blocking=True)), ("D2H", lambda: h.copy_(d, non_blocking=True))):
        fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(5): fn()
        torch.cuda.synchronize(); print(f"{name} contiguous 256 MiB: {5 * 256 / 1024 / (time.perf_counter() - t0):.1f} GiB/s")
    s2 = torch.cuda.Stream()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(5):
        d.copy_(h, non_blocking=True)
        with torch.cuda.stream(s2): h.copy_(d, non_blocking=True)
    torch.cuda.synchronize(); print(f"both directions at once: {2 * 5 * 256 / 1024 / (time.perf_counter() - t0):.1f} GiB/s total")
    sys.exit(0)
import koala_b200 as kb
from koala_b200 import spec
m = "gpurun_out/r.kpv"; os.makedirs("gpurun_out", exist_ok=True); spec.save_model(m, spec.random_model())
n, steps = 8192, int(sys.argv[1]) if len(sys.argv) > 1 else 64
eng = kb.BatchKoala(n, model_path=m, precision="bf16")
h_in = torch.from_numpy((np.random.default_rng(0).standard_normal((n, steps, 256)) * 2000).astype(np.int16)).pin_memory()
h_out = torch.empty_like(h_in).pin_memory()
eng.process(h_in[:, :8].contiguous().pin_memory(), out=torch.empty(n, 8, 256, dtype=torch.int16).pin_memory())
for rep in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter(); eng.process(h_in, out=h_out); dt = time.perf_counter() - t0
print(f"chunk {os.environ.get('KOALA_HOST_CHUNK', 'default')} steps {steps}: {dt * 1e3:.2f} ms, {dt / steps * 1e6:.1f} us/step, {n * steps / dt / 1e6:.1f} M frames/s")
