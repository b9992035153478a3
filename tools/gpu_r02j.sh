#!/bin/bash
# Round 2: cheap tile decode + fp32 mode on the tensor cores (three bf16 planes): tests, variant timing against the previous build, cfg2.
mkdir -p gpurun_out
L=koala_b200/lib/libpv_koala_b200.so
echo "== smoke";   timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3
echo "== pytest";  timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
echo "== variants (8192 streams bf16, 16 frames per call)"
timeout 200 python tools/variant_bench.py gpurun_lib_HEAD.so 8192 100
timeout 200 python tools/variant_bench.py $L 8192 100
timeout 200 python tools/variant_bench.py gpurun_lib_HEAD.so 8192 100
timeout 200 python tools/variant_bench.py $L 8192 100
echo "== cfg5 / cfg2"
for w in cfg5_128_per_gpu_bf16 cfg2_256_fp32; do
  timeout 300 python bench.py --workload $w --steps 512 --no-cpu-baseline --no-others --e2e-steps 64 2>gpurun_out/err_$w.txt | tail -1 > gpurun_out/bench_r02j_$w.json
  python -c "import json,sys; d=json.load(open('gpurun_out/bench_r02j_$w.json')); print('$w', d['value'], d['ms_per_step'], d['e2e']['value'], d['kernel_ms_per_step'])" || tail -5 gpurun_out/err_$w.txt
done
KOALA_FP32_CUDA_CORES=1 timeout 300 python bench.py --workload cfg2_256_fp32 --steps 512 --no-cpu-baseline --no-others --e2e-steps 64 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('cfg2 CUDA cores', d['value'], d['ms_per_step'])"
echo "== trace"; timeout 100 python tools/gpu_trace.py 8192 gpurun_lib_TRACE.so 16 40 2>&1 | head -24 > gpurun_out/trace_r02j.txt; head -24 gpurun_out/trace_r02j.txt
