#!/usr/bin/env python
"""Copy the round-2 evidence out of gpurun_out/ (scratch) into profiles/ (tracked): bench lines,
sweeps, the ncu launch list, and a trimmed table of the `ncu --set full` captures."""
import csv
import json
import shutil
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
OUT, PROF = ROOT / "gpurun_out", ROOT / "profiles"

COPIES = {  # profiles name -> candidates under gpurun_out (first that exists wins)
    "r02_bench_n1.json": ["f1_bench.json"],
    "r02_bench_n1_steps20.json": ["f1_bench_steps20.json"],
    "r02_bench_reference_arm.json": ["f1_bench_ref.json"],
    "r02_bench_sweep.json": ["f1_bench_sweep.json"],
    "r02_bench_group.json": ["f1_bench_group.json"],
    "r02_bench_single_process.json": ["f1_bench_single.json"],
    "r02_bench_bank_fused.json": ["f1_bench_bank_fused.json"],
    "r02_bench_bank_fused_graph.json": ["f1_bench_bank_fused_graph.json"],
    "r02_bench_bank_fused_graph_external.json": ["f1_bench_bank_fused_graph_external.json"],
    "r02_bench_bank_two_calls.json": ["f1_bench_bank_two_calls.json"],
    "r02_bench_n2.json": ["m2_bench.json"],
    "r02_bench_n2_single_process.json": ["m2_bench_single.json"],
    "r02_bench_n2_reference_arm.json": ["m2_bench_ref.json"],
    "r02_bench_n4.json": ["m4_bench.json"],
    "r02_bench_n8.json": ["m8_bench.json"],
    "r02_bench_n8_single_process.json": ["m8_bench_single.json"],
    "r02_bench_n8_reference_arm.json": ["m8_bench_ref.json"],
    "r02_bench_n8_sweep.json": ["m8_bench_sweep.json"],
    "r02_launches.csv": ["f1_launches.csv"],
    "r02_sweep_host_path_and_plugin_pairs.json": ["f1_sweep_host_plugin.json", "s3_sweep.json"],
    "r02_sweep_bank_and_extensions.json": ["f1_sweep_bank_ext.json"],
    "r02_sweep_batched_loopback.json": ["s4_sweep.json"],
    "r02_sweep_pipeline_modes.json": ["s3_sweep.json"],
    "r02_probe_batch_per_call.log": ["s4_probe_batch.log"],
    "r02_probe_batch_with_default_mempool.log": ["probe_batch.log"],
    "r02_probe_batch_with_default_mempool_ncu.csv": ["probe_batch_ncu.csv"],
    "r02_sanitizer_memcheck.log": ["f1_sanitizer_memcheck.log"],
    "r02_sanitizer_racecheck.log": ["f1_sanitizer_racecheck.log"],
    "r02_final_gpu_pytest.log": ["f1_pytest.log"],
}

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes.sum.per_second",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_shared_mem", "smsp__inst_executed.sum",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio"]


def raw_rows(rep: Path):
    text = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(text.splitlines()))
    if len(rows) < 3:
        return [], []
    hdr, units = rows[0], rows[1]
    return [dict(zip(hdr, r)) for r in rows[2:]], dict(zip(hdr, units))


def main():
    PROF.mkdir(exist_ok=True)
    for dst, srcs in COPIES.items():
        for src in srcs:
            if (OUT / src).exists():
                shutil.copyfile(OUT / src, PROF / dst)
                break
        else:
            print("missing:", srcs, file=sys.stderr)
    table, traffic = [], {}
    for rep in sorted(list(OUT.glob("f1_ncu_*.ncu-rep")) + list(OUT.glob("s9a_ncu_*.ncu-rep"))):
        rows, units = raw_rows(rep)
        seen = {}
        for d in rows:
            name = d["Kernel Name"]
            seen[name] = seen.get(name, 0) + 1
            if seen[name] > 2:
                continue
            row = {"capture": rep.name, "kernel": name, "grid": d.get("Grid Size"), "block": d.get("Block Size")}
            for k in KEYS:
                if k in d:
                    row[k + (" [" + units.get(k, "") + "]" if units.get(k) else "")] = d[k]
            table.append(row)
    if table:
        cols = []
        for r in table:
            for k in r:
                if k not in cols:
                    cols.append(k)
        with open(PROF / "r02_ncu_full_metrics.csv", "w", newline="") as f:
            w = csv.DictWriter(f, fieldnames=cols)
            w.writeheader()
            w.writerows(table)
        print(f"r02_ncu_full_metrics.csv: {len(table)} launches from {len(set(r['capture'] for r in table))} captures")


if __name__ == "__main__":
    main()
