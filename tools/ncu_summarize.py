"""Turn the raw ncu exports of tools/gpu_ncu.sh into the small summaries committed under profiles/:

    python tools/ncu_summarize.py gpurun_out/launches.csv gpurun_out/prof_level_raw.csv profiles/r01

writes <prefix>_ncu_launch_list_summary.csv (per kernel: launches, total us, share of GPU time) and
<prefix>_ncu_mlp_level_pair_summary.json (selected `--set full` metrics, one value per captured launch).
"""
import csv
import json
import re
import sys

KEEP = {
    'LTS.TriageCompute.lts__throughput.avg.pct_of_peak_sustained_elapsed',
    'TPC.TriageCompute.sm__cycles_active.avg',
    'TPC.TriageCompute.sm__inst_executed_pipe_alu_realtime.avg.pct_of_peak_sustained_elapsed',
    'TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed',
    'TPC.TriageCompute.sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg',
    'dram__bytes_read.sum',
    'dram__bytes_read.sum.pct_of_peak_sustained_elapsed',
    'dram__bytes_read.sum.per_second',
    'dram__bytes_write.sum',
    'dram__bytes_write.sum.pct_of_peak_sustained_elapsed',
    'dram__bytes_write.sum.per_second',
    'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
    'gpu__time_duration.sum',
    'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
    'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
    'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
    'launch__block_size',
    'launch__cluster_dim_x',
    'launch__grid_size',
    'launch__registers_per_thread',
    'launch__registers_per_thread_allocated',
    'launch__shared_mem_per_block_dynamic',
    'lts__throughput.avg.pct_of_peak_sustained_elapsed',
    'sm__cycles_active.avg',
    'sm__cycles_elapsed.avg',
    'sm__cycles_elapsed.avg.per_second',
    'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_alu.max.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_alu.min.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_fma.max.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_fma.min.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_fma_type_fp16.avg.pct_of_peak_sustained_active',
    'sm__throughput.avg.pct_of_peak_sustained_elapsed',
    'sm__warps_active.avg.pct_of_peak_sustained_active',
    'smsp__inst_executed.sum',
    'smsp__issue_active.avg.pct_of_peak_sustained_active',
}


def launch_list(path, out):
    rows = [r for r in csv.reader(l for l in open(path, errors="replace") if l.startswith('"'))]
    head = rows[0]
    ki, mi, vi = head.index("Kernel Name"), head.index("Metric Name"), head.index("Metric Value")
    ui = head.index("Metric Unit")
    agg = {}
    for r in rows[1:]:
        if r[mi] != "gpu__time_duration.sum":
            continue
        us = float(r[vi].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[ui], 1e-3)
        name = re.sub(r"\(.*", "", r[ki]).replace("mipnerf::", "").replace("(anonymous namespace)::", "").strip()
        n, t = agg.get(name, (0, 0.0))
        agg[name] = (n + 1, t + us)
    total = sum(t for _, t in agg.values())
    with open(out, "w") as f:
        f.write("kernel,launches,total_us,share\n")
        for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{name},{n},{t:.1f},{t / total:.4f}\n")
    print(open(out).read())


def full_capture(path, out):
    rows = [r for r in csv.reader(l for l in open(path, errors="replace") if l.startswith('"'))]
    head, units, data = rows[0], rows[1], rows[2:]
    ki = head.index("Kernel Name")
    metrics = {}
    for j, name in enumerate(head):
        if name in KEEP:
            metrics[name] = {"unit": units[j], "launches": [r[j].replace(",", "") for r in data]}
    json.dump({"kernel": sorted({r[ki] for r in data}), "metrics": metrics}, open(out, "w"), indent=1)
    for k in ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread"):
        if k in metrics:
            print(k, metrics[k])


if __name__ == "__main__":
    lst, raw, prefix = sys.argv[1:4]
    launch_list(lst, prefix + "_ncu_launch_list_summary.csv")
    full_capture(raw, prefix + "_ncu_mlp_level_pair_summary.json")
