"""Selected metrics of an `ncu --set full` report as a small JSON (one entry per captured launch).

    python tools/ncu_extract.py gpurun_out/prof.ncu-rep profiles/r02_ncu_<name>.json
"""
import csv
import io
import json
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "sm__cycles_active.avg", "sm__cycles_elapsed.avg",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
    "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "TPC.TriageCompute.sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__cluster_dim_x",
    "launch__shared_mem_per_block_dynamic", "smsp__cycles_active.avg",
]
rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
launches = []
for row in rows[2:]:
    entry = {"kernel": row[idx["Kernel Name"]]}
    for k in KEEP:
        if k in idx:
            entry[k] = {"value": row[idx[k]], "unit": units[idx[k]]}
    launches.append(entry)
json.dump({"report": rep, "how": "ncu --set full --clock-control none --import-source on (single capture, tools/gpu_r2_c.sh)",
           "launches": launches}, open(out, "w"), indent=1)
print(out, len(launches), "launches")
