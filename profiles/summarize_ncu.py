#!/usr/bin/env python
"""Summarise .ncu-rep captures (ncu --set full) into the text files committed under profiles/.
    python profiles/summarize_ncu.py gpurun_out/r1_k4_v3.ncu-rep profiles/r1_k4_v3_ncu_summary.txt "header text"
Reads the report with `ncu -i ... --page raw --csv` (works without a GPU)."""
import csv
import io
import subprocess
import sys

KEEP = (
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_active",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.avg",
    "sm__cycles_elapsed.avg.per_second", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_fp64.sum",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
)


def main():
    rep, out, header = sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else ""
    if rep.endswith(".csv"):
        raw = open(rep).read()          # already exported on the GPU box (profiles/capture_r1.sh)
    else:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    with open(out, "w") as fh:
        for ln in header.split("\\n"):
            fh.write("# " + ln + "\n")
        for r in rows[2:]:
            name = r[hdr.index("Kernel Name")]
            fh.write(f"\n== {name}  grid {r[hdr.index('Grid Size')]} block {r[hdr.index('Block Size')]}\n")
            for i, h in enumerate(hdr):
                if h in KEEP or "issue_stalled" in h and h.endswith("per_issue_active.ratio"):
                    fh.write(f"{h:86s}{units[i]:17s}{r[i]}\n")


if __name__ == "__main__":
    main()
