#!/usr/bin/env python
"""Summarise an .ncu-rep: headline metrics per kernel (markdown) and, with --source, the hottest
SASS lines by stall samples.  Usage: tools/ncu_summary.py rep.ncu-rep [--source KERNEL_SUBSTR]"""
import csv
import io
import subprocess
import sys

METRICS = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__block_size", "launch__grid_size",
           "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
           "smsp__thread_inst_executed_per_inst_executed.ratio",
           "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
           "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "smsp__warps_eligible.avg.per_cycle_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
           "sm__icc_request_hit_rate.pct",
           # the L1 data pipe: what bounds the BVH bounce kernels (divergent node / leaf fetches)
           "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
           "l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
           "l1tex__t_output_wavefronts_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_output_wavefronts_pipe_lsu_mem_local_op_st.sum",
           "lts__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed"]
STALLS = "smsp__pcsamp_warps_issue_stalled_"


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]


def main():
    rep = sys.argv[1]
    hdr, units, rows = raw(rep)
    col = {h: i for i, h in enumerate(hdr)}
    names = [r[col["Kernel Name"]][:44] for r in rows]
    print("| metric | " + " | ".join(names) + " |")
    print("|---|" + "---|" * len(names))
    for m in METRICS:
        if m in col:
            print(f"| {m} ({units[col[m]]}) | " + " | ".join(r[col[m]] for r in rows) + " |")
    for h in hdr:
        if h.startswith(STALLS) and "not_issued" not in h:
            vals = [r[col[h]] for r in rows]
            if any(v and float(v.replace(",", "")) > 0 for v in vals):
                print(f"| {h[len('smsp__pcsamp_'):]} | " + " | ".join(vals) + " |")
    if "--source" in sys.argv:
        sub = sys.argv[sys.argv.index("--source") + 1]
        out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + sub],
                             capture_output=True, text=True).stdout
        print(out[:200000])


if __name__ == "__main__":
    main()
