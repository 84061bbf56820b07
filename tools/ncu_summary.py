#!/usr/bin/env python
"""Summarise an .ncu-rep (read on the CPU box): per-kernel duration, DRAM bytes, throughput, stall reasons.

    python tools/ncu_summary.py gpurun_out/full_x.ncu-rep > profiles/r01_x.txt
"""
import csv
import io
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__block_size", "launch__grid_size", "launch__shared_mem_per_block_dynamic",
    "launch__waves_per_multiprocessor", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
    "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_write.sum",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    print(f"# ncu summary of {path} (ncu --set full --clock-control none; per-launch values, cold cache, serialised)")
    for r in rows[2:]:
        print(f"\n## {r[idx['Kernel Name']]}")
        for w in WANT:
            if w in idx and r[idx[w]] != "":
                print(f"{w:80s} {r[idx[w]]} {units[idx[w]]}")
        rd = float(r[idx["dram__bytes_read.sum"]]) * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}[units[idx["dram__bytes_read.sum"]]]
        wr = float(r[idx["dram__bytes_write.sum"]]) * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}[units[idx["dram__bytes_write.sum"]]]
        us = float(r[idx["gpu__time_duration.sum"]]) * {"us": 1, "ms": 1e3, "ns": 1e-3}[units[idx["gpu__time_duration.sum"]]]
        print(f"{'derived: dram traffic per launch (bytes)':80s} {rd + wr:.0f}")
        print(f"{'derived: dram GB/s under ncu':80s} {(rd + wr) / us / 1e3:.1f}")
        print("stall reasons (warps per issue):")
        for h in hdr:
            if "smsp__average_warps_issue_stalled" in h and "per_issue_active" in h:
                v = float(r[idx[h]])
                if v >= 0.2:
                    print(f"    {h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''):28s} {v:.2f}")


if __name__ == "__main__":
    main(sys.argv[1])
