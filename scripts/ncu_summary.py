#!/usr/bin/env python
"""Key counters of one ncu report (ncu -i REP --page raw --csv): the lines quoted in profiles/*.txt.
usage: ncu_summary.py report.ncu-rep [kmers_per_launch]"""
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__thread_inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "sm__icc_request_hit_rate.pct", "gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "launch__shared_mem_per_block_dynamic",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum",
    "sm__cycles_elapsed.avg", "sm__cycles_active.avg",
]


def main():
    rep = sys.argv[1]
    nk = float(sys.argv[2]) if len(sys.argv) > 2 else 0.0
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
    print("# kernel:", d.get("Kernel Name", ("?",))[0])
    for k in KEYS:
        if k in d:
            print("%-86s %-16s %s" % (k, d[k][1], d[k][0]))
    for h in hdr:
        if h.startswith("smsp__average_warp") and "issue_stalled" in h and h.endswith("_per_warp_active.pct") is False and h.endswith(".ratio"):
            try:
                v = float(d[h][0])
            except ValueError:
                continue
            if v >= 0.15:
                print("stall:%-60s %.3f" % (h.split("issue_stalled_")[1].replace("_per_warp_active.ratio", "").replace(".ratio", ""), v))
    if nk:
        def f(k):
            return float(d[k][0].replace(",", "")) if k in d else float("nan")
        scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}
        rd = f("dram__bytes_read.sum") * scale.get(d["dram__bytes_read.sum"][1], 1.0)
        wr = f("dram__bytes_write.sum") * scale.get(d["dram__bytes_write.sum"][1], 1.0)
        print("# per k-mer: %.2f warp-instr, %.1f thread-instr, %.2f DRAM bytes (read %.3g + write %.3g per launch)"
              % (f("smsp__inst_executed.sum") / nk, f("smsp__thread_inst_executed.sum") / nk, (rd + wr) / nk, rd, wr))


if __name__ == "__main__":
    main()
