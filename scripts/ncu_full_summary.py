#!/usr/bin/env python
"""Summarise an `ncu --set full` report (.ncu-rep) into a small JSON under profiles/: per captured launch the headline metrics,
stall shares, pipe utilisation and the executed-opcode histogram (top 24).   python scripts/ncu_full_summary.py in.ncu-rep out.json "note" """
import csv
import io
import json
import subprocess
import sys

METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
           "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
           "smsp__warps_eligible.avg.per_cycle_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
           "sm__icc_request_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed_op_branch.sum",
           "smsp__sass_inst_executed_op_shared_ld.sum", "smsp__sass_inst_executed_op_shared_st.sum", "launch__occupancy_limit_registers",
           "launch__occupancy_limit_shared_mem", "smsp__thread_inst_executed_per_inst_executed.ratio"]


def ncu_csv(rep, extra):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"] + extra, capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep, dst = sys.argv[1], sys.argv[2]
    note = sys.argv[3] if len(sys.argv) > 3 else ""
    rows = ncu_csv(rep, [])
    hdr, units = rows[0], rows[1]
    ops = ncu_csv(rep, ["--print-metric-instances", "details", "--metrics", "sass__inst_executed_per_opcode"])
    oi = ops[0].index("sass__inst_executed_per_opcode") if "sass__inst_executed_per_opcode" in ops[0] else None
    launches = []
    for k, r in enumerate(rows[2:]):
        d = {"Kernel Name": r[hdr.index("Kernel Name")]}
        for m in METRICS:
            if m in hdr:
                d[m] = r[hdr.index(m)]
        st = {}
        for i, h in enumerate(hdr):
            if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
                try:
                    st[h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]] = float(r[i])
                except ValueError:
                    pass
        tot = sum(st.values()) or 1.0
        d["stall_pct"] = {a: round(100 * b / tot, 1) for a, b in sorted(st.items(), key=lambda kv: -kv[1])[:9]}
        if oi is not None and k + 2 < len(ops):
            txt = ops[k + 2][oi]
            body = txt[txt.find("(") + 1: txt.rfind(")")]
            hist = []
            for part in body.split(";"):
                if ":" in part:
                    a, b = part.split(":")
                    hist.append((a.strip(), int(b)))
            d["opcode_histogram_top24"] = {a: b for a, b in hist[:24]}
        launches.append(d)
    json.dump({"report": rep + " (ncu --set full --clock-control none --import-source on)", "note": note,
               "units": {m: units[hdr.index(m)] for m in METRICS if m in hdr}, "launches": launches}, open(dst, "w"), indent=1)
    print("wrote", dst, len(launches), "launches")


if __name__ == "__main__":
    main()
