#!/usr/bin/env python3
"""Summarise an .ncu-rep (ncu --set full) into a small text file for profiles/.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r1_x.txt ["note"]
"""
import csv
import io
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__inst_executed.sum", "smsp__inst_executed_pipe_fp64.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__cycles_elapsed.avg",
]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    note = sys.argv[3] if len(sys.argv) > 3 else ""
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    lines = [f"# ncu --set full summary of {rep}", f"# {note}", ""]
    for r in rows[2:]:
        lines.append("kernel: " + r[hdr.index("Kernel Name")])
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                lines.append(f"  {w} = {r[i]} {units[i]}")
        try:
            rd = float(r[hdr.index("dram__bytes_read.sum")]); wr = float(r[hdr.index("dram__bytes_write.sum")])
            lines.append(f"  traffic (dram read+write, unit as above) = {rd + wr}")
        except Exception:
            pass
        st = [(h, float(r[hdr.index(h)])) for h in hdr if h.startswith("smsp__average_warp") and h.endswith("_per_issue_active.ratio")]
        st.sort(key=lambda t: -t[1])
        lines.append("  top warp stall reasons (per issue-active): " + ", ".join(
            f"{h.split('issue_stalled_')[-1].replace('_per_issue_active.ratio', '')}={v:.2f}" for h, v in st[:6]))
        lines.append("")
    open(out, "w").write("\n".join(lines))
    print("\n".join(lines))


if __name__ == "__main__":
    main()
