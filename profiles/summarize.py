#!/usr/bin/env python3
"""Summarise gpurun_out/*.ncu-rep + launches csv into profiles/<tag>_summary.md (run here, no GPU needed)."""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "launch__registers_per_thread", "sm__inst_executed.avg.per_cycle_active",
        "smsp__average_warp_latency_per_inst_issued.ratio", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem"]


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    return {h: (v, u) for h, u, v in zip(hdr, units, vals)}


def main(tag):
    lines = [f"# ncu summary {tag}", "", "Source: `gpurun_out/prof_<kernel>_%s.ncu-rep` (`ncu --set full --clock-control none --import-source on`, one launch per kernel," % tag,
             "workload = config C2b, 8192 frames, kernels launched one after the other; k_exec_cta: config C2a, 64 frames of 128 chained blocks) and `launches_%s.csv` (`--metrics gpu__time_duration.sum`, cold-cache, serialised: compare shares)." % tag, ""]
    for k in ["k_setup", "k_huf", "k_fse", "k_exec", "k_xxh64", "k_exec_cta"]:
        try:
            m = raw(f"gpurun_out/prof_{k}_{tag}.ncu-rep")
        except Exception as e:
            lines.append(f"## {k}: missing ({e})"); continue
        lines.append(f"## {k}")
        lines.append("| metric | value | unit |"); lines.append("|---|---|---|")
        for key in KEYS:
            if key in m:
                lines.append(f"| {key} | {m[key][0]} | {m[key][1]} |")
        lines.append("")
    # launch list
    try:
        txt = open(f"gpurun_out/launches_{tag}.csv").read()
        rows = [r for r in csv.reader(io.StringIO(txt[txt.index('"ID"'):]))]
        hdr = rows[0]
        ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
        agg = {}
        for r in rows[1:]:
            if len(r) <= vi:
                continue
            v = float(r[vi].replace(",", ""))
            u = r[ui]
            ns = v * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(u, 1)
            name = r[ki].split("(")[0]
            a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += ns
        tot = sum(a[1] for n, a in agg.items() if "b200z" in n or n.startswith("k_"))
        lines.append("## launch list (whole bench process under ncu)")
        lines.append("| kernel | launches | total ms | avg ms | share of b200z kernels |"); lines.append("|---|---|---|---|---|")
        for n, a in sorted(agg.items(), key=lambda x: -x[1][1])[:12]:
            sh = f"{a[1] / tot:.3f}" if ("b200z" in n or n.startswith("k_")) and tot else ""
            lines.append(f"| {n} | {a[0]} | {a[1] / 1e6:.3f} | {a[1] / 1e6 / a[0]:.3f} | {sh} |")
    except Exception as e:
        lines.append(f"(launch list missing: {e})")
    open(f"profiles/{tag}_summary.md", "w").write("\n".join(lines) + "\n")
    print("\n".join(lines[-16:]))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "r02")
