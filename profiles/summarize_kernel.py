"""One-kernel ncu summary: python profiles/summarize_kernel.py <title> <rep> <out.txt>"""
import csv
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.per_cycle_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum", "sm__cycles_elapsed.max",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio"]
title, rep, out = sys.argv[1:4]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
m = dict(zip(rows[0], zip(rows[1], rows[2])))
with open(out, "w") as f:
    f.write(f"# {title}\n# ncu --set full --clock-control none, one launch\nkernel: {m.get('Kernel Name', ('', ''))[1]}\n")
    for k in WANT:
        if k in m:
            f.write(f"{k:90s} {m[k][0]:14s} {m[k][1]}\n")
    t = m.get("gpu__time_duration.sum")
    rd, wr = m.get("dram__bytes_read.sum"), m.get("dram__bytes_write.sum")
    if t and rd and wr:
        def to_bytes(v, u):
            return float(v) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
        def to_s(v, u):
            return float(v) * {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}.get(u, 1e-9)
        tot = to_bytes(rd[1], rd[0]) + to_bytes(wr[1], wr[0])
        f.write(f"derived: dram traffic {tot / 1e6:.1f} MB per launch -> {tot / to_s(t[1], t[0]) / 1e9:.0f} GB/s under ncu "
                f"(6446.9 GB/s measured copy peak)\n")
