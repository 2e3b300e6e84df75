"""Turns the ncu artefacts a gpurun call brought back (gpurun_out/) into the small text summaries
committed under profiles/.  Usage: python profiles/summarize.py <tag> <launches.csv> <full.ncu-rep> <out.txt>"""
import collections
import csv
import re
import subprocess
import sys


def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        name = re.sub(r"\(.*", "", row["Kernel Name"])[:70]
        try:
            agg.setdefault(name, []).append(float(row["Metric Value"].replace(",", "")))
        except ValueError:
            pass
    return agg


def raw_metrics(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    return dict(zip(rows[0], zip(rows[1], rows[2])))


WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.per_cycle_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum",
        "sm__cycles_elapsed.max",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio"]


def main():
    tag, csv_path, rep, out = sys.argv[1:5]
    with open(out, "w") as f:
        f.write(f"# {tag}\n\n## per-kernel device time, all launches of the run (ncu --metrics gpu__time_duration.sum,\n"
                "## --clock-control none; cold-cache + serialised: compare SHARES, not absolutes)\n")
        agg = launches(csv_path)
        ours = {k: v for k, v in agg.items() if "b2r::" in k}
        step_total = sum(v[-1] for v in ours.values())
        for k, v in agg.items():
            share = f"{100 * v[-1] / step_total:5.1f}% of a step" if k in ours else ""
            f.write(f"{k:72s} n={len(v):3d} last={v[-1] / 1000:10.1f} us  mean={sum(v) / len(v) / 1000:10.1f} us  {share}\n")
        f.write(f"\nsum of one launch of each b2r kernel (one search step): {step_total / 1000:.1f} us\n")
        f.write("\n## ncu --set full, dominant kernel (filter scan), one launch\n")
        m = raw_metrics(rep)
        f.write(f"kernel: {m.get('Kernel Name', ('', ''))[1]}\n")
        for key in WANT:
            if key in m:
                f.write(f"{key:90s} {m[key][0]:12s} {m[key][1]}\n")


if __name__ == "__main__":
    main()
