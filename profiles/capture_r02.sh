#!/bin/bash
# Round-2 profile capture (run under gpurun from the repo root).  Every ncu pass follows the same command
# without ncu (&&); bench values are never taken from a profiled run.  Outputs land in gpurun_out/ and are
# summarised into profiles/ by profiles/summarize.py / summarize_kernel.py.
set -u
O=gpurun_out
# headline bench command: launch list + full capture of the dominant kernel (filter scan) + the select kernel
B="python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e --no-extra --no-anchor"
$B > $O/r02_b4096.json 2> $O/r02_b4096.err && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r02_launches_bench_q4096.csv $B > $O/ncu_a.log 2>&1
$B > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:scan_tc -s 7 -c 1 -o $O/r02_scan_filter_q4096 $B > $O/ncu_b.log 2>&1
B="python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e --no-extra --no-anchor --batch 64"
$B > $O/r02_b64.json 2> $O/r02_b64.err && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r02_launches_bench_q64.csv $B > $O/ncu_c.log 2>&1
# fused tower (26 x 10M-row tables, batch 65536)
T="python tests/bench_extra.py tower"
$T > $O/r02_tower.jsonl 2>&1 && ncu --set full --clock-control none --import-source on -k regex:tower_fused -s 5 -c 1 -o $O/r02_tower_fused $T > $O/ncu_d.log 2>&1
# IVF-Flat filter scan (fused path) and IVF-PQ scan, 10M x 256, nlist 4096, nprobe 32, Q = 4096
I="python tests/prof_ivf.py IVF 4096 10000000 4096 2"
$I > $O/r02_ivf.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:ivf_scan_kernel -s 5 -c 1 -o $O/r02_ivf_filter_scan $I > $O/ncu_e.log 2>&1
# CTA-pair scan (option), Q = 4096
P="python tests/prof_variants.py 4096 1000000 3 pair_scan=1"
$P > $O/r02_pair.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:scan_pair -s 3 -c 1 -o $O/r02_scan_pair_q4096 $P > $O/ncu_f.log 2>&1
# Stage-2 ranker launch list (500 rows)
R="python tests/bench_extra.py ranker"
$R > $O/r02_ranker.jsonl 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv -c 400 --log-file $O/r02_launches_ranker.csv $R > $O/ncu_g.log 2>&1
ls -la $O/*.ncu-rep
