#!/bin/bash
# Round-1 profile capture (run under gpurun from the repo root).  Every ncu run is preceded by the
# same command without ncu (&&).  Outputs land in gpurun_out/ and are summarised by profiles/summarize.py.
set -u
O=gpurun_out
B="python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e"
$B > $O/r01_b4096.json 2> $O/r01_b4096.err && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r01_launches_bench_q4096.csv $B > $O/ncu1.log 2>&1
$B > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:scan_tc -s 7 -c 1 -o $O/r01_scan_filter_q4096 $B > $O/ncu2.log 2>&1
B="python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e --batch 64"
$B > $O/r01_b64.json 2> $O/r01_b64.err && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r01_launches_bench_q64.csv $B > $O/ncu3.log 2>&1
$B > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:scan_tc -s 7 -c 1 -o $O/r01_scan_filter_q64 $B > $O/ncu4.log 2>&1
# select kernel at Q=4096 (second-largest share of the step)
B="python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e"
$B > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:select_rescore -s 4 -c 1 -o $O/r01_select_q4096 $B > $O/ncu5.log 2>&1
# tower: gather (HBM-bound random 64-byte rows) and the first GEMM (tcgen05)
T="python tests/bench_extra.py tower"
$T > $O/r01_tower.jsonl 2>&1 && ncu --set full --clock-control none --import-source on -k regex:gather_concat -s 3 -c 1 -o $O/r01_gather_concat $T > $O/ncu6.log 2>&1
$T > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:gemm_bias_act -s 9 -c 1 -o $O/r01_tower_gemm $T > $O/ncu7.log 2>&1
# IVF list scan and IVF-PQ ADC scan (4M vectors, nlist 2048, nprobe 32, Q=4096)
I="python tests/prof_ivf.py IVF 4096 4000000 2048 2"
$I > $O/r01_ivf.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:ivf_scan -s 2 -c 1 -o $O/r01_ivf_scan $I > $O/ncu8.log 2>&1
I="python tests/prof_ivf.py IVFPQ 4096 4000000 2048 2"
$I > $O/r01_ivfpq.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:ivfpq_scan -s 2 -c 1 -o $O/r01_ivfpq_scan $I > $O/ncu9.log 2>&1
ls -la $O/*.ncu-rep
