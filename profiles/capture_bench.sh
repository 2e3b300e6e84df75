# Round-1 (final) capture of the bench command's launch list and the dominant kernel's full profile.
# Run under gpurun from the repo root; every ncu run follows the same command without ncu (&&).
set -u
O=gpurun_out
B="python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e"
$B > $O/r01_b4096.json 2> $O/r01_b4096.err && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r01_launches_bench_q4096.csv $B > $O/ncu1.log 2>&1
$B > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:scan_tc -s 7 -c 1 -o $O/r01_scan_filter_q4096 $B > $O/ncu2.log 2>&1
ls -la $O/*.ncu-rep
