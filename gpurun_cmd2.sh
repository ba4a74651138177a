set -x
CMD="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_r01.csv $CMD > gpurun_out/ncu_l.log 2>&1
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k 'regex:^k_update_win$' -s 3 -c 1 -f -o gpurun_out/r01_k_update_win $CMD > gpurun_out/ncu_a.log 2>&1
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 200 gpurun_out/bench_n1.json
