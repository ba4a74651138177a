set -x
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/tests_full.log 2>&1; tail -5 gpurun_out/tests_full.log
CMD="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_r01.csv $CMD > gpurun_out/ncu_l.log 2>&1
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_update_win -s 3 -c 1 -o gpurun_out/r01_k_update_win $CMD > gpurun_out/ncu_a.log 2>&1
export SCGPU_BENCH_WORKLOAD=flat
$CMD > gpurun_out/plain3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_update_flat -s 3 -c 1 -o gpurun_out/r01_k_update_flat $CMD > gpurun_out/ncu_b.log 2>&1
unset SCGPU_BENCH_WORKLOAD
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 300 gpurun_out/bench_n1.json
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; tail -c 300 gpurun_out/bench_ref.json
