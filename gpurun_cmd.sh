timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/tests_full.log 2>&1; tail -2 gpurun_out/tests_full.log
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 600 gpurun_out/bench_n1.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; tail -c 300 gpurun_out/bench_ref.json
SCGPU_BENCH_WORKLOAD=flat python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/flat5.json 2> gpurun_out/flat5.err; python -c "
import json; d=json.load(open('gpurun_out/flat5.json')); print('flat', d['ms_per_step'], d['roofline']['kernel_ms_avg'], d['roofline']['frac'])"
CMD="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_r01.csv $CMD > gpurun_out/ncu_l.log 2>&1; tail -1 gpurun_out/ncu_l.log | cut -c1-200
ncu --set full --clock-control none --import-source on -k 'regex:^k_update_win$' -s 3 -c 1 -f -o gpurun_out/r01_k_update_win $CMD > gpurun_out/ncu_w.log 2>&1; tail -1 gpurun_out/ncu_w.log
