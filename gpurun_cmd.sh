timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/tests_full.log 2>&1; tail -4 gpurun_out/tests_full.log
CMD="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -c 24 --csv --log-file gpurun_out/launches_bw.csv $CMD > gpurun_out/ncu_l.log 2>&1
grep -E "k_build_windows|k_resolve_parents|k_flatten" gpurun_out/launches_bw.csv | awk -F'","' '{print $5, $NF}' | cut -c1-80
