CMD="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline"
ncu --set full --clock-control none --import-source on -k 'regex:^k_build_windows$' -c 1 -f -o gpurun_out/r01_k_build_windows $CMD > gpurun_out/ncu_bw.log 2>&1; tail -2 gpurun_out/ncu_bw.log
