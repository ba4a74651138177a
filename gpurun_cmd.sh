CMD="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_update -s 3 -c 1 -o gpurun_out/prof_update6 $CMD > gpurun_out/ncu6.log 2>&1
