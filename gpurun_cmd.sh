timeout 600 python scratch/dbg_flat2.py 100 > gpurun_out/dbg_flat3.log 2>&1; tail -3 gpurun_out/dbg_flat3.log | cut -c1-300
CMD="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline"
ncu --set full --clock-control none --import-source on -k regex:k_update_win -s 3 -c 1 -o gpurun_out/r01_k_update_win_v2 $CMD > gpurun_out/ncu_a.log 2>&1
SCGPU_BENCH_WORKLOAD=flat python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/f.json; python -c "
import json;d=json.load(open('gpurun_out/f.json'));print(d['roofline']['kernel_ms_avg'])"
