timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/tests_full.log 2>&1; tail -4 gpurun_out/tests_full.log
python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu-baseline | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ms/step',d['ms_per_step'],'kernel',d['roofline']['kernel_ms_avg'],'value',d['value'])"
CMD="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_bw.csv $CMD > gpurun_out/ncu_l.log 2>&1
grep -E "k_scan_tiles|k_scatter" gpurun_out/launches_bw.csv | tail -4 | awk -F'","' '{print substr($5,1,20), $NF}'
