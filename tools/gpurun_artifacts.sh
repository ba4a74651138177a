# the round-end artefact call: full bench line at N = 1 and the ncu launch list that profiles/make_summaries_r02.py reads
timeout 400 python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "bench rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --churn-frames 4 > gpurun_out/ncu_l.log 2>&1; echo "ncu list rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/r02_bench_n1.json').read().strip().splitlines()[-1])
print('ms_per_step',d['ms_per_step'],'value',d['value'],'kernel',d['roofline']['kernel_ms_avg'],'frac',d['roofline']['frac'],'e2e',d['e2e']['ms_per_step'],'cpu',d['cpu_baseline']['value'], 'churn', d['churn']['e2e_frame_ms_median'], d['churn']['fused_kernel_ms']['first10_mean'], d['churn']['fused_kernel_ms']['last10_mean'], d['churn']['device_update_ms_median'])
PY
