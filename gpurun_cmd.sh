T=${TAG:-a}
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/t_$T.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/t_$T.log
timeout 300 python bench.py --no-cpu-baseline --churn-frames 100 > gpurun_out/b_$T.json 2> gpurun_out/b_$T.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/b_$T.json').read().strip().splitlines()[-1])
print('ms_per_step',d['ms_per_step'],'kernel',d['roofline']['kernel_ms_avg'],'frac',d['roofline']['frac'],'partial',d.get('partial_dirty',{}).get('kernel_ms_avg'),'clean',d.get('clean_frame',{}).get('kernel_ms_avg'), d['visible_per_view'])
print('e2e', d['e2e']['ms_per_step'], {k:v['ms_per_step'] for k,v in d['e2e']['variants'].items()})
c=d.get('churn',{}); print('churn', {k:c[k] for k in c if not isinstance(c[k],(list,dict))})
PY
timeout 200 ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum -k regex:k_update_win$ -s 2 -c 1 --clock-control none --csv --log-file gpurun_out/ncu_$T.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-churn --no-partial > gpurun_out/ncu_$T.log 2>&1; echo "ncu rc=$?"; tail -6 gpurun_out/ncu_$T.csv | cut -d, -f13-
