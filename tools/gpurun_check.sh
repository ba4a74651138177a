# what a `gpurun -- bash tools/gpurun_check.sh` call of round 2 ran after a kernel change: GPU parity suite, one bench line, (optionally) an ncu metric pass
T=${TAG:-a}
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/t_$T.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/t_$T.log
timeout 300 python bench.py --no-cpu-baseline --no-e2e --churn-frames 30 > gpurun_out/b_$T.json 2> gpurun_out/b_$T.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/b_$T.json').read().strip().splitlines()[-1])
print('ms_per_step',d['ms_per_step'],'kernel',d['roofline']['kernel_ms_avg'],'frac',d['roofline']['frac'],'partial',d.get('partial_dirty',{}).get('kernel_ms_avg'),'clean',d.get('clean_frame',{}).get('kernel_ms_avg'), d['visible_per_view'])
c=d['churn']; print('churn fused', c['fused_kernel_ms']['first10_mean'], c['fused_kernel_ms']['last10_mean'], c['fused_kernel_ms']['series'][:3], c['fused_kernel_ms']['series'][-3:], 'update', c['device_update_ms_median'], 'frame', c['e2e_frame_ms_median'], c['calls_ms_median'])
PY
