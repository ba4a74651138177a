timeout 80 python tools/bench_next_rows.py > gpurun_out/next_rows.json 2> gpurun_out/next_rows.err; echo "next rows rc=$?"; tail -2 gpurun_out/next_rows.err; python -c "
import json; d=json.load(open('gpurun_out/next_rows.json'))
for k,v in d.items(): print(k, v.get('ms'))"
timeout 70 python tools/bench_churn.py > gpurun_out/churn_final.json 2> gpurun_out/churn.err; echo "churn rc=$?"; tail -3 gpurun_out/churn.err; python -c "
import json; d=json.load(open('gpurun_out/churn_final.json')); print(d['host_ms_median'], d['pool_replay_alone_ms_median'], d['device_update_ms_median'])"
