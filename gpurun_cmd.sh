timeout 100 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "churn or forest or empty" > gpurun_out/tests_churn.log 2>&1; echo "parity rc=$?"; tail -2 gpurun_out/tests_churn.log
for t in 1 4 8 16; do
SCGPU_HOST_THREADS=$t timeout 70 python tools/bench_churn.py > gpurun_out/churn_t$t.json 2> gpurun_out/churn.err; echo "churn threads=$t rc=$?"; tail -3 gpurun_out/churn.err; python -c "
import json; d=json.load(open('gpurun_out/churn_t$t.json')); print(d['host_ms_median'], d['pool_replay_alone_ms_median'], d['device_update_ms_median'])"
done
