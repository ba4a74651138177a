show() { python - "$1" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], "ms/step %.3f"%d["ms_per_step"], "kernel ms %.3f"%d["roofline"]["kernel_ms_avg"], "frac %.3f"%d["roofline"]["frac"], d["visible_per_view"])
PY
}
B="timeout 300 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu-baseline"
$B > gpurun_out/h.json 2>> gpurun_out/b.err; show gpurun_out/h.json
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/tests_full.log 2>&1; tail -5 gpurun_out/tests_full.log
