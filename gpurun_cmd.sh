timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/tests_full.log 2>&1; tail -2 gpurun_out/tests_full.log
python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/b_pdl.json 2> gpurun_out/b_pdl.err; wc -l gpurun_out/b_pdl.json; python -c "
import json; d=json.load(open('gpurun_out/b_pdl.json')); print(d['ms_per_step'], d['roofline']['kernel_ms_avg'], d['roofline']['update_ms_avg'], d['roofline']['frac'])"
