timeout 600 python -m pytest tests/test_gpu_traffic.py -x -q > gpurun_out/tests_traffic.log 2>&1; tail -15 gpurun_out/tests_traffic.log
timeout 300 python tools/bench_traffic.py > gpurun_out/traffic.json 2> gpurun_out/traffic.err; tail -3 gpurun_out/traffic.err; cat gpurun_out/traffic.json
