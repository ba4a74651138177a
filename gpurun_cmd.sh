timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "forest_large or churn" > gpurun_out/tests_churn.log 2>&1; tail -25 gpurun_out/tests_churn.log
