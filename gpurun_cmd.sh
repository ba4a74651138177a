timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "overflow or shell" > gpurun_out/t.log 2>&1; tail -12 gpurun_out/t.log
