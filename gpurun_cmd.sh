timeout 100 python -m pytest tests -m gpu -x -q > gpurun_out/tests_full.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/tests_full.log
