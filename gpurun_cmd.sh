GLIBC_TUNABLES=glibc.cpu.hwcaps=-FMA,-AVX2 timeout 120 oracle/_ref/sc_dropin_test 2>&1 | tail -14
timeout 600 python -m pytest tests/test_gpu_golden_and_scale.py -m gpu -x -q -k dropin 2>&1 | tail -4
