timeout 900 python tools/bench_churn.py > gpurun_out/churn.json 2> gpurun_out/churn.err; cat gpurun_out/churn.json; tail -5 gpurun_out/churn.err
