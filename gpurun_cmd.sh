timeout 600 python tools/bench_next_rows.py > gpurun_out/next_rows.json 2> gpurun_out/next_rows.err; cat gpurun_out/next_rows.json; tail -5 gpurun_out/next_rows.err
