timeout 600 python -m pytest tests/test_gpu_golden_and_scale.py -m gpu -x -q -k "gpus or multi or shard or gather or rank" > gpurun_out/tests_mgpu.log 2>&1; tail -3 gpurun_out/tests_mgpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; tail -c 900 gpurun_out/bench_n2.json | head -c 500; python -c "
import json; d=json.load(open('gpurun_out/bench_n2.json')); print(d['n_gpus'], d['ms_per_step'], d['value']/1e9)"
