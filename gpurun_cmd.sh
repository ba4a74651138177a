timeout 600 python -m pytest tests/test_gpu_golden_and_scale.py -m gpu -x -q -k "nccl" > gpurun_out/tests_nccl.log 2>&1; tail -25 gpurun_out/tests_nccl.log
for mode in peer nccl; do
SCGPU_GATHER=$mode timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 --no-e2e > gpurun_out/bench_n2_$mode.json 2> gpurun_out/bench_n2.err; python -c "
import json;d=json.loads(open('gpurun_out/bench_n2_$mode.json').read().strip().splitlines()[-1]);print('$mode',d['value'],d['ms_per_step'])"
done
