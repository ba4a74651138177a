for mode in peer nccl; do
SCGPU_GATHER=$mode timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 3 --no-e2e > gpurun_out/bench_n8_$mode.json 2> gpurun_out/bench_n8.err; python -c "
import json;d=json.loads(open('gpurun_out/bench_n8_$mode.json').read().strip().splitlines()[-1]);print('$mode',d['value'],d['ms_per_step'])"
done
tail -3 gpurun_out/bench_n8.err
