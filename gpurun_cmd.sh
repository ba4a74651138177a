timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 50 --warmup 5 --no-e2e > gpurun_out/bench_n8_peer50.json 2> gpurun_out/bench_n8.err
python -c "
import json;d=json.loads(open('gpurun_out/bench_n8_peer50.json').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step']);[print(r) for r in d['per_rank']]"
nvidia-smi --query-gpu=index,clocks.sm,power.draw,temperature.gpu --format=csv
