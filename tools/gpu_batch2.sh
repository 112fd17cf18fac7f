mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/dp_check_gpu.py 2>&1 | grep -E "dp_check|DP_CHECK|Error|error|assert" | head
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_n2.json
python -c "
import json; d=json.loads(open('gpurun_out/bench_n2.json').read()); print('N=2', d['ms_per_step'], d['value'], d['e2e']['value'], d['loss'])"
echo "=== 1 GPU on the same box"
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_n1.json
python -c "
import json; d=json.loads(open('gpurun_out/bench_n1.json').read()); print('N=1', d['ms_per_step'], d['value'], d['e2e']['value'], d['loss'])"
