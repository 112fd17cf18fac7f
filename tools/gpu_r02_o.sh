mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l); echo "gpus: $N"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 tools/dp_check_gpu.py 2>&1 | grep -E "dp_check|DP_CHECK|Error|error" | head -5
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29552 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_o_n$N.json 2> gpurun_out/r02_bench_o_n$N.err; python -c "
import json; d=json.loads(open('gpurun_out/r02_bench_o_n$N.json').read().strip().splitlines()[-1]); print('N=$N', d['ms_per_step'], d['value'], 'e2e', d['e2e']['ms_per_step'], 'loss', d['loss'])" || tail -20 gpurun_out/r02_bench_o_n$N.err
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('N=1 same box', d['ms_per_step'], d['value'])"
