# One gpurun call = the round-end sequence on a fresh box: GPU tests, smoke, the reference arm and our arm of bench.py.
#   gpurun --timeout 1800 -- 'bash tools/gpu_batch.sh'
# 2 GPUs (data-parallel check + N=2 bench):
#   gpurun --gpus 2 -- 'python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/dp_check_gpu.py;
#                       python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2'
mkdir -p gpurun_out
timeout 600 python -m pytest tests/ -x -q -m gpu > gpurun_out/t_gpu.log 2>&1; tail -3 gpurun_out/t_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 2>&1 | tail -1 | cut -c1-400
timeout 600 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
python -c "
import json; d=json.loads(open('gpurun_out/bench_default.json').read().strip().splitlines()[-1])
print('ours', d['ms_per_step'], d['value'], 'e2e', d['e2e']['value'], 'roofline', d['roofline']['achieved'], d['roofline']['frac'], 'launches', d['gpu_launches'], 'clocks', d['clocks'], 'cpu', d['cpu_baseline']['value'])"
