mkdir -p gpurun_out
nvidia-smi -L | head -3
timeout 900 python -m pytest tests/test_e2e_more_gpu.py -m gpu -q -x -p no:cacheprovider -k "data_parallel or second_device" 2>&1 | tail -8
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_z_n2.json 2> gpurun_out/r02_bench_z_n2.err; python -c "
import json; d=json.loads(open('gpurun_out/r02_bench_z_n2.json').read().strip().splitlines()[-1]); print('N=2', d['ms_per_step'], d['value'], 'e2e', d['e2e']['ms_per_step'], 'loss', d['loss'])" || tail -20 gpurun_out/r02_bench_z_n2.err
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('N=1 same box', d['ms_per_step'], d['value'])"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29545 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 2>/dev/null | tail -1 | cut -c1-300
