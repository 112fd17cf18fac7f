mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py > gpurun_out/r02_bench_z_base.json 2> gpurun_out/r02_bench_z_base.err; python -c "
import json; d=json.loads(open('gpurun_out/r02_bench_z_base.json').read().strip().splitlines()[-1]); print('train', d['steps'], d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['frac'], d['step_tensor_frac'], d['clocks'], d['cpu_baseline']['value'])"
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_z_reference.json 2>/dev/null; tail -c 300 gpurun_out/r02_bench_z_reference.json
