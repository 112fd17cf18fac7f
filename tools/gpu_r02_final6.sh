mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/r02_bench_z_base.json 2> gpurun_out/r02_bench_z_base.err; python -c "
import json; d=json.loads(open('gpurun_out/r02_bench_z_base.json').read().strip().splitlines()[-1]); print('train', d['steps'], d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['frac'], d['step_tensor_frac'], d['clocks'], d['cpu_baseline']['value'])"
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
