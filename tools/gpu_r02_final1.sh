mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_z_base.json 2> gpurun_out/r02_bench_z_base.err; tail -c 1200 gpurun_out/r02_bench_z_base.json; tail -3 gpurun_out/r02_bench_z_base.err
timeout 600 python bench.py --model t5-large --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_z_large.json 2> gpurun_out/r02_bench_z_large.err; python -c "
import json; d=json.loads(open('gpurun_out/r02_bench_z_large.json').read().strip().splitlines()[-1]); print('t5-large', d['ms_per_step'], d['value'], d.get('step_tensor_frac'), d['roofline']['frac'], d['e2e'])" || tail -5 gpurun_out/r02_bench_z_large.err
timeout 600 python bench.py --mode decode --steps 3 --warmup 3 > gpurun_out/r02_bench_z_decode.json 2> gpurun_out/r02_bench_z_decode.err; tail -c 700 gpurun_out/r02_bench_z_decode.json; tail -3 gpurun_out/r02_bench_z_decode.err
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_z_reference.json 2>gpurun_out/r02_bench_z_reference.err; tail -c 600 gpurun_out/r02_bench_z_reference.json
python tools/attn_timeline.py > gpurun_out/r02_attn_bwd_timeline.txt 2>&1; tail -5 gpurun_out/r02_attn_bwd_timeline.txt
python tools/attn_timeline.py fwd > gpurun_out/r02_attn_fwd_timeline.txt 2>&1; tail -3 gpurun_out/r02_attn_fwd_timeline.txt
python tools/attn_bwd_cta_life.py cross > gpurun_out/r02_attn_bwd_cta_life_cross.txt 2>&1
