cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r1_bench_v3.json 2> gpurun_out/r1_bench_v3.err
python -c "
import json; d=json.load(open('gpurun_out/r1_bench_v3.json')); print(d['ms_per_step'], d['value'], d['roofline']['frac'], d['e2e']['value'], d['gpu_launches']); print(d['roofline']['kernels_ms_per_step'])"
python -c "import __graft_entry__ as g; g.smoke()"
