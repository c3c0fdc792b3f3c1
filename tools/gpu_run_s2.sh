cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r1_bench_v3.json 2> gpurun_out/r1_bench_v3.err
python -c "
import json; d=json.load(open('gpurun_out/r1_bench_v3.json')); print(d['ms_per_step'], d['value'], d['roofline']['frac'], d['roofline']['dominant_kernel'], d['roofline']['dominant_kernel_frac'], d['roofline']['traffic'], d['e2e']['value'], d['gpu_launches']); print(d['roofline']['kernels_ms_per_step']); print(d['cpu_baseline'])"
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 | cut -c1-300
