cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
echo "== smoke"
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== full GPU suite"
timeout 900 python -m pytest tests -m gpu -q -rfE 2>&1 | tail -4 | tee gpurun_out/r2_gputests_final.log
echo "== default bench (what the driver runs)"
timeout 600 python bench.py 2>gpurun_out/r2_bench_default.err > gpurun_out/r2_bench_default.json; python -c "
import json
d=json.loads(open('gpurun_out/r2_bench_default.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['roofline']['frac'], d['roofline']['dram_util'], d['cpu_baseline']['value'], d['cpu_baseline']['kind'])
"; tail -1 gpurun_out/r2_bench_default.err | cut -c1-200
