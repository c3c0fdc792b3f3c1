cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
B="python bench.py --no-cpu-baseline --no-e2e"
show() { python -c "
import sys,json
t=open('$1').read().strip()
l=t.splitlines()[-1] if t else ''
try:
    d=json.loads(l); print('ms/step %.4f frac %.3f ok=%s'%(d['ms_per_step'],d['roofline']['frac'],d['state_after_timed_steps_ok']), d['roofline']['kernels_ms_per_step'])
except Exception as e: print('NOJSON', l[:300])
"; }
for m in 1 0; do
echo "== bench config 3 SSB_ISS_PREFETCH=$m"
SSB_ISS_PREFETCH=$m timeout 200 $B --config 3 --steps 5 --warmup 3 2>/dev/null > gpurun_out/r2v_c3_$m.json; show gpurun_out/r2v_c3_$m.json
done
echo "== ISS tests"
timeout 300 python -m pytest tests -m gpu -q -k "iss or ISS" 2>&1 | tail -2
echo "== final: config 2 (full line: e2e + reference arm on the host cores)"
timeout 600 python bench.py --steps 20 --warmup 3 2>gpurun_out/r2_bench_c2_final.err > gpurun_out/r2_bench_c2_final.json; cut -c1-1500 gpurun_out/r2_bench_c2_final.json | tail -1; tail -2 gpurun_out/r2_bench_c2_final.err | cut -c1-300
echo "== final: reference arm"
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 2>gpurun_out/r2_bench_c2_reference.err > gpurun_out/r2_bench_c2_reference.json; cut -c1-700 gpurun_out/r2_bench_c2_reference.json | tail -1
