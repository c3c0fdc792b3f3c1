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
echo "== MNMF tests"
timeout 400 python -m pytest tests -m gpu -q -s -k "mnmf or MNMF" 2>&1 | grep -E "relerr.*MNMF|passed|failed|Error|assert|FAILED" | sed 's/^\.*//' | cut -c1-200 | tail -8
echo "== bench config 5"
timeout 300 $B --config 5 --steps 3 --warmup 3 2>gpurun_out/r2x_c5.err > gpurun_out/r2x_c5.json; show gpurun_out/r2x_c5.json; tail -1 gpurun_out/r2x_c5.err | cut -c1-200
