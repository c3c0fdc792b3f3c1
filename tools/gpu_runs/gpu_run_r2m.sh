cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 500 python -m pytest tests -m gpu -q -s -k "8-IP or n8" 2>&1 | grep -E "relerr|passed|failed|Error|assert" | cut -c1-220 | tail -12
echo "== bench N=8"
timeout 150 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --sources 8 2>/dev/null | tail -1 | python -c "
import sys,json
l=sys.stdin.read().strip()
d=json.loads(l); print('ms/step %.4f frac %.3f ok=%s'%(d['ms_per_step'],d['roofline']['frac'],d['state_after_timed_steps_ok']), d['roofline']['kernels_ms_per_step'])
"
