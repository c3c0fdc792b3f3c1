cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
echo "== N=8 tests, SSB_COV_MMA=2 (tensor-core covariance also for IP2, flushed accumulators)"
SSB_COV_MMA=2 timeout 600 python -m pytest tests -m gpu -q -x -s -k "(8-IP2) or (8-IP-) or n8 or (fused_tensor_core and 8-)" 2>&1 | grep -E "relerr|passed|failed|Error|assert" | cut -c1-220 | tail -14
for m in 1 2; do
  echo "== bench config 4 SSB_COV_MMA=$m"
  SSB_COV_MMA=$m timeout 200 python bench.py --config 4 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>gpurun_out/r2n_c4_$m.err | tail -1 > gpurun_out/r2n_c4_$m.json
  python -c "
import sys,json
l=open('gpurun_out/r2n_c4_$m.json').read().strip()
try:
    d=json.loads(l); print('ms/step %.4f frac %.3f ok=%s'%(d['ms_per_step'],d['roofline']['frac'],d['state_after_timed_steps_ok']), d['roofline']['kernels_ms_per_step'])
except Exception as e: print('NOJSON', l[:300])
"
  tail -2 gpurun_out/r2n_c4_$m.err | cut -c1-200
done
echo "== bench N=8 IP (config-2 shape)"
timeout 150 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --sources 8 2>/dev/null | tail -1 > gpurun_out/r2n_n8.json
python -c "
import json
d=json.loads(open('gpurun_out/r2n_n8.json').read()); print('ms/step %.4f frac %.3f'%(d['ms_per_step'],d['roofline']['frac']), d['roofline']['kernels_ms_per_step'])
"
