cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
echo "== tests: tma + N=8 (mma covariance)"
timeout 500 python -m pytest tests -m gpu -q -x -s -k "tma_tile or n8 or 8-IP or (fused_tensor_core and 8-) or test_batched_input" 2>&1 | grep -E "relerr|passed|failed|Error|assert" | cut -c1-220 | tail -12
for m in 0 3; do
  echo "== bench N=2 SSB_TMA=$m"
  SSB_TMA=$m timeout 150 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | tail -1 | python -c "
import sys,json
l=sys.stdin.read().strip()
try:
    d=json.loads(l); print('ms/step %.4f frac %.3f ok=%s'%(d['ms_per_step'],d['roofline']['frac'],d['state_after_timed_steps_ok']), d['roofline']['kernels_ms_per_step'])
except Exception as e: print('NOJSON', l[:300])
"
done
for m in 0 1; do
  echo "== bench N=8 SSB_COV_MMA=$m"
  SSB_COV_MMA=$m timeout 150 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --sources 8 2>gpurun_out/r2l_n8_$m.err | tail -1 | python -c "
import sys,json
l=sys.stdin.read().strip()
try:
    d=json.loads(l); print('ms/step %.4f frac %.3f ok=%s'%(d['ms_per_step'],d['roofline']['frac'],d['state_after_timed_steps_ok']), d['roofline']['kernels_ms_per_step'])
except Exception as e: print('NOJSON', l[:300])
"
  tail -2 gpurun_out/r2l_n8_$m.err | cut -c1-200
done
echo "== config 4 single GPU"
timeout 200 python bench.py --config 4 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>gpurun_out/r2l_c4.err | tail -1 | python -c "
import sys,json
l=sys.stdin.read().strip()
try:
    d=json.loads(l); print('ms/step %.4f frac %.3f ok=%s'%(d['ms_per_step'],d['roofline']['frac'],d['state_after_timed_steps_ok']), d['roofline']['kernels_ms_per_step'])
except Exception as e: print('NOJSON', l[:300])
"
tail -2 gpurun_out/r2l_c4.err | cut -c1-200
