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
echo "== MNMF / stft / linalg / n4 tests"
timeout 600 python -m pytest tests -m gpu -q -s -k "mnmf or MNMF or stft or linalg_operators or tensor_core_covariance_n4" 2>&1 | grep -E "relerr.*MNMF|passed|failed|Error|assert|FAILED" | sed 's/^\.*//' | cut -c1-200 | tail -16
for m in 1 0; do
echo "== bench config 5 SSB_MNMF_FUSED=$m"
SSB_MNMF_FUSED=$m timeout 300 $B --config 5 --steps 3 --warmup 3 2>gpurun_out/r2s_c5_$m.err > gpurun_out/r2s_c5_$m.json; show gpurun_out/r2s_c5_$m.json; tail -1 gpurun_out/r2s_c5_$m.err | cut -c1-200
done
