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
for m in 3 2; do
echo "== bench config 5 SSB_MNMF_OCC=$m"
SSB_MNMF_OCC=$m timeout 300 $B --config 5 --steps 3 --warmup 3 2>/dev/null > gpurun_out/r2w_c5_$m.json; show gpurun_out/r2w_c5_$m.json
done
echo "== MNMF + linalg tests, SSB_MNMF_OCC=3"
SSB_MNMF_OCC=3 timeout 400 python -m pytest tests -m gpu -q -k "mnmf or MNMF or linalg_operators" 2>&1 | tail -3
