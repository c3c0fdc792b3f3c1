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
echo "== MNMF + N=4 tests"
timeout 600 python -m pytest tests -m gpu -q -s -x -k "mnmf or MNMF or (4-IP) or (4- and fused_tensor_core) or batched" 2>&1 | grep -E "relerr.*(MNMF|N=4)|passed|failed|Error|assert" | sed 's/^\.*//' | cut -c1-160 | tail -14
echo "== bench config 5"
timeout 300 $B --config 5 --steps 3 --warmup 3 2>gpurun_out/r2r_c5.err > gpurun_out/r2r_c5.json; show gpurun_out/r2r_c5.json; tail -1 gpurun_out/r2r_c5.err | cut -c1-200
for m in 1 0; do
echo "== bench N=4 IP SSB_COV_MMA4=$m"
SSB_COV_MMA4=$m timeout 150 $B --steps 20 --warmup 3 --sources 4 2>/dev/null > gpurun_out/r2r_n4_$m.json; show gpurun_out/r2r_n4_$m.json
done
