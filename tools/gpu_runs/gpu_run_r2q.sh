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
echo "== full GPU suite"
timeout 900 python -m pytest tests -m gpu -q -s -rfE 2>&1 > gpurun_out/r2q_gputests.log
grep -E "passed|failed|^FAILED|^ERROR" gpurun_out/r2q_gputests.log | cut -c1-200 | tail -12
grep -E "relerr" gpurun_out/r2q_gputests.log | sed 's/^\.*//' | grep -E "N=4|MNMF" | cut -c1-150
for m in 0 1; do
echo "== bench N=4 IP SSB_COV_MMA4=$m"
SSB_COV_MMA4=$m timeout 150 $B --steps 20 --warmup 3 --sources 4 2>gpurun_out/r2q_n4_$m.err > gpurun_out/r2q_n4_$m.json; show gpurun_out/r2q_n4_$m.json; tail -1 gpurun_out/r2q_n4_$m.err | cut -c1-200
echo "== bench config 3 SSB_ISS_OCC=$m"
SSB_ISS_OCC=$m timeout 200 $B --config 3 --steps 5 --warmup 3 2>/dev/null > gpurun_out/r2q_c3_$m.json; show gpurun_out/r2q_c3_$m.json
done
echo "== bench config 5"
timeout 300 $B --config 5 --steps 3 --warmup 3 2>gpurun_out/r2q_c5.err > gpurun_out/r2q_c5.json; show gpurun_out/r2q_c5.json; tail -1 gpurun_out/r2q_c5.err | cut -c1-200
echo "== bench config 4"
timeout 200 $B --config 4 --steps 5 --warmup 3 2>/dev/null > gpurun_out/r2q_c4.json; show gpurun_out/r2q_c4.json
