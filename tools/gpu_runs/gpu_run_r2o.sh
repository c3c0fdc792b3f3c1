cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
B="python bench.py --no-cpu-baseline --no-e2e"
show() { python -c "
import sys,json
l=open('$1').read().strip().splitlines()[-1] if open('$1').read().strip() else ''
try:
    d=json.loads(l); print('ms/step %.4f frac %.3f ok=%s'%(d['ms_per_step'],d['roofline']['frac'],d['state_after_timed_steps_ok']), d['roofline']['kernels_ms_per_step'])
except Exception as e: print('NOJSON', l[:300])
"; }
echo "== full GPU suite, SSB_COV_MMA=2 (rn splits everywhere, 3-way phi in the tensor-core covariance)"
SSB_COV_MMA=2 timeout 900 python -m pytest tests -m gpu -q -s -rfE 2>&1 > gpurun_out/r2o_gputests.log
grep -E "^GaussILRMA|^Aux|^FastGauss|passed|failed|^FAILED|^ERROR" gpurun_out/r2o_gputests.log | cut -c1-200 | tail -45
echo "== bench config 4 SSB_COV_MMA=2"
SSB_COV_MMA=2 timeout 200 $B --config 4 --steps 5 --warmup 3 2>gpurun_out/r2o_c4.err > gpurun_out/r2o_c4.json; show gpurun_out/r2o_c4.json
echo "== bench config 2 default"
timeout 150 $B --steps 20 --warmup 3 2>/dev/null > gpurun_out/r2o_c2.json; show gpurun_out/r2o_c2.json
for cs in "8 4" "8 8" "16 2" "32 2" "16 8" "4 8"; do
  set -- $cs
  echo "== bench config 2 chunk=$1 streams=$2"
  timeout 150 $B --steps 20 --warmup 3 --chunk $1 --streams $2 2>/dev/null > gpurun_out/r2o_c2_$1_$2.json; show gpurun_out/r2o_c2_$1_$2.json | cut -c1-40
done
echo "== bench N=4 IP (config-2 shape)"
timeout 150 $B --steps 20 --warmup 3 --sources 4 2>/dev/null > gpurun_out/r2o_n4.json; show gpurun_out/r2o_n4.json
