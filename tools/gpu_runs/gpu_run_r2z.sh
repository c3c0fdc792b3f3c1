cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
B="python bench.py --no-cpu-baseline --no-e2e"
show() { python -c "
import sys,json
t=open('$1').read().strip()
l=t.splitlines()[-1] if t else ''
try:
    d=json.loads(l); print('ms/step %.4f frac %.3f ok=%s launches=%d'%(d['ms_per_step'],d['roofline']['frac'],d['state_after_timed_steps_ok'],d['gpu_launches']), d['roofline']['kernels_ms_per_step'])
except Exception as e: print('NOJSON', l[:300])
"; }
for m in 1 0; do
echo "== bench config 5 SSB_MNMF_Z2EMIT=$m"
SSB_MNMF_Z2EMIT=$m timeout 300 $B --config 5 --steps 5 --warmup 3 2>gpurun_out/r2z_c5_$m.err > gpurun_out/r2z_c5_$m.json; show gpurun_out/r2z_c5_$m.json; tail -1 gpurun_out/r2z_c5_$m.err | cut -c1-200
done
