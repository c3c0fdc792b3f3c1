cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for m in 0 3 7; do
  echo "== bench SSB_TMA=$m"
  SSB_TMA=$m timeout 150 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e 2>gpurun_out/r2j_bench_$m.err | tail -1 > gpurun_out/r2j_bench_$m.json
  python -c "
import sys,json
l=open('gpurun_out/r2j_bench_$m.json').read().strip()
try:
    d=json.loads(l); print('ms/step %.4f frac %.3f ok=%s regions %s'%(d['ms_per_step'],d['roofline']['frac'],d['state_after_timed_steps_ok'],[round(x,2) for x in d['timed_regions_ms']]), d['roofline']['kernels_ms_per_step'])
except Exception as e: print('NOJSON', l[:300]); print(open('gpurun_out/r2j_bench_$m.err').read()[-600:])
"
done
for n in 4 8; do
  for m in 0 3; do
  echo "== bench N=$n SSB_TMA=$m"
  SSB_TMA=$m timeout 150 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --sources $n 2>/dev/null | tail -1 | python -c "
import sys,json
l=sys.stdin.read().strip()
try:
    d=json.loads(l); print('ms/step %.4f frac %.3f ok=%s'%(d['ms_per_step'],d['roofline']['frac'],d['state_after_timed_steps_ok']), d['roofline']['kernels_ms_per_step'])
except Exception as e: print('NOJSON', l[:300])
"
  done
done
echo "== config 3"
timeout 200 python bench.py --config 3 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>gpurun_out/r2j_c3.err | tail -1 | python -c "
import sys,json
l=sys.stdin.read().strip()
try:
    d=json.loads(l); print('ms/step %.4f frac %.3f ok=%s'%(d['ms_per_step'],d['roofline']['frac'],d['state_after_timed_steps_ok']), d['roofline']['kernels_ms_per_step'])
except Exception as e: print('NOJSON', l[:300])
"
tail -3 gpurun_out/r2j_c3.err
echo "== full gpu tests"
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
