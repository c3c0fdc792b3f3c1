cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
echo "== gpu tests (tma + host pipeline + chunked)"
timeout 400 python -m pytest tests -m gpu -q -x -k "tma_tile or host_tensor or chunked or cuda_tensor_io or singular" 2>&1 | tail -4
echo "== bench default with e2e"
timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>gpurun_out/r2k_bench.err | tail -1 > gpurun_out/r2k_bench.json
python -c "
import json
d=json.load(open('gpurun_out/r2k_bench.json')); print('ms/step %.4f frac %.3f e2e %.0f (%s) ok=%s'%(d['ms_per_step'],d['roofline']['frac'],d['e2e']['value'],d['e2e']['calls_s'],d['state_after_timed_steps_ok']), d['roofline']['kernels_ms_per_step'])
" || tail -5 gpurun_out/r2k_bench.err
echo "== e2e sweep"
timeout 300 python tools/e2e_sweep.py 2>&1 | tail -9
