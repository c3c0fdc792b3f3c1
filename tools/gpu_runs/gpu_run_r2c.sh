set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -rfEs -s -x > gpurun_out/r2c_gputests.log 2>&1
grep -n "relerr Y\|^FAILED\|passed\|failed\|Error\|error" gpurun_out/r2c_gputests.log | cut -c1-220 | tail -50
for cfg in "default:" "fuse0:SSB_FUSE_ITER=0" "tma0:SSB_TMA=0"; do
  name=${cfg%%:*}; envs=${cfg#*:}
  env $envs timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2c_bench_$name.json 2> gpurun_out/r2c_bench_$name.err
  python - <<PY
import json
d=json.load(open("gpurun_out/r2c_bench_$name.json"))
print("$name", "ms/step %.4f"%d["ms_per_step"], "frac %.3f"%d["roofline"]["frac"], "e2e %.0f"%d["e2e"]["value"], d["roofline"]["kernels_ms_per_step"])
PY
done
for n in 4 8; do
  timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --sources $n > gpurun_out/r2c_bench_n$n.json 2> gpurun_out/r2c_bench_n$n.err
  SSB_TMA=0 timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --sources $n > gpurun_out/r2c_bench_n${n}_tma0.json 2>> gpurun_out/r2c_bench_n$n.err
  python - <<PY
import json
for f in ("gpurun_out/r2c_bench_n$n.json","gpurun_out/r2c_bench_n${n}_tma0.json"):
    d=json.load(open(f)); print(f, "ms/step %.4f"%d["ms_per_step"], "frac %.3f"%d["roofline"]["frac"], d["roofline"]["kernels_ms_per_step"])
PY
done
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_gpu_parity.py -q -x -k "fused_iteration_kernel or (fused_tensor_core and (2-37 or 8-17 or 4-130))" > gpurun_out/r2c_memcheck.log 2>&1
tail -5 gpurun_out/r2c_memcheck.log
