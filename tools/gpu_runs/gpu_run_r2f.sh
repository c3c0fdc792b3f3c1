set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x -k "baseline_shapes and (IP- or 2-IP2) or fused_iteration or fused_tensor_core or chunked" > gpurun_out/r2f_gputests.log 2>&1
tail -5 gpurun_out/r2f_gputests.log
for cfg in "default:" "fuse0:SSB_FUSE_ITER=0" "tma0:SSB_TMA=0"; do
  name=${cfg%%:*}; envs=${cfg#*:}
  env $envs timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2f_bench_$name.json 2> gpurun_out/r2f_bench_$name.err
  python - <<PY
import json
d=json.load(open("gpurun_out/r2f_bench_$name.json"))
print("$name", "ms/step %.4f"%d["ms_per_step"], "frac %.3f"%d["roofline"]["frac"], d["roofline"]["kernels_ms_per_step"])
PY
done
for n in 4 8; do
  timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --sources $n > gpurun_out/r2f_bench_n$n.json 2> gpurun_out/r2f_bench_n$n.err
  python - <<PY
import json
for f in ("gpurun_out/r2f_bench_n$n.json",):
    d=json.load(open(f)); print(f, "ms/step %.4f"%d["ms_per_step"], "frac %.3f"%d["roofline"]["frac"], d["roofline"]["kernels_ms_per_step"])
PY
done
