cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 500 python tools/bench_configs.py --steps 10 > gpurun_out/r1_bench_configs_h.jsonl 2> gpurun_out/r1_bench_configs_h.err
python - <<PY
import json
for l in open('gpurun_out/r1_bench_configs_h.jsonl'):
    d=json.loads(l); print(d['config'][:58], d['ms_per_step'], d['hbm_frac'])
PY
tail -2 gpurun_out/r1_bench_configs_h.err
