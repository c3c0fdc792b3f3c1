set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/r1s2_gputests.log
cat gpurun_out/r1s2_gputests.log
for g in 4 2; do
SSB_COV_G=$g timeout 900 python tools/bench_configs.py --steps 10 --only "N=8" > gpurun_out/r1s2_configs_n8.jsonl 2> gpurun_out/r1s2_configs.err
python - <<PY
import json
for l in open('gpurun_out/r1s2_configs_n8.jsonl'):
    d=json.loads(l); print('G=$g', d['config'][:60], d['ms_per_step'], d['hbm_frac'], d['kernels_ms_per_step'])
PY
done
