set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 > gpurun_out/r1s2_gputests.log
cat gpurun_out/r1s2_gputests.log
for sh in 1 0; do
SSB_ACT_SHAPE=$sh timeout 900 python tools/bench_configs.py --steps 20 --only "c2 GaussILRMA-IP" > gpurun_out/r1s2_configs_act.jsonl 2> gpurun_out/r1s2_configs.err
python - <<PY
import json
for l in open('gpurun_out/r1s2_configs_act.jsonl'):
    d=json.loads(l); print('shape=$sh', d['config'][:60], d['ms_per_step'], d['hbm_frac'], d['kernels_ms_per_step'])
PY
done
