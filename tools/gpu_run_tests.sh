cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r1_gputests_final.log 2>&1
tail -4 gpurun_out/r1_gputests_final.log; grep -n "^E \|^FAILED" gpurun_out/r1_gputests_final.log | head -8
python tools/bench_configs.py --steps 10 --only "ISS" | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print(d['config'][:50], d['ms_per_step'], d['hbm_frac'], d['kernels_ms_per_step'])"
