cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
python tools/e2e_profile.py 2>&1 | tail -3
python bench.py --steps 20 --warmup 3 --no-cpu-baseline | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['e2e'], d['roofline']['frac'], d['gpu_launches'])"
python tools/e2e_sweep.py 2>&1 | head -8
