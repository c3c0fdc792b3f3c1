cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for m in 1 2 4 7; do
  echo "== SSB_TMA=$m"
  SSB_TMA=$m timeout 120 python -m pytest tests -m gpu -q -x -s -k "baseline_shapes and 2-IP-1025 or fused_iteration" 2>&1 | grep -E "relerr|passed|failed|Error" | cut -c1-200 | tail -6
done
echo "== bench tma=1 (basis only)"
SSB_TMA=1 timeout 90 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | cut -c1-200
echo "== bench tma=2"
SSB_TMA=2 timeout 90 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | cut -c1-200
echo "== bench tma=4"
SSB_TMA=4 timeout 90 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | cut -c1-200
