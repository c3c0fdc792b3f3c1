set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -rfEs -s > gpurun_out/r2b_gputests.log 2>&1
grep -n "relerr Y\|^FAILED\|passed\|failed" gpurun_out/r2b_gputests.log | cut -c1-220 | tail -60
python tools/debug_substeps.py 2>&1 | tail -20
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err
cut -c1-300 gpurun_out/r2b_bench.json
timeout 900 python tools/ip2_error_probe.py > gpurun_out/r2b_ip2_probe.jsonl 2> gpurun_out/r2b_ip2_probe.err
cat gpurun_out/r2b_ip2_probe.jsonl
