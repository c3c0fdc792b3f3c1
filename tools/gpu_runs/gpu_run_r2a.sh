set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
timeout 2400 python -m pytest tests -m gpu -q -rfEs -s > gpurun_out/r2a_gputests.log 2>&1
tail -40 gpurun_out/r2a_gputests.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
cut -c1-400 gpurun_out/r2a_bench.json
timeout 900 python tools/bench_configs.py --steps 10 > gpurun_out/r2a_bench_configs.jsonl 2> gpurun_out/r2a_configs.err
timeout 900 python tools/ip2_error_probe.py > gpurun_out/r2a_ip2_probe.jsonl 2> gpurun_out/r2a_ip2_probe.err
cat gpurun_out/r2a_ip2_probe.jsonl
