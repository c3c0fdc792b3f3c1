set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -5 > gpurun_out/r1s2_gputests.log
cat gpurun_out/r1s2_gputests.log
timeout 900 python tools/bench_configs.py --steps 10 > gpurun_out/r1_bench_configs_g.jsonl 2> gpurun_out/r1s2_configs.err
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r1_bench_v3.json 2> gpurun_out/r1_bench_v3.err
cut -c1-300 gpurun_out/r1_bench_v3.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1_ncu_launches_v3.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/r1s2_ncu_launch.log 2>&1
tail -2 gpurun_out/r1s2_ncu_launch.log | cut -c1-200
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_gpu_parity.py -q -x -k "fused_tensor_core or iss2_seeded or fast_gauss_mnmf_batched or ilrma_ipa or iva_laplace_ipa or full_size and not IP2" > gpurun_out/r1_compute_sanitizer_memcheck_v3.log 2>&1
tail -5 gpurun_out/r1_compute_sanitizer_memcheck_v3.log
