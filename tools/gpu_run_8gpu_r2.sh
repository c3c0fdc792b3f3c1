cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
nvidia-smi --query-gpu=index,name --format=csv | head -9
echo "== NCCL sharding test (2 GPUs)"
timeout 300 python -m pytest tests/test_multigpu_nccl.py -m gpu -q 2>&1 | tail -3
echo "== pcie probe 8 ranks"
timeout 200 $TR --nproc-per-node 8 --master-port 29541 tools/pcie_probe_multi.py 2>/dev/null | tail -1 | tee gpurun_out/r2_pcie_probe_8gpu.json | cut -c1-900
echo "== config 4 (BASELINE configs[3]) on 8 GPUs"
timeout 500 $TR --nproc-per-node 8 --master-port 29542 bench.py --gpus 8 --config 4 --steps 10 --warmup 3 --no-e2e > gpurun_out/r2_bench_c4_8gpu.json 2> gpurun_out/r2_bench_c4_8gpu.err
tail -1 gpurun_out/r2_bench_c4_8gpu.json | cut -c1-500; tail -2 gpurun_out/r2_bench_c4_8gpu.err | cut -c1-300
echo "== config 4 on 1 GPU"
timeout 300 python bench.py --config 4 --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r2_bench_c4_1gpu.json 2> gpurun_out/r2_bench_c4_1gpu.err
tail -1 gpurun_out/r2_bench_c4_1gpu.json | cut -c1-500; tail -2 gpurun_out/r2_bench_c4_1gpu.err | cut -c1-300
echo "== config 2 on 8 GPUs with e2e"
timeout 400 $TR --nproc-per-node 8 --master-port 29543 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/r2_bench_c2_8gpu.json 2> gpurun_out/r2_bench_c2_8gpu.err
tail -1 gpurun_out/r2_bench_c2_8gpu.json | cut -c1-700; tail -2 gpurun_out/r2_bench_c2_8gpu.err | cut -c1-300
