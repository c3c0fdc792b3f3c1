cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
echo "== NCCL sharding tests (2 GPUs)"
timeout 300 python -m pytest tests/test_multigpu_nccl.py -m gpu -q 2>&1 | tail -2
echo "== config 4 (BASELINE configs[3]) on 2 GPUs, final kernels"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --config 4 --steps 10 --warmup 3 --no-e2e --no-cpu-baseline 2>gpurun_out/r2_bench_c4_2gpu_final.err | tail -1 > gpurun_out/r2_bench_c4_2gpu_final.json
python -c "
import json
d=json.loads(open('gpurun_out/r2_bench_c4_2gpu_final.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','n_gpus','gpu_launches')}, d['roofline']['frac'])
"
echo "== config 2 on 2 GPUs (with e2e)"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu-baseline 2>gpurun_out/r2_bench_c2_2gpu_final.err | tail -1 > gpurun_out/r2_bench_c2_2gpu_final.json
python -c "
import json
d=json.loads(open('gpurun_out/r2_bench_c2_2gpu_final.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, d['e2e']['value'])
"
