cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 240 python -m pytest tests -m gpu -x -q > gpurun_out/r1s3_gputests.log 2>&1
tail -5 gpurun_out/r1s3_gputests.log
timeout 120 python tools/fuse_iter_ab.py > gpurun_out/r1s3_fuse_ab.jsonl 2> gpurun_out/r1s3_fuse_ab.err
cat gpurun_out/r1s3_fuse_ab.jsonl; tail -3 gpurun_out/r1s3_fuse_ab.err
