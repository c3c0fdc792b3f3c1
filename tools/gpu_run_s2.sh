set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x -k "fdica or permutation" 2>&1 | tail -40 > gpurun_out/r1s2_gputests.log
cat gpurun_out/r1s2_gputests.log
