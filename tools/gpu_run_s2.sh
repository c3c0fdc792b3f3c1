cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r1s2_gputests.log 2>&1
tail -4 gpurun_out/r1s2_gputests.log; grep -n "^E " gpurun_out/r1s2_gputests.log | head -5
