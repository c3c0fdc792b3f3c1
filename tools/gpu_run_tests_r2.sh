cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -rfEs > gpurun_out/r2_gputests_full.log 2>&1
tail -12 gpurun_out/r2_gputests_full.log | cut -c1-250
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
