# First measurement of the fused-iteration kernel: does its second pass hit L2?
#   gpurun --timeout 900 -- 'bash tools/gpu_run_fuse_iter_profile.sh'
# Writes gpurun_out/r2_fuse_iter.ncu-rep (read with: ncu -i ... --page raw --csv | grep -E "dram__bytes|lts__t_sector_hit|issue_active"),
# the opt-in parity shapes and a racecheck of the kernel.
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
SSB_TEST_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_gpu_parity.py -q -k "fused_iteration or substeps" > gpurun_out/r2_fuse_iter_tests.log 2>&1
tail -3 gpurun_out/r2_fuse_iter_tests.log
SSB_FUSE_ITER=1 SSB_CHUNK=64 STEPS=3 timeout 600 ncu --set full --import-source on --clock-control none \
  -k regex:'kf_cov_ip1_basis|kf_activation_coop|kf_normalize' -c 4 -f -o gpurun_out/r2_fuse_iter \
  python tools/fuse_iter_ab.py 1 > gpurun_out/r2_fuse_iter_ncu.log 2>&1
tail -2 gpurun_out/r2_fuse_iter_ncu.log
SSB_TEST_EXPERIMENTAL=1 timeout 600 compute-sanitizer --tool racecheck --error-exitcode 3 python -m pytest tests/test_gpu_parity.py -q -x \
  -k "fused_iteration and (37-48 or 20-16 or 33-64)" > gpurun_out/r2_fuse_iter_racecheck.log 2>&1
tail -4 gpurun_out/r2_fuse_iter_racecheck.log
timeout 200 python tools/fuse_iter_ab.py > gpurun_out/r2_fuse_iter_ab.jsonl 2>&1
cat gpurun_out/r2_fuse_iter_ab.jsonl
