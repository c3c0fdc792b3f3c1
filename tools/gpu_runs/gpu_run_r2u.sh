cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
echo "== full GPU suite"
timeout 900 python -m pytest tests -m gpu -q -rfE 2>&1 | tail -4 | tee gpurun_out/r2_gputests_final.log
echo "== compute-sanitizer memcheck on the round-2 kernels (small cases)"
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_gpu_parity.py -q -x -k "tensor_core_covariance_n4 or (fused_tensor_core and (8-17-64 or 8-9-48 or 8-21-96)) or stft or linalg_operators or (fast_gauss_mnmf and batched) or mnmf_ip1" > gpurun_out/r2_compute_sanitizer_memcheck.log 2>&1
tail -4 gpurun_out/r2_compute_sanitizer_memcheck.log
echo "== ncu --set full: kc_cov_mma8 (N = 8, config-2 shape)"
timeout 400 ncu --set full --import-source on --clock-control none -k regex:'kc_cov_mma8' -s 1 -c 1 -f -o gpurun_out/r2_ncu_cov_mma8 python tools/ncu_target.py --steps 2 --sources 8 > /dev/null 2>&1
echo "== ncu --set full: kf_mnmf_update (config 5)"
timeout 400 ncu --set full --import-source on --clock-control none -k regex:'kf_mnmf_update' -s 2 -c 2 -f -o gpurun_out/r2_ncu_mnmf_update python tools/ncu_target.py --config 5 --steps 2 > /dev/null 2>&1
ls -la gpurun_out | grep -E "r2_ncu_cov_mma8|r2_ncu_mnmf_update"
