cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
SSB_CHUNK=64 timeout 600 ncu --set full --import-source on --clock-control none -k regex:'kf_basis_coop|kf_activation_coop|kf_phi_cov|kf_ip1_n2|kf_normalize' -c 5 -f -o gpurun_out/r1_final_coop python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu-baseline > gpurun_out/ncu_final.log 2>&1
tail -2 gpurun_out/ncu_final.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 3 python -m pytest tests/test_gpu_parity.py -q -x -k "fused_tensor_core and (2-37-48 or 8-17-64 or 4-130-96 or 2-33-64)" > gpurun_out/r1_compute_sanitizer_racecheck_v3.log 2>&1
tail -4 gpurun_out/r1_compute_sanitizer_racecheck_v3.log
