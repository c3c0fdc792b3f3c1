set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none -k regex:'kf_basis_coop|kf_activation_coop|kf_phi_cov|kf_vsplit' -c 4 -f -o gpurun_out/r1s2_coop python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu-baseline > gpurun_out/r1s2_ncu.log 2>&1
tail -3 gpurun_out/r1s2_ncu.log
