set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
SSB_COOP_COV=1 timeout 600 ncu --set full --import-source on --clock-control none -k regex:'kf_cov_coop' -c 1 -f -o gpurun_out/r1s2_cov4_coop python tools/bench_configs.py --steps 1 --only "GaussILRMA-IP N=4" > gpurun_out/r1s2_ncu.log 2>&1
SSB_COOP_COV=0 timeout 600 ncu --set full --import-source on --clock-control none -k regex:'kf_phi_cov' -c 1 -f -o gpurun_out/r1s2_cov4_old python tools/bench_configs.py --steps 1 --only "GaussILRMA-IP N=4" >> gpurun_out/r1s2_ncu.log 2>&1
SSB_COOP_COV=1 timeout 600 ncu --set full --import-source on --clock-control none -k regex:'kf_cov_coop' -c 1 -f -o gpurun_out/r1s2_cov8_coop python tools/bench_configs.py --steps 1 --only "GaussILRMA-IP N=8" >> gpurun_out/r1s2_ncu.log 2>&1
tail -3 gpurun_out/r1s2_ncu.log
