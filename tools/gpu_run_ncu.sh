set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
SSB_ISS_COV=1 timeout 600 ncu --set full --import-source on --clock-control none -k regex:'k_iss1' -c 1 -f -o gpurun_out/r1s2_iss_cov python tools/bench_configs.py --steps 1 --only "c3 AuxLaplaceIVA-ISS" > gpurun_out/r1s2_ncu.log 2>&1
SSB_ISS_COV=0 timeout 600 ncu --set full --import-source on --clock-control none -k regex:'k_iss1' -c 1 -f -o gpurun_out/r1s2_iss_old python tools/bench_configs.py --steps 1 --only "c3 AuxLaplaceIVA-ISS" >> gpurun_out/r1s2_ncu.log 2>&1
tail -3 gpurun_out/r1s2_ncu.log
