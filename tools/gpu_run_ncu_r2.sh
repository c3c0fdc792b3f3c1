cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
NCU="ncu --clock-control none"
# (1) launch list with device times, default path, single plan, 3 steps
timeout 300 $NCU --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv --log-file gpurun_out/r2_ncu_launches_c2.csv python tools/ncu_target.py --steps 3 > /dev/null 2>&1
# (2) the same for the chunked engine path (4 plans on 4 streams): DRAM bytes per step of the run bench.py times
timeout 300 $NCU --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv --log-file gpurun_out/r2_ncu_launches_c2_chunked.csv python tools/ncu_target.py --steps 3 --chunked > /dev/null 2>&1
# (3) TMA path launch list (all TMA kernels, fused)
SSB_TMA=7 timeout 300 $NCU --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct --csv --log-file gpurun_out/r2_ncu_launches_c2_tma7.csv python tools/ncu_target.py --steps 3 > /dev/null 2>&1
SSB_TMA=3 SSB_FUSE_ITER=0 timeout 300 $NCU --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum --csv --log-file gpurun_out/r2_ncu_launches_c2_tma3.csv python tools/ncu_target.py --steps 3 > /dev/null 2>&1
# (4) full captures: dominant kernels of the default path and the TMA kernels
timeout 400 $NCU --set full --import-source on -k regex:'kf_basis_coop|kf_phi_cov|kf_activation_coop' -s 3 -c 3 -o gpurun_out/r2_ncu_c2_default python tools/ncu_target.py --steps 3 > /dev/null 2>&1
SSB_TMA=7 timeout 400 $NCU --set full --import-source on -k regex:'kt_tile' -s 2 -c 2 -o gpurun_out/r2_ncu_c2_tma python tools/ncu_target.py --steps 3 > /dev/null 2>&1
SSB_TMA=3 SSB_FUSE_ITER=0 timeout 400 $NCU --set full --import-source on -k regex:'kt_tile' -s 2 -c 2 -o gpurun_out/r2_ncu_c2_tma_fs1 python tools/ncu_target.py --steps 3 > /dev/null 2>&1
# (5) N = 8 launch list
timeout 300 $NCU --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv --log-file gpurun_out/r2_ncu_launches_n8.csv python tools/ncu_target.py --steps 2 --sources 8 > /dev/null 2>&1
ls -la gpurun_out | grep r2_ncu
