cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
B="python bench.py --no-cpu-baseline --no-e2e"
show() { python -c "
import sys,json
t=open('$1').read().strip()
l=t.splitlines()[-1] if t else ''
try:
    d=json.loads(l); print('ms/step %.4f frac %.3f ok=%s'%(d['ms_per_step'],d['roofline']['frac'],d['state_after_timed_steps_ok']), d['roofline']['kernels_ms_per_step'])
except Exception as e: print('NOJSON', l[:300])
"; }
echo "== ISS diag + ISS / N=8 tests"
python tools/diag_iss.py 2>&1 | grep "n_iter [25]" | cut -c1-150
timeout 600 python -m pytest tests -m gpu -q -x -k "full_size or ISS or iss or (8-IP2) or (8-IP-) or n8 or phi or mnmf or MNMF" 2>&1 | tail -3
echo "== bench config 4 default (TC covariance for IP2)"
timeout 200 $B --config 4 --steps 5 --warmup 3 2>gpurun_out/r2p_c4.err > gpurun_out/r2p_c4.json; show gpurun_out/r2p_c4.json
echo "== bench config 4 SSB_IP2_OCC=1"
SSB_IP2_OCC=1 timeout 200 $B --config 4 --steps 5 --warmup 3 2>/dev/null > gpurun_out/r2p_c4_occ.json; show gpurun_out/r2p_c4_occ.json
echo "== bench N=8 IP"
timeout 150 $B --steps 20 --warmup 3 --sources 8 2>/dev/null > gpurun_out/r2p_n8.json; show gpurun_out/r2p_n8.json
echo "== bench config 3"
timeout 200 $B --config 3 --steps 5 --warmup 3 2>gpurun_out/r2p_c3.err > gpurun_out/r2p_c3.json; show gpurun_out/r2p_c3.json
echo "== bench config 5"
timeout 300 $B --config 5 --steps 3 --warmup 3 2>gpurun_out/r2p_c5.err > gpurun_out/r2p_c5.json; show gpurun_out/r2p_c5.json; tail -2 gpurun_out/r2p_c5.err
echo "== ncu launch list config 3 and 5 (dram bytes)"
NCU="ncu --clock-control none --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv"
timeout 300 $NCU --log-file gpurun_out/r2_ncu_launches_c3.csv python tools/ncu_target.py --config 3 --steps 2 > /dev/null 2>&1
timeout 400 $NCU --log-file gpurun_out/r2_ncu_launches_c5.csv python tools/ncu_target.py --config 5 --steps 2 > /dev/null 2>&1
ls -la gpurun_out | grep "r2_ncu_launches_c[35]"
