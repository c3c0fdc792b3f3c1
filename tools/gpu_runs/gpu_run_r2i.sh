cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for m in 3 4; do
  echo "== SSB_TMA=$m"
  SSB_TMA=$m timeout 200 python -m pytest tests -m gpu -q -x -s -k "baseline_shapes and 2-IP-1025 or fused_iteration or fused_tensor_core" 2>&1 | grep -E "relerr|passed|failed|Error" | cut -c1-200 | tail -4
done
echo "== multi-tile fused check (B=6, I=1025)"
SSB_TMA=7 timeout 120 python - <<'PY'
import numpy as np, os
from ssspy_b200.bss import GaussILRMA
from ssspy_b200.utils.synth import make_batch, make_nmf_init
B,N,I,J,K=6,2,1025,512,16
X=make_batch(B,N,I,J,config_id=77); T,V=make_nmf_init(N,I,J,K,seed=5)
out={}
for flag in ("0","1"):
    os.environ["SSB_FUSE_ITER"]=flag
    m=GaussILRMA(n_basis=K, record_loss=False); m.chunk_size=B
    try:
        Y=m(X,n_iter=4,basis=T,activation=V); out[flag]=(Y,m.basis.copy())
    except Exception as e:
        print("flag",flag,"raised",repr(e)); out[flag]=None
if out["0"] is not None and out["1"] is not None:
    for a,b,nm in zip(out["1"],out["0"],("Y","T")):
        print(nm,"relerr fused vs unfused %.2e"%(np.linalg.norm(a-b)/np.linalg.norm(b)), "finite", np.isfinite(a).all())
    e=np.linalg.norm(out["1"][0]-out["0"][0],axis=(1,2,3))/np.linalg.norm(out["0"][0],axis=(1,2,3)); print("per mixture",e)
PY
for m in 0 3 7; do
  echo "== bench SSB_TMA=$m"
  SSB_TMA=$m timeout 100 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "
import sys,json
l=sys.stdin.read().strip()
try:
    d=json.loads(l); print('ms/step %.4f frac %.3f'%(d['ms_per_step'],d['roofline']['frac']), d['roofline']['kernels_ms_per_step'])
except Exception as e: print('NOJSON', l[:300])
"
done
