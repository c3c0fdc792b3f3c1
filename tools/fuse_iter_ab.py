"""A/B timing of SSB_FUSE_ITER variants at the headline configuration (device-resident batch, ssb_run)."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ssspy_b200.bss import GaussILRMA  # noqa: E402
from ssspy_b200.utils.synth import make_batch, make_nmf_init  # noqa: E402


def main():
    B, N, I, J, K = 64, 2, 1025, 512, 16
    steps = int(os.environ.get("STEPS", "20"))
    X = torch.from_numpy(make_batch(B, N, I, J, config_id=2, mode="mix").astype(np.complex64)).cuda()
    T0, V0 = make_nmf_init(N, I, J, K, seed=0)
    for mode in sys.argv[1:] or ["0", "1", "9", "3", "5", "0"]:
        os.environ["SSB_FUSE_ITER"] = mode
        sep = GaussILRMA(n_basis=K, spatial_algorithm="IP", record_loss=False)
        sep(X, n_iter=0, basis=T0, activation=V0)
        sep.run_iterations(3)
        torch.cuda.synchronize()
        best = []
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            sep.run_iterations(steps)
            e1.record()
            torch.cuda.synchronize()
            best.append(e0.elapsed_time(e1) / steps)
        print(json.dumps({"SSB_FUSE_ITER": mode, "chunks": len(sep._chunks), "ms_per_step": [round(t, 4) for t in best]}),
              flush=True)
        del sep


if __name__ == "__main__":
    main()
