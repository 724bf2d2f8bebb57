"""KKT certificate (tests/test_gpu_fullsize.py) on larger horizon-80 batches of both models, fresh seed:
    python scripts/soak_kkt_horizon80.py"""
import sys, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import fsae_mpc_b200 as fm
import test_gpu_fullsize as T
from fsae_mpc_b200 import workload as wl
mpc = fm.FsaeMpc(0)
for tid, (n, t) in enumerate(wl.load_tracks().items()): mpc.set_track(tid, t[0], t[1], t[2])
mpc.set_params(3, fm.default_params(fm.DYNAMIC))
for args in (("kinematic", "fsg2019", 8192, 512, 0, 0, 8192 // 20, 80), ("dynamic", "fss2019", 4096, 256, 1, 3, 4096 // 20, 80)):
    w, ninf = T._certify(mpc, *args[:7], N=args[7])
    bad = w["primal"] > 1e-7 or w["stat"] > 1e-7 or w["dual"] > 1e-6 or w["active"] > 1e-7
    print(args[0], args[1], args[2], "N", args[7], "infeasible", ninf, {k: f"{v:.1e}" for k, v in w.items()}, "BAD" if bad else "ok", flush=True)
