"""Tiny run of every fused-kernel family for compute-sanitizer (memcheck / racecheck / synccheck):
    compute-sanitizer --tool racecheck python scripts/sanitize_small.py"""
import sys, numpy as np
sys.path.insert(0, '.')
import fsae_mpc_b200 as fm
from fsae_mpc_b200 import workload as wl
B = 6
mpc = fm.FsaeMpc(0)
for tid, (n, t) in enumerate(wl.load_tracks().items()):
    mpc.set_track(tid, t[0], t[1], t[2])
mpc.set_params(1, fm.default_params(fm.DYNAMIC))
x0, xr, xl, ul = wl.perturbed_batch("kinematic", "fsg2019", B, 3)
r = mpc.ltvmpc_kinetmatic_curvilinear(x0, xr, 0.05, xl, ul)
print("kin40", r.exitflag, r.iters)
c = lambda a: np.ascontiguousarray(a[:, :20])
r = mpc.ltvmpc_kinetmatic_curvilinear(x0, c(xr), 0.05, c(xl), c(ul))
print("kin20", r.exitflag, r.iters)
g = dict(np.load("tests/golden/kinematic_lap_fsg2019_N80.npz"))
tr = lambda a: np.ascontiguousarray(a.transpose(0, 2, 1)[:B])
r = mpc.ltvmpc_kinetmatic_curvilinear(g["x0"][:B], tr(g["x_ref"]), 0.05, tr(g["x_lin"]), tr(g["u_lin"]))
print("kin80", r.exitflag, r.iters)
x0, xr, xl, ul = wl.perturbed_batch("dynamic", "fss2019", B, 3)
r = mpc.ltvmpc_dynamic_curvilinear(x0, xr, 0.05, xl, ul, track_id=np.ones(B, np.int32), param_id=np.ones(B, np.int32))
print("dyn40", r.exitflag, r.iters)
