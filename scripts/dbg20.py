import sys, numpy as np
sys.path.insert(0,'.'); sys.path.insert(0,'tests')
from conftest import load_golden, c_layout
import fsae_mpc_b200 as fm
from fsae_mpc_b200 import workload as wl
mpc = fm.FsaeMpc(0)
for tid,(n,t) in enumerate(wl.load_tracks().items()): mpc.set_track(tid,t[0],t[1],t[2])
g = load_golden("kinematic_lap_fsg2019_N20.npz")
b=15
sl=slice(b,b+1)
for kv in (2,1):
    mpc.set_kernel_version(kv); mpc.counters(reset=True)
    r = mpc.ltvmpc_kinetmatic_curvilinear(g["x0"][sl], c_layout(g["x_ref"][sl]), 0.05, c_layout(g["x_lin"][sl]), c_layout(g["u_lin"][sl]))
    print('kv',kv,'exit', r.exitflag, 'iters', r.iters, 'counters', mpc.counters(), 'slack', r.slack_opt, 'fval', r.fval, g['fval'][b])
    print(' wsB', np.nonzero(r.workingSetB[0])[0], r.workingSetB[0][np.nonzero(r.workingSetB[0])[0]])
    print(' wsC', np.nonzero(r.workingSetC[0])[0], r.workingSetC[0][np.nonzero(r.workingSetC[0])[0]])
    print(' du', np.abs(r.u_opt[0]-g['u_opt'][b]).round(4))
print('gold wsB', np.nonzero(g['wsB'][b])[0], 'wsC', np.nonzero(g['wsC'][b])[0], g['wsC'][b][np.nonzero(g['wsC'][b])[0]])
