"""Golden fixture for the dynamic model at horizon 80 (BASELINE.json configs[4] names 20/40/80; the reference takes
any N_steps = length(x_ref), ltvmpc_dynamic_curvilinear.m:17): an oracle closed-loop run of main.m with the
dynamic model and N_steps = 80 on fss2019, every third step kept.
    PYTHONPATH=. python scripts/make_golden_dyn80.py"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import spline as sp, closed_loop as cl

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
N, n_sim, every = 80, int(sys.argv[1]) if len(sys.argv) > 1 else 75, 3
tr = sp.Track.from_csv("/root/reference/data/fss2019.csv")
recs = []
def rec(i, inp, out):
    sol = out[5]
    recs.append(dict(x0=inp["x0"], x_ref=inp["x_ref"], x_lin=inp["x_lin"], u_lin=inp["u_lin"], u_opt=out[0], x_opt=out[1],
                     exitflag=out[2], fval=out[3], slack=np.asarray(out[4]), wsB=sol.workingSetB.astype(np.int8),
                     wsC=sol.workingSetC.astype(np.int8), iters=sol.iter))
    print(i, out[2], sol.iter, f"{time.time()-t:.0f}s", flush=True)
t = time.time()
h = cl.run(tr, "DYNAMIC", n_sim=n_sim, N_steps=N, record=rec)
sel = recs[::every]
out = {k: np.stack([np.asarray(r[k]) for r in sel]) for k in sel[0]}
np.savez_compressed(os.path.join(OUT, f"dynamic_lap_fss2019_N{N}.npz"), **out)
print(f"N={N}: {h['steps'] if 'steps' in h else len(recs)} steps in {time.time()-t:.0f}s, kept {len(sel)}, exit!=0 {sum(r['exitflag']!=0 for r in recs)}, iters mean {np.mean([r['iters'] for r in recs]):.1f}")
