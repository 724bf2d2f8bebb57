"""Golden fixtures for the other horizons of BASELINE.json configs[4] (N_steps = 20, 80):
oracle closed-loop runs with that horizon, every few steps kept.
    PYTHONPATH=. python scripts/make_golden_horizons.py"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import spline as sp, closed_loop as cl

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
for N, n_sim, every in ((20, 160, 8), (80, 130, 10)):
    tr = sp.Track.from_csv("/root/reference/data/fsg2019.csv")
    recs = []
    def rec(i, inp, out):
        sol = out[5]
        recs.append(dict(x0=inp["x0"], x_ref=inp["x_ref"], x_lin=inp["x_lin"], u_lin=inp["u_lin"], u_opt=out[0], x_opt=out[1],
                         exitflag=out[2], fval=out[3], slack=np.asarray(out[4]), wsB=sol.workingSetB.astype(np.int8),
                         wsC=sol.workingSetC.astype(np.int8), iters=sol.iter))
    t = time.time()
    h = cl.run(tr, "KINEMATIC", n_sim=n_sim, N_steps=N, record=rec)
    sel = recs[::every]
    out = {k: np.stack([np.asarray(r[k]) for r in sel]) for k in sel[0]}
    np.savez_compressed(os.path.join(OUT, f"kinematic_lap_fsg2019_N{N}.npz"), **out)
    print(f"N={N}: {h['steps']} steps in {time.time()-t:.0f}s, kept {len(sel)}, exit!=0 {sum(r['exitflag']!=0 for r in recs)}, iters mean {np.mean([r['iters'] for r in recs]):.1f}")
