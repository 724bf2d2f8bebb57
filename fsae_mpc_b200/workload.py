"""Synthetic batch generator for BASELINE.json's configs: lap samples of committed closed-loop runs
(fsae_mpc_b200/data/*_lap_*.npz; the tests reach the same files through links in tests/golden/) with seeded
perturbations of the initial state.  Product-side (bench.py, examples); does not touch oracle/ or tests/."""
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "fsae_mpc_b200", "data")


def load_tracks():
    t = dict(np.load(os.path.join(GOLDEN, "tracks.npz")))
    return {n: (t[n + "_x"], t[n + "_y"], float(t[n + "_dl"]), float(t[n + "_L"]))
            for n in ("fsg2019", "fss2019", "fso2020")}


def perturbed_batch(model, track, B, seed=0):
    """B problems: (x0, x_ref, x_lin, u_lin) in the C-ABI layout (B,NX) (B,N,NX) (B,N,NX) (B,N,NU).
    Perturbation: n +-0.3 m, mu +-0.08 rad, v +-1.5 m/s, delta +-0.05 rad around lap samples."""
    g = dict(np.load(os.path.join(GOLDEN, f"{model}_lap_{track}.npz")))
    rng = np.random.default_rng(seed)
    n = g["x0"].shape[0]
    pick = rng.integers(n, size=B)
    x0 = g["x0"][pick].copy()
    x0[:, 1] += rng.uniform(-0.3, 0.3, B)
    x0[:, 2] += rng.uniform(-0.08, 0.08, B)
    x0[:, 3] = np.maximum(0.5, x0[:, 3] + rng.uniform(-1.5, 1.5, B))
    x0[:, -1] += rng.uniform(-0.05, 0.05, B)
    tr = lambda a: np.ascontiguousarray(a.transpose(0, 2, 1)[pick])
    return x0, tr(g["x_ref"]), tr(g["x_lin"]), tr(g["u_lin"])


def horizon_batch(model, track, B, N, seed=0):
    """BASELINE.json configs[4] (horizons 20 / 40 / 80): N = 40 the lap problems, N = 20 their first 20 steps,
    N = 80 perturbed samples of the committed horizon-80 laps (kinematic: fsg2019, dynamic: fss2019 -- `track` is
    ignored there).  Returns (x0, x_ref, x_lin, u_lin, track_name)."""
    if N == 80:
        track = "fsg2019" if model == "kinematic" else "fss2019"
        g = dict(np.load(os.path.join(GOLDEN, f"{model}_lap_{track}_N80.npz")))
        rng = np.random.default_rng(seed)
        pick = rng.integers(g["x0"].shape[0], size=B)
        x0 = g["x0"][pick].copy()
        x0[:, 1] += rng.uniform(-0.3, 0.3, B)
        x0[:, 2] += rng.uniform(-0.08, 0.08, B)
        x0[:, 3] = np.maximum(0.5, x0[:, 3] + rng.uniform(-1.5, 1.5, B))
        x0[:, -1] += rng.uniform(-0.05, 0.05, B)
        tr = lambda a: np.ascontiguousarray(a.transpose(0, 2, 1)[pick])
        return x0, tr(g["x_ref"]), tr(g["x_lin"]), tr(g["u_lin"]), track
    x0, xr, xl, ul = perturbed_batch(model, track, B, seed)
    c = lambda a: np.ascontiguousarray(a[:, :N])
    return x0, c(xr), c(xl), c(ul), track
