"""CPU: the C restatement (oracle/ltvmpc_oracle.c, the bench's CPU baseline) against the
numpy oracle's committed results."""
import os
import time

import numpy as np

from conftest import load_golden, c_layout, DT, ROOT


def _baseline():
    import sys
    sys.path.insert(0, ROOT)
    import cpu_baseline
    t = load_golden("tracks.npz")
    return cpu_baseline.Baseline((t["fsg2019_x"], t["fsg2019_y"], float(t["fsg2019_dl"])))


def test_c_port_matches_numpy_oracle_lap():
    g = load_golden("kinematic_lap_fsg2019.npz")
    b = _baseline()
    b.run(g["x0"], c_layout(g["x_ref"]), c_layout(g["x_lin"]), c_layout(g["u_lin"]), DT)
    o = b.last
    assert np.array_equal(o["exitflag"], g["exitflag"].astype(np.int32))
    scale = np.maximum(1.0, np.abs(g["u_opt"]).max(axis=1))
    assert (np.abs(o["u_opt"] - g["u_opt"]).max(axis=1) / scale).max() < 1e-7
    assert np.abs(o["x_opt"] - g["x_opt"]).max() < 1e-6
    assert np.abs(o["slack"] - g["slack"]).max() < 1e-8
    assert (np.abs(o["fval"] - g["fval"]) / (1 + np.abs(g["fval"]))).max() < 1e-8


def test_c_port_matches_numpy_oracle_perturbed():
    g = load_golden("kinematic_perturbed_fsg2019.npz")
    b = _baseline()
    b.run(g["x0"], c_layout(g["x_ref"]), c_layout(g["x_lin"]), c_layout(g["u_lin"]), DT)
    o = b.last
    assert np.array_equal(o["exitflag"], g["exitflag"].astype(np.int32))
    scale = np.maximum(1.0, np.abs(g["u_opt"]).max(axis=1))
    assert (np.abs(o["u_opt"] - g["u_opt"]).max(axis=1) / scale).max() < 1e-7
    assert np.abs(o["slack"] - g["slack"]).max() < 1e-8
