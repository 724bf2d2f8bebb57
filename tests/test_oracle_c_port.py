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


def _dyn_baseline(track):
    import sys
    sys.path.insert(0, ROOT)
    import cpu_baseline
    t = load_golden("tracks.npz")
    return cpu_baseline.DynamicBaseline((t[track + "_x"], t[track + "_y"], float(t[track + "_dl"])))


def test_c_port_dynamic_matches_numpy_oracle_and_reference_vectors():
    """The dynamic model of the C port (bench.py's CPU baseline for configs[2]) against the numpy oracle's lap
    and against the post-processed outputs of the reference's own ltvmpc_dynamic_curvilinear.m (wide fixture)."""
    g = load_golden("dynamic_lap_fss2019.npz")
    b = _dyn_baseline("fss2019")
    sel = slice(0, None, 3)
    b.run(g["x0"][sel], c_layout(g["x_ref"][sel]), c_layout(g["x_lin"][sel]), c_layout(g["u_lin"][sel]), DT)
    o = b.last
    assert np.array_equal(o["exitflag"], g["exitflag"][sel].astype(np.int32))
    scale = np.maximum(1.0, np.abs(g["u_opt"][sel]).max(axis=1))
    assert (np.abs(o["u_opt"] - g["u_opt"][sel]).max(axis=1) / scale).max() < 1e-6
    assert np.abs(o["x_opt"] - g["x_opt"][sel]).max() < 1e-5
    assert (np.abs(o["fval"] - g["fval"][sel]) / (1 + np.abs(g["fval"][sel]))).max() < 1e-7
    w = load_golden("reference_m_wide.npz")
    fss = w["dynamic_N40_track"] == 1
    b.run(w["dynamic_N40_x0"][fss], c_layout(w["dynamic_N40_x_ref"][fss]), c_layout(w["dynamic_N40_x_lin"][fss]),
          c_layout(w["dynamic_N40_u_lin"][fss]), DT)
    o = b.last
    ok = w["dynamic_N40_exitflag"][fss] == 0
    assert np.array_equal(o["exitflag"] == 0, ok)
    ref_u = w["dynamic_N40_u_opt"][fss][ok]
    assert (np.abs(o["u_opt"][ok] - ref_u).max(axis=1) / np.maximum(1.0, np.abs(ref_u).max(axis=1))).max() < 1e-6
    assert np.abs(o["slack"][ok] - w["dynamic_N40_slack"][fss][ok]).max() < 1e-6
