import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")
DT = 0.05


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu)")


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name)))


def c_layout(a):
    """MATLAB-shaped per-problem (B, NX, N) -> C-ABI layout (B, N, NX) contiguous."""
    return np.ascontiguousarray(np.asarray(a).transpose(0, 2, 1))


class GoldenTrack:
    """Committed spline coefficients (derived from the reference's data/*.csv by
    scripts/make_golden.py) with the oracle's kappa(s)."""

    def __init__(self, name):
        from oracle import spline as sp
        t = load_golden("tracks.npz")
        self.name = name
        self.track = sp.Track(t[name + "_x"], t[name + "_y"], float(t[name + "_dl"]), float(t[name + "_L"]))
        self.x_spline, self.y_spline = self.track.x_spline, self.track.y_spline
        self.dl, self.L = self.track.dl, self.track.L
        self.kappa = self.track.kappa


@pytest.fixture(scope="session")
def fsg():
    return GoldenTrack("fsg2019")


@pytest.fixture(scope="session")
def fss():
    return GoldenTrack("fss2019")


@pytest.fixture(scope="session")
def fso():
    return GoldenTrack("fso2020")


@pytest.fixture(scope="session")
def mpc_x():
    """Context on the CROSS-CHECK build of the library (libfsae_mpc_b200_xcheck.so: product kernels plus the
    shared-memory operator kernel and the warp-count / block-size variants) for the variant-agreement tests."""
    import fsae_mpc_b200 as fm
    from fsae_mpc_b200 import build
    ctx = fm.FsaeMpc(0, lib_path=build.LIB_XCHECK)
    t = load_golden("tracks.npz")
    for tid, name in enumerate(("fsg2019", "fss2019", "fso2020")):
        ctx.set_track(tid, t[name + "_x"], t[name + "_y"], float(t[name + "_dl"]))
    yield ctx
    ctx.close()


@pytest.fixture(scope="session")
def mpc():
    """The CUDA context.  Fails (does not skip) when the extension or the GPU is missing:
    -m gpu tests must never pass on a fallback."""
    import fsae_mpc_b200 as fm
    ctx = fm.FsaeMpc(0)
    t = load_golden("tracks.npz")
    for tid, name in enumerate(("fsg2019", "fss2019", "fso2020")):
        ctx.set_track(tid, t[name + "_x"], t[name + "_y"], float(t[name + "_dl"]))
    yield ctx
    ctx.close()
