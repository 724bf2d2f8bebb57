"""ORACLE (test infrastructure only) -- numpy restatement of reference spline/*.m.

Nothing under oracle/ is product code: only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may import it, and only as the
checker.  PARITY STATUS: see oracle/README.md ("parity unpinned" for the QP
solve, reference-executed for everything the .m interpreter can run).

All arrays follow the MATLAB shapes of the reference: spline coefficient
matrices are (N_seg, 4) = [P0 P1 P2 P3] cubic Bezier control values.
"""
import numpy as np
from scipy import integrate as _integrate


def make_spline_periodic(P):
    """spline/make_spline_periodic.m:1-34 -- periodic cubic spline, Bezier form."""
    P = np.asarray(P, dtype=np.float64).reshape(-1)
    N = P.size
    A = np.zeros((N, N))
    idx = np.arange(N)
    A[idx, idx] = 4.0
    A[idx[1:], idx[:-1]] = 1.0
    A[idx[:-1], idx[1:]] = 1.0
    A[N - 1, 0] = 1.0
    A[0, N - 1] = 1.0
    b = np.zeros(N)
    b[: N - 1] = 4 * P[: N - 1] + 2 * P[1:N]
    b[N - 1] = 4 * P[N - 1] + 2 * P[0]
    P1 = np.linalg.solve(A, b)
    P2 = np.zeros(N)
    P2[: N - 1] = 2 * P[1:N] - P1[1:N]
    P2[N - 1] = 2 * P[0] - P1[0]
    P0 = P.copy()
    P3 = np.concatenate([P[1:N], P[:1]])
    return np.stack([P0, P1, P2, P3], axis=1)


def make_spline(P):
    """spline/make_spline.m:1-39 -- open cubic spline, Bezier form.

    Follows the reference literally, including its boundary rows
    (diag(1,:)=[1,2,0], diag(N-1,1)=2, diag(N,:)=[0,7,1]) fed to spdiags,
    which places column k of `diag` on diagonal k-1 with MATLAB's
    "drop from the top for super-diagonals / bottom for sub-diagonals" rule.
    """
    P = np.asarray(P, dtype=np.float64).reshape(-1)
    N = P.size - 1
    diag = np.zeros((N, 3))
    diag[0, :] = [1, 2, 0]
    diag[1 : N - 1, :] = [1, 4, 1]
    diag[N - 2, 0] = 2
    diag[N - 1, :] = [0, 7, 1]
    # spdiags(B, -1:1, N, N) for square N: A(i,j) on diagonal d=j-i takes B(j, d+2)
    # (for m>=n the element comes from row j of B).
    A = np.zeros((N, N))
    for j in range(N):
        A[j, j] = diag[j, 1]
        if j + 1 < N:
            A[j + 1, j] = diag[j, 0]      # sub-diagonal element in column j
            A[j, j + 1] = diag[j + 1, 2]  # super-diagonal element in column j+1
    b = np.zeros(N)
    b[0] = P[0] + 2 * P[1]
    b[1 : N - 1] = 4 * P[1 : N - 1] + 2 * P[2:N]
    b[N - 1] = 8 * P[N - 1] + P[N]
    P1 = np.linalg.solve(A, b)
    P2 = np.zeros(N)
    P2[0] = 2 * P1[0] - P[0]
    P2[1 : N - 1] = 2 * P[2:N] - P1[2:N]
    P2[N - 1] = (P[N] + P1[N - 1]) / 2
    return np.stack([P[:N], P1, P2, P[1 : N + 1]], axis=1)


def _segment(t, P, dl):
    """Shared prologue of interpolate_spline*.m:10-14 (mod, floor, local t)."""
    t = np.asarray(t, dtype=np.float64).reshape(-1)
    n_seg = max(P.shape)  # MATLAB length(P)
    t = np.mod(t, dl * n_seg)
    i = np.floor(t / dl).astype(np.int64)  # zero-based segment (MATLAB i-1)
    tl = t / dl - i
    return i, tl


def interpolate_spline(t, P, dl):
    """spline/interpolate_spline.m:1-20."""
    i, t = _segment(t, P, dl)
    return (P[i, 0] * (1 - t) ** 3 + 3 * P[i, 1] * (1 - t) ** 2 * t
            + 3 * P[i, 2] * (1 - t) * t ** 2 + P[i, 3] * t ** 3)


def interpolate_spline_d(t, P, dl):
    """spline/interpolate_spline_d.m:1-23."""
    i, t = _segment(t, P, dl)
    x_d = (-3 * (1 - t) ** 2 * P[i, 0] + 3 * (3 * t ** 2 - 4 * t + 1) * P[i, 1]
           + 3 * (2 * t - 3 * t ** 2) * P[i, 2] + 3 * t ** 2 * P[i, 3])
    return x_d / dl


def interpolate_spline_dd(t, P, dl):
    """spline/interpolate_spline_dd.m:1-23."""
    i, t = _segment(t, P, dl)
    x_dd = (6 * (1 - t) * P[i, 0] + 6 * (3 * t - 2) * P[i, 1]
            + 6 * (1 - 3 * t) * P[i, 2] + 6 * t * P[i, 3])
    return x_dd / dl ** 2


def interpolate_spline_ddd(t, P, dl):
    """spline/interpolate_spline_ddd.m:1-21."""
    i, _ = _segment(t, P, dl)
    x_ddd = -6 * P[i, 0] + 18 * P[i, 1] - 18 * P[i, 2] + 6 * P[i, 3]
    return x_ddd / dl ** 3


def interpolate_curvature(s, x_P, y_P, dl):
    """spline/interpolate_curvature.m:1-20."""
    X_d = interpolate_spline_d(s, x_P, dl)
    Y_d = interpolate_spline_d(s, y_P, dl)
    X_dd = interpolate_spline_dd(s, x_P, dl)
    Y_dd = interpolate_spline_dd(s, y_P, dl)
    return (X_d * Y_dd - X_dd * Y_d) / (X_d ** 2 + Y_d ** 2) ** 1.5


def interpolate_curvature_d(s, x_P, y_P, dl):
    """spline/interpolate_curvature_d.m:1-19 (central difference, delta = dl)."""
    s = np.asarray(s, dtype=np.float64)
    delta = dl
    kl = interpolate_curvature(s - delta, x_P, y_P, dl)
    ku = interpolate_curvature(s + delta, x_P, y_P, dl)
    return (ku - kl) / (2 * delta)


def interpolate_angle(s, x_P, y_P, dl):
    """spline/interpolate_angle.m:1-18."""
    return np.arctan2(interpolate_spline_d(s, y_P, dl), interpolate_spline_d(s, x_P, dl))


def closest_point(x0, y0, x_P, y_P, dl, s, epsilon):
    """spline/closest_point.m:1-34 -- Newton-Raphson on squared distance."""
    s = float(s)
    delta = epsilon * 2
    it = 0
    while abs(delta) > epsilon:
        X = interpolate_spline(s, x_P, dl)[0]
        Y = interpolate_spline(s, y_P, dl)[0]
        X_d = interpolate_spline_d(s, x_P, dl)[0]
        Y_d = interpolate_spline_d(s, y_P, dl)[0]
        X_dd = interpolate_spline_dd(s, x_P, dl)[0]
        Y_dd = interpolate_spline_dd(s, y_P, dl)[0]
        dist_d = 2 * (X - x0) * X_d + 2 * (Y - y0) * Y_d
        dist_dd = 2 * (X - x0) * X_dd + 2 * X_d ** 2 + 2 * (Y - y0) * Y_dd + 2 * Y_d ** 2
        delta = dist_d / dist_dd
        s = s - delta
        it += 1
        if it > 1000:
            raise RuntimeError("closest_point did not converge")
    return s


def _bisection(xl, xu, f, epsilon):
    """spline/arclength_reparam.m:68-97."""
    while True:
        x = (xl + xu) / 2
        fx = f(x)
        if abs(fx) <= epsilon:
            return x
        elif fx < 0:
            xl = x
        else:
            xu = x


def arclength_reparam(x_P, y_P, M, periodic):
    """spline/arclength_reparam.m:1-66.

    The reference's speed integrand uses P(i,1) where the Bezier derivative has
    P(i,2) (lines 20-23, 43-46); that is kept verbatim -- it is what defines the
    reference's dl, L and knot placement.  MATLAB `integral` (adaptive
    Gauss-Kronrod, AbsTol 1e-10, RelTol 1e-6) is restated with scipy.quad at
    tighter tolerance; the smooth polynomial integrand makes both agree to
    ~1e-12, far inside the 0.01 bisection threshold that consumes the value.
    """
    x_P = np.asarray(x_P, dtype=np.float64)
    y_P = np.asarray(y_P, dtype=np.float64)
    N = max(x_P.shape)

    def speed(i):
        def f(t):
            x_d = (-3 * (1 - t) ** 2 * x_P[i, 0] + 3 * (3 * t ** 2 - 4 * t + 1) * x_P[i, 0]
                   + 3 * (2 * t - 3 * t ** 2) * x_P[i, 2] + 3 * t ** 2 * x_P[i, 3])
            y_d = (-3 * (1 - t) ** 2 * y_P[i, 0] + 3 * (3 * t ** 2 - 4 * t + 1) * y_P[i, 0]
                   + 3 * (2 * t - 3 * t ** 2) * y_P[i, 2] + 3 * t ** 2 * y_P[i, 3])
            return np.sqrt(x_d ** 2 + y_d ** 2)
        return f

    def quad(f, a, b):
        return _integrate.quad(f, a, b, epsabs=1e-13, epsrel=1e-13, limit=200)[0]

    l = np.array([quad(speed(i), 0.0, 1.0) for i in range(N)])
    l_cum = np.concatenate([[0.0], np.cumsum(l)])
    dl = l_cum[N] / M

    Px = np.zeros(M + 1)
    Py = np.zeros(M + 1)
    Px[0] = x_P[0, 0]
    Py[0] = y_P[0, 0]
    Px[M] = x_P[N - 1, 3]
    Py[M] = y_P[N - 1, 3]
    for i in range(1, M):
        j1 = int(np.argmax(l_cum >= i * dl))  # MATLAB find(...,1), one-based = j1+1
        j = j1 - 1                             # zero-based segment (MATLAB j-1)
        sp = speed(j)
        f = lambda T, sp=sp, j=j, i=i: quad(sp, 0.0, T) + l_cum[j] - i * dl
        t_i = _bisection(0.0, 1.0, f, 0.01)
        Px[i] = interpolate_spline(t_i + j, x_P, 1.0)[0]
        Py[i] = interpolate_spline(t_i + j, y_P, 1.0)[0]

    if periodic:
        x_new = make_spline_periodic(Px[:M])
        y_new = make_spline_periodic(Py[:M])
    else:
        x_new = make_spline(Px)
        y_new = make_spline(Py)
    L = l_cum[-1]
    return x_new, y_new, dl, L


def read_raceline_csv(filename):
    """util/read_raceline_csv.m:1-21 (readmatrix skips the header row)."""
    A = np.loadtxt(filename, delimiter=",", skiprows=1)
    return tuple(A[:, k] for k in range(11))


class Track:
    """Bundle (x_spline, y_spline, dl, L) -- what main.m:15-18 builds -- and expose
    kappa(s) like the anonymous function main.m:18 hands to the MPC."""

    def __init__(self, x_spline, y_spline, dl, L):
        self.x_spline = np.ascontiguousarray(x_spline, dtype=np.float64)
        self.y_spline = np.ascontiguousarray(y_spline, dtype=np.float64)
        self.dl = float(dl)
        self.L = float(L)

    @classmethod
    def from_csv(cls, filename, M=100):
        x, y = read_raceline_csv(filename)[:2]
        xs = make_spline_periodic(x)
        ys = make_spline_periodic(y)
        xs, ys, dl, L = arclength_reparam(xs, ys, M, True)
        return cls(xs, ys, dl, L)

    def kappa(self, s):
        k = interpolate_curvature(s, self.x_spline, self.y_spline, self.dl)
        return k[0] if np.ndim(s) == 0 else k
