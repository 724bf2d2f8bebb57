"""ORACLE (test infrastructure only) -- CPU stand-in for the reference's QP solver call

    [x,fval,exitflag,iter,lambda,auxOutput] = qpOASES(H,g,A,lb,ub,lbA,ubA)

(reference call sites: mpc/ltv/kinematic/ltvmpc_kinetmatic_curvilinear.m:52,
mpc/ltv/dynamic/ltvmpc_dynamic_curvilinear.m:52; interface contract:
optimizers/matlab/qpOASES/qpOASES.m:14-65).

PARITY UNPINNED for this function: qpOASES is a third-party dependency that the
reference ships only as Windows MEX binaries (optimizers/matlab/qpOASES/*.mexw64,
qpOASES 3.2, (C) 2007-2017 per qpOASES.m:2) -- no source, nothing runnable
here.  What this file restates is the *contract* of that call: return the
minimiser of  1/2 x'Hx + x'g  s.t. lb<=x<=ub, lbA<=Ax<=ubA, the optimal
objective, the multipliers and the working set, with exitflag 0, or a non-zero
exitflag when infeasible.  The MPC QPs are strictly convex in the control
variables and have a linear exact-penalty cost on the slack variables, so their
minimiser is unique; any solver that returns a point passing the KKT test
below returns qpOASES's point up to round-off.  The solver is therefore
certified, not trusted: `qpoases()` refuses to return exitflag 0 unless
`kkt_residuals()` passes on the *unregularised* problem.

Method: Goldfarb-Idnani dual active set (Math. Prog. 27, 1983) on H+eps*I_null
to identify the working set, followed by one exact equality-constrained KKT
solve on the original H ("polish").  This is deliberately a different code
shape from the CUDA kernel (dense re-solves with numpy, no factor updates).
"""
from dataclasses import dataclass, field
import numpy as np

QPOASES_INFTY = 1.0e20   # bounds beyond this are "no bound" (qpOASES Constants: INFTY)


@dataclass
class QPSolution:
    x: np.ndarray
    fval: float
    exitflag: int
    iter: int
    lam: np.ndarray                 # multipliers, qpOASES sign: >=0 at lower, <=0 at upper
    workingSetB: np.ndarray         # -1 / 0 / +1 per variable bound
    workingSetC: np.ndarray         # -1 / 0 / +1 per general constraint
    kkt: dict = field(default_factory=dict)


def _stack(n, A, lb, ub, lbA, ubA):
    A = np.zeros((0, n)) if A is None else np.asarray(A, dtype=np.float64)
    C = np.vstack([np.eye(n), A])
    lo = np.concatenate([np.asarray(lb, float).reshape(-1), np.asarray(lbA, float).reshape(-1)])
    up = np.concatenate([np.asarray(ub, float).reshape(-1), np.asarray(ubA, float).reshape(-1)])
    lo = np.where(lo <= -QPOASES_INFTY, -np.inf, lo)
    up = np.where(up >= QPOASES_INFTY, np.inf, up)
    return C, lo, up


def kkt_residuals(H, g, A, lb, ub, lbA, ubA, x, lam):
    """Scaled KKT residuals of (x, lam) for the ORIGINAL problem.
    lam uses the qpOASES convention: H x + g = [I; A]' lam."""
    n = g.size
    C, lo, up = _stack(n, A, lb, ub, lbA, ubA)
    Cx = C @ x
    scale_c = 1.0 + np.minimum(np.abs(np.where(np.isfinite(lo), lo, 0)),
                               np.abs(np.where(np.isfinite(up), up, 0)))
    prim = max(0.0, float(np.max((lo - Cx) / scale_c)), float(np.max((Cx - up) / scale_c)))
    grad = H @ x + g
    stat = float(np.max(np.abs(grad - C.T @ lam)) / (1.0 + np.max(np.abs(grad)) + np.max(np.abs(lam), initial=0.0)))
    lam_scale = 1.0 + np.max(np.abs(lam), initial=0.0)
    # complementarity: lam>0 needs Cx==lo, lam<0 needs Cx==up
    comp_lo = np.where(lam > 0, np.abs(Cx - lo), 0.0)
    comp_up = np.where(lam < 0, np.abs(Cx - up), 0.0)
    comp_lo = np.where(np.isfinite(comp_lo), comp_lo, np.inf)
    comp = float(max(np.max(comp_lo / scale_c), np.max(comp_up / scale_c)))
    return dict(primal=prim, stationarity=stat, complementarity=comp, lam_scale=lam_scale)


def _gi_active_set(H, g, C, lo, up, max_iter, tol, W0=()):
    """Goldfarb-Idnani dual active-set iteration with dense re-solves.
    Returns (x, W, u, it, status) with W a list of (row, sign), sign=-1 lower / +1 upper.
    W0 is an initial working set that must be dual feasible (see qpoases())."""
    n = g.size
    Hinv = np.linalg.inv(H)
    Hinv = 0.5 * (Hinv + Hinv.T)
    x = -Hinv @ g
    W = list(W0)    # active (row, sign)
    u = np.zeros(0)
    if W:
        N0 = np.stack([-s * C[r] for (r, s) in W], axis=1)
        b0 = np.array([lo[r] if s < 0 else -up[r] for (r, s) in W])
        G0 = Hinv @ N0
        u = np.linalg.solve(N0.T @ G0, b0 - N0.T @ x)
        x = x + G0 @ u
        assert (u >= 0).all(), "initial working set is not dual feasible"
    scale = 1.0 + np.minimum(np.abs(np.where(np.isfinite(lo), lo, 0)), np.abs(np.where(np.isfinite(up), up, 0)))
    it = 0

    def normal(row, sign):
        return -sign * C[row]          # constraint written  nrm'x >= b

    def rhs(row, sign):
        return lo[row] if sign < 0 else -up[row]

    while True:
        Cx = C @ x
        viol_lo = (Cx - lo) / scale      # negative = violated
        viol_up = (up - Cx) / scale
        for (row, sign) in W:
            viol_lo[row] = np.inf
            viol_up[row] = np.inf
        vl = np.where(np.isnan(viol_lo), np.inf, viol_lo)
        vu = np.where(np.isnan(viol_up), np.inf, viol_up)
        il, iu = int(np.argmin(vl)), int(np.argmin(vu))
        if min(vl[il], vu[iu]) >= -tol:
            return x, W, u, it, 0
        p, psign = (il, -1) if vl[il] <= vu[iu] else (iu, +1)
        n_p = normal(p, psign)
        u = np.append(u, 0.0)
        while True:
            it += 1
            if it > max_iter:
                return x, W, u[:-1], it, 1
            s_p = n_p @ x - rhs(p, psign)
            q = len(W)
            Hn = Hinv @ n_p
            if q > 0:
                N = np.stack([normal(r, s) for (r, s) in W], axis=1)   # n x q
                G = Hinv @ N
                M = N.T @ G
                r = np.linalg.solve(M, N.T @ Hn)
                z = Hn - G @ r
            else:
                r = np.zeros(0)
                z = Hn
            zn = z @ n_p
            lin_dep = zn <= 1e-13 * max(1.0, Hn @ n_p)
            # partial (dual) step length
            t1, l = np.inf, -1
            for j in range(q):
                if r[j] > 1e-14:
                    tj = u[j] / r[j]
                    if tj < t1:
                        t1, l = tj, j
            t2 = np.inf if lin_dep else -s_p / zn
            t = min(t1, t2)
            if not np.isfinite(t):
                return x, W, u[:-1], it, -2      # infeasible
            if not np.isfinite(t2):
                # dual step only, drop l
                u[:q] -= t * r
                u[q] += t
                W.pop(l)
                u = np.delete(u, l)
                continue
            x = x + t * z
            u[:q] -= t * r
            u[q] += t
            if t2 <= t1:
                W.append((p, psign))
                # Re-derive (x, u) from the working set alone: the minimiser of the
                # equality-constrained problem.  Removes the round-off that long
                # sequences of partial steps (t ~ 1e4 on near-degenerate constraints)
                # leave in x; costs nothing that matters in an oracle.
                Nw = np.stack([normal(r_, s_) for (r_, s_) in W], axis=1)
                bw = np.array([rhs(r_, s_) for (r_, s_) in W])
                Gw = Hinv @ Nw
                x_unc = -Hinv @ g
                u_new = np.linalg.solve(Nw.T @ Gw, bw - Nw.T @ x_unc)
                if np.all(u_new >= -1e-9 * (1.0 + np.abs(u_new).max())):
                    u = np.maximum(u_new, 0.0)
                    x = x_unc + Gw @ u_new
                break
            W.pop(l)
            u = np.delete(u, l)


def qpoases(H, g, A, lb, ub, lbA, ubA, max_iter=2000, tol=1e-10, kkt_tol=1e-7):
    """Stand-in for qpOASES(H,g,A,lb,ub,lbA,ubA) -- see module docstring."""
    H = np.asarray(H, dtype=np.float64)
    g = np.asarray(g, dtype=np.float64).reshape(-1)
    n = g.size
    C, lo, up = _stack(n, A, lb, ub, lbA, ubA)
    m_all = C.shape[0]

    # regularise only the (numerically) flat directions of H for the active-set search
    w, V = np.linalg.eigh(0.5 * (H + H.T))
    eps = 1e-9 * max(1.0, w.max())
    flat = w < eps
    Hreg = H + (V[:, flat] * eps) @ V[:, flat].T if flat.any() else H

    # Variables on which the objective is purely linear (zero row of H, g_i != 0): the
    # problem restricted to their bounds has them AT the bound g pushes towards, with
    # multiplier |g_i| -- a dual-feasible start that keeps the 1e8 penalty gradients out of
    # the unconstrained minimiser (x_i would be -g_i/eps there).
    W0 = []
    for i in range(n):
        if np.all(H[i] == 0) and g[i] != 0:
            if g[i] > 0 and np.isfinite(lo[i]):
                W0.append((i, -1))
            elif g[i] < 0 and np.isfinite(up[i]):
                W0.append((i, +1))
    x, W, u, it, status = _gi_active_set(Hreg, g, C, lo, up, max_iter, tol, W0)

    wsB = np.zeros(n, dtype=np.int64)
    wsC = np.zeros(m_all - n, dtype=np.int64)
    lam = np.zeros(m_all)
    if status != 0:
        return QPSolution(x, float(0.5 * x @ H @ x + g @ x), status, it, lam, wsB, wsC)

    # Phase 2 -- textbook primal active-set clean-up from the (feasible) GI end point.
    # GI carries x through long chains of partial steps; near-degenerate vertices can leave
    # it with one constraint too many/few.  Re-solving the equality-constrained problem on W
    # from scratch and walking to it (adding blockers, dropping wrong-signed multipliers)
    # removes any dependence on that history.
    def eqp(Hm, Wl):
        q_ = len(Wl)
        if q_ == 0:
            return np.linalg.solve(Hm, -g), np.zeros(0)
        rows_ = np.array([r for r, _ in Wl])
        signs_ = np.array([s_ for _, s_ in Wl])
        Cw = C[rows_]
        bw = np.where(signs_ < 0, lo[rows_], up[rows_])
        K = np.block([[Hm, -Cw.T], [Cw, np.zeros((q_, q_))]])
        rhs_ = np.concatenate([-g, bw])
        try:
            sol_ = np.linalg.solve(K, rhs_)
            # iterative refinement with extended-precision residuals (the KKT matrix mixes
            # 1e6-size Hessian entries with unit constraint rows; plain LU leaves ~1e-8)
            Kl, rl = K.astype(np.longdouble), rhs_.astype(np.longdouble)
            for _ in range(3):
                res = (rl - Kl @ sol_.astype(np.longdouble)).astype(np.float64)
                sol_ = sol_ + np.linalg.solve(K, res)
        except np.linalg.LinAlgError:
            sol_ = np.linalg.lstsq(K, rhs_, rcond=None)[0]
        return sol_[:n], sol_[n:]

    flat_var = np.array([np.all(H[i] == 0) for i in range(n)])
    W = list(W)
    for _ in range(4 * (n + 10)):
        x_eq, lw = eqp(Hreg, W)
        pdir = x_eq - x
        if np.max(np.abs(pdir)) <= 1e-11 * (1.0 + np.max(np.abs(x))):
            x = x_eq
            grad_u = np.abs((H @ x + g)[~flat_var]).max() if (~flat_var).any() else 1.0
            dtol = 1e-9 * (1.0 + grad_u)
            sgn = np.array([-lw[k] if W[k][1] < 0 else lw[k] for k in range(len(W))])
            if len(W) == 0 or sgn.max() <= dtol:
                break
            W.pop(int(np.argmax(sgn)))
            it += 1
            continue
        Cx = C @ x
        Cp = C @ pdir
        alpha, blk = 1.0, None
        inW = {r for r, _ in W}
        for r_ in range(m_all):
            if r_ in inW:
                continue
            if Cp[r_] < -1e-14 and np.isfinite(lo[r_]):
                a_ = max(0.0, (lo[r_] - Cx[r_]) / Cp[r_])
                if a_ < alpha:
                    alpha, blk = a_, (r_, -1)
            elif Cp[r_] > 1e-14 and np.isfinite(up[r_]):
                a_ = max(0.0, (up[r_] - Cx[r_]) / Cp[r_])
                if a_ < alpha:
                    alpha, blk = a_, (r_, +1)
        x = x + alpha * pdir
        if blk is not None:
            W.append(blk)
            it += 1

    # Phase 3 -- polish: exact KKT solve on the ORIGINAL (unregularised) H
    xp, lw = eqp(H, W)
    lam = np.zeros(m_all)
    if W:
        lam[np.array([r for r, _ in W])] = lw
    x = xp
    for (r, s) in W:
        if r < n:
            wsB[r] = s
        else:
            wsC[r - n] = s
    kkt = kkt_residuals(H, g, A, lb, ub, lbA, ubA, x, lam)
    # dual feasibility of the polished multipliers
    # (scaled by the gradient of the curved variables, NOT by the 1e8 penalty multipliers)
    flat_v = np.array([np.all(H[i] == 0) for i in range(n)])
    dscale = 1.0 + (np.abs((H @ x + g)[~flat_v]).max() if (~flat_v).any() else 0.0)
    dual = 0.0
    for (r, s) in W:
        dual = max(dual, (-lam[r] if s < 0 else lam[r]) / dscale)
    kkt["dual"] = dual
    ok = (kkt["primal"] <= kkt_tol and kkt["stationarity"] <= kkt_tol
          and kkt["complementarity"] <= kkt_tol and dual <= kkt_tol)
    fval = float(0.5 * x @ H @ x + g @ x)
    return QPSolution(x, fval, 0 if ok else -1, it, lam, wsB, wsC, kkt)
