"""Prototype (numpy) of the K-form Goldfarb-Idnani iteration planned for the CUDA kernel.
Not product code; used to validate numerics/iteration counts before writing CUDA."""
import sys, pickle, numpy as np
from oracle import spline as sp, ltv, qp as oqp

def gi_kform(H, g, C, lo, up, W0, eps_flat=1e-8, tol=1e-9, max_iter=500, rule="raw", refine=True):
    n = g.size
    Hr = H.copy()
    for i in range(n):
        if H[i, i] == 0: Hr[i, i] = eps_flat
    L = np.linalg.cholesky(Hr)
    Linv = np.linalg.inv(L)
    M = Linv.T.copy()            # J = L^-T, all columns null-space (q=0)
    x = -M @ (M.T @ g)
    act = []                     # (row, sign)
    lam = []
    q = 0
    nadd = ndrop = 0
    def normal(r, s): return -s * C[r]
    def rhs(r, s): return lo[r] if s < 0 else -up[r]
    def add(nrm, y):
        nonlocal M, q
        r = y[:q].copy(); d2 = y[q:].copy()
        z = M[:, q:] @ d2
        delta2 = d2 @ d2
        delta = np.sqrt(delta2)
        sgn = 1.0 if d2[0] >= 0 else -1.0
        v = d2.copy(); v[0] += sgn * delta
        beta = 2.0 / (v @ v)
        w = M[:, q:] @ v
        M[:, q:] -= beta * np.outer(w, v)
        k = z / delta2
        M[:, :q] -= np.outer(k, r)
        M[:, q] = k
        q += 1
    def drop(l):
        nonlocal M, q
        k = M[:, l].copy()
        w = Hr @ k
        kHk = k @ w
        cols = [j for j in range(q) if j != l]
        rp = -(M[:, cols].T @ w) / kHk
        K1 = M[:, cols] + np.outer(k, rp)
        j = k / np.sqrt(kHk)
        M[:, :q-1] = K1
        # shift: new null column goes at position q-1
        M[:, q-1] = j
        q -= 1
    # initial working set
    for (r, s) in W0:
        nrm = normal(r, s)
        y = M.T @ nrm
        z = M[:, q:] @ y[q:]
        t = -(nrm @ x - rhs(r, s)) / (y[q:] @ y[q:])
        rr = y[:q]
        x = x + t * z
        lam = [lam[j] - t * rr[j] for j in range(q)] + [t]
        add(nrm, y); act.append((r, s))
    scale = np.ones(C.shape[0])
    if rule == "norm":
        # steepest-edge like: scale by sqrt(a H^-1 a)
        Y = C @ M
        scale = np.sqrt((Y * Y).sum(1))
    if rule == "eucl":
        scale = np.sqrt((C * C).sum(1))
    if callable(rule):
        scale = rule(C, M)
    it = 0
    nrefresh = 0
    while True:
        Cx = C @ x
        vlo = (Cx - lo) / scale; vup = (up - Cx) / scale
        for (r, s) in act: vlo[r] = np.inf; vup[r] = np.inf
        il, iu = int(np.argmin(vlo)), int(np.argmin(vup))
        if min(vlo[il], vup[iu]) >= -tol:
            if not refine or nrefresh >= 3:
                break
            # refresh: Newton step on the active manifold + multipliers from stationarity
            nrefresh += 1
            grad = Hr @ x + g
            y = M.T @ grad
            dx = M[:, q:] @ y[q:]
            x = x - dx
            if q:
                N = np.stack([normal(r, s) for (r, s) in act], axis=1)
                b = np.array([rhs(r, s) for (r, s) in act])
                x = x - M[:, :q] @ (N.T @ x - b)
            grad = Hr @ x + g
            lam = list(np.maximum(M[:, :q].T @ grad, 0.0))
            if np.max(np.abs(dx)) < 1e-12 * max(1.0, np.max(np.abs(x))):
                break
            continue
        p, ps = (il, -1) if vlo[il] <= vup[iu] else (iu, +1)
        nrm = normal(p, ps)
        lam_p = 0.0
        while True:
            it += 1
            if it > max_iter: return x, act, lam, it, 1, (nadd, ndrop)
            y = M.T @ nrm
            r = y[:q]; d2 = y[q:]
            delta2 = d2 @ d2
            s_p = nrm @ x - rhs(p, ps)
            t1, l = np.inf, -1
            for j in range(q):
                if r[j] > 1e-13:
                    tj = lam[j] / r[j]
                    if tj < t1: t1, l = tj, j
            lin_dep = delta2 <= 1e-13 * max(1.0, nrm @ nrm)
            t2 = np.inf if lin_dep else -s_p / delta2
            t = min(t1, t2)
            if not np.isfinite(t): return x, act, lam, it, -2, (nadd, ndrop)
            if np.isfinite(t2):
                z = M[:, q:] @ d2
                x = x + t * z
            lam = [lam[j] - t * r[j] for j in range(q)]
            lam_p += t
            if np.isfinite(t2) and t2 <= t1:
                add(nrm, y); act.append((p, ps)); lam.append(lam_p); nadd += 1
                break
            drop(l); act.pop(l); lam.pop(l); ndrop += 1
    return x, act, lam, it, 0, (nadd, ndrop)

if __name__ == "__main__":
    saved = pickle.load(open(sys.argv[1], 'rb'))
    rule = sys.argv[2] if len(sys.argv) > 2 else "raw"
    model = 'KINEMATIC' if 'KINEMATIC' in sys.argv[1] else 'DYNAMIC'
    trk = sys.argv[1].split('_')[-1].split('.')[0]
    tr = sp.Track.from_csv(f'/root/reference/data/{trk}.csv')
    errs = []; its = []; drops = []; asd = 0
    for i in sorted(saved):
        inp = saved[i]
        build = ltv.build_kinematic_qp if model == 'KINEMATIC' else ltv.build_dynamic_qp
        Q = build(inp['x0'], inp['x_ref'], tr.kappa, 0.05, inp['x_lin'], inp['u_lin'])
        n = Q['f'].size
        C, lo, up = oqp._stack(n, Q['xA'], Q['lb'], Q['ub'], Q['lbA'], Q['ubA'])
        ns = Q['n_soft']
        W0 = [(n - ns + k, -1) for k in range(ns)]
        x, act, lam, it, st, (na, nd) = gi_kform(Q['H'], Q['f'], C, lo, up, W0, rule=rule)
        uo = inp['u_opt']
        err = np.max(np.abs(x[:n-ns] - uo)) / max(1.0, np.max(np.abs(uo)))
        ws = np.zeros(C.shape[0], int)
        for (r, s) in act: ws[r] = s
        same = np.array_equal(ws[:n], inp['wsB']) and np.array_equal(ws[n:], inp['wsC'])
        asd += (not same)
        errs.append(err); its.append(it); drops.append(nd)
        if st != 0 or err > 1e-7: print('step', i, 'status', st, 'err', err, 'it', it, 'same', same)
    print('max err %.3g  mean it %.1f max it %d  drops total %d  active-set mismatches %d / %d' % (max(errs), np.mean(its), max(its), sum(drops), asd, len(errs)))
