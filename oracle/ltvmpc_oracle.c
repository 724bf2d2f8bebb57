/* ORACLE (test infrastructure only) -- plain-C restatement of the reference's kinematic
 * (and, further down, dynamic) LTV-MPC step, used (1) as the CPU baseline bench.py times on the host cores and (2) as a
 * second checker beside oracle/*.py.  The product never links or calls this file.
 *
 * It follows the reference's DENSE formulation step by step, the way the MATLAB code runs:
 *   rk2_kinematic_curvilinear.m:25-50      -> lin_rk2()
 *   sequential_integration.m:16-47         -> seq_int()      (dense A_bar, B_bar, D)
 *   kinematic_state_constraints.m:10-48    -> constraints()  (dense xA 6N x nV)
 *   kinematic_tyre_linearise_constraints.m
 *   generate_qp.m:23-33                    -> gen_qp()       (dense H = 2(B'QB+R))
 *   qpOASES(H,f,xA,lb,ub,lbA,ubA)          -> qp_solve()     (dense dual active set;
 *                                             PARITY UNPINNED vs the qpOASES binary, see
 *                                             oracle/qp.py header)
 *   ltvmpc_kinetmatic_curvilinear.m:57-60  -> outputs
 * QUIRKs of the reference are kept (B(:,:,1) on every diagonal block of B_bar).
 *
 * Build: make -C oracle   (gcc -O3 -fopenmp -shared)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define NX 5
#define NU 2
#define NS 1
#define LR 0.6183
#define LF 0.8672

typedef struct {
    const double *xs, *ys; /* [nseg x 4] column-major (MATLAB) */
    int nseg;
    double dl;
} track_t;

/* spline/interpolate_curvature.m:12-18 */
static double kappa(const track_t* tr, double s) {
    const double period = tr->dl * tr->nseg;
    double t = s - floor(s / period) * period;
    if (t < 0) t += period;
    if (t >= period) t -= period;
    int i = (int)floor(t / tr->dl);
    if (i < 0) i = 0;
    if (i >= tr->nseg) i = tr->nseg - 1;
    const double u = t / tr->dl - i, n = tr->nseg;
    (void)n;
    const double *X = tr->xs + i, *Y = tr->ys + i;
    const int st = tr->nseg;
    const double b0 = -3 * (1 - u) * (1 - u), b1 = 3 * (3 * u * u - 4 * u + 1), b2 = 3 * (2 * u - 3 * u * u), b3 = 3 * u * u;
    const double e0 = 6 * (1 - u), e1 = 6 * (3 * u - 2), e2 = 6 * (1 - 3 * u), e3 = 6 * u;
    const double Xd = (b0 * X[0] + b1 * X[st] + b2 * X[2 * st] + b3 * X[3 * st]) / tr->dl;
    const double Yd = (b0 * Y[0] + b1 * Y[st] + b2 * Y[2 * st] + b3 * Y[3 * st]) / tr->dl;
    const double Xdd = (e0 * X[0] + e1 * X[st] + e2 * X[2 * st] + e3 * X[3 * st]) / (tr->dl * tr->dl);
    const double Ydd = (e0 * Y[0] + e1 * Y[st] + e2 * Y[2 * st] + e3 * Y[3 * st]) / (tr->dl * tr->dl);
    return (Xd * Ydd - Xdd * Yd) / pow(Xd * Xd + Yd * Yd, 1.5);
}

/* f_curv_kin.m:17-29 */
static void f_kin(const double* x, const double* u, const track_t* tr, double* f) {
    const double lr_ratio = LR / (LR + LF);
    const double k = kappa(tr, x[0]);
    const double beta = atan(lr_ratio * tan(x[4]));
    const double smb = sin(x[2] + beta), cmb = cos(x[2] + beta);
    const double den = 1.0 / (1.0 - x[1] * k);
    f[0] = x[3] * cmb * den;
    f[1] = x[3] * smb;
    f[2] = x[3] * sin(beta) / LR - x[3] * cmb * den * k;
    f[3] = u[0];
    f[4] = u[1];
}

/* A_curv_kin.m:19-55 (3-argument form: no kappa_d terms), row-major 5x5 */
static void A_kin(const double* x, const track_t* tr, double* A) {
    const double lr_ratio = LR / (LR + LF);
    const double k = kappa(tr, x[0]);
    const double td = tan(x[4]);
    const double beta = atan(lr_ratio * td);
    const double smb = sin(x[2] + beta), cmb = cos(x[2] + beta);
    const double sec = 1.0 / cos(x[4]);
    const double beta_d = lr_ratio * sec * sec / (1.0 + (lr_ratio * td) * (lr_ratio * td));
    const double den = 1.0 / (1.0 - x[1] * k);
    memset(A, 0, 25 * sizeof(double));
    const double s_n = x[3] * cmb * den * den * k, s_mu = -x[3] * smb * den, s_v = cmb * den;
    const double s_d = -x[3] * smb * den * beta_d;
    A[1] = s_n; A[2] = s_mu; A[3] = s_v; A[4] = s_d;
    A[5 + 2] = x[3] * cmb; A[5 + 3] = smb; A[5 + 4] = x[3] * cmb * beta_d;
    A[10 + 1] = -s_n * k; A[10 + 2] = -s_mu * k; A[10 + 3] = sin(beta) / LR - s_v * k;
    A[10 + 4] = x[3] * cos(beta) * beta_d / LR - s_d * k;
}

/* rk2_kinematic_curvilinear.m:25-50, one step; A row-major, B row-major 5x2 */
static void lin_rk2(const double* x, const double* u, const track_t* tr, double dt, double* A, double* B, double* d) {
    double k1[NX], k2[NX], xt[NX], A1[25], A2[25];
    f_kin(x, u, tr, k1);
    for (int i = 0; i < NX; ++i) xt[i] = x[i] + k1[i] * dt / 2;
    f_kin(xt, u, tr, k2);
    A_kin(x, tr, A1);
    A_kin(xt, tr, A2);
    for (int r = 0; r < NX; ++r)
        for (int c = 0; c < NX; ++c) {
            double acc = A2[r * 5 + c];
            for (int k = 0; k < NX; ++k) acc += A2[r * 5 + k] * (A1[k * 5 + c] * dt / 2);
            A[r * 5 + c] = acc;
        }
    for (int r = 0; r < NX; ++r) {
        B[r * 2 + 0] = (r == 3 ? 1.0 : 0.0) + A2[r * 5 + 3] * dt / 2;
        B[r * 2 + 1] = (r == 4 ? 1.0 : 0.0) + A2[r * 5 + 4] * dt / 2;
    }
    for (int r = 0; r < NX; ++r) {
        double acc = k2[r];
        for (int c = 0; c < NX; ++c) acc -= A[r * 5 + c] * x[c];
        acc -= B[r * 2] * u[0] + B[r * 2 + 1] * u[1];
        d[r] = acc;
    }
}

/* ---- dense dual active-set QP (Goldfarb-Idnani, operator form M = [K1 | J2]) ----------
 * min 1/2 x'Hx + g'x  s.t. lo <= Cx <= up with C = [I; A].  n vars, m general rows.
 * Returns exitflag (0 / 1 / -2).  ws[n+m] gets -1/0/+1. */
static int qp_solve(int n, int m, const double* H /*n x n col-major*/, const double* g, const double* A /*m x n col-major*/,
                    const double* lb, const double* ub, const double* lbA, const double* ubA, int nflat, const int* flat,
                    double* x, double* fval, int* iters, int8_t* ws) {
    const int K = n + m, max_iter = 2000;
    const double tol = 1e-9, flat_eps = 1e-8;
    double* M = (double*)calloc((size_t)n * n, sizeof(double)); /* row-major */
    double* Hr = (double*)malloc((size_t)n * n * sizeof(double));
    double* L = (double*)malloc((size_t)n * n * sizeof(double));
    double* y = (double*)malloc(8 * (size_t)n * sizeof(double));
    double *z = y + n, *nv = y + 2 * n, *kv = y + 3 * n, *wv = y + 4 * n, *lam = y + 5 * n, *rp = y + 6 * n, *vv = y + 7 * n;
    int* act = (int*)malloc(n * sizeof(int));
    double* Cx = (double*)malloc(K * sizeof(double));
    int q = 0, it = 0, status = 0, nrefresh = 0;
    memset(ws, 0, K);
    memcpy(Hr, H, (size_t)n * n * sizeof(double));
    for (int i = 0; i < nflat; ++i) Hr[flat[i] * n + flat[i]] = flat_eps;
    /* Cholesky Hr = L L' (lower, row-major), then M = L^-T */
    memcpy(L, Hr, (size_t)n * n * sizeof(double));
    for (int j = 0; j < n; ++j) {
        double s = L[j * n + j];
        for (int k = 0; k < j; ++k) s -= L[j * n + k] * L[j * n + k];
        if (!(s > 0)) { status = -1; goto done; }
        const double dj = sqrt(s);
        L[j * n + j] = dj;
        for (int i = j + 1; i < n; ++i) {
            double t = L[i * n + j];
            for (int k = 0; k < j; ++k) t -= L[i * n + k] * L[j * n + k];
            L[i * n + j] = t / dj;
        }
    }
    /* Linv (lower) by forward substitution per column; M[r][c] = Linv[c][r] */
    for (int c = 0; c < n; ++c) {
        for (int i = 0; i < n; ++i) z[i] = 0.0;
        for (int i = c; i < n; ++i) {
            double t = (i == c) ? 1.0 : 0.0;
            for (int k = c; k < i; ++k) t -= L[i * n + k] * z[k];
            z[i] = t / L[i * n + i];
        }
        for (int i = c; i < n; ++i) M[c * n + i] = z[i]; /* M = Linv^T: M[c][i] = Linv[i][c] */
    }
    /* Flat variables (zero Hessian row, e.g. the slack): decoupled, so their column of
       L^-T is e_f/sqrt(eps).  Start with their bound ACTIVE (the one g pushes towards,
       multiplier |g_f|): K1 columns e_f first, J2 = the remaining columns.  Nothing of size
       g_f/eps ever enters x. */
    if (nflat > 0) {
        double* M2 = (double*)calloc((size_t)n * n, sizeof(double));
        int col = nflat;
        for (int j = 0; j < n; ++j) {
            int isflat = -1;
            for (int k = 0; k < nflat; ++k) if (flat[k] == j) isflat = k;
            if (isflat >= 0) { M2[(size_t)j * n + isflat] = 1.0; continue; }
            for (int i = 0; i < n; ++i) M2[(size_t)i * n + col] = M[(size_t)i * n + j];
            col++;
        }
        memcpy(M, M2, (size_t)n * n * sizeof(double));
        free(M2);
        for (int k = 0; k < nflat; ++k) {
            const int fv = flat[k], side = g[fv] > 0 ? -1 : +1;
            act[k] = fv * 2 + (side > 0); lam[k] = fabs(g[fv]); ws[fv] = (int8_t)side;
        }
        q = nflat;
    }
    /* x = -J2 J2' g ; flat variables on their bound */
    for (int j = 0; j < n; ++j) { double s = 0; for (int i = 0; i < n; ++i) s += M[i * n + j] * g[i]; y[j] = s; }
    for (int i = 0; i < n; ++i) { double s = 0; for (int j = q; j < n; ++j) s += M[i * n + j] * y[j]; x[i] = -s; }
    for (int k = 0; k < nflat; ++k) x[flat[k]] = ws[flat[k]] < 0 ? lb[flat[k]] : ub[flat[k]];

#define LO(c) ((c) < n ? lb[c] : lbA[(c) - n])
#define UP(c) ((c) < n ? ub[c] : ubA[(c) - n])
    for (;;) {
        int p = -1, ps = 0;
        double sp = 0;
        {
            for (int c = 0; c < n; ++c) Cx[c] = x[c];
            for (int r = 0; r < m; ++r) { double s = 0; for (int j = 0; j < n; ++j) s += A[(size_t)j * m + r] * x[j]; Cx[n + r] = s; }
            double best = -tol;
            for (int c = 0; c < K; ++c) {
                if (ws[c]) continue;
                const double vl = Cx[c] - LO(c), vu = UP(c) - Cx[c];
                if (vl < best) { best = vl; p = c; ps = -1; }
                if (vu < best) { best = vu; p = c; ps = +1; }
            }
            if (p < 0) {
                if (nrefresh >= 2) break;
                /* refresh: Newton step on the active manifold, multipliers from stationarity */
                nrefresh++;
                for (int i = 0; i < n; ++i) { double s = g[i]; for (int j = 0; j < n; ++j) s += Hr[i * n + j] * x[j]; wv[i] = s; }
                for (int j = 0; j < n; ++j) { double s = 0; for (int i = 0; i < n; ++i) s += M[i * n + j] * wv[i]; y[j] = s; }
                double dmax = 0;
                for (int i = 0; i < n; ++i) { double s = 0; for (int j = q; j < n; ++j) s += M[i * n + j] * y[j]; x[i] -= s; if (fabs(s) > dmax) dmax = fabs(s); }
                for (int j = 0; j < q; ++j) lam[j] = y[j] > 0 ? y[j] : 0;
                if (dmax < 1e-13) break;
                continue;
            }
        }
        for (int j = 0; j < n; ++j) nv[j] = (p < n) ? (j == p ? -ps : 0.0) : -ps * A[(size_t)j * m + (p - n)];
        const double bnd = ps < 0 ? LO(p) : -UP(p);
        double lam_p = 0;
        for (;;) {
            if (++it > max_iter) { status = 1; goto done; }
            double d2 = 0, nn = 0;
            sp = -bnd;
            for (int j = 0; j < n; ++j) { double s = 0; for (int i = 0; i < n; ++i) s += M[i * n + j] * nv[i]; y[j] = s; }
            for (int j = 0; j < n; ++j) { nn += nv[j] * nv[j]; sp += nv[j] * x[j]; if (j >= q) d2 += y[j] * y[j]; }
            double t1 = INFINITY; int l = -1;
            for (int j = 0; j < q; ++j) if (y[j] > 1e-13) { const double tj = lam[j] / y[j]; if (tj < t1) { t1 = tj; l = j; } }
            const int dep = !(d2 > 1e-13 * (nn > 1 ? nn : 1));
            const double t2 = dep ? INFINITY : (sp < 0 ? -sp / d2 : 0.0);
            if (isinf(t1) && isinf(t2)) { status = -2; goto done; }
            const double t = t2 <= t1 ? t2 : t1;
            if (!isinf(t2)) {
                for (int i = 0; i < n; ++i) { double s = 0; for (int j = q; j < n; ++j) s += M[i * n + j] * y[j]; z[i] = s; x[i] += t * s; }
            }
            for (int j = 0; j < q; ++j) lam[j] -= t * y[j];
            lam_p += t;
            if (t2 <= t1) { /* add p */
                const double delta = sqrt(d2), yq = y[q], sgd = yq >= 0 ? delta : -delta;
                const double beta = 1.0 / (d2 + fabs(yq) * delta);
                for (int i = 0; i < n; ++i) {
                    const double ki = z[i] / d2, wi = z[i] + sgd * M[i * n + q];
                    double* Mi = M + (size_t)i * n;
                    for (int j = 0; j < q; ++j) Mi[j] -= ki * y[j];
                    for (int j = q + 1; j < n; ++j) Mi[j] -= beta * wi * y[j];
                    Mi[q] = ki;
                }
                act[q] = p * 2 + (ps > 0); lam[q] = lam_p; ws[p] = (int8_t)ps; q++;
                break;
            }
            /* drop l */
            for (int i = 0; i < n; ++i) kv[i] = M[i * n + l];
            double kHk = 0;
            for (int i = 0; i < n; ++i) { double s = 0; for (int j = 0; j < n; ++j) s += Hr[i * n + j] * kv[j]; wv[i] = s; kHk += kv[i] * s; }
            for (int j = 0; j < q; ++j) { double s = 0; for (int i = 0; i < n; ++i) s += M[i * n + j] * wv[i]; rp[j] = -s / kHk; }
            const double rs = 1.0 / sqrt(kHk);
            for (int i = 0; i < n; ++i) {
                double* Mi = M + (size_t)i * n;
                for (int j = 0; j < q; ++j) if (j != l) Mi[j] += kv[i] * rp[j];
                if (l != q - 1) Mi[l] = Mi[q - 1];
                Mi[q - 1] = kv[i] * rs;
            }
            ws[act[l] >> 1] = 0; act[l] = act[q - 1]; lam[l] = lam[q - 1]; q--;
        }
    }
done:
    {
        double f = 0;
        for (int i = 0; i < n; ++i) { double s = 0; for (int j = 0; j < n; ++j) s += H[i * n + j] * x[j]; f += x[i] * (0.5 * s + g[i]); }
        *fval = f;
    }
    *iters = it;
    (void)vv;
    free(M); free(Hr); free(L); free(y); free(act); free(Cx);
    return status;
}

/* One kinematic LTV-MPC step (ltvmpc_kinetmatic_curvilinear.m:1-62), dense like the reference.
 * x_ref/x_lin [NX x N] column-major, u_lin [NU x N]. */
int oracle_ltvmpc_kinematic(const double* xs, const double* ys, int nseg, double dl, int N, double dt,
                            const double* x0, const double* x_ref, const double* x_lin, const double* u_lin,
                            double* u_opt, double* x_opt, int32_t* exitflag, double* fval, double* slack,
                            int32_t* iters, int8_t* wsB, int8_t* wsC) {
    const track_t tr = {xs, ys, nseg, dl};
    const int nU = NU * N, nV = nU + NS, nXN = NX * N, nC = 6 * N;
    const double Q[NX] = {5, 250, 2000, 0, 0}, R[NU] = {10, 10}, R_soft = 1e8;
    double* Ad = (double*)malloc((size_t)N * 25 * sizeof(double));
    double* dd = (double*)malloc((size_t)nXN * sizeof(double));
    double B1[NX * NU];
    double* Abar = (double*)calloc((size_t)nXN * NX, sizeof(double));   /* row-major nXN x NX */
    double* Bbar = (double*)calloc((size_t)nXN * nV, sizeof(double));   /* row-major nXN x nV */
    double* D = (double*)calloc((size_t)nXN * nXN, sizeof(double));     /* row-major */
    double* dbar = (double*)calloc(nXN, sizeof(double));
    double* e = (double*)malloc(nXN * sizeof(double));
    double* H = (double*)calloc((size_t)nV * nV, sizeof(double));
    double* f = (double*)calloc(nV, sizeof(double));
    double* xA = (double*)calloc((size_t)nC * nV, sizeof(double));      /* col-major nC x nV */
    double* lbA = (double*)malloc(nC * sizeof(double));
    double* ubA = (double*)malloc(nC * sizeof(double));
    double* lb = (double*)malloc(nV * sizeof(double));
    double* ub = (double*)malloc(nV * sizeof(double));
    double* z = (double*)malloc(nV * sizeof(double));
    int8_t* ws = (int8_t*)malloc(nV + nC);
    /* linearise + Euler discretise (sequential_integration.m:16-18) */
    for (int k = 0; k < N; ++k) {
        double A[25], B[10], d[NX];
        lin_rk2(x_lin + k * NX, u_lin + k * NU, &tr, dt, A, B, d);
        for (int r = 0; r < NX; ++r)
            for (int c = 0; c < NX; ++c) Ad[k * 25 + r * 5 + c] = A[r * 5 + c] * dt + (r == c);
        for (int r = 0; r < NX; ++r) dd[k * NX + r] = d[r] * dt;
        if (k == 0) for (int i = 0; i < 10; ++i) B1[i] = B[i] * dt;
    }
    /* A_bar (sequential_integration.m:21-26) */
    for (int r = 0; r < NX; ++r) for (int c = 0; c < NX; ++c) Abar[r * NX + c] = Ad[r * 5 + c];
    for (int k = 1; k < N; ++k)
        for (int r = 0; r < NX; ++r)
            for (int c = 0; c < NX; ++c) {
                double s = 0;
                for (int j = 0; j < NX; ++j) s += Ad[k * 25 + r * 5 + j] * Abar[((k - 1) * NX + j) * NX + c];
                Abar[(k * NX + r) * NX + c] = s;
            }
    /* B_bar (sequential_integration.m:28-36), QUIRK B(:,:,1) */
    for (int i = 0; i < N; ++i) {
        for (int r = 0; r < NX; ++r) for (int c = 0; c < NU; ++c) Bbar[(size_t)(i * NX + r) * nV + i * NU + c] = B1[r * 2 + c];
        for (int j = i + 1; j < N; ++j)
            for (int r = 0; r < NX; ++r)
                for (int c = 0; c < NU; ++c) {
                    double s = 0;
                    for (int k = 0; k < NX; ++k) s += Ad[j * 25 + r * 5 + k] * Bbar[(size_t)((j - 1) * NX + k) * nV + i * NU + c];
                    Bbar[(size_t)(j * NX + r) * nV + i * NU + c] = s;
                }
    }
    /* D and d_bar (sequential_integration.m:38-47) */
    for (int i = 0; i < N; ++i) {
        for (int r = 0; r < NX; ++r) D[(size_t)(i * NX + r) * nXN + i * NX + r] = 1.0;
        for (int j = i + 1; j < N; ++j)
            for (int r = 0; r < NX; ++r)
                for (int c = 0; c < NX; ++c) {
                    double s = 0;
                    for (int k = 0; k < NX; ++k) s += Ad[j * 25 + r * 5 + k] * D[(size_t)((j - 1) * NX + k) * nXN + i * NX + c];
                    D[(size_t)(j * NX + r) * nXN + i * NX + c] = s;
                }
    }
    for (int r = 0; r < nXN; ++r) { double s = 0; for (int c = 0; c <= r / NX * NX + NX - 1 && c < nXN; ++c) s += D[(size_t)r * nXN + c] * dd[c]; dbar[r] = s; }
    /* free response and tracking error */
    for (int r = 0; r < nXN; ++r) {
        double s = dbar[r];
        for (int c = 0; c < NX; ++c) s += Abar[r * NX + c] * x0[c];
        x_opt[r] = s;              /* holds A_bar x0 + d_bar for now */
        e[r] = s - x_ref[r];
    }
    /* generate_qp.m:23-33 */
    for (int i = 0; i < nU; ++i)
        for (int j = 0; j <= i; ++j) {
            double s = 0;
            for (int r = (i / NU) * NX; r < nXN; ++r) {
                const int k = r / NX, c = r - k * NX;
                const double qq = (k == N - 1 ? 10 : 1) * Q[c];
                if (qq != 0) s += qq * Bbar[(size_t)r * nV + i] * Bbar[(size_t)r * nV + j];
            }
            if (i == j) s += R[i % NU];
            H[i * nV + j] = H[j * nV + i] = 2 * s;
        }
    double cconst = 0;
    for (int r = 0; r < nXN; ++r) { const int k = r / NX, c = r - k * NX; cconst += (k == N - 1 ? 10 : 1) * Q[c] * e[r] * e[r]; }
    for (int j = 0; j < nU; ++j) {
        double s = 0;
        for (int r = (j / NU) * NX; r < nXN; ++r) { const int k = r / NX, c = r - k * NX; s += (k == N - 1 ? 10 : 1) * Q[c] * Bbar[(size_t)r * nV + j] * e[r]; }
        f[j] = 2 * s;
    }
    f[nU] = R_soft;
    /* kinematic_state_constraints.m + kinematic_tyre_linearise_constraints.m */
    for (int k = 0; k < N; ++k) {
        const double* xl = x_lin + k * NX;
        const double c4 = 2 * xl[3] * xl[4] / (LF + LR), c5 = xl[3] * xl[3] / (LF + LR), g0 = xl[3] * xl[3] * xl[4] / (LR + LF);
        const double* xf = x_opt + k * NX;
        for (int j = 0; j < nU; ++j) {
            const double bv = Bbar[(size_t)(k * NX + 3) * nV + j], bd = Bbar[(size_t)(k * NX + 4) * nV + j], bn = Bbar[(size_t)(k * NX + 1) * nV + j];
            xA[(size_t)j * nC + k] = bv;
            xA[(size_t)j * nC + N + k] = bd;
            xA[(size_t)j * nC + 2 * N + k] = bn;
            xA[(size_t)j * nC + 3 * N + k] = bn;
            xA[(size_t)j * nC + 4 * N + k] = c4 * bv + c5 * bd;
            xA[(size_t)j * nC + 5 * N + k] = c4 * bv + c5 * bd;
        }
        xA[(size_t)nU * nC + 2 * N + k] = 1; xA[(size_t)nU * nC + 3 * N + k] = -1;
        xA[(size_t)nU * nC + 4 * N + k] = 1; xA[(size_t)nU * nC + 5 * N + k] = -1;
        lbA[k] = 0 - xf[3]; ubA[k] = INFINITY;
        lbA[N + k] = -0.4 - xf[4]; ubA[N + k] = 0.4 - xf[4];
        lbA[2 * N + k] = -0.75 - xf[1]; ubA[2 * N + k] = 1e10;
        lbA[3 * N + k] = -1e10; ubA[3 * N + k] = 0.75 - xf[1];
        const double cc = g0 + c4 * (xf[3] - xl[3]) + c5 * (xf[4] - xl[4]);
        lbA[4 * N + k] = -5.0 - cc; ubA[4 * N + k] = INFINITY;
        lbA[5 * N + k] = -INFINITY; ubA[5 * N + k] = 5.0 - cc;
    }
    for (int j = 0; j < nU; ++j) { lb[j] = (j % 2) ? -0.4 : -10; ub[j] = (j % 2) ? 0.4 : 10; }
    lb[nU] = 0; ub[nU] = INFINITY;
    const int flat[1] = {nU};
    double fv; int it;
    const int st = qp_solve(nV, nC, H, f, xA, lb, ub, lbA, ubA, 1, flat, z, &fv, &it, ws);
    /* ltvmpc_kinetmatic_curvilinear.m:57-60 */
    for (int r = 0; r < nXN; ++r) { double s = x_opt[r]; for (int j = 0; j < nU; ++j) s += Bbar[(size_t)r * nV + j] * z[j]; x_opt[r] = s; }
    memcpy(u_opt, z, nU * sizeof(double));
    slack[0] = z[nU];
    *fval = fv + cconst;
    *exitflag = st;
    if (iters) *iters = it;
    if (wsB) memcpy(wsB, ws, nV);
    if (wsC) memcpy(wsC, ws + nV, nC);
    free(Ad); free(dd); free(Abar); free(Bbar); free(D); free(dbar); free(e); free(H); free(f); free(xA);
    free(lbA); free(ubA); free(lb); free(ub); free(z); free(ws);
    return st;
}

/* ======================================================================================
 * Dynamic (tyre-force) model: ltvmpc_dynamic_curvilinear.m:1-62, dense like the reference.
 *   f_curv_dyn.m:20-63, A_curv_dyn.m:20-106   -> f_dyn(), A_dyn()
 *   rk4_dynamic_curvilinear.m:25-57           -> lin_rk4_dyn()   (QUIRK: dt/2 in the 4th control sensitivity)
 *   dynamic_state_constraints.m:10-57 + dynamic_slip_linearise_constraints.m + dynamic_tyre_linearise_constraints.m
 * ====================================================================================== */
#define DX 7
#define DS 4
#define MASS 280.0
#define INERTIA 200.0
#define GRAV 9.81
#define PB 12.56
#define PC 1.38
#define PD 1.60
#define PE (-0.58)

static double pacejka(double alpha, double Fz) {
    const double bt = PB * alpha, inner = bt - PE * (bt - atan(bt));
    return Fz * PD * sin(PC * atan(inner));
}
static double pacejka_d(double alpha, double Fz) {
    const double bt = PB * alpha, inner = bt - PE * (bt - atan(bt));
    return Fz * PD * cos(PC * atan(inner)) * PC / (1 + inner * inner) * (PB - PE * (PB - PB / (1 + PB * PB * alpha * alpha)));
}
typedef struct { double Fcr, Fcr_d, vr, dvr2, xdh, xdhd, vf, dvf2; } dyn_aux_t;

/* f_curv_dyn.m:20-63 */
static void f_dyn(const double* x, const double* u, const track_t* tr, double* f) {
    const double n = x[1], mu = x[2], x_d = x[3], y_d = x[4], th = x[5], delta = x[6];
    const double Fx = u[0] * MASS;
    const double xdh = x_d + 5 * exp(-x_d / 5);
    const double k = kappa(tr, x[0]), den = 1 / (1 - n * k);
    const double af = delta - atan((y_d + LF * th) / xdh), ar = -atan((y_d - LR * th) / xdh);
    const double Fzf = MASS * GRAV * LR / (LR + LF), Fzr = MASS * GRAV * LF / (LR + LF);
    const double Fcf = pacejka(af, Fzf), Fcr = pacejka(ar, Fzr);
    const double sd_ = (x_d * cos(mu) - y_d * sin(mu)) * den;
    f[0] = sd_;
    f[1] = x_d * sin(mu) + y_d * cos(mu);
    f[2] = th - sd_ * k;
    f[3] = (Fx - Fcf * sin(delta) + MASS * y_d * th) / MASS;
    f[4] = (Fcr + Fcf * cos(delta) - MASS * x_d * th) / MASS;
    f[5] = (LF * Fcf * cos(delta) - LR * Fcr) / INERTIA;
    f[6] = u[1];
}

/* A_curv_dyn.m:20-106, row-major 7x7 (+ the auxiliaries the constraint linearisations reuse) */
static void A_dyn(const double* x, const track_t* tr, double* A, dyn_aux_t* ax) {
    const double m = MASS, I = INERTIA, lr = LR, lf = LF;
    const double n = x[1], mu = x[2], x_d = x[3], y_d = x[4], th = x[5], delta = x[6];
    const double xdh = x_d + 5 * exp(-x_d / 5), xdhd = 1 - exp(-x_d / 5);
    const double af = delta - atan((y_d + lf * th) / xdh), ar = -atan((y_d - lr * th) / xdh);
    const double Fzf = m * GRAV * lr / (lr + lf), Fzr = m * GRAV * lf / (lr + lf);
    const double Fcf = pacejka(af, Fzf), Fcr = pacejka(ar, Fzr);
    const double Fcf_d = pacejka_d(af, Fzf), Fcr_d = pacejka_d(ar, Fzr);
    const double k = kappa(tr, x[0]), den = 1 / (1 - n * k);
    const double vf = (y_d + lf * th) / xdh, vr = (y_d - lr * th) / xdh;
    const double dvf2 = 1 / (1 + vf * vf), dvr2 = 1 / (1 + vr * vr);
    const double sm = sin(mu), cm = cos(mu), sd = sin(delta), cd = cos(delta);
    memset(A, 0, 49 * sizeof(double));
    const double s_n = (x_d * cm - y_d * sm) * den * den * k, s_mu = (-x_d * sm - y_d * cm) * den;
    const double s_xd = cm * den, s_yd = -sm * den;
    A[1] = s_n; A[2] = s_mu; A[3] = s_xd; A[4] = s_yd;
    A[7 + 2] = x_d * cm - y_d * sm; A[7 + 3] = sm; A[7 + 4] = cm;
    A[14 + 1] = -s_n * k; A[14 + 2] = -s_mu * k; A[14 + 3] = -s_xd * k; A[14 + 4] = -s_yd * k; A[14 + 5] = 1.0;
    A[21 + 3] = -Fcf_d * dvf2 * vf * sd * xdhd / (m * xdh);
    A[21 + 4] = (Fcf_d * dvf2 * sd / xdh + m * th) / m;
    A[21 + 5] = (Fcf_d * dvf2 * lf * sd / xdh + m * y_d) / m;
    A[21 + 6] = (-Fcf * cd - Fcf_d * sd) / m;
    A[28 + 3] = (Fcr_d * dvr2 * vr * xdhd / xdh + Fcf_d * dvf2 * vf * cd * xdhd / xdh - m * th) / m;
    A[28 + 4] = (-Fcr_d * dvr2 / xdh - Fcf_d * dvf2 / xdh * cd) / m;
    A[28 + 5] = (Fcr_d * dvr2 * lr / xdh - Fcf_d * dvf2 * lf / xdh * cd - m * xdh) / m;
    A[28 + 6] = (-Fcf * sd + Fcf_d * cd) / m;
    A[35 + 3] = (lf * Fcf_d * dvf2 * vf * cd * xdhd / xdh - lr * Fcr_d * dvr2 * vr * xdhd / xdh) / I;
    A[35 + 4] = (-lf * Fcf_d * dvf2 * cd / xdh + lr * Fcr_d * dvr2 / xdh) / I;
    A[35 + 5] = (-lf * Fcf_d * dvf2 * lf * cd / xdh - lr * Fcr_d * dvr2 * lr / xdh) / I;
    A[35 + 6] = (-lf * Fcf * sd + lf * Fcf_d * cd) / I;
    if (ax) { ax->Fcr = Fcr; ax->Fcr_d = Fcr_d; ax->vr = vr; ax->dvr2 = dvr2; ax->xdh = xdh; ax->xdhd = xdhd; ax->vf = vf; ax->dvf2 = dvf2; }
}

static void mm7(const double* P, const double* Q, double s, double* out /* = P (I + s Q) */) {
    for (int r = 0; r < DX; ++r)
        for (int c = 0; c < DX; ++c) {
            double acc = P[r * DX + c];
            for (int k = 0; k < DX; ++k) acc += P[r * DX + k] * (Q[k * DX + c] * s);
            out[r * DX + c] = acc;
        }
}
/* rk4_dynamic_curvilinear.m:25-57; A row-major 7x7, B row-major 7x2 */
static void lin_rk4_dyn(const double* x, const double* u, const track_t* tr, double dt, double* A, double* B, double* d) {
    double k1[DX], k2[DX], k3[DX], k4[DX], xt[DX], f[DX];
    double F1[49], F2[49], F3[49], F4[49], K2[49], K3[49], K4[49];
    const double B0[DX * NU] = {0, 0, 0, 0, 0, 0, 1, 0, 0, 0, 0, 0, 0, 1};      /* B_curv_dyn.m: rows 4 and 7 */
    double U2[DX * NU], U3[DX * NU], U4[DX * NU];
    f_dyn(x, u, tr, k1);
    A_dyn(x, tr, F1, 0);
    for (int i = 0; i < DX; ++i) xt[i] = x[i] + k1[i] * dt / 2;
    f_dyn(xt, u, tr, k2); A_dyn(xt, tr, F2, 0);
    for (int i = 0; i < DX; ++i) xt[i] = x[i] + k2[i] * dt / 2;
    f_dyn(xt, u, tr, k3); A_dyn(xt, tr, F3, 0);
    for (int i = 0; i < DX; ++i) xt[i] = x[i] + k3[i] * dt;
    f_dyn(xt, u, tr, k4); A_dyn(xt, tr, F4, 0);
    mm7(F2, F1, dt / 2, K2);
    mm7(F3, K2, dt / 2, K3);
    mm7(F4, K3, dt, K4);
    for (int r = 0; r < DX; ++r)
        for (int c = 0; c < NU; ++c) {
            double a2 = B0[r * NU + c], a3, a4;
            for (int k = 0; k < DX; ++k) a2 += F2[r * DX + k] * B0[k * NU + c] * dt / 2;
            U2[r * NU + c] = a2;
            (void)a3; (void)a4;
        }
    for (int r = 0; r < DX; ++r)
        for (int c = 0; c < NU; ++c) {
            double a3 = B0[r * NU + c];
            for (int k = 0; k < DX; ++k) a3 += F3[r * DX + k] * U2[k * NU + c] * dt / 2;
            U3[r * NU + c] = a3;
        }
    for (int r = 0; r < DX; ++r)
        for (int c = 0; c < NU; ++c) {
            double a4 = B0[r * NU + c];
            for (int k = 0; k < DX; ++k) a4 += F4[r * DX + k] * U3[k * NU + c] * dt / 2;     /* QUIRK rk4_*:52 */
            U4[r * NU + c] = a4;
        }
    for (int i = 0; i < 49; ++i) A[i] = (F1[i] + 2 * K2[i] + 2 * K3[i] + K4[i]) / 6;
    for (int i = 0; i < DX * NU; ++i) B[i] = (B0[i] + 2 * U2[i] + 2 * U3[i] + U4[i]) / 6;
    for (int r = 0; r < DX; ++r) {
        f[r] = (k1[r] + 2 * k2[r] + 2 * k3[r] + k4[r]) / 6;
        double acc = f[r];
        for (int c = 0; c < DX; ++c) acc -= A[r * DX + c] * x[c];
        acc -= B[r * NU] * u[0] + B[r * NU + 1] * u[1];
        d[r] = acc;
    }
}

/* One dynamic LTV-MPC step.  x_ref/x_lin [7 x N] column-major, u_lin [2 x N]. */
int oracle_ltvmpc_dynamic(const double* xs, const double* ys, int nseg, double dl, int N, double dt,
                          const double* x0, const double* x_ref, const double* x_lin, const double* u_lin,
                          double* u_opt, double* x_opt, int32_t* exitflag, double* fval, double* slack,
                          int32_t* iters, int8_t* wsB, int8_t* wsC) {
    const track_t tr = {xs, ys, nseg, dl};
    const int nU = NU * N, nV = nU + DS, nXN = DX * N, nC = 20 * N, NP = 12;
    const double Q[DX] = {5, 250, 2000, 0, 0, 0, 0}, R[NU] = {10, 10}, R_soft[DS] = {1e8, 1e6, 1e6, 1e4};
    double* Ad = (double*)malloc((size_t)N * 49 * sizeof(double));
    double* dd = (double*)malloc((size_t)nXN * sizeof(double));
    double B1[DX * NU];
    double* Abar = (double*)calloc((size_t)nXN * DX, sizeof(double));
    double* Bbar = (double*)calloc((size_t)nXN * nV, sizeof(double));
    double* xf = (double*)calloc(nXN, sizeof(double));
    double* e = (double*)malloc(nXN * sizeof(double));
    double* H = (double*)calloc((size_t)nV * nV, sizeof(double));
    double* f = (double*)calloc(nV, sizeof(double));
    double* xA = (double*)calloc((size_t)nC * nV, sizeof(double));
    double* lbA = (double*)malloc(nC * sizeof(double));
    double* ubA = (double*)malloc(nC * sizeof(double));
    double* lb = (double*)malloc(nV * sizeof(double));
    double* ub = (double*)malloc(nV * sizeof(double));
    double* z = (double*)malloc(nV * sizeof(double));
    int8_t* ws = (int8_t*)malloc(nV + nC);
    for (int k = 0; k < N; ++k) {
        double A[49], B[DX * NU], d[DX];
        lin_rk4_dyn(x_lin + k * DX, u_lin + k * NU, &tr, dt, A, B, d);
        for (int r = 0; r < DX; ++r)
            for (int c = 0; c < DX; ++c) Ad[k * 49 + r * DX + c] = A[r * DX + c] * dt + (r == c);
        for (int r = 0; r < DX; ++r) dd[k * DX + r] = d[r] * dt;
        if (k == 0) for (int i = 0; i < DX * NU; ++i) B1[i] = B[i] * dt;
    }
    /* A_bar, free response (A_bar x0 + d_bar by the recursion x_f,k = A_k x_f,k-1 + d_k: the same numbers as
       D * d of sequential_integration.m:38-47 up to round-off) */
    for (int r = 0; r < DX; ++r) for (int c = 0; c < DX; ++c) Abar[r * DX + c] = Ad[r * DX + c];
    for (int k = 1; k < N; ++k)
        for (int r = 0; r < DX; ++r)
            for (int c = 0; c < DX; ++c) {
                double s = 0;
                for (int j = 0; j < DX; ++j) s += Ad[k * 49 + r * DX + j] * Abar[((k - 1) * DX + j) * DX + c];
                Abar[(k * DX + r) * DX + c] = s;
            }
    for (int k = 0; k < N; ++k)
        for (int r = 0; r < DX; ++r) {
            double s = dd[k * DX + r];
            for (int c = 0; c < DX; ++c) s += Ad[k * 49 + r * DX + c] * (k ? xf[(k - 1) * DX + c] : x0[c]);
            xf[k * DX + r] = s;
        }
    for (int i = 0; i < N; ++i) {
        for (int r = 0; r < DX; ++r) for (int c = 0; c < NU; ++c) Bbar[(size_t)(i * DX + r) * nV + i * NU + c] = B1[r * NU + c];
        for (int j = i + 1; j < N; ++j)
            for (int r = 0; r < DX; ++r)
                for (int c = 0; c < NU; ++c) {
                    double s = 0;
                    for (int k = 0; k < DX; ++k) s += Ad[j * 49 + r * DX + k] * Bbar[(size_t)((j - 1) * DX + k) * nV + i * NU + c];
                    Bbar[(size_t)(j * DX + r) * nV + i * NU + c] = s;
                }
    }
    for (int r = 0; r < nXN; ++r) e[r] = xf[r] - x_ref[r];
    /* generate_qp.m:23-33 */
    for (int i = 0; i < nU; ++i)
        for (int j = 0; j <= i; ++j) {
            double s = 0;
            for (int r = (i / NU) * DX; r < nXN; ++r) {
                const int k = r / DX, c = r - k * DX;
                const double qq = (k == N - 1 ? 10 : 1) * Q[c];
                if (qq != 0) s += qq * Bbar[(size_t)r * nV + i] * Bbar[(size_t)r * nV + j];
            }
            if (i == j) s += R[i % NU];
            H[i * nV + j] = H[j * nV + i] = 2 * s;
        }
    double cconst = 0;
    for (int r = 0; r < nXN; ++r) { const int k = r / DX, c = r - k * DX; cconst += (k == N - 1 ? 10 : 1) * Q[c] * e[r] * e[r]; }
    for (int j = 0; j < nU; ++j) {
        double s = 0;
        for (int r = (j / NU) * DX; r < nXN; ++r) { const int k = r / DX, c = r - k * DX; s += (k == N - 1 ? 10 : 1) * Q[c] * Bbar[(size_t)r * nV + j] * e[r]; }
        f[j] = 2 * s;
    }
    for (int j = 0; j < DS; ++j) f[nU + j] = R_soft[j];
    /* polygon of dynamic_tyre_linearise_constraints.m:18-23 */
    double acl[13], all_[13];
    for (int j = 0; j <= NP; ++j) { const double th = 2 * M_PI * j / NP; acl[j] = 9.163 * sin(th); all_[j] = 10.0 * cos(th); }
    for (int k = 0; k < N; ++k) {
        const double* xl = x_lin + k * DX;
        const double* ul = u_lin + k * NU;
        const double* xk = xf + k * DX;
        double Atmp[49];
        dyn_aux_t ax;
        A_dyn(xl, &tr, Atmp, &ax);
        const double cr[3] = {ax.dvr2 * ax.vr * ax.xdhd / ax.xdh, -ax.dvr2 / ax.xdh, ax.dvr2 * LR / ax.xdh};
        const double cf[3] = {ax.dvf2 * ax.vf * ax.xdhd / ax.xdh, -ax.dvf2 / ax.xdh, -ax.dvf2 * LF / ax.xdh};
        const double ct[3] = {-ax.Fcr_d * ax.dvr2 * ax.vr * ax.xdhd / ax.xdh / MASS, ax.Fcr_d * ax.dvr2 / ax.xdh / MASS,
                              -ax.Fcr_d * ax.dvr2 * LR / ax.xdh / MASS};
        for (int j = 0; j < nU; ++j) {
            const double bn = Bbar[(size_t)(k * DX + 1) * nV + j], bx = Bbar[(size_t)(k * DX + 3) * nV + j];
            const double by = Bbar[(size_t)(k * DX + 4) * nV + j], bt = Bbar[(size_t)(k * DX + 5) * nV + j];
            const double bd = Bbar[(size_t)(k * DX + 6) * nV + j];
            xA[(size_t)j * nC + k] = bx;
            xA[(size_t)j * nC + N + k] = bd;
            xA[(size_t)j * nC + 2 * N + k] = bn;
            xA[(size_t)j * nC + 3 * N + k] = bn;
            const double ar_ = cr[0] * bx + cr[1] * by + cr[2] * bt, af_ = cf[0] * bx + cf[1] * by + cf[2] * bt + bd;
            xA[(size_t)j * nC + 4 * N + 2 * k] = ar_; xA[(size_t)j * nC + 4 * N + 2 * k + 1] = af_;
            xA[(size_t)j * nC + 6 * N + 2 * k] = ar_; xA[(size_t)j * nC + 6 * N + 2 * k + 1] = af_;
            const double tb = ct[0] * bx + ct[1] * by + ct[2] * bt;
            for (int p = 0; p < NP; ++p)
                xA[(size_t)j * nC + 8 * N + NP * k + p] = (all_[p + 1] - all_[p]) * tb + ((j == NU * k) ? (acl[p + 1] - acl[p]) : 0.0);
        }
        xA[(size_t)nU * nC + 2 * N + k] = 1; xA[(size_t)nU * nC + 3 * N + k] = -1;
        xA[(size_t)(nU + 1) * nC + 4 * N + 2 * k] = 1; xA[(size_t)(nU + 2) * nC + 4 * N + 2 * k + 1] = 1;
        xA[(size_t)(nU + 1) * nC + 6 * N + 2 * k] = -1; xA[(size_t)(nU + 2) * nC + 6 * N + 2 * k + 1] = -1;
        for (int p = 0; p < NP; ++p) xA[(size_t)(nU + 3) * nC + 8 * N + NP * k + p] = -1;
        lbA[k] = 0 - xk[3]; ubA[k] = INFINITY;
        lbA[N + k] = -0.4 - xk[6]; ubA[N + k] = 0.4 - xk[6];
        lbA[2 * N + k] = -0.75 - xk[1]; ubA[2 * N + k] = 1e10;
        lbA[3 * N + k] = -1e10; ubA[3 * N + k] = 0.75 - xk[1];
        const double d3 = xk[3] - xl[3], d4 = xk[4] - xl[4], d5 = xk[5] - xl[5];
        const double c_r = -atan(ax.vr) + cr[0] * d3 + cr[1] * d4 + cr[2] * d5;
        const double c_f = xl[6] - atan(ax.vf) + cf[0] * d3 + cf[1] * d4 + cf[2] * d5 + (xk[6] - xl[6]);
        lbA[4 * N + 2 * k] = -0.1 - c_r; ubA[4 * N + 2 * k] = INFINITY;
        lbA[4 * N + 2 * k + 1] = -0.1 - c_f; ubA[4 * N + 2 * k + 1] = INFINITY;
        lbA[6 * N + 2 * k] = -INFINITY; ubA[6 * N + 2 * k] = 0.1 - c_r;
        lbA[6 * N + 2 * k + 1] = -INFINITY; ubA[6 * N + 2 * k + 1] = 0.1 - c_f;
        for (int p = 0; p < NP; ++p) {
            const double dac = acl[p + 1] - acl[p], dal = all_[p + 1] - all_[p];
            const double g0 = (ul[0] - all_[p]) * dac - (ax.Fcr / 280 - acl[p]) * dal;
            const double cc = g0 + dal * (ct[0] * d3 + ct[1] * d4 + ct[2] * d5) - dac * ul[0];
            lbA[8 * N + NP * k + p] = -INFINITY; ubA[8 * N + NP * k + p] = 0.0 - cc;
        }
    }
    for (int j = 0; j < nU; ++j) { lb[j] = (j % 2) ? -0.4 : -10; ub[j] = (j % 2) ? 0.4 : 10; }
    for (int j = 0; j < DS; ++j) { lb[nU + j] = 0; ub[nU + j] = INFINITY; }
    const int flat[DS] = {nU, nU + 1, nU + 2, nU + 3};
    double fv; int it;
    const int st = qp_solve(nV, nC, H, f, xA, lb, ub, lbA, ubA, DS, flat, z, &fv, &it, ws);
    for (int r = 0; r < nXN; ++r) { double s = xf[r]; for (int j = 0; j < nU; ++j) s += Bbar[(size_t)r * nV + j] * z[j]; x_opt[r] = s; }
    memcpy(u_opt, z, nU * sizeof(double));
    for (int j = 0; j < DS; ++j) slack[j] = z[nU + j];
    *fval = fv + cconst;
    *exitflag = st;
    if (iters) *iters = it;
    if (wsB) memcpy(wsB, ws, nV);
    if (wsC) memcpy(wsC, ws + nV, nC);
    free(Ad); free(dd); free(Abar); free(Bbar); free(xf); free(e); free(H); free(f); free(xA);
    free(lbA); free(ubA); free(lb); free(ub); free(z); free(ws);
    return st;
}

int oracle_ltvmpc_dynamic_batch(const double* xs, const double* ys, int nseg, double dl, int B, int N, double dt,
                                const double* x0, const double* x_ref, const double* x_lin, const double* u_lin,
                                double* u_opt, double* x_opt, int32_t* exitflag, double* fval, double* slack,
                                int32_t* iters, int nthreads) {
    int used = 1;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
    used = omp_get_max_threads();
#endif
#pragma omp parallel for schedule(dynamic, 2)
    for (int b = 0; b < B; ++b)
        oracle_ltvmpc_dynamic(xs, ys, nseg, dl, N, dt, x0 + (size_t)b * DX, x_ref + (size_t)b * DX * N,
                              x_lin + (size_t)b * DX * N, u_lin + (size_t)b * NU * N, u_opt + (size_t)b * NU * N,
                              x_opt + (size_t)b * DX * N, exitflag + b, fval + b, slack + (size_t)b * DS, iters ? iters + b : 0, 0, 0);
    return used;
}

/* Batch over B problems in the C-ABI layout (batch trailing), OpenMP over problems. */
int oracle_ltvmpc_kinematic_batch(const double* xs, const double* ys, int nseg, double dl, int B, int N, double dt,
                                  const double* x0, const double* x_ref, const double* x_lin, const double* u_lin,
                                  double* u_opt, double* x_opt, int32_t* exitflag, double* fval, double* slack,
                                  int32_t* iters, int nthreads) {
    int used = 1;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
    used = omp_get_max_threads();
#endif
#pragma omp parallel for schedule(dynamic, 4)
    for (int b = 0; b < B; ++b)
        oracle_ltvmpc_kinematic(xs, ys, nseg, dl, N, dt, x0 + (size_t)b * NX, x_ref + (size_t)b * NX * N,
                                x_lin + (size_t)b * NX * N, u_lin + (size_t)b * NU * N, u_opt + (size_t)b * NU * N,
                                x_opt + (size_t)b * NX * N, exitflag + b, fval + b, slack + b, iters ? iters + b : 0, 0, 0);
    return used;
}
