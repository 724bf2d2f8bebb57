"""ORACLE (test infrastructure only) -- numpy restatement of reference mpc/ltv/*.

Shapes follow MATLAB: x is (N_x, N_steps), A is (N_x, N_x, N_steps), stacked
vectors are column-major flattenings (`x(:)` == x.flatten(order='F')).
Known quirks of the reference are KEPT and flagged "QUIRK" -- the CUDA path has
to reproduce the reference's numbers, not a corrected model.
"""
import math
import numpy as np

from . import vehicle as vm
from . import qp as _qp

INF = np.inf


# --------------------------------------------------------------------------- linearise
def _rk_linearise(f_fun, A_fun, B_fun, x, u, kappa, dt, order):
    """mpc/ltv/kinematic/{euler,rk2,rk4}_kinematic_curvilinear.m and
    mpc/ltv/dynamic/{euler,rk2,rk4}_dynamic_curvilinear.m share this body."""
    N_x, N_steps = x.shape
    N_u = u.shape[0]
    I = np.eye(N_x)
    A = np.zeros((N_x, N_x, N_steps))
    B = np.zeros((N_x, N_u, N_steps))
    d = np.zeros((N_x, N_steps))
    for i in range(N_steps):
        x_i = x[:, i].copy()
        u_i = u[:, i].copy()
        if order == 1:          # euler_*_curvilinear.m:24-30
            Ai = A_fun(x_i, u_i, kappa)
            Bi = B_fun(x_i, u_i, kappa)
            f = f_fun(x_i, u_i, kappa)
        elif order == 2:        # rk2_*_curvilinear.m:25-48
            k1 = f_fun(x_i, u_i, kappa)
            k2 = f_fun(x_i + k1 * dt / 2, u_i, kappa)
            f = k2
            dfdx1 = A_fun(x_i, u_i, kappa)
            dfdx2 = A_fun(x_i + k1 * dt / 2, u_i, kappa)
            dkdx1 = dfdx1
            dkdx2 = dfdx2 @ (I + dkdx1 * dt / 2)
            dkdu1 = B_fun(x_i, u_i, kappa)
            dkdu2 = B_fun(x_i + k1 * dt / 2, u_i, kappa) + dfdx2 @ dkdu1 * dt / 2
            Ai, Bi = dkdx2, dkdu2
        elif order == 4:        # rk4_*_curvilinear.m:25-57
            k1 = f_fun(x_i, u_i, kappa)
            k2 = f_fun(x_i + k1 * dt / 2, u_i, kappa)
            k3 = f_fun(x_i + k2 * dt / 2, u_i, kappa)
            k4 = f_fun(x_i + k3 * dt, u_i, kappa)
            f = (k1 + 2 * k2 + 2 * k3 + k4) / 6
            dfdx1 = A_fun(x_i, u_i, kappa)
            dfdx2 = A_fun(x_i + k1 * dt / 2, u_i, kappa)
            dfdx3 = A_fun(x_i + k2 * dt / 2, u_i, kappa)
            dfdx4 = A_fun(x_i + k3 * dt, u_i, kappa)
            dkdx1 = dfdx1
            dkdx2 = dfdx2 @ (I + dkdx1 * dt / 2)
            dkdx3 = dfdx3 @ (I + dkdx2 * dt / 2)
            dkdx4 = dfdx4 @ (I + dkdx3 * dt)
            dkdu1 = B_fun(x_i, u_i, kappa)
            dkdu2 = B_fun(x_i + k1 * dt / 2, u_i, kappa) + dfdx2 @ dkdu1 * dt / 2
            dkdu3 = B_fun(x_i + k2 * dt / 2, u_i, kappa) + dfdx3 @ dkdu2 * dt / 2
            # QUIRK rk4_*_curvilinear.m:52: dt/2 (not dt) in the 4th control sensitivity
            dkdu4 = B_fun(x_i + k3 * dt, u_i, kappa) + dfdx4 @ dkdu3 * dt / 2
            Ai = (dkdx1 + 2 * dkdx2 + 2 * dkdx3 + dkdx4) / 6
            Bi = (dkdu1 + 2 * dkdu2 + 2 * dkdu3 + dkdu4) / 6
        else:
            raise ValueError(order)
        A[:, :, i] = Ai
        B[:, :, i] = Bi
        d[:, i] = f - Ai @ x_i - Bi @ u_i
    return A, B, d


def _f_dyn(x, u, kappa):
    return vm.f_curv_dyn(x, u, kappa)[0]


def _A_dyn(x, u, kappa):
    return vm.A_curv_dyn(x, u, kappa)[0]


def euler_kinematic_curvilinear(x, u, kappa, dt=None):
    return _rk_linearise(vm.f_curv_kin, vm.A_curv_kin, vm.B_curv_kin, x, u, kappa, dt, 1)


def rk2_kinematic_curvilinear(x, u, kappa, dt):
    return _rk_linearise(vm.f_curv_kin, vm.A_curv_kin, vm.B_curv_kin, x, u, kappa, dt, 2)


def rk4_kinematic_curvilinear(x, u, kappa, dt):
    return _rk_linearise(vm.f_curv_kin, vm.A_curv_kin, vm.B_curv_kin, x, u, kappa, dt, 4)


def euler_dynamic_curvilinear(x, u, kappa, dt=None):
    return _rk_linearise(_f_dyn, _A_dyn, vm.B_curv_dyn, x, u, kappa, dt, 1)


def rk2_dynamic_curvilinear(x, u, kappa, dt):
    return _rk_linearise(_f_dyn, _A_dyn, vm.B_curv_dyn, x, u, kappa, dt, 2)


def rk4_dynamic_curvilinear(x, u, kappa, dt):
    return _rk_linearise(_f_dyn, _A_dyn, vm.B_curv_dyn, x, u, kappa, dt, 4)


# --------------------------------------------------------------------------- condense
def sequential_integration(A, B, d, dt):
    """mpc/ltv/sequential_integration.m:1-49."""
    N_x, N_steps = d.shape
    N_u = B.shape[1]
    A = A * dt + np.eye(N_x)[:, :, None]
    B = B * dt
    d = d * dt

    A_bar = np.zeros((N_x * N_steps, N_x))
    A_bar[:N_x, :] = A[:, :, 0]
    for i in range(1, N_steps):
        A_bar[i * N_x:(i + 1) * N_x, :] = A[:, :, i] @ A_bar[(i - 1) * N_x:i * N_x, :]

    B_bar = np.zeros((N_x * N_steps, N_u * N_steps))
    for i in range(N_steps):
        # QUIRK sequential_integration.m:30: every diagonal block is B(:,:,1)
        B_bar[i * N_x:(i + 1) * N_x, i * N_u:(i + 1) * N_u] = B[:, :, 0]
        for j in range(i + 1, N_steps):
            B_bar[j * N_x:(j + 1) * N_x, i * N_u:(i + 1) * N_u] = (
                A[:, :, j] @ B_bar[(j - 1) * N_x:j * N_x, i * N_u:(i + 1) * N_u])

    D = np.zeros((N_x * N_steps, N_x * N_steps))
    for i in range(N_steps):
        D[i * N_x:(i + 1) * N_x, i * N_x:(i + 1) * N_x] = np.eye(N_x)
        for j in range(i + 1, N_steps):
            D[j * N_x:(j + 1) * N_x, i * N_x:(i + 1) * N_x] = (
                A[:, :, j] @ D[(j - 1) * N_x:j * N_x, i * N_x:(i + 1) * N_x])
    d_bar = D @ d.flatten(order="F")
    return A_bar, B_bar, d_bar


def generate_qp(A_bar, B_bar, d_bar, x0, x_ref, Q, Q_terminal, R, R_soft):
    """mpc/ltv/generate_qp.m:1-35."""
    N_x, N_steps = x_ref.shape
    R = np.asarray(R, dtype=np.float64).reshape(-1)
    R_soft = np.asarray(R_soft, dtype=np.float64).reshape(-1)
    N_soft = R_soft.size
    q = np.concatenate([np.tile(np.asarray(Q, float).reshape(-1), N_steps - 1),
                        np.asarray(Q_terminal, float).reshape(-1)])
    r = np.concatenate([np.tile(R, N_steps), np.zeros(N_soft)])
    e = A_bar @ x0 + d_bar - x_ref.flatten(order="F")
    H = 2 * (B_bar.T @ (q[:, None] * B_bar) + np.diag(r))
    f = 2 * B_bar.T @ (q * e)
    f[f.size - N_soft:] = R_soft
    const = float(e @ (q * e))
    return H, f, const


def _state_index_sets(N_x, N_steps, state_idx, soft_idx):
    """kinematic_state_constraints.m:14-22 (one-based -> zero-based)."""
    st = np.concatenate([np.arange(i - 1, N_x * N_steps, N_x) for i in state_idx]).astype(int)
    so = np.concatenate([np.arange(i - 1, N_x * N_steps, N_x) for i in soft_idx]).astype(int)
    return st, so


def kinematic_tyre_linearise_constraints(A_bar, B_bar, d_bar, x_lin, x0):
    """mpc/ltv/kinematic/kinematic_tyre_linearise_constraints.m:1-35."""
    lr, lf = vm.LR, vm.LF
    N_x = A_bar.shape[1]
    N_steps = A_bar.shape[0] // N_x
    C_bar = np.zeros((N_steps, N_steps * N_x))
    g_bar = np.zeros(N_steps)
    for i in range(N_steps):
        x = x_lin[:, i]
        g0 = x[3] ** 2 * x[4] / (lr + lf)
        C = np.array([0, 0, 0, 2 * x[3] * x[4], x[3] ** 2]) / (lf + lr)
        C_bar[i, i * N_x:(i + 1) * N_x] = C
        g_bar[i] = g0
    A = C_bar @ B_bar
    const = g_bar + C_bar @ (A_bar @ x0 + d_bar - x_lin.flatten(order="F"))
    return A, -5.0 - const, 5.0 - const


def kinematic_state_constraints(A_bar, B_bar, d_bar, x0, lb, ub, state_idx, soft_idx, x_lin):
    """mpc/ltv/kinematic/kinematic_state_constraints.m:1-50."""
    N_x = A_bar.shape[1]
    N_steps = A_bar.shape[0] // N_x
    N_state, N_soft = len(state_idx), len(soft_idx)
    B_bar = np.hstack([B_bar, np.zeros((N_x * N_steps, 1))])
    st, so = _state_index_sets(N_x, N_steps, state_idx, soft_idx)
    cidx = np.concatenate([st, so])
    xA = B_bar[np.concatenate([cidx, so]), :].copy()
    const = A_bar[cidx, :] @ x0 + d_bar[cidx]
    lbA = lb - const
    ubA = ub - const
    ns = N_soft * N_steps
    lbA = np.concatenate([lbA, -np.ones(ns) * 1e10])
    ubA = np.concatenate([ubA[:N_state * N_steps], np.ones(ns) * 1e10, ubA[N_state * N_steps:]])
    xA[xA.shape[0] - 2 * ns:, -1] = np.concatenate([np.ones(ns), -np.ones(ns)])

    A_ay, lb_ay, ub_ay = kinematic_tyre_linearise_constraints(A_bar, B_bar, d_bar, x_lin, x0)
    xA = np.vstack([xA, A_ay, A_ay])
    lbA = np.concatenate([lbA, lb_ay, -INF * np.ones(N_steps)])
    ubA = np.concatenate([ubA, INF * np.ones(N_steps), ub_ay])
    xA[xA.shape[0] - 2 * N_steps:, -1] = np.concatenate([np.ones(N_steps), -np.ones(N_steps)])
    return B_bar, xA, lbA, ubA


def dynamic_slip_linearise_constraints(A_bar, B_bar, d_bar, x_lin, u_lin, x0, kappa):
    """mpc/ltv/dynamic/dynamic_slip_linearise_constraints.m:1-47."""
    lr, lf = vm.LR, vm.LF
    N_x = A_bar.shape[1]
    N_u = u_lin.shape[0]
    N_steps = A_bar.shape[0] // N_x
    C_bar = np.zeros((N_steps * 2, N_steps * N_x))
    D_bar = np.zeros((N_steps * 2, N_steps * N_u))
    g_bar = np.zeros(N_steps * 2)
    for i in range(N_steps):
        x = x_lin[:, i]
        _, _, _, vr, dvr2, xdh, xdhd, vf, dvf2 = vm.A_curv_dyn(x, u_lin[:, i], kappa)
        g0 = [-math.atan(vr), x[6] - math.atan(vf)]
        C = np.array([[0, 0, 0, dvr2 * vr * xdhd / xdh, -dvr2 / xdh, dvr2 * lr / xdh, 0],
                      [0, 0, 0, dvf2 * vf * xdhd / xdh, -dvf2 / xdh, -dvf2 * lf / xdh, 1]])
        C_bar[2 * i:2 * i + 2, i * N_x:(i + 1) * N_x] = C
        g_bar[2 * i:2 * i + 2] = g0
    A = C_bar @ B_bar + np.hstack([D_bar, np.zeros((N_steps * 2, 4))])
    const = (g_bar + C_bar @ (A_bar @ x0 + d_bar - x_lin.flatten(order="F"))
             - D_bar @ u_lin.flatten(order="F"))
    lb = np.tile([-0.1, -0.1], N_steps) - const
    ub = np.tile([0.1, 0.1], N_steps) - const
    return A, lb, ub


def dynamic_tyre_linearise_constraints(A_bar, B_bar, d_bar, x_lin, u_lin, x0, kappa):
    """mpc/ltv/dynamic/dynamic_tyre_linearise_constraints.m:1-64 (12-gon friction ellipse)."""
    ac_max, al_max, lr = 9.163, 10.0, vm.LR
    N_x = A_bar.shape[1]
    N_u = u_lin.shape[0]
    N_steps = A_bar.shape[0] // N_x
    N = 12
    theta = np.linspace(0, 2 * np.pi, N + 1)
    ac_list = ac_max * np.sin(theta)
    al_list = al_max * np.cos(theta)
    dac = ac_list[1:] - ac_list[:N]
    dal = al_list[1:] - al_list[:N]
    C_bar = np.zeros((N_steps * N, N_steps * N_x))
    D_bar = np.zeros((N_steps * N, N_steps * N_u))
    g_bar = np.zeros(N_steps * N)
    for i in range(N_steps):
        x = x_lin[:, i]
        u = u_lin[:, i]
        _, Fcr, Fcr_d, vr, dvr2, xdh, xdhd = vm.A_curv_dyn(x, u, kappa)[:7]
        g0 = np.zeros(N)
        C = np.zeros((N, N_x))
        D = np.zeros((N, N_u))
        for j in range(N):
            g0[j] = (u[0] - al_list[j]) * dac[j] - (Fcr / 280 - ac_list[j]) * dal[j]
            C[j, :] = [0, 0, 0,
                       -dal[j] * Fcr_d * dvr2 * vr * xdhd / xdh / 280,
                       dal[j] * Fcr_d * dvr2 / xdh / 280,
                       -dal[j] * Fcr_d * dvr2 * lr / xdh / 280,
                       0]
            D[j, :] = [dac[j], 0]
        C_bar[N * i:N * (i + 1), i * N_x:(i + 1) * N_x] = C
        D_bar[N * i:N * (i + 1), i * N_u:(i + 1) * N_u] = D
        g_bar[N * i:N * (i + 1)] = g0
    A = C_bar @ B_bar + np.hstack([D_bar, np.zeros((N_steps * N, 4))])
    const = (g_bar + C_bar @ (A_bar @ x0 + d_bar - x_lin.flatten(order="F"))
             - D_bar @ u_lin.flatten(order="F"))
    lb = -INF * np.ones(N_steps * N)
    ub = np.zeros(N_steps * N) - const
    return A, lb, ub


def dynamic_state_constraints(A_bar, B_bar, d_bar, x0, lb, ub, state_idx, soft_idx, x_lin, u_lin, kappa):
    """mpc/ltv/dynamic/dynamic_state_constraints.m:1-58."""
    N_x = A_bar.shape[1]
    N_steps = A_bar.shape[0] // N_x
    N_state, N_soft = len(state_idx), len(soft_idx)
    B_bar = np.hstack([B_bar, np.zeros((N_x * N_steps, 4))])
    st, so = _state_index_sets(N_x, N_steps, state_idx, soft_idx)
    cidx = np.concatenate([st, so])
    xA = B_bar[np.concatenate([cidx, so]), :].copy()
    const = A_bar[cidx, :] @ x0 + d_bar[cidx]
    lbA = lb - const
    ubA = ub - const
    ns = N_soft * N_steps
    lbA = np.concatenate([lbA, -np.ones(ns) * 1e10])
    ubA = np.concatenate([ubA[:N_state * N_steps], np.ones(ns) * 1e10, ubA[N_state * N_steps:]])
    nV = xA.shape[1]
    xA[xA.shape[0] - 2 * ns:, nV - 4] = np.concatenate([np.ones(ns), -np.ones(ns)])

    A_ay, lb_ay, ub_ay = dynamic_slip_linearise_constraints(A_bar, B_bar, d_bar, x_lin, u_lin, x0, kappa)
    xA = np.vstack([xA, A_ay, A_ay])
    lbA = np.concatenate([lbA, lb_ay, -INF * np.ones(N_steps * 2)])
    ubA = np.concatenate([ubA, INF * np.ones(N_steps * 2), ub_ay])
    eye2 = np.tile(np.eye(2), (N_steps, 1))
    xA[xA.shape[0] - 4 * N_steps:, nV - 3:nV - 1] = np.vstack([eye2, -eye2])

    A_ay, lb_ay, ub_ay = dynamic_tyre_linearise_constraints(A_bar, B_bar, d_bar, x_lin, u_lin, x0, kappa)
    xA = np.vstack([xA, A_ay])
    lbA = np.concatenate([lbA, lb_ay])
    ubA = np.concatenate([ubA, ub_ay])
    # length(A_ay) = max(size) = 12*N_steps rows (dynamic_state_constraints.m:57)
    xA[xA.shape[0] - A_ay.shape[0]:, nV - 1] = -1.0
    return B_bar, xA, lbA, ubA


# --------------------------------------------------------------------------- the MPC step
def build_kinematic_qp(x0, x_ref, kappa, dt, x_lin, u_lin, order=2):
    """Everything of ltvmpc_kinetmatic_curvilinear.m:16-41 (up to the qpOASES call)."""
    N_steps = max(x_ref.shape)
    state_idx = [4, 5]
    soft_idx = [2]
    x_lb = np.tile([0, -0.4, -0.75], (N_steps, 1)).flatten(order="F")
    x_ub = np.tile([INF, 0.4, 0.75], (N_steps, 1)).flatten(order="F")
    u_lb = np.concatenate([np.tile([-10, -0.4], N_steps), [0]])
    u_ub = np.concatenate([np.tile([10, 0.4], N_steps), [INF]])
    Q = np.array([5, 250, 2000, 0, 0], dtype=np.float64)
    Q_terminal = Q * 10
    R = [10, 10]
    R_soft = [1e8]
    lin = {1: euler_kinematic_curvilinear, 2: rk2_kinematic_curvilinear, 4: rk4_kinematic_curvilinear}[order]
    A, B, d = lin(x_lin, u_lin, kappa, dt)
    A_bar, B_bar, d_bar = sequential_integration(A, B, d, dt)
    B_bar, xA, lbA, ubA = kinematic_state_constraints(A_bar, B_bar, d_bar, x0, x_lb, x_ub,
                                                      state_idx, soft_idx, x_lin)
    H, f, const = generate_qp(A_bar, B_bar, d_bar, x0, x_ref, Q, Q_terminal, R, R_soft)
    return dict(A=A, B=B, d=d, A_bar=A_bar, B_bar=B_bar, d_bar=d_bar, H=H, f=f, const=const,
                xA=xA, lbA=lbA, ubA=ubA, lb=u_lb, ub=u_ub, n_soft=1)


def build_dynamic_qp(x0, x_ref, kappa, dt, x_lin, u_lin, order=4):
    """Everything of ltvmpc_dynamic_curvilinear.m:16-41 (up to the qpOASES call)."""
    N_steps = max(x_ref.shape)
    state_idx = [4, 7]
    soft_idx = [2]
    x_lb = np.tile([0, -0.4, -0.75], (N_steps, 1)).flatten(order="F")
    x_ub = np.tile([INF, 0.4, 0.75], (N_steps, 1)).flatten(order="F")
    u_lb = np.concatenate([np.tile([-10, -0.4], N_steps), [0, 0, 0, 0]])
    u_ub = np.concatenate([np.tile([10, 0.4], N_steps), [INF] * 4])
    Q = np.array([5, 250, 2000, 0, 0, 0, 0], dtype=np.float64)
    Q_terminal = Q * 10
    R = [10, 10]
    R_soft = [1e8, 1e6, 1e6, 1e4]
    lin = {1: euler_dynamic_curvilinear, 2: rk2_dynamic_curvilinear, 4: rk4_dynamic_curvilinear}[order]
    A, B, d = lin(x_lin, u_lin, kappa, dt)
    A_bar, B_bar, d_bar = sequential_integration(A, B, d, dt)
    B_bar, xA, lbA, ubA = dynamic_state_constraints(A_bar, B_bar, d_bar, x0, x_lb, x_ub,
                                                    state_idx, soft_idx, x_lin, u_lin, kappa)
    H, f, const = generate_qp(A_bar, B_bar, d_bar, x0, x_ref, Q, Q_terminal, R, R_soft)
    return dict(A=A, B=B, d=d, A_bar=A_bar, B_bar=B_bar, d_bar=d_bar, H=H, f=f, const=const,
                xA=xA, lbA=lbA, ubA=ubA, lb=u_lb, ub=u_ub, n_soft=4)


def _finish(qp, x0, sol):
    """ltvmpc_*_curvilinear.m:52-60."""
    n_soft = qp["n_soft"]
    z = sol.x
    slack_opt = z[z.size - n_soft:].copy()
    x_opt = qp["A_bar"] @ x0 + qp["B_bar"] @ z + qp["d_bar"]
    u_opt = z[:z.size - n_soft].copy()
    fval = sol.fval + qp["const"]
    return u_opt, x_opt, sol.exitflag, fval, slack_opt, sol


def ltvmpc_kinetmatic_curvilinear(x0, x_ref, kappa, dt, x_lin, u_lin, order=2):
    """mpc/ltv/kinematic/ltvmpc_kinetmatic_curvilinear.m:1-62.
    Returns (u_opt, x_opt, exitflag, fval, slack_opt, sol)."""
    x0 = np.asarray(x0, dtype=np.float64).reshape(-1)
    qp = build_kinematic_qp(x0, x_ref, kappa, dt, x_lin, u_lin, order)
    sol = _qp.qpoases(qp["H"], qp["f"], qp["xA"], qp["lb"], qp["ub"], qp["lbA"], qp["ubA"])
    return _finish(qp, x0, sol)


def ltvmpc_dynamic_curvilinear(x0, x_ref, kappa, dt, x_lin, u_lin, order=4):
    """mpc/ltv/dynamic/ltvmpc_dynamic_curvilinear.m:1-62."""
    x0 = np.asarray(x0, dtype=np.float64).reshape(-1)
    qp = build_dynamic_qp(x0, x_ref, kappa, dt, x_lin, u_lin, order)
    sol = _qp.qpoases(qp["H"], qp["f"], qp["xA"], qp["lb"], qp["ub"], qp["lbA"], qp["ubA"])
    return _finish(qp, x0, sol)
