"""ORACLE (test infrastructure only) -- numpy restatement of reference vehicle_models/*.

`kappa` is a callable s -> curvature, like the anonymous function of main.m:18.
State/control vectors are 1-D numpy arrays in the reference's order.
"""
import math
import numpy as np

from . import spline as sp

# vehicle constants as hard-coded in the reference (f_curv_kin.m:13-15, f_curv_dyn.m:13-18)
LR = 0.6183
LF = 0.8672
MASS = 280.0
INERTIA = 200.0
GRAV = 9.81
PAC_B, PAC_C, PAC_D, PAC_E = 12.56, 1.38, 1.60, -0.58


def f_curv_kin(x, u, kappa):
    """vehicle_models/curvilinear_kinematic/f_curv_kin.m:1-31."""
    lr_ratio = LR / (LR + LF)
    k = kappa(x[0])
    beta = math.atan(lr_ratio * math.tan(x[4]))
    s_mu_beta = math.sin(x[2] + beta)
    c_mu_beta = math.cos(x[2] + beta)
    denom_nk = 1 / (1 - x[1] * k)
    return np.array([x[3] * c_mu_beta * denom_nk,
                     x[3] * s_mu_beta,
                     x[3] * math.sin(beta) / LR - x[3] * c_mu_beta * denom_nk * k,
                     u[0],
                     u[1]])


def A_curv_kin(x, u, kappa, kappa_d=None):
    """vehicle_models/curvilinear_kinematic/A_curv_kin.m:1-57."""
    lr_ratio = LR / (LR + LF)
    k = kappa(x[0])
    beta = math.atan(lr_ratio * math.tan(x[4]))
    s_mu_beta = math.sin(x[2] + beta)
    c_mu_beta = math.cos(x[2] + beta)
    sec = 1 / math.cos(x[4])
    beta_d = lr_ratio * sec ** 2 / (1 + (lr_ratio * math.tan(x[4])) ** 2)
    denom_nk = 1 / (1 - x[1] * k)

    s_s = 0.0
    s_n = x[3] * c_mu_beta * denom_nk ** 2 * k
    s_mu = -x[3] * s_mu_beta * denom_nk
    s_v = c_mu_beta * denom_nk
    s_delta = -x[3] * s_mu_beta * denom_nk * beta_d

    n_mu = x[3] * c_mu_beta
    n_v = s_mu_beta
    n_delta = x[3] * c_mu_beta * beta_d

    mu_s = 0.0
    mu_n = -s_n * k
    mu_mu = -s_mu * k
    mu_v = math.sin(beta) / LR - s_v * k
    mu_delta = x[3] * math.cos(beta) * beta_d / LR - s_delta * k

    if kappa_d is not None:
        k_d = kappa_d(x[0])
        s_s = x[3] * c_mu_beta * denom_nk ** 2 * k_d * x[1]
        mu_s = -x[3] * c_mu_beta * denom_nk * k_d - s_s * k

    return np.array([[s_s, s_n, s_mu, s_v, s_delta],
                     [0, 0, n_mu, n_v, n_delta],
                     [mu_s, mu_n, mu_mu, mu_v, mu_delta],
                     [0, 0, 0, 0, 0],
                     [0, 0, 0, 0, 0]], dtype=np.float64)


def B_curv_kin(x=None, u=None, kappa=None):
    """vehicle_models/curvilinear_kinematic/B_curv_kin.m:1-18."""
    return np.array([[0, 0], [0, 0], [0, 0], [1, 0], [0, 1]], dtype=np.float64)


def _pacejka(alpha, Fz):
    bt = PAC_B * alpha
    inner = bt - PAC_E * (bt - math.atan(bt))
    return Fz * PAC_D * math.sin(PAC_C * math.atan(inner))


def _pacejka_d(alpha, Fz):
    bt = PAC_B * alpha
    inner = bt - PAC_E * (bt - math.atan(bt))
    return (Fz * PAC_D * math.cos(PAC_C * math.atan(inner))
            * PAC_C / (1 + inner ** 2)
            * (PAC_B - PAC_E * (PAC_B - PAC_B / (1 + PAC_B ** 2 * alpha ** 2))))


def f_curv_dyn(x, u, kappa):
    """vehicle_models/curvilinear_dynamic/f_curv_dyn.m:1-64.  Returns (f, Fcr)."""
    m, I, lr, lf, g = MASS, INERTIA, LR, LF, GRAV
    s, n, mu, x_d, y_d, theta_d, delta = x
    Fx = u[0] * m
    delta_d = u[1]
    x_d_hat = x_d + 5 * math.exp(-x_d / 5)
    k = kappa(s)
    denom_nk = 1 / (1 - n * k)
    alpha_f = delta - math.atan((y_d + lf * theta_d) / x_d_hat)
    alpha_r = -math.atan((y_d - lr * theta_d) / x_d_hat)
    Fzf = m * g * lr / (lr + lf)
    Fzr = m * g * lf / (lr + lf)
    Fcf = _pacejka(alpha_f, Fzf)
    Fcr = _pacejka(alpha_r, Fzr)
    f = np.array([(x_d * math.cos(mu) - y_d * math.sin(mu)) * denom_nk,
                  x_d * math.sin(mu) + y_d * math.cos(mu),
                  theta_d - (x_d * math.cos(mu) - y_d * math.sin(mu)) * denom_nk * k,
                  (Fx - Fcf * math.sin(delta) + m * y_d * theta_d) / m,
                  (Fcr + Fcf * math.cos(delta) - m * x_d * theta_d) / m,
                  (lf * Fcf * math.cos(delta) - lr * Fcr) / I,
                  delta_d])
    return f, Fcr


def A_curv_dyn(x, u, kappa):
    """vehicle_models/curvilinear_dynamic/A_curv_dyn.m:1-107.

    Returns (A, Fcr, Fcr_d, vr, denom_vr2, x_d_hat, x_d_hat_d, vf, denom_vf2).
    """
    m, I, lr, lf, g = MASS, INERTIA, LR, LF, GRAV
    s, n, mu, x_d, y_d, theta_d, delta = x
    x_d_hat = x_d + 5 * math.exp(-x_d / 5)
    x_d_hat_d = 1 - math.exp(-x_d / 5)
    alpha_f = delta - math.atan((y_d + lf * theta_d) / x_d_hat)
    alpha_r = -math.atan((y_d - lr * theta_d) / x_d_hat)
    Fzf = m * g * lr / (lr + lf)
    Fzr = m * g * lf / (lr + lf)
    Fcf = _pacejka(alpha_f, Fzf)
    Fcr = _pacejka(alpha_r, Fzr)
    Fcf_d = _pacejka_d(alpha_f, Fzf)
    Fcr_d = _pacejka_d(alpha_r, Fzr)

    k = kappa(s)
    denom_nk = 1 / (1 - n * k)
    vf = (y_d + lf * theta_d) / x_d_hat
    vr = (y_d - lr * theta_d) / x_d_hat
    denom_vf2 = 1 / (1 + vf ** 2)
    denom_vr2 = 1 / (1 + vr ** 2)
    sm, cm = math.sin(mu), math.cos(mu)
    sd, cd = math.sin(delta), math.cos(delta)

    s_n = (x_d * cm - y_d * sm) * denom_nk ** 2 * k
    s_mu = (-x_d * sm - y_d * cm) * denom_nk
    s_xd = cm * denom_nk
    s_yd = -sm * denom_nk

    n_mu = x_d * cm - y_d * sm
    n_xd = sm
    n_yd = cm

    mu_n = -s_n * k
    mu_mu = -s_mu * k
    mu_xd = -s_xd * k
    mu_yd = -s_yd * k
    mu_thetad = 1.0

    xd_xd = -Fcf_d * denom_vf2 * vf * sd * x_d_hat_d / (m * x_d_hat)
    xd_yd = (Fcf_d * denom_vf2 * sd / x_d_hat + m * theta_d) / m
    xd_thetad = (Fcf_d * denom_vf2 * lf * sd / x_d_hat + m * y_d) / m
    xd_delta = (-Fcf * cd - Fcf_d * sd) / m

    yd_xd = (Fcr_d * denom_vr2 * vr * x_d_hat_d / x_d_hat
             + Fcf_d * denom_vf2 * vf * cd * x_d_hat_d / x_d_hat - m * theta_d) / m
    yd_yd = (-Fcr_d * denom_vr2 / x_d_hat - Fcf_d * denom_vf2 / x_d_hat * cd) / m
    yd_thetad = (Fcr_d * denom_vr2 * lr / x_d_hat - Fcf_d * denom_vf2 * lf / x_d_hat * cd
                 - m * x_d_hat) / m
    yd_delta = (-Fcf * sd + Fcf_d * cd) / m

    t_xd = (lf * Fcf_d * denom_vf2 * vf * cd * x_d_hat_d / x_d_hat
            - lr * Fcr_d * denom_vr2 * vr * x_d_hat_d / x_d_hat) / I
    t_yd = (-lf * Fcf_d * denom_vf2 * cd / x_d_hat + lr * Fcr_d * denom_vr2 / x_d_hat) / I
    t_thetad = (-lf * Fcf_d * denom_vf2 * lf * cd / x_d_hat
                - lr * Fcr_d * denom_vr2 * lr / x_d_hat) / I
    t_delta = (-lf * Fcf * sd + lf * Fcf_d * cd) / I

    A = np.array([[0, s_n, s_mu, s_xd, s_yd, 0, 0],
                  [0, 0, n_mu, n_xd, n_yd, 0, 0],
                  [0, mu_n, mu_mu, mu_xd, mu_yd, mu_thetad, 0],
                  [0, 0, 0, xd_xd, xd_yd, xd_thetad, xd_delta],
                  [0, 0, 0, yd_xd, yd_yd, yd_thetad, yd_delta],
                  [0, 0, 0, t_xd, t_yd, t_thetad, t_delta],
                  [0, 0, 0, 0, 0, 0, 0]], dtype=np.float64)
    return A, Fcr, Fcr_d, vr, denom_vr2, x_d_hat, x_d_hat_d, vf, denom_vf2


def B_curv_dyn(x=None, u=None, kappa=None):
    """vehicle_models/curvilinear_dynamic/B_curv_dyn.m:1-20."""
    B = np.zeros((7, 2))
    B[3, 0] = 1.0
    B[6, 1] = 1.0
    return B


def f_cart_dyn(x, u):
    """vehicle_models/cartesian_dynamic/f_cart_dyn.m:1-56 (the simulated plant)."""
    m, I, lr, lf, g = MASS, INERTIA, LR, LF, GRAV
    theta, x_d, y_d, theta_d, delta = x[2], x[3], x[4], x[5], x[6]
    Fx, delta_d = u[0], u[1]
    alpha_f = delta - math.atan((y_d + lf * theta_d) / (x_d + 0.01))
    alpha_r = -math.atan((y_d - lr * theta_d) / (x_d + 0.01))
    Fzf = m * g * lr / (lr + lf)
    Fzr = m * g * lf / (lr + lf)
    Fcf = _pacejka(alpha_f, Fzf)
    Fcr = _pacejka(alpha_r, Fzr)
    return np.array([x_d * math.cos(theta) - y_d * math.sin(theta),
                     x_d * math.sin(theta) + y_d * math.cos(theta),
                     theta_d,
                     (Fx - Fcf * math.sin(delta) + m * y_d * theta_d) / m,
                     (Fcr + Fcf * math.cos(delta) - m * x_d * theta_d) / m,
                     (lf * Fcf * math.cos(delta) - lr * Fcr) / I,
                     delta_d])


def integrate_cart_dyn(x0, u, dt):
    """vehicle_models/cartesian_dynamic/integrate_cart_dyn.m:1-24 (6-stage RK, as written)."""
    k1 = f_cart_dyn(x0, u)
    k2 = f_cart_dyn(x0 + k1 * dt / 2, u)
    k3 = f_cart_dyn(x0 + k1 * dt / 4 + k2 * dt / 8, u)
    k4 = f_cart_dyn(x0 - k2 * dt + 2 * k3 * dt, u)
    k5 = f_cart_dyn(x0 + 7 / 27 * k2 * dt + 10 / 27 * k2 * dt + k4 * dt / 27, u)
    k6 = f_cart_dyn(x0 + 28 / 625 * k1 * dt - k2 * dt / 5 + 546 / 625 * k3 * dt
                    + 54 / 625 * k4 * dt - 378 / 625 * k5 * dt, u)
    f = k1 / 24 + 5 / 48 * k4 + 27 / 56 * k5 + 125 / 336 * k6
    return x0 + dt * f


def pid_controller(target, current, settings, status):
    """vehicle_models/pid_controller.m:1-20."""
    kp, ki, kd, max_output = settings
    error = target - current
    integral_error = status[0] + error
    derivative_error = error - status[1]
    output = kp * error + ki * integral_error + kd * derivative_error
    output = max(min(output, max_output), -max_output)
    return output, (integral_error, error)


def _wrap_to_pi(a):
    return (a + math.pi) % (2 * math.pi) - math.pi if not (-math.pi <= a <= math.pi) else a


def cartesian_to_curvilinear(x, y, theta, x_P, y_P, dl, s0):
    """vehicle_models/cartesian_to_curvilinear.m:1-28 (angdiff(a,b) = wrapToPi(b-a))."""
    s = sp.closest_point(x, y, x_P, y_P, dl, s0, 0.01)
    car = np.array([x - sp.interpolate_spline(s, x_P, dl)[0],
                    y - sp.interpolate_spline(s, y_P, dl)[0]])
    tangent = np.array([-sp.interpolate_spline_d(s, y_P, dl)[0],
                        sp.interpolate_spline_d(s, x_P, dl)[0]])
    tangent = tangent / np.linalg.norm(tangent)
    n = float(car @ tangent)
    mu = _wrap_to_pi(theta - sp.interpolate_angle(s, x_P, y_P, dl)[0])
    return s, n, mu


def curvilinear_to_cartesian(s, n, mu, x_P, y_P, dl):
    """vehicle_models/curvilinear_to_cartesian.m:1-30."""
    s = np.atleast_1d(np.asarray(s, dtype=np.float64))
    x_track = sp.interpolate_spline(s, x_P, dl)
    y_track = sp.interpolate_spline(s, y_P, dl)
    x_t = -sp.interpolate_spline_d(s, y_P, dl)
    y_t = sp.interpolate_spline_d(s, x_P, dl)
    nrm = np.sqrt(x_t ** 2 + y_t ** 2)
    x_t, y_t = x_t / nrm, y_t / nrm
    return (x_track + n * x_t, y_track + n * y_t,
            sp.interpolate_angle(s, x_P, y_P, dl) + mu)
