"""ORACLE (test infrastructure only) -- numpy restatement of reference main.m
(closed-loop LTV-MPC on a raceline, plant = Cartesian dynamic bicycle + PIDs).

`mpc_step` is injectable so tests can run the same loop with the CUDA step and
compare trajectories against the oracle step.
"""
import numpy as np

from . import vehicle as vm
from . import ltv


def make_reference(x0, x_vel, N_x, N_steps, dt, target_vel=20.0):
    """main.m:99-107 -- constant-acceleration speed ramp to TARGET_VEL and its arc length."""
    x_ref = np.zeros((N_x, N_steps))
    k = np.arange(1, N_steps + 1)
    if x_vel < target_vel:
        x_ref[3, :] = np.minimum(x0[3] + 10 * dt * k, target_vel)
    else:
        x_ref[3, :] = np.maximum(x0[3] - 10 * dt * k, target_vel)
    x_ref[0, :] = x0[0] + np.cumsum(x_ref[3, :] * dt)
    return x_ref


def initial_guess(N_x, N_u, N_steps, dt):
    """main.m:42-53 -- quadratic arc length, linear speed, constant acceleration."""
    x_mpc = np.zeros((N_x + N_u, N_steps))
    t = dt * np.arange(1, N_steps + 1)
    x_mpc[0, :] = 10 * t ** 2 / 2
    x_mpc[3, :] = 10 * t
    x_mpc[N_x + N_u - 2, :] = 10
    return x_mpc[:N_x, :].copy(), x_mpc[N_x:, :].copy()


def run(track, model="KINEMATIC", n_sim=1000, N_steps=40, dt=0.05, mpc_step=None, record=None):
    """main.m:56-170.  Returns a dict of histories.  `record(i, inputs, outputs)` is
    called once per MPC step with the exact operands of the ltvmpc_* call."""
    N_x = 5 if model == "KINEMATIC" else 7
    N_u = 2
    kappa = track.kappa
    if mpc_step is None:
        mpc_step = (ltv.ltvmpc_kinetmatic_curvilinear if model == "KINEMATIC"
                    else ltv.ltvmpc_dynamic_curvilinear)
    x_opt, u_opt = initial_guess(N_x, N_u, N_steps, dt)
    x = np.zeros(7)
    vel_pid_settings = (16000.0, 0, 0, 2800)
    steer_pid_settings = (80.0, 0, 0, 0.8)
    vel_pid_status = (0, 0)
    steer_pid_status = (0, 0)
    hist = dict(n=[], exitflag=[], fval=[], slack=[], x=[], u0=[], x0=[], iters=[])
    for i in range(n_sim):
        s, n, mu = vm.cartesian_to_curvilinear(x[0], x[1], x[2], track.x_spline, track.y_spline,
                                               track.dl, x_opt[0, 0])
        if model == "KINEMATIC":
            x0 = np.array([s, n, mu, np.linalg.norm(x[3:5]), x[6]])
        else:
            x0 = np.array([s, n, mu, x[3], x[4], x[5], x[6]])
        hist["n"].append(n)
        if s >= track.L:
            break
        x_ref = make_reference(x0, x[3], N_x, N_steps, dt)
        out = mpc_step(x0, x_ref, kappa, dt, x_opt.reshape(N_x, N_steps, order="F"),
                       u_opt.reshape(N_u, N_steps, order="F"))
        u_new, x_new, exitflag, fval, slack = out[:5]
        if record is not None:
            record(i, dict(x0=x0, x_ref=x_ref, x_lin=x_opt.copy(), u_lin=u_opt.copy()), out)
        x_opt = np.asarray(x_new).reshape(N_x, N_steps, order="F")
        u_opt = np.asarray(u_new).reshape(N_u, N_steps, order="F")
        hist["exitflag"].append(exitflag)
        hist["fval"].append(fval)
        hist["slack"].append(np.array(slack))
        hist["x0"].append(x0)
        v_ref = x_opt[3, 0]
        delta_ref = x_opt[N_x - 1, 0]
        for _ in range(10):
            vel_rate, vel_pid_status = vm.pid_controller(v_ref, x[3], vel_pid_settings, vel_pid_status)
            steer_rate, steer_pid_status = vm.pid_controller(delta_ref, x[6], steer_pid_settings,
                                                             steer_pid_status)
            x = vm.integrate_cart_dyn(x, np.array([vel_rate, steer_rate]), dt / 10)
        hist["x"].append(x.copy())
        hist["u0"].append(u_opt[:, 0].copy())
    hist["steps"] = len(hist["exitflag"])
    return hist
