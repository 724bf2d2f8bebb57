"""ORACLE (test infrastructure only) -- util/obtain_reference.m: re-parameterise a planned lap
(state samples every `ds` metres of arclength with the time `t` spent in each segment) from
arclength to time, giving the MPC its reference over the horizon (call site: main.m:115).

Pinned: tests/golden/reference_m_obtain_reference.npz holds what the reference's unmodified
.m file computes (executed by oracle/mlab, scripts/make_reference_fixtures.py);
tests/test_oracle_golden.py checks this restatement against it.
"""
import numpy as np


def _nxt(i, N):
    """obtain_reference.m:58-60 (one-based)."""
    return int(np.mod(i, N)) + 1


def obtain_reference(x, ds, N_s, t, s0, dt, N_t):
    """util/obtain_reference.m:1-50.  x: plan [8*N_s] (n, mu, x_d, y_d, theta_d, delta, a, delta_d per
    sample), t: [N_s] segment times.  Returns x_ref [7 x N_t] (s, n, mu, x_d, y_d, theta_d, delta)."""
    x = np.asarray(x, dtype=np.float64).reshape(-1)
    t = np.asarray(t, dtype=np.float64).reshape(-1)
    N_s, N_t = int(N_s), int(N_t)
    L = ds * N_s                                                   # :5
    cols = [x[k::8] for k in range(6)]                              # :7-12  n, mu, x_d, y_d, theta_d, delta
    idx = np.zeros(N_t + 1, dtype=np.int64)
    rto = np.zeros(N_t + 1)
    idx[0] = int(np.floor(np.mod(s0, L) / ds)) + 1                  # :21
    rto[0] = np.mod(np.mod(s0, L) / ds, 1)                          # :22
    for i in range(1, N_t + 1):                                     # :24-35
        t_remaining = dt
        idx[i] = idx[i - 1]
        rto[i] = rto[i - 1] + t_remaining / t[idx[i] - 1]
        t_remaining = t_remaining - t[idx[i - 1] - 1] * (1 - rto[i - 1])
        while rto[i] > 1:
            idx[i] = _nxt(idx[i], N_s)
            rto[i] = t_remaining / t[idx[i] - 1]
            t_remaining = t_remaining - t[idx[i] - 1]
    x_ref = np.zeros((7, N_t))
    for i in range(1, N_t + 1):                                     # :40-48
        x_ref[0, i - 1] = s0 + np.mod(idx[i] + rto[i] - idx[0] - rto[0], N_s) * ds
        a, b = idx[i] - 1, _nxt(idx[i], N_s) - 1
        for k in range(6):
            x_ref[k + 1, i - 1] = cols[k][a] + (cols[k][b] - cols[k][a]) * rto[i]
    return x_ref
