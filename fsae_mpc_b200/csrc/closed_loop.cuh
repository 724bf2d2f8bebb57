// Batch driver: main.m's closed loop (main.m:87-170) for B vehicles at once.  Per simulation
// step: one fused LTV-MPC launch + one element-wise "advance" kernel (one thread per vehicle)
// that (a) applies the first predicted state through the actuator PIDs to the Cartesian
// dynamic plant (main.m:146-160, pid_controller.m, integrate_cart_dyn.m, f_cart_dyn.m) and
// (b) prepares the next MPC call: projection onto the track (cartesian_to_curvilinear.m,
// closest_point.m), x0 assembly (main.m:89-94) and the speed-ramp reference (main.m:99-107).
#pragma once
#include "models.cuh"

namespace fsae {

struct SimArgs {
    int B, N, model;            // model: FSAE_MODEL_*
    double dt, target_vel, ramp;
    const int32_t* track_id;
    const int32_t* param_id;
    const DevTrack* tracks;
    const fsae_params* params;
    double* plant;              // [7 x B] Cartesian plant state
    double* pid;                // [4 x B] vel integral, vel last error, steer integral, steer last error
    int32_t* alive;             // [B] 1 while the lap is not finished
    const double* x_opt;        // [NX*N x B] last MPC prediction (also the next linearisation point)
    const int32_t* exitflag;    // [B] of the MPC step just done (may be null on the first call)
    double* x0;                 // [NX x B]     next MPC inputs
    double* x_ref;              // [NX x N x B]
    double* n_hist;             // [n_sim x B] lateral deviation per step (optional)
    double* plant_hist;         // [7 x n_sim x B] (optional)
    int32_t* exit_hist;         // [n_sim x B] (optional)
    int32_t* steps;             // [B] MPC steps taken
    double track_len[FSAE_MAX_TRACKS];
    int step, n_sim, do_plant;  // do_plant = 0 on the very first call (no prediction yet)
};

__device__ __forceinline__ void spline_eval(const DevTrack& tr, double s, double* X, double* Xd, double* Xdd,
                                            double* Y, double* Yd, double* Ydd) {
    const double period = tr.dl * (double)tr.n_seg;
    double t = s - floor(s / period) * period;
    if (t < 0.0) t += period;
    if (t >= period) t -= period;
    int i = (int)floor(t / tr.dl);
    i = i < 0 ? 0 : (i >= tr.n_seg ? tr.n_seg - 1 : i);
    const double u = t / tr.dl - (double)i, o = 1.0 - u;
    const double* c = tr.coef + 8 * i;
    const double a0 = o * o * o, a1 = 3.0 * o * o * u, a2 = 3.0 * o * u * u, a3 = u * u * u;
    const double b0 = -3.0 * o * o, b1 = 3.0 * (3.0 * u * u - 4.0 * u + 1.0), b2 = 3.0 * (2.0 * u - 3.0 * u * u), b3 = 3.0 * u * u;
    const double e0 = 6.0 * o, e1 = 6.0 * (3.0 * u - 2.0), e2 = 6.0 * (1.0 - 3.0 * u), e3 = 6.0 * u;
    *X = a0 * c[0] + a1 * c[1] + a2 * c[2] + a3 * c[3];
    *Y = a0 * c[4] + a1 * c[5] + a2 * c[6] + a3 * c[7];
    *Xd = (b0 * c[0] + b1 * c[1] + b2 * c[2] + b3 * c[3]) / tr.dl;
    *Yd = (b0 * c[4] + b1 * c[5] + b2 * c[6] + b3 * c[7]) / tr.dl;
    *Xdd = (e0 * c[0] + e1 * c[1] + e2 * c[2] + e3 * c[3]) / (tr.dl * tr.dl);
    *Ydd = (e0 * c[4] + e1 * c[5] + e2 * c[6] + e3 * c[7]) / (tr.dl * tr.dl);
}

// vehicle_models/cartesian_dynamic/f_cart_dyn.m:20-54
__device__ __forceinline__ void f_cart_dyn(const double* x, double Fx, double delta_d, const fsae_params& p, double* f) {
    const double m = p.mass, I = p.inertia, lr = p.lr, lf = p.lf, g = p.grav;
    const double theta = x[2], x_d = x[3], y_d = x[4], theta_d = x[5], delta = x[6];
    const double alpha_f = delta - atan((y_d + lf * theta_d) / (x_d + 0.01));
    const double alpha_r = -atan((y_d - lr * theta_d) / (x_d + 0.01));
    const double Fzf = m * g * lr / (lr + lf), Fzr = m * g * lf / (lr + lf);
    double Fcf, Fcr, dmy;
    DynModel::pacejka(p, alpha_f, Fzf, Fcf, dmy);
    DynModel::pacejka(p, alpha_r, Fzr, Fcr, dmy);
    double st, ct, sd, cd;
    sincos(theta, &st, &ct);
    sincos(delta, &sd, &cd);
    f[0] = x_d * ct - y_d * st;
    f[1] = x_d * st + y_d * ct;
    f[2] = theta_d;
    f[3] = (Fx - Fcf * sd + m * y_d * theta_d) / m;
    f[4] = (Fcr + Fcf * cd - m * x_d * theta_d) / m;
    f[5] = (lf * Fcf * cd - lr * Fcr) / I;
    f[6] = delta_d;
}

// vehicle_models/cartesian_dynamic/integrate_cart_dyn.m:11-22 (six-stage scheme, as written)
__device__ void integrate_cart_dyn(double* x, double Fx, double dd, double h, const fsae_params& p) {
    double k1[7], k2[7], k3[7], k4[7], k5[7], k6[7], xt[7];
    f_cart_dyn(x, Fx, dd, p, k1);
    for (int i = 0; i < 7; ++i) xt[i] = x[i] + k1[i] * h / 2;
    f_cart_dyn(xt, Fx, dd, p, k2);
    for (int i = 0; i < 7; ++i) xt[i] = x[i] + k1[i] * h / 4 + k2[i] * h / 8;
    f_cart_dyn(xt, Fx, dd, p, k3);
    for (int i = 0; i < 7; ++i) xt[i] = x[i] - k2[i] * h + 2 * k3[i] * h;
    f_cart_dyn(xt, Fx, dd, p, k4);
    for (int i = 0; i < 7; ++i) xt[i] = x[i] + 7.0 / 27 * k2[i] * h + 10.0 / 27 * k2[i] * h + k4[i] * h / 27;
    f_cart_dyn(xt, Fx, dd, p, k5);
    for (int i = 0; i < 7; ++i)
        xt[i] = x[i] + 28.0 / 625 * k1[i] * h - k2[i] * h / 5 + 546.0 / 625 * k3[i] * h + 54.0 / 625 * k4[i] * h - 378.0 / 625 * k5[i] * h;
    f_cart_dyn(xt, Fx, dd, p, k6);
    for (int i = 0; i < 7; ++i) x[i] += h * (k1[i] / 24 + 5.0 / 48 * k4[i] + 27.0 / 56 * k5[i] + 125.0 / 336 * k6[i]);
}

__global__ void sim_advance_kernel(SimArgs a) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= a.B) return;
    if (!a.alive[b]) return;
    const int NX = a.model == FSAE_MODEL_DYNAMIC ? 7 : 5, N = a.N;
    const int tid_ = a.track_id ? a.track_id[b] : 0;
    const fsae_params& P = a.params[a.param_id ? a.param_id[b] : 0];
    const DevTrack tr = a.tracks[tid_];
    double x[7];
    for (int i = 0; i < 7; ++i) x[i] = a.plant[(size_t)b * 7 + i];
    const double* xo = a.x_opt + (size_t)b * NX * N;
    if (a.do_plant) {
        // main.m:146-160: track the first predicted speed / steering angle with the PIDs
        const double v_ref = xo[3], delta_ref = xo[NX - 1];
        double vi = a.pid[(size_t)b * 4 + 0], ve = a.pid[(size_t)b * 4 + 1];
        double si = a.pid[(size_t)b * 4 + 2], se = a.pid[(size_t)b * 4 + 3];
        for (int j = 0; j < 10; ++j) {
            // pid_controller.m:5-18 with {kp,ki,kd,max} = {16000,0,0,2800} and {80,0,0,0.8} (main.m:79-83)
            double e = v_ref - x[3];
            vi += e;
            double vel_rate = 16000.0 * e + 0.0 * vi + 0.0 * (e - ve);
            vel_rate = fmax(fmin(vel_rate, 2800.0), -2800.0);
            ve = e;
            e = delta_ref - x[6];
            si += e;
            double steer_rate = 80.0 * e + 0.0 * si + 0.0 * (e - se);
            steer_rate = fmax(fmin(steer_rate, 0.8), -0.8);
            se = e;
            integrate_cart_dyn(x, vel_rate, steer_rate, a.dt / 10, P);
        }
        a.pid[(size_t)b * 4 + 0] = vi; a.pid[(size_t)b * 4 + 1] = ve;
        a.pid[(size_t)b * 4 + 2] = si; a.pid[(size_t)b * 4 + 3] = se;
        for (int i = 0; i < 7; ++i) a.plant[(size_t)b * 7 + i] = x[i];
        if (a.plant_hist)
            for (int i = 0; i < 7; ++i) a.plant_hist[((size_t)b * a.n_sim + (a.step - 1)) * 7 + i] = x[i];
        if (a.exit_hist && a.exitflag) a.exit_hist[(size_t)b * a.n_sim + (a.step - 1)] = a.exitflag[b];
        a.steps[b] = a.step;
    }
    if (a.step >= a.n_sim) return;
    // cartesian_to_curvilinear.m:17-26 with closest_point.m:15-32 (Newton, epsilon 0.01)
    double s = xo[0];
    double X, Xd, Xdd, Y, Yd, Ydd;
    double delta = 0.02;
    int it = 0;
    while (fabs(delta) > 0.01 && it < 200) {
        spline_eval(tr, s, &X, &Xd, &Xdd, &Y, &Yd, &Ydd);
        const double dist_d = 2 * (X - x[0]) * Xd + 2 * (Y - x[1]) * Yd;
        const double dist_dd = 2 * (X - x[0]) * Xdd + 2 * Xd * Xd + 2 * (Y - x[1]) * Ydd + 2 * Yd * Yd;
        delta = dist_d / dist_dd;
        s -= delta;
        ++it;
    }
    spline_eval(tr, s, &X, &Xd, &Xdd, &Y, &Yd, &Ydd);
    const double tn = sqrt(Xd * Xd + Yd * Yd);
    const double n = ((x[0] - X) * (-Yd) + (x[1] - Y) * Xd) / tn;
    double mu = x[2] - atan2(Yd, Xd);                      // angdiff(angle, theta) = wrapToPi(theta - angle)
    if (mu > M_PI || mu < -M_PI) { mu = fmod(mu + M_PI, 2 * M_PI); if (mu < 0) mu += 2 * M_PI; mu -= M_PI; }
    if (a.n_hist) a.n_hist[(size_t)b * a.n_sim + a.step] = n;
    if (s >= a.track_len[tid_]) { a.alive[b] = 0; return; }   // main.m:97-99 lap finished
    double* x0 = a.x0 + (size_t)b * NX;
    x0[0] = s; x0[1] = n; x0[2] = mu;
    if (NX == 5) { x0[3] = sqrt(x[3] * x[3] + x[4] * x[4]); x0[4] = x[6]; }
    else { x0[3] = x[3]; x0[4] = x[4]; x0[5] = x[5]; x0[6] = x[6]; }
    // main.m:101-108: speed ramp towards TARGET_VEL and its arc length
    double* xr = a.x_ref + (size_t)b * NX * N;
    double cs = 0.0;
    const bool up = x[3] < a.target_vel;
    for (int k = 0; k < N; ++k) {
        double v = up ? fmin(x0[3] + a.ramp * a.dt * (k + 1), a.target_vel) : fmax(x0[3] - a.ramp * a.dt * (k + 1), a.target_vel);
        cs += v * a.dt;
        for (int i = 0; i < NX; ++i) xr[k * NX + i] = 0.0;
        xr[k * NX + 3] = v;
        xr[k * NX + 0] = x0[0] + cs;
    }
}

}  // namespace fsae
