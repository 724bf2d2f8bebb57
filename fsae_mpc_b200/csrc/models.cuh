// Device-side vehicle models and spline curvature for the batched LTV-MPC step.
//
// Reference behaviour reproduced here (NOT ported code -- the reference is MATLAB):
//   spline/interpolate_curvature.m:1-20, interpolate_spline_d.m, interpolate_spline_dd.m
//   vehicle_models/curvilinear_kinematic/{f,A,B}_curv_kin.m
//   vehicle_models/curvilinear_dynamic/{f,A,B}_curv_dyn.m
//   mpc/ltv/{kinematic,dynamic}/{euler,rk2,rk4}_*_curvilinear.m
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include "../../include/fsae_mpc_b200.h"

namespace fsae {

// Track in device memory: per segment 8 doubles [x P0..P3 | y P0..P3] so one segment's
// control values are one 64-byte line.
struct DevTrack {
    const double* coef;
    int n_seg;
    double dl;
};

// spline/interpolate_curvature.m:12-18 with the segment lookup of interpolate_spline_d.m:11-14
__device__ __forceinline__ double curvature(const DevTrack& tr, double s) {
    const double period = tr.dl * (double)tr.n_seg;
    double t = s - floor(s / period) * period;          // mod(t, dl*length(P))
    if (t < 0.0) t += period;
    if (t >= period) t -= period;
    int i = (int)floor(t / tr.dl);
    i = i < 0 ? 0 : (i >= tr.n_seg ? tr.n_seg - 1 : i);
    const double u = t / tr.dl - (double)i;
    const double* c = tr.coef + 8 * i;
    const double omu = 1.0 - u;
    const double b0 = -3.0 * omu * omu, b1 = 3.0 * (3.0 * u * u - 4.0 * u + 1.0);
    const double b2 = 3.0 * (2.0 * u - 3.0 * u * u), b3 = 3.0 * u * u;
    const double e0 = 6.0 * omu, e1 = 6.0 * (3.0 * u - 2.0), e2 = 6.0 * (1.0 - 3.0 * u), e3 = 6.0 * u;
    const double Xd = (b0 * c[0] + b1 * c[1] + b2 * c[2] + b3 * c[3]) / tr.dl;
    const double Yd = (b0 * c[4] + b1 * c[5] + b2 * c[6] + b3 * c[7]) / tr.dl;
    const double Xdd = (e0 * c[0] + e1 * c[1] + e2 * c[2] + e3 * c[3]) / (tr.dl * tr.dl);
    const double Ydd = (e0 * c[4] + e1 * c[5] + e2 * c[6] + e3 * c[7]) / (tr.dl * tr.dl);
    const double v2 = Xd * Xd + Yd * Yd;
    return (Xd * Ydd - Xdd * Yd) / (v2 * sqrt(v2));
}

// ------------------------------------------------------------------ kinematic bicycle
struct KinModel {
    struct Aux {};                          // (nothing of the model evaluation is reused by the constraint linearisation)
    static constexpr int NX = 5;
    static constexpr int NU = 2;
    static constexpr int NS = 1;            // slack variables
    static constexpr int DEFAULT_LIN = FSAE_LIN_RK2;

    // f_curv_kin.m:17-29 and A_curv_kin.m:19-55 in one pass (shared sub-expressions).
    // A is row-major NX x NX.  B is constant (B_curv_kin.m:12-16): columns e_3, e_4.
    __device__ static void eval(const double* x, const double* u, const DevTrack& tr,
                                const fsae_params& p, double* f, double* A) {
        const double lr = p.lr, lf = p.lf;
        const double lr_ratio = lr / (lr + lf);
        const double k = curvature(tr, x[0]);
        const double tan_d = tan(x[4]);
        const double lt = lr_ratio * tan_d;
        const double beta = atan(lt);
        double s_mb, c_mb, s_b, c_b;
        sincos(x[2] + beta, &s_mb, &c_mb);
        sincos(beta, &s_b, &c_b);
        const double cd = cos(x[4]);
        const double beta_d = lr_ratio * (1.0 / (cd * cd)) / (1.0 + lt * lt);
        const double denom = 1.0 / (1.0 - x[1] * k);
        const double v = x[3];

        f[0] = v * c_mb * denom;
        f[1] = v * s_mb;
        f[2] = v * s_b / lr - v * c_mb * denom * k;
        f[3] = u[0];
        f[4] = u[1];

        const double s_n = v * c_mb * denom * denom * k;
        const double s_mu = -v * s_mb * denom;
        const double s_v = c_mb * denom;
        const double s_delta = -v * s_mb * denom * beta_d;
#pragma unroll
        for (int i = 0; i < 25; ++i) A[i] = 0.0;
        A[0 * 5 + 1] = s_n;
        A[0 * 5 + 2] = s_mu;
        A[0 * 5 + 3] = s_v;
        A[0 * 5 + 4] = s_delta;
        A[1 * 5 + 2] = v * c_mb;
        A[1 * 5 + 3] = s_mb;
        A[1 * 5 + 4] = v * c_mb * beta_d;
        A[2 * 5 + 1] = -s_n * k;
        A[2 * 5 + 2] = -s_mu * k;
        A[2 * 5 + 3] = s_b / lr - s_v * k;
        A[2 * 5 + 4] = v * c_b * beta_d / lr - s_delta * k;
    }
    // column of the constant continuous-time B that control j drives
    __device__ __forceinline__ static void eval_first(const double* x, const double* u, const DevTrack& tr,
                                                      const fsae_params& p, double* f, double* A, Aux*) {
        eval(x, u, tr, p, f, A);
    }
    __device__ __forceinline__ static int bcol_state(int j) { return j == 0 ? 3 : 4; }
};

// ------------------------------------------------------------------ dynamic bicycle
struct DynAux {     // extra outputs of A_curv_dyn.m:1 used by the constraint linearisations
    double Fcr, Fcr_d, vr, denom_vr2, x_d_hat, x_d_hat_d, vf, denom_vf2;
};

struct DynModel {
    using Aux = DynAux;                     // tyre-force terms at the linearisation point, reused by Cons<DynModel>::step_coefs
    static constexpr int NX = 7;
    static constexpr int NU = 2;
    static constexpr int NS = 4;
    static constexpr int DEFAULT_LIN = FSAE_LIN_RK4;

    __device__ __forceinline__ static void pacejka(const fsae_params& p, double alpha, double Fz,
                                                   double& F, double& Fd) {
        const double B = p.pac_B, C = p.pac_C, D = p.pac_D, E = p.pac_E;
        const double bt = B * alpha;
        const double inner = bt - E * (bt - atan(bt));
        const double ang = C * atan(inner);
        double sn, cs;
        sincos(ang, &sn, &cs);
        F = Fz * D * sn;
        Fd = Fz * D * cs * C / (1.0 + inner * inner) * (B - E * (B - B / (1.0 + B * B * alpha * alpha)));
    }

    // f_curv_dyn.m:20-62 and A_curv_dyn.m:22-105
    __device__ static void eval_aux(const double* x, const double* u, const DevTrack& tr,
                                    const fsae_params& p, double* f, double* A, DynAux* aux) {
        const double m = p.mass, I = p.inertia, lr = p.lr, lf = p.lf, g = p.grav;
        const double s = x[0], n = x[1], mu = x[2], x_d = x[3], y_d = x[4], theta_d = x[5], delta = x[6];
        const double ex = exp(-x_d / 5.0);
        const double x_d_hat = x_d + 5.0 * ex;
        const double x_d_hat_d = 1.0 - ex;
        const double vf = (y_d + lf * theta_d) / x_d_hat;
        const double vr = (y_d - lr * theta_d) / x_d_hat;
        const double alpha_f = delta - atan(vf);
        const double alpha_r = -atan(vr);
        const double Fzf = m * g * lr / (lr + lf);
        const double Fzr = m * g * lf / (lr + lf);
        double Fcf, Fcf_d, Fcr, Fcr_d;
        pacejka(p, alpha_f, Fzf, Fcf, Fcf_d);
        pacejka(p, alpha_r, Fzr, Fcr, Fcr_d);
        const double k = curvature(tr, s);
        const double denom_nk = 1.0 / (1.0 - n * k);
        const double denom_vf2 = 1.0 / (1.0 + vf * vf);
        const double denom_vr2 = 1.0 / (1.0 + vr * vr);
        double sm, cm, sd, cd;
        sincos(mu, &sm, &cm);
        sincos(delta, &sd, &cd);
        if (f) {
            const double Fx = u[0] * m;
            f[0] = (x_d * cm - y_d * sm) * denom_nk;
            f[1] = x_d * sm + y_d * cm;
            f[2] = theta_d - (x_d * cm - y_d * sm) * denom_nk * k;
            f[3] = (Fx - Fcf * sd + m * y_d * theta_d) / m;
            f[4] = (Fcr + Fcf * cd - m * x_d * theta_d) / m;
            f[5] = (lf * Fcf * cd - lr * Fcr) / I;
            f[6] = u[1];
        }
        if (A) {
            const double s_n = (x_d * cm - y_d * sm) * denom_nk * denom_nk * k;
            const double s_mu = (-x_d * sm - y_d * cm) * denom_nk;
            const double s_xd = cm * denom_nk;
            const double s_yd = -sm * denom_nk;
#pragma unroll
            for (int i = 0; i < 49; ++i) A[i] = 0.0;
            A[0 * 7 + 1] = s_n;
            A[0 * 7 + 2] = s_mu;
            A[0 * 7 + 3] = s_xd;
            A[0 * 7 + 4] = s_yd;
            A[1 * 7 + 2] = x_d * cm - y_d * sm;
            A[1 * 7 + 3] = sm;
            A[1 * 7 + 4] = cm;
            A[2 * 7 + 1] = -s_n * k;
            A[2 * 7 + 2] = -s_mu * k;
            A[2 * 7 + 3] = -s_xd * k;
            A[2 * 7 + 4] = -s_yd * k;
            A[2 * 7 + 5] = 1.0;
            A[3 * 7 + 3] = -Fcf_d * denom_vf2 * vf * sd * x_d_hat_d / (m * x_d_hat);
            A[3 * 7 + 4] = (Fcf_d * denom_vf2 * sd / x_d_hat + m * theta_d) / m;
            A[3 * 7 + 5] = (Fcf_d * denom_vf2 * lf * sd / x_d_hat + m * y_d) / m;
            A[3 * 7 + 6] = (-Fcf * cd - Fcf_d * sd) / m;
            A[4 * 7 + 3] = (Fcr_d * denom_vr2 * vr * x_d_hat_d / x_d_hat
                            + Fcf_d * denom_vf2 * vf * cd * x_d_hat_d / x_d_hat - m * theta_d) / m;
            A[4 * 7 + 4] = (-Fcr_d * denom_vr2 / x_d_hat - Fcf_d * denom_vf2 / x_d_hat * cd) / m;
            A[4 * 7 + 5] = (Fcr_d * denom_vr2 * lr / x_d_hat - Fcf_d * denom_vf2 * lf / x_d_hat * cd
                            - m * x_d_hat) / m;
            A[4 * 7 + 6] = (-Fcf * sd + Fcf_d * cd) / m;
            A[5 * 7 + 3] = (lf * Fcf_d * denom_vf2 * vf * cd * x_d_hat_d / x_d_hat
                            - lr * Fcr_d * denom_vr2 * vr * x_d_hat_d / x_d_hat) / I;
            A[5 * 7 + 4] = (-lf * Fcf_d * denom_vf2 * cd / x_d_hat + lr * Fcr_d * denom_vr2 / x_d_hat) / I;
            A[5 * 7 + 5] = (-lf * Fcf_d * denom_vf2 * lf * cd / x_d_hat
                            - lr * Fcr_d * denom_vr2 * lr / x_d_hat) / I;
            A[5 * 7 + 6] = (-lf * Fcf * sd + lf * Fcf_d * cd) / I;
        }
        if (aux) {
            aux->Fcr = Fcr; aux->Fcr_d = Fcr_d; aux->vr = vr; aux->denom_vr2 = denom_vr2;
            aux->x_d_hat = x_d_hat; aux->x_d_hat_d = x_d_hat_d; aux->vf = vf; aux->denom_vf2 = denom_vf2;
        }
    }
    __device__ static void eval(const double* x, const double* u, const DevTrack& tr,
                                const fsae_params& p, double* f, double* A) {
        eval_aux(x, u, tr, p, f, A, nullptr);
    }
    // the evaluation AT the linearisation point (first stage of every RK scheme), handing out the shared terms
    __device__ static void eval_first(const double* x, const double* u, const DevTrack& tr,
                                      const fsae_params& p, double* f, double* A, Aux* aux) {
        eval_aux(x, u, tr, p, f, A, aux);
    }
    __device__ __forceinline__ static int bcol_state(int j) { return j == 0 ? 3 : 6; }
};

// ------------------------------------------------------------------ RK linearisation
// One horizon step: continuous-time (A, B, d) as {euler,rk2,rk4}_*_curvilinear.m return them.
// A row-major NX*NX, Bm row-major NX*NU.
template <class Model>
__device__ void linearise_step(int scheme, const double* x, const double* u, double dt,
                               const DevTrack& tr, const fsae_params& p,
                               double* A, double* Bm, double* d, typename Model::Aux* aux0 = nullptr) {
    constexpr int NX = Model::NX, NU = Model::NU;
    double f[NX];
    auto matmul_IpA = [&](const double* F, const double* K, double h, double* out) {
        // out = F * (I + K*h)
#pragma unroll
        for (int r = 0; r < NX; ++r)
#pragma unroll
            for (int c = 0; c < NX; ++c) {
                double acc = F[r * NX + c];
#pragma unroll
                for (int k = 0; k < NX; ++k) acc += F[r * NX + k] * (K[k * NX + c] * h);
                out[r * NX + c] = acc;
            }
    };
    auto ctrl_sens = [&](const double* F, const double* Kp, double h, double* out) {
        // out = Bc + F*Kp*h, Bc = constant continuous B (unit entries)
#pragma unroll
        for (int r = 0; r < NX; ++r)
#pragma unroll
            for (int c = 0; c < NU; ++c) {
                double acc = (r == Model::bcol_state(c)) ? 1.0 : 0.0;
#pragma unroll
                for (int k = 0; k < NX; ++k) acc += F[r * NX + k] * Kp[k * NU + c] * h;
                out[r * NU + c] = acc;
            }
    };
    if (scheme == FSAE_LIN_EULER) {
        Model::eval_first(x, u, tr, p, f, A, aux0);
#pragma unroll
        for (int r = 0; r < NX; ++r)
#pragma unroll
            for (int c = 0; c < NU; ++c) Bm[r * NU + c] = (r == Model::bcol_state(c)) ? 1.0 : 0.0;
    } else if (scheme == FSAE_LIN_RK2) {
        double k1[NX], A1[NX * NX], A2[NX * NX], xt[NX], B1[NX * NU];
        Model::eval_first(x, u, tr, p, k1, A1, aux0);
#pragma unroll
        for (int i = 0; i < NX; ++i) xt[i] = x[i] + k1[i] * dt / 2;
        Model::eval(xt, u, tr, p, f, A2);
        matmul_IpA(A2, A1, dt / 2, A);
#pragma unroll
        for (int r = 0; r < NX; ++r)
#pragma unroll
            for (int c = 0; c < NU; ++c) B1[r * NU + c] = (r == Model::bcol_state(c)) ? 1.0 : 0.0;
        ctrl_sens(A2, B1, dt / 2, Bm);
    } else {
        double k1[NX], k2[NX], k3[NX], k4[NX], xt[NX];
        double F1[NX * NX], F2[NX * NX], F3[NX * NX], F4[NX * NX], K2[NX * NX], K3[NX * NX], K4[NX * NX];
        double U1[NX * NU], U2[NX * NU], U3[NX * NU], U4[NX * NU];
        Model::eval_first(x, u, tr, p, k1, F1, aux0);
#pragma unroll
        for (int i = 0; i < NX; ++i) xt[i] = x[i] + k1[i] * dt / 2;
        Model::eval(xt, u, tr, p, k2, F2);
#pragma unroll
        for (int i = 0; i < NX; ++i) xt[i] = x[i] + k2[i] * dt / 2;
        Model::eval(xt, u, tr, p, k3, F3);
#pragma unroll
        for (int i = 0; i < NX; ++i) xt[i] = x[i] + k3[i] * dt;
        Model::eval(xt, u, tr, p, k4, F4);
#pragma unroll
        for (int i = 0; i < NX; ++i) f[i] = (k1[i] + 2 * k2[i] + 2 * k3[i] + k4[i]) / 6;
        matmul_IpA(F2, F1, dt / 2, K2);
        matmul_IpA(F3, K2, dt / 2, K3);
        matmul_IpA(F4, K3, dt, K4);
#pragma unroll
        for (int i = 0; i < NX * NX; ++i) A[i] = (F1[i] + 2 * K2[i] + 2 * K3[i] + K4[i]) / 6;
#pragma unroll
        for (int r = 0; r < NX; ++r)
#pragma unroll
            for (int c = 0; c < NU; ++c) U1[r * NU + c] = (r == Model::bcol_state(c)) ? 1.0 : 0.0;
        ctrl_sens(F2, U1, dt / 2, U2);
        ctrl_sens(F3, U2, dt / 2, U3);
        ctrl_sens(F4, U3, dt / 2, U4);   // reference rk4_*_curvilinear.m:52 uses dt/2 here too
#pragma unroll
        for (int i = 0; i < NX * NU; ++i) Bm[i] = (U1[i] + 2 * U2[i] + 2 * U3[i] + U4[i]) / 6;
    }
    // d = f - A x - B u   ({euler,rk2,rk4}_*_curvilinear.m last line of the loop)
#pragma unroll
    for (int r = 0; r < NX; ++r) {
        double acc = f[r];
#pragma unroll
        for (int c = 0; c < NX; ++c) acc -= A[r * NX + c] * x[c];
#pragma unroll
        for (int c = 0; c < NU; ++c) acc -= Bm[r * NU + c] * u[c];
        d[r] = acc;
    }
}

}  // namespace fsae
