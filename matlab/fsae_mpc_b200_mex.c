/* fsae_mpc_b200_mex.c -- thin MEX gateway from the reference's MATLAB host code to the
 * C-ABI in include/fsae_mpc_b200.h.  One gateway, command string first (the same shape as
 * the reference's qpOASES_sequence MEX, optimizers/matlab/qpOASES/qpOASES_sequence.m):
 *
 *   h = fsae_mpc_b200_mex('create', device)
 *   fsae_mpc_b200_mex('set_track', h, track_id, x_spline, y_spline, dl)
 *   [u_opt, x_opt, exitflag, fval, slack_opt, iters] = fsae_mpc_b200_mex('ltvmpc', h, model,
 *           x0, x_ref, dt, x_lin, u_lin [, track_id, param_id])
 *   [A, B, d] = fsae_mpc_b200_mex('linearise', h, model, x_lin, u_lin, dt [, track_id, param_id])
 *   [H, f, xA, lbA, ubA, lb, ub, A_bar, B_bar, d_bar, const] = fsae_mpc_b200_mex('condense', h, model,
 *           x0, x_ref, dt, x_lin, u_lin [, track_id, param_id])
 *   [x, fval, exitflag, iter, lambda, workingSetB, workingSetC] = fsae_mpc_b200_mex('qpoases', h,
 *           H, g, A, lb, ub, lbA, ubA)                       (qpOASES.m:22; see matlab/qpOASES_b200.m)
 *   [u_opt, x_opt, exitflag, fval, slack_opt, iters] = fsae_mpc_b200_mex('sqp', h, model, n_sqp,
 *           x0, x_ref, dt, x_lin, u_lin [, track_id, param_id])
 *   [plant_final, steps, n_hist, plant_hist, exit_hist] = fsae_mpc_b200_mex('closed_loop', h, model,
 *           N_steps, dt, n_sim, target_vel, plant0, x_opt0, u_opt0 [, track_id, param_id])   (main.m:90-190)
 *   fsae_mpc_b200_mex('destroy', h)
 *
 * The handle is a DEVICE POOL (fsae_pool_create): h = fsae_mpc_b200_mex('create') takes every visible GPU,
 * 'create', d one device, 'create', [d0 d1 ..] a list.  'ltvmpc' splits the batch over the pool's devices from
 * this one host thread (fsae_ltvmpc_host_pool); the other commands run on the pool's first device.
 *
 * Batched arrays carry the batch as the TRAILING dimension, so MATLAB's column-major
 * storage is exactly the C-ABI layout and mxGetPr() pointers are passed straight through:
 *   x0 [N_x x B], x_ref/x_lin [N_x x N x B], u_lin [N_u x N x B]
 *   -> u_opt [N_u*N x B], x_opt [N_x*N x B], exitflag [1 x B], fval [1 x B], slack_opt [N_s x B]
 *
 * Build (on a machine with MATLAB):  mex -I../include fsae_mpc_b200_mex.c -L../fsae_mpc_b200 -lfsae_mpc_b200
 * This image has no MATLAB; tests/test_cabi.py compiles this file against matlab/mex_stub.h
 * (declarations only) to keep it honest.
 */
#include <stdint.h>
#include <string.h>
#include "mex.h"
#include "fsae_mpc_b200.h"

static fsae_pool* pool_of(const mxArray* a) {
    if (!mxIsUint64(a) || mxGetNumberOfElements(a) != 1) mexErrMsgTxt("fsae_mpc_b200: bad handle");
    return (fsae_pool*)(uintptr_t)(*(uint64_t*)mxGetData(a));
}

static void check(fsae_ctx* c, int rc, const char* what) {
    if (rc != FSAE_OK) {
        char msg[512];
        snprintf(msg, sizeof(msg), "fsae_mpc_b200 %s failed (%d): %s", what, rc, c ? fsae_last_error(c) : "");
        mexErrMsgTxt(msg);
    }
}

static void check_pool(fsae_pool* p, int rc, const char* what) {
    if (rc != FSAE_OK) {
        char msg[512];
        snprintf(msg, sizeof(msg), "fsae_mpc_b200 %s failed (%d): %s", what, rc, p ? fsae_pool_last_error(p) : "");
        mexErrMsgTxt(msg);
    }
}

/* model dimensions and the batch size of an [N_x x N x B] array */
static void dims_of(int model, const mxArray* x_traj, int* NX, int* NS, int* N, int* B) {
    const mwSize* dr = mxGetDimensions(x_traj);
    const int nd = (int)mxGetNumberOfDimensions(x_traj);
    *NX = model == FSAE_MODEL_DYNAMIC ? 7 : 5;
    *NS = model == FSAE_MODEL_DYNAMIC ? 4 : 1;
    *N = (int)dr[1];
    *B = nd > 2 ? (int)dr[2] : 1;
    if ((int)dr[0] != *NX) mexErrMsgTxt("fsae_mpc_b200: trajectory has the wrong N_x for this model");
}

static const int32_t* ids_of(const mxArray* a, int B) {
    if (!a || mxIsEmpty(a)) return NULL;
    if (!mxIsInt32(a) || (int)mxGetNumberOfElements(a) != B) mexErrMsgTxt("fsae_mpc_b200: ids must be int32 [B]");
    return (const int32_t*)mxGetData(a);
}

void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
    char cmd[32];
    if (nrhs < 1 || mxGetString(prhs[0], cmd, sizeof(cmd))) mexErrMsgTxt("fsae_mpc_b200: command string expected");

    if (!strcmp(cmd, "create")) {
        /* () all visible GPUs; (d) one device; ([d0 d1 ..]) a list of devices */
        fsae_pool* p = NULL;
        int devs[64], nd = 0;
        if (nrhs > 1 && !mxIsEmpty(prhs[1])) {
            nd = (int)mxGetNumberOfElements(prhs[1]);
            if (nd > 64) mexErrMsgTxt("create: at most 64 devices");
            for (int i = 0; i < nd; ++i) devs[i] = (int)mxGetPr(prhs[1])[i];
        }
        check_pool(NULL, fsae_pool_create(&p, nd ? devs : NULL, nd), "create (needs sm_100 GPUs; there is no CPU fallback)");
        plhs[0] = mxCreateNumericMatrix(1, 1, mxUINT64_CLASS, mxREAL);
        *(uint64_t*)mxGetData(plhs[0]) = (uint64_t)(uintptr_t)p;
        return;
    }
    if (nrhs < 2) mexErrMsgTxt("fsae_mpc_b200: handle expected");
    fsae_pool* pool = pool_of(prhs[1]);
    fsae_ctx* c = fsae_pool_ctx(pool, 0);

    if (!strcmp(cmd, "destroy")) {
        check_pool(pool, fsae_pool_destroy(pool), "destroy");
    } else if (!strcmp(cmd, "n_devices")) {
        plhs[0] = mxCreateDoubleMatrix(1, 1, mxREAL);
        mxGetPr(plhs[0])[0] = (double)fsae_pool_size(pool);
    } else if (!strcmp(cmd, "set_track")) {
        /* (h, track_id, x_spline [n x 4], y_spline [n x 4], dl)  -- main.m:15-17 outputs */
        if (nrhs != 6) mexErrMsgTxt("set_track: 5 arguments");
        const int n = (int)mxGetM(prhs[3]);
        check_pool(pool, fsae_pool_set_track(pool, (int)mxGetScalar(prhs[2]), mxGetPr(prhs[3]), mxGetPr(prhs[4]), n,
                                             mxGetScalar(prhs[5])), "set_track");
    } else if (!strcmp(cmd, "obtain_reference")) {
        /* (h, x [8*N_s], ds, N_s, t [N_s], s0 [1 x B], dt, N_t) -> x_ref [7 x N_t x B]  -- util/obtain_reference.m:1 */
        if (nrhs != 9) mexErrMsgTxt("obtain_reference: 8 arguments");
        const int N_s = (int)mxGetScalar(prhs[4]), N_t = (int)mxGetScalar(prhs[8]);
        const int B = (int)mxGetNumberOfElements(prhs[6]);
        mwSize dims[3] = {7, (mwSize)N_t, (mwSize)B};
        plhs[0] = mxCreateNumericArray(3, dims, mxDOUBLE_CLASS, mxREAL);
        check(c, fsae_obtain_reference_host(c, mxGetPr(prhs[2]), mxGetPr(prhs[5]), N_s, mxGetScalar(prhs[3]),
                                            mxGetPr(prhs[6]), B, mxGetScalar(prhs[7]), N_t, mxGetPr(plhs[0])),
              "obtain_reference");
    } else if (!strcmp(cmd, "ltvmpc")) {
        /* (h, model, x0, x_ref, dt, x_lin, u_lin [, track_id, param_id]) */
        if (nrhs < 8) mexErrMsgTxt("ltvmpc: 7+ arguments");
        const int model = (int)mxGetScalar(prhs[2]);
        const int NX = model == FSAE_MODEL_DYNAMIC ? 7 : 5, NU = 2, NS = model == FSAE_MODEL_DYNAMIC ? 4 : 1;
        const mwSize* dr = mxGetDimensions(prhs[4]);
        const int nd = (int)mxGetNumberOfDimensions(prhs[4]);
        const int N = (int)dr[1], B = nd > 2 ? (int)dr[2] : 1;
        if ((int)dr[0] != NX || (int)mxGetM(prhs[3]) != NX) mexErrMsgTxt("ltvmpc: x0 / x_ref have the wrong N_x");
        plhs[0] = mxCreateDoubleMatrix(NU * N, B, mxREAL);
        mxArray* x_opt = mxCreateDoubleMatrix(NX * N, B, mxREAL);
        mxArray* ef = mxCreateNumericMatrix(1, B, mxINT32_CLASS, mxREAL);
        mxArray* fv = mxCreateDoubleMatrix(1, B, mxREAL);
        mxArray* sl = mxCreateDoubleMatrix(NS, B, mxREAL);
        mxArray* it = mxCreateNumericMatrix(1, B, mxINT32_CLASS, mxREAL);
        check_pool(pool, fsae_ltvmpc_host_pool(pool, model, B, N, mxGetScalar(prhs[5]),
                                  ids_of(nrhs > 8 ? prhs[8] : NULL, B), ids_of(nrhs > 9 ? prhs[9] : NULL, B),
                                  mxGetPr(prhs[3]), mxGetPr(prhs[4]), mxGetPr(prhs[6]), mxGetPr(prhs[7]),
                                  mxGetPr(plhs[0]), mxGetPr(x_opt), (int32_t*)mxGetData(ef), mxGetPr(fv),
                                  mxGetPr(sl), (int32_t*)mxGetData(it), NULL, NULL), "ltvmpc");
        if (nlhs > 1) plhs[1] = x_opt;
        if (nlhs > 2) plhs[2] = ef;
        if (nlhs > 3) plhs[3] = fv;
        if (nlhs > 4) plhs[4] = sl;
        if (nlhs > 5) plhs[5] = it;
    } else if (!strcmp(cmd, "linearise")) {
        /* (h, model, x_lin [N_x x N x B], u_lin [N_u x N x B], dt [, track_id, param_id]) -> A, B, d
         * rk2_kinematic_curvilinear.m:1 / rk4_dynamic_curvilinear.m:1 (the scheme is a field of the parameter set) */
        if (nrhs < 6) mexErrMsgTxt("linearise: 5+ arguments");
        const int model = (int)mxGetScalar(prhs[2]);
        int NX, NS, N, B;
        dims_of(model, prhs[3], &NX, &NS, &N, &B);
        mwSize dA[4] = {(mwSize)NX, (mwSize)NX, (mwSize)N, (mwSize)B}, dB[4] = {(mwSize)NX, 2, (mwSize)N, (mwSize)B};
        mwSize dd[3] = {(mwSize)NX, (mwSize)N, (mwSize)B};
        mxArray* A = mxCreateNumericArray(4, dA, mxDOUBLE_CLASS, mxREAL);
        mxArray* Bm = mxCreateNumericArray(4, dB, mxDOUBLE_CLASS, mxREAL);
        mxArray* d = mxCreateNumericArray(3, dd, mxDOUBLE_CLASS, mxREAL);
        check(c, fsae_linearise_host(c, model, B, N, mxGetScalar(prhs[5]),
                                     ids_of(nrhs > 6 ? prhs[6] : NULL, B), ids_of(nrhs > 7 ? prhs[7] : NULL, B),
                                     mxGetPr(prhs[3]), mxGetPr(prhs[4]), mxGetPr(A), mxGetPr(Bm), mxGetPr(d)), "linearise");
        plhs[0] = A;
        if (nlhs > 1) plhs[1] = Bm;
        if (nlhs > 2) plhs[2] = d;
    } else if (!strcmp(cmd, "condense")) {
        /* (h, model, x0, x_ref, dt, x_lin, u_lin [, track_id, param_id]) -> the QP of ltvmpc_*_curvilinear.m:38-41 */
        if (nrhs < 8) mexErrMsgTxt("condense: 7+ arguments");
        const int model = (int)mxGetScalar(prhs[2]);
        int NX, NS, N, B;
        dims_of(model, prhs[4], &NX, &NS, &N, &B);
        const int nV = 2 * N + NS, nC = (model == FSAE_MODEL_DYNAMIC ? 20 : 6) * N, nXN = NX * N;
        mwSize d3[3];
        mxArray* o[11];
        d3[0] = nV; d3[1] = nV; d3[2] = B;   o[0] = mxCreateNumericArray(3, d3, mxDOUBLE_CLASS, mxREAL);   /* H */
        o[1] = mxCreateDoubleMatrix(nV, B, mxREAL);                                                          /* f */
        d3[0] = nC; d3[1] = nV;              o[2] = mxCreateNumericArray(3, d3, mxDOUBLE_CLASS, mxREAL);   /* xA */
        o[3] = mxCreateDoubleMatrix(nC, B, mxREAL);
        o[4] = mxCreateDoubleMatrix(nC, B, mxREAL);
        o[5] = mxCreateDoubleMatrix(nV, B, mxREAL);
        o[6] = mxCreateDoubleMatrix(nV, B, mxREAL);
        d3[0] = nXN; d3[1] = NX;             o[7] = mxCreateNumericArray(3, d3, mxDOUBLE_CLASS, mxREAL);   /* A_bar */
        d3[0] = nXN; d3[1] = nV;             o[8] = mxCreateNumericArray(3, d3, mxDOUBLE_CLASS, mxREAL);   /* B_bar */
        o[9] = mxCreateDoubleMatrix(nXN, B, mxREAL);
        o[10] = mxCreateDoubleMatrix(1, B, mxREAL);
        check(c, fsae_condense_host(c, model, B, N, mxGetScalar(prhs[5]),
                                    ids_of(nrhs > 8 ? prhs[8] : NULL, B), ids_of(nrhs > 9 ? prhs[9] : NULL, B),
                                    mxGetPr(prhs[3]), mxGetPr(prhs[4]), mxGetPr(prhs[6]), mxGetPr(prhs[7]),
                                    mxGetPr(o[0]), mxGetPr(o[1]), mxGetPr(o[2]), mxGetPr(o[3]), mxGetPr(o[4]), mxGetPr(o[5]),
                                    mxGetPr(o[6]), mxGetPr(o[7]), mxGetPr(o[8]), mxGetPr(o[9]), mxGetPr(o[10])), "condense");
        for (int i = 0; i < 11 && (i == 0 || i < nlhs); ++i) plhs[i] = o[i];
    } else if (!strcmp(cmd, "qpoases")) {
        /* (h, H [nV x nV x B], g [nV x B], A [nC x nV x B], lb, ub [nV x B], lbA, ubA [nC x B])
         * -> x [nV x B], fval, exitflag, iter [1 x B], lambda [(nV+nC) x B], workingSetB [nV x B], workingSetC [nC x B]
         * optimizers/matlab/qpOASES/qpOASES.m:22-24 */
        if (nrhs != 9) mexErrMsgTxt("qpoases: H, g, A, lb, ub, lbA, ubA expected");
        const int nV = (int)mxGetM(prhs[3]), B = (int)mxGetN(prhs[3]);
        const int nC = mxIsEmpty(prhs[4]) ? 0 : (int)mxGetDimensions(prhs[4])[0];
        mxArray* x = mxCreateDoubleMatrix(nV, B, mxREAL);
        mxArray* fv = mxCreateDoubleMatrix(1, B, mxREAL);
        mxArray* ef = mxCreateNumericMatrix(1, B, mxINT32_CLASS, mxREAL);
        mxArray* it = mxCreateNumericMatrix(1, B, mxINT32_CLASS, mxREAL);
        mxArray* lam = mxCreateDoubleMatrix(nV + nC, B, mxREAL);
        mxArray* wb = mxCreateNumericMatrix(nV, B, mxINT8_CLASS, mxREAL);
        mxArray* wc = mxCreateNumericMatrix(nC, B, mxINT8_CLASS, mxREAL);
        check(c, fsae_qpoases_host(c, B, nV, nC, mxGetPr(prhs[2]), mxGetPr(prhs[3]), nC ? mxGetPr(prhs[4]) : NULL,
                                   mxGetPr(prhs[5]), mxGetPr(prhs[6]), nC ? mxGetPr(prhs[7]) : NULL, nC ? mxGetPr(prhs[8]) : NULL,
                                   mxGetPr(x), mxGetPr(fv), (int32_t*)mxGetData(ef), (int32_t*)mxGetData(it), mxGetPr(lam),
                                   (int8_t*)mxGetData(wb), (int8_t*)mxGetData(wc)), "qpoases");
        plhs[0] = x;
        if (nlhs > 1) plhs[1] = fv;
        if (nlhs > 2) plhs[2] = ef;
        if (nlhs > 3) plhs[3] = it;
        if (nlhs > 4) plhs[4] = lam;
        if (nlhs > 5) plhs[5] = wb;
        if (nlhs > 6) plhs[6] = wc;
    } else if (!strcmp(cmd, "sqp")) {
        /* (h, model, n_sqp, x0, x_ref, dt, x_lin, u_lin [, track_id, param_id]): repeated relinearise + QP passes */
        if (nrhs < 9) mexErrMsgTxt("sqp: 8+ arguments");
        const int model = (int)mxGetScalar(prhs[2]), n_sqp = (int)mxGetScalar(prhs[3]);
        int NX, NS, N, B;
        dims_of(model, prhs[5], &NX, &NS, &N, &B);
        plhs[0] = mxCreateDoubleMatrix(2 * N, B, mxREAL);
        mxArray* x_opt = mxCreateDoubleMatrix(NX * N, B, mxREAL);
        mxArray* ef = mxCreateNumericMatrix(1, B, mxINT32_CLASS, mxREAL);
        mxArray* fv = mxCreateDoubleMatrix(1, B, mxREAL);
        mxArray* sl = mxCreateDoubleMatrix(NS, B, mxREAL);
        mxArray* it = mxCreateNumericMatrix(1, B, mxINT32_CLASS, mxREAL);
        check(c, fsae_ltvmpc_sqp_host(c, model, B, N, mxGetScalar(prhs[6]), n_sqp,
                                      ids_of(nrhs > 9 ? prhs[9] : NULL, B), ids_of(nrhs > 10 ? prhs[10] : NULL, B),
                                      mxGetPr(prhs[4]), mxGetPr(prhs[5]), mxGetPr(prhs[7]), mxGetPr(prhs[8]),
                                      mxGetPr(plhs[0]), mxGetPr(x_opt), (int32_t*)mxGetData(ef), mxGetPr(fv), mxGetPr(sl),
                                      (int32_t*)mxGetData(it)), "sqp");
        if (nlhs > 1) plhs[1] = x_opt;
        if (nlhs > 2) plhs[2] = ef;
        if (nlhs > 3) plhs[3] = fv;
        if (nlhs > 4) plhs[4] = sl;
        if (nlhs > 5) plhs[5] = it;
    } else if (!strcmp(cmd, "closed_loop")) {
        /* (h, model, N_steps, dt, n_sim, target_vel, plant0 [7 x B], x_opt0 [N_x*N x B], u_opt0 [N_u*N x B]
         *  [, track_id, param_id]) -> plant_final [7 x B], steps [1 x B], n_hist [n_sim x B],
         *  plant_hist [7 x n_sim x B], exit_hist [n_sim x B]            main.m:90-190 for B vehicles */
        if (nrhs < 10) mexErrMsgTxt("closed_loop: 9+ arguments");
        const int model = (int)mxGetScalar(prhs[2]), N = (int)mxGetScalar(prhs[3]), n_sim = (int)mxGetScalar(prhs[5]);
        const int B = (int)mxGetN(prhs[7]);
        mxArray* pf = mxCreateDoubleMatrix(7, B, mxREAL);
        mxArray* st = mxCreateNumericMatrix(1, B, mxINT32_CLASS, mxREAL);
        mxArray* nh = nlhs > 2 ? mxCreateDoubleMatrix(n_sim, B, mxREAL) : NULL;
        mwSize dh[3] = {7, (mwSize)n_sim, (mwSize)B};
        mxArray* ph = nlhs > 3 ? mxCreateNumericArray(3, dh, mxDOUBLE_CLASS, mxREAL) : NULL;
        mxArray* eh = nlhs > 4 ? mxCreateNumericMatrix(n_sim, B, mxINT32_CLASS, mxREAL) : NULL;
        check(c, fsae_closed_loop_host(c, model, B, N, mxGetScalar(prhs[4]), n_sim, mxGetScalar(prhs[6]),
                                       ids_of(nrhs > 10 ? prhs[10] : NULL, B), ids_of(nrhs > 11 ? prhs[11] : NULL, B),
                                       mxGetPr(prhs[7]), mxGetPr(prhs[8]), mxGetPr(prhs[9]),
                                       mxGetPr(pf), (int32_t*)mxGetData(st), nh ? mxGetPr(nh) : NULL, ph ? mxGetPr(ph) : NULL,
                                       eh ? (int32_t*)mxGetData(eh) : NULL), "closed_loop");
        plhs[0] = pf;
        if (nlhs > 1) plhs[1] = st;
        if (nlhs > 2) plhs[2] = nh;
        if (nlhs > 3) plhs[3] = ph;
        if (nlhs > 4) plhs[4] = eh;
    } else {
        mexErrMsgTxt("fsae_mpc_b200: unknown command");
    }
}
