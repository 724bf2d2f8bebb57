/* fsae_mpc_b200_mex.c -- thin MEX gateway from the reference's MATLAB host code to the
 * C-ABI in include/fsae_mpc_b200.h.  One gateway, command string first (the same shape as
 * the reference's qpOASES_sequence MEX, optimizers/matlab/qpOASES/qpOASES_sequence.m):
 *
 *   h = fsae_mpc_b200_mex('create', device)
 *   fsae_mpc_b200_mex('set_track', h, track_id, x_spline, y_spline, dl)
 *   [u_opt, x_opt, exitflag, fval, slack_opt, iters] = fsae_mpc_b200_mex('ltvmpc', h, model,
 *           x0, x_ref, dt, x_lin, u_lin [, track_id, param_id])
 *   fsae_mpc_b200_mex('destroy', h)
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

static fsae_ctx* ctx_of(const mxArray* a) {
    if (!mxIsUint64(a) || mxGetNumberOfElements(a) != 1) mexErrMsgTxt("fsae_mpc_b200: bad handle");
    return (fsae_ctx*)(uintptr_t)(*(uint64_t*)mxGetData(a));
}

static void check(fsae_ctx* c, int rc, const char* what) {
    if (rc != FSAE_OK) {
        char msg[512];
        snprintf(msg, sizeof(msg), "fsae_mpc_b200 %s failed (%d): %s", what, rc, c ? fsae_last_error(c) : "");
        mexErrMsgTxt(msg);
    }
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
        fsae_ctx* c = NULL;
        const int dev = nrhs > 1 ? (int)mxGetScalar(prhs[1]) : 0;
        check(NULL, fsae_create(&c, dev), "create (needs an sm_100 GPU; there is no CPU fallback)");
        plhs[0] = mxCreateNumericMatrix(1, 1, mxUINT64_CLASS, mxREAL);
        *(uint64_t*)mxGetData(plhs[0]) = (uint64_t)(uintptr_t)c;
        return;
    }
    if (nrhs < 2) mexErrMsgTxt("fsae_mpc_b200: handle expected");
    fsae_ctx* c = ctx_of(prhs[1]);

    if (!strcmp(cmd, "destroy")) {
        check(c, fsae_destroy(c), "destroy");
    } else if (!strcmp(cmd, "set_track")) {
        /* (h, track_id, x_spline [n x 4], y_spline [n x 4], dl)  -- main.m:15-17 outputs */
        if (nrhs != 6) mexErrMsgTxt("set_track: 5 arguments");
        const int n = (int)mxGetM(prhs[3]);
        check(c, fsae_set_track(c, (int)mxGetScalar(prhs[2]), mxGetPr(prhs[3]), mxGetPr(prhs[4]), n,
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
        check(c, fsae_ltvmpc_host(c, model, B, N, mxGetScalar(prhs[5]),
                                  ids_of(nrhs > 8 ? prhs[8] : NULL, B), ids_of(nrhs > 9 ? prhs[9] : NULL, B),
                                  mxGetPr(prhs[3]), mxGetPr(prhs[4]), mxGetPr(prhs[6]), mxGetPr(prhs[7]),
                                  mxGetPr(plhs[0]), mxGetPr(x_opt), (int32_t*)mxGetData(ef), mxGetPr(fv),
                                  mxGetPr(sl), (int32_t*)mxGetData(it), NULL, NULL), "ltvmpc");
        if (nlhs > 1) plhs[1] = x_opt;
        if (nlhs > 2) plhs[2] = ef;
        if (nlhs > 3) plhs[3] = fv;
        if (nlhs > 4) plhs[4] = sl;
        if (nlhs > 5) plhs[5] = it;
    } else {
        mexErrMsgTxt("fsae_mpc_b200: unknown command");
    }
}
