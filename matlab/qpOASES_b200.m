function [x, fval, exitflag, iter, lambda, auxOutput] = qpOASES_b200(H, g, A, lb, ub, lbA, ubA, options, auxInput) %#ok<INUSD>
%QPOASES_B200 Drop-in for optimizers/matlab/qpOASES/qpOASES.m:22-24 on a B200:
%
%    [x,fval,exitflag,iter,lambda,auxOutput] = qpOASES_b200(H,g,A,lb,ub,lbA,ubA)
%    [x,fval,exitflag,iter,lambda,auxOutput] = qpOASES_b200(H,g,lb,ub)            (no general constraints)
%
%same arguments, same outputs, same encodings (exitflag 0 / 1 / -1 / -2 / -3, lambda in qpOASES's sign
%convention, auxOutput.workingSetB / workingSetC with -1 / 0 / +1, qpOASES.m:40-60), solved by the batched
%dual active-set kernel of the fsae_mpc_b200 library (fsae_qpoases_host, nV <= 191: every QP the reference forms up to horizon 80).  This is the literal
%replacement for the call in mpc/ltv/kinematic/ltvmpc_kinetmatic_curvilinear.m:52 and
%mpc/ltv/dynamic/ltvmpc_dynamic_curvilinear.m:52; the fused drop-ins ltvmpc_*_curvilinear_b200.m never form H.
%
%BATCH: H [nV x nV x B], A [nC x nV x B], vectors [nV x B] / [nC x B] solve B independent QPs in one call.
%qpOASES's own "sequence of QPs" form (one H and A, matrices of vectors, qpOASES.m:62-64) is accepted too: H and
%A are then replicated.  `options` / `auxInput` are accepted for call compatibility and ignored: the reference
%calls with qpOASES's defaults (see DESIGN.md "qpOASES options the parity argument relies on").
    if nargin == 4 || (nargin >= 4 && nargin <= 6 && ~isnumeric(ub))      % qpOASES(H,g,lb,ub{,options{,auxInput}})
        ub = lb; lb = A; A = []; lbA = []; ubA = [];
    end
    nV = size(g, 1);
    B = size(g, 2);
    if size(H, 3) == 1 && B > 1, H = repmat(H, 1, 1, B); end
    if ~isempty(A) && size(A, 3) == 1 && B > 1, A = repmat(A, 1, 1, B); end
    if isempty(lb), lb = -inf(nV, B); end
    if isempty(ub), ub = inf(nV, B); end
    h = fsae_mpc_b200_handle();
    [x, fval, exitflag, iter, lambda, wsB, wsC] = fsae_mpc_b200_mex('qpoases', h, full(H), g, full(A), lb, ub, lbA, ubA);
    exitflag = double(exitflag);
    iter = double(iter);
    auxOutput = struct('workingSetB', double(wsB), 'workingSetC', double(wsC), 'cpuTime', NaN);
end
