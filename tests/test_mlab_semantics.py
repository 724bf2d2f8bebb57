"""Hand-computed MATLAB answers for the language semantics the parity fixtures rely on.  The
interpreter (oracle/mlab) and the numpy oracle were written by the same builder, so a shared
misreading of MATLAB would go unnoticed by comparing them with each other; every expected value
below is a literal worked out from the MATLAB language reference, not computed by either."""
import math
import os
import tempfile

import numpy as np
import pytest

from oracle.mlab.interp import Matlab


@pytest.fixture(scope="module")
def ml():
    d = tempfile.mkdtemp()
    src = {
        "t_mod": "function r = t_mod(a, b)\n r = mod(a, b);\nend\n",
        "t_len": "function r = t_len()\n A = zeros(3, 7); B = zeros(7, 2); r = [length(A), length(B), length(zeros(1,5)), numel(A)];\nend\n",
        "t_rep": "function r = t_rep()\n A = [1 2; 3 4]; R = repmat(A, 2, 1); r = R(:)';\nend\n",
        "t_colon": "function r = t_colon()\n A = [1 2 3; 4 5 6]; r = A(:)';\nend\n",
        "t_range": "function r = t_range()\n r = 0.05:0.05:0.2;\nend\n",
        "t_range2": "function r = t_range2()\n r = [numel(1:0), numel(5:-1:3), numel(0:0.1:0.3)];\nend\n",
        "t_reshape": "function r = t_reshape()\n r = reshape(1:6, 2, 3);\nend\n",
        "t_end": "function r = t_end()\n v = [10 20 30 40]; A = [1 2 3; 4 5 6]; r = [v(end), v(end-1), A(end, 1), A(1, end), A(end)];\nend\n",
        "t_minmax": "function r = t_minmax()\n r = [min([3 1 2]), max(min(5, 2), -1), min([4 9], 6)];\nend\n",
        "t_cumsum": "function r = t_cumsum()\n r = cumsum([1 2 3 4] * 0.5);\nend\n",
        "t_prec": "function r = t_prec()\n r = [-2^2, 2^-1, 1:3 + 1, ~0 + 1];\nend\n",
        "t_div": "function r = t_div()\n A = [2 0; 0 4]; b = [2; 4]; r = [(A \\ b)', ([2 4] / A)];\nend\n",
        "t_step": "function r = t_step()\n x = zeros(1, 10); x(2:3:end) = 1; r = x;\nend\n",
        "t_idx": "function r = t_idx()\n x = 1:12; y = x([1:4:end; 2:4:end]); r = y(:)';\nend\n",
        "t_cell": "function r = t_cell()\n c = {5, 7}; d = c; d{1} = d{1} + c{2}; r = [c{1}, d{1}, d{2}];\nend\n",
        "t_str": "function r = t_str()\n M = \"DYNAMIC\"; r = [M == \"DYNAMIC\", M == \"KINEMATIC\"];\nend\n",
        "t_grow": "function r = t_grow()\n for i = 1:3\n  q(i) = i * i;\n end\n r = q;\nend\n",
        "t_ang": "function r = t_ang()\n r = [angdiff(0.1, 0.3), angdiff(3, -3), angdiff(-3, 3), angdiff(0, 7)];\nend\n",
        "t_norm": "function r = t_norm()\n r = [norm([3; 4]), dot([1 2 3], [4; 5; 6])];\nend\n",
        "t_diagq": "function r = t_diagq()\n q = [1; 2]; Q = spdiags(repmat(q, 2, 1), 0, 4, 4); r = full(Q);\nend\n",
    }
    for k, v in src.items():
        with open(os.path.join(d, k + ".m"), "w") as f:
            f.write(v)
    return Matlab([d])


def test_mod_follows_the_sign_of_the_divisor(ml):
    # MATLAB: mod(-1,3) = 2, mod(1,-3) = -2, mod(5,3) = 2, mod(-7.5, 2) = 0.5, mod(x,0) n/a
    assert ml.call("t_mod", -1.0, 3.0).item() == 2.0
    assert ml.call("t_mod", 1.0, -3.0).item() == -2.0
    assert ml.call("t_mod", 5.0, 3.0).item() == 2.0
    assert ml.call("t_mod", -7.5, 2.0).item() == 0.5


def test_length_is_the_largest_dimension(ml):
    assert ml.call("t_len").ravel().tolist() == [7.0, 7.0, 5.0, 21.0]


def test_repmat_and_colon_are_column_major(ml):
    # repmat([1 2;3 4],2,1) = [1 2;3 4;1 2;3 4]; (:) stacks columns
    assert ml.call("t_rep").ravel().tolist() == [1, 3, 1, 3, 2, 4, 2, 4]
    assert ml.call("t_colon").ravel().tolist() == [1, 4, 2, 5, 3, 6]
    assert ml.call("t_reshape").tolist() == [[1, 3, 5], [2, 4, 6]]
    assert ml.call("t_idx").ravel().tolist() == [1, 2, 5, 6, 9, 10]


def test_ranges(ml):
    r = ml.call("t_range").ravel()
    assert r.shape == (4,) and abs(r[-1] - 0.2) < 1e-15 and r[0] == 0.05
    assert ml.call("t_range2").ravel().tolist() == [0, 3, 4]
    assert ml.call("t_step").ravel().tolist() == [0, 1, 0, 0, 1, 0, 0, 1, 0, 0]


def test_end_min_max_cumsum_precedence(ml):
    assert ml.call("t_end").ravel().tolist() == [40, 30, 4, 3, 6]
    assert ml.call("t_minmax").ravel().tolist() == [1, 2, 4, 6]
    assert ml.call("t_cumsum").ravel().tolist() == [0.5, 1.5, 3.0, 5.0]
    # -2^2 = -4; 2^-1 = 0.5; 1:3+1 = 1:4; ~0+1 = 2
    assert ml.call("t_prec").ravel().tolist() == [-4, 0.5, 1, 2, 3, 4, 2]


def test_matrix_division(ml):
    assert ml.call("t_div").ravel().tolist() == [1, 1, 1, 1]


def test_cells_strings_growth(ml):
    assert ml.call("t_cell").ravel().tolist() == [5, 12, 7]          # value semantics: c unchanged
    assert ml.call("t_str").ravel().tolist() == [1, 0]
    assert ml.call("t_grow").ravel().tolist() == [1, 4, 9]


def test_angdiff_wraps_to_pi(ml):
    r = ml.call("t_ang").ravel()
    assert abs(r[0] - 0.2) < 1e-15
    assert abs(r[1] - (-6 + 2 * math.pi)) < 1e-15                      # -3 - 3 = -6 -> +0.283..
    assert abs(r[2] - (6 - 2 * math.pi)) < 1e-15
    assert abs(r[3] - (7 - 2 * math.pi)) < 1e-15


def test_norm_dot_spdiags(ml):
    assert ml.call("t_norm").ravel().tolist() == [5.0, 32.0]
    assert ml.call("t_diagq").tolist() == np.diag([1.0, 2.0, 1.0, 2.0]).tolist()
