"""CPU check of the algorithm of the diagonal-block factorization kernel (k_potrf_inv_small, voxelfem_b200/csrc/vf_dense.cu): the
elimination on [A | I] WITHOUT normalising the pivot rows, on the lower triangle only, followed by one scaling with 1 / sqrt(d),
restated in numpy step for step and compared with numpy's Cholesky factor and its inverse.  (The kernel replaces the coarsest-level
CHOLMOD factorization of the reference, TensorProductSimulator.hh:1198-1230 / SparseMatrices.hh:1984-2131.)"""
import numpy as np
import pytest


def potrf_inv_unnormalised(A):
    m = A.shape[0]
    L = np.tril(A).astype(float)          # left half: lower triangle of A, becomes the columns of Lt D
    X = np.eye(m)                         # right half: becomes Lt^-1 (unit lower triangular)
    piv = np.ones(m)
    for j in range(m):
        col = L[:, j].copy()              # column j of the left half as published to shared memory
        rowx = X[j, :].copy()             # row j of the right half
        a = col[j]
        assert a > 0.0, "not positive definite"
        piv[j] = a
        rd = 1.0 / a
        for i in range(j + 1, m):
            f = -col[i] * rd              # -a_ij / d_j
            for c in range(j + 1, i + 1): # columns right of j, lower triangle only: a_ic -= a_ij a_cj / d_j
                L[i, c] += f * col[c]
            for c in range(0, j + 1):     # columns <= j of the inverse half: x_ic -= a_ij x_jc / d_j
                X[i, c] += f * rowx[c]
    rs = 1.0 / np.sqrt(piv)
    return np.tril(L) * rs[None, :], np.tril(X) * rs[:, None]


@pytest.mark.parametrize("m", [1, 2, 7, 33, 64])
def test_unnormalised_elimination_gives_cholesky_and_inverse(m):
    rng = np.random.default_rng(m)
    B = rng.normal(size=(m, m + 3))
    A = B @ B.T + 0.1 * np.eye(m)
    L, X = potrf_inv_unnormalised(A)
    Lref = np.linalg.cholesky(A)
    assert np.abs(L - Lref).max() < 1e-12 * np.abs(Lref).max()
    assert np.abs(X @ Lref - np.eye(m)).max() < 1e-10
    assert np.abs(X.T @ X @ A - np.eye(m)).max() < 1e-8      # L^-T L^-1 = A^-1: what the coarse solve applies
    assert np.all(np.triu(L, 1) == 0) and np.all(np.triu(X, 1) == 0)


def test_reciprocal_refinement_reaches_double_precision():
    """pivot_reciprocal: a 20-bit seed and one cubically convergent step y (1 + e + e^2), e = 1 - a y."""
    rng = np.random.default_rng(0)
    a = np.exp(rng.uniform(-20, 20, 10000))
    y = (1.0 / a).astype(np.float32).astype(np.float64) * (1 + rng.uniform(-1, 1, a.size) * 2.0 ** -20)   # seed no better than 2^-20
    e = 1.0 - a * y
    y = y + y * (e + e * e)
    assert np.abs(a * y - 1.0).max() < 4e-16
