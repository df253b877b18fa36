"""CPU tests of the host logic of the Newton-Krylov driver (no device needed)"""
import numpy as np


def test_krylov_basis_coeffs_least_squares():
    """krylov_solver.py:168-182: per (module, region) min || beta e1 - H y ||"""
    from nk_ooc_b200.solver import comp_krylov_basis_coeffs

    rng = np.random.default_rng(0)
    n_mod, j, R = 2, 3, 4
    h_mat = np.zeros((n_mod, j + 2, j + 1, R))
    for m in range(n_mod):
        for r in range(R):
            h = np.triu(rng.normal(size=(j + 2, j + 1)), -1)  # upper Hessenberg
            h_mat[m, :, :, r] = h
    beta = np.abs(rng.normal(size=(n_mod, R))) + 0.1
    coeff = comp_krylov_basis_coeffs(beta, h_mat)
    assert coeff.shape == (n_mod, j + 1, R)
    for m in range(n_mod):
        for r in range(R):
            h = h_mat[m, :, :, r]
            rhs = np.zeros(j + 2)
            rhs[0] = beta[m, r]
            want = np.linalg.solve(h.T @ h, h.T @ rhs)
            np.testing.assert_allclose(coeff[m, :, r], want, rtol=1e-9)
    # one iteration: y = beta h00 / (h00^2 + h10^2)
    h1 = np.zeros((1, 2, 1, 1))
    h1[0, :, 0, 0] = [3.0, 4.0]
    np.testing.assert_allclose(comp_krylov_basis_coeffs(np.array([[5.0]]), h1)[0, 0, 0], 5.0 * 3.0 / 25.0)
