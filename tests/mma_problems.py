"""Test problems for MMA (shared by the oracle tests and the GPU parity tests)."""
import numpy as np


def svanberg_toy():
    """K. Svanberg's 3-variable, 2-constraint "toy problem" distributed with his MMA/GCMMA codes (toy2.m): known optimum
    x* = (2.0175, 1.7800, 1.2375), f0* = 8.7702 -- a published known answer for any MMA implementation (the 8-digit value
    returned here was cross-checked with scipy SLSQP, see tests/test_oracle_mma_lbl.py)."""
    def f(x):
        return np.array([x[0] ** 2 + x[1] ** 2 + x[2] ** 2,
                         (x[0] - 5) ** 2 + (x[1] - 2) ** 2 + (x[2] - 1) ** 2 - 9,
                         (x[0] - 3) ** 2 + (x[1] - 4) ** 2 + (x[2] - 3) ** 2 - 9])

    def df(x):
        return np.array([[2 * x[0], 2 * x[1], 2 * x[2]],
                         [2 * (x[0] - 5), 2 * (x[1] - 2), 2 * (x[2] - 1)],
                         [2 * (x[0] - 3), 2 * (x[1] - 4), 2 * (x[2] - 3)]])
    return 3, 2, np.zeros(3), 5 * np.ones(3), f, df, np.array([4.0, 3.0, 2.0]), np.array([2.01751859, 1.78001142, 1.23750717])


def compliance_like(n, seed=0):
    """Separable topopt-like problem: minimise sum c_j / (eps + x_j^3) subject to mean(x) <= V (m = 1), the structure of
    the reference's drivers (python/LayerByLayerOptimization.py:88-139: one volume constraint)."""
    rng = np.random.default_rng(seed)
    c = rng.uniform(0.5, 2.0, n)
    V = 0.4

    def f(x):
        return np.array([np.sum(c / (1e-3 + x ** 3)) / n, np.mean(x) / V - 1.0])

    def df(x):
        return np.stack([-3 * c * x ** 2 / (1e-3 + x ** 3) ** 2 / n, np.full(n, 1.0 / (V * n))])
    return n, 1, np.full(n, 1e-3), np.ones(n), f, df, np.full(n, V), None
