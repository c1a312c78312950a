"""CPU check of the algorithm behind the residual-emitting sweep of the stored-stencil levels (k_stencil_tile<GS, RES>,
voxelfem_b200/csrc/vf_stencil.cu): an independent numpy restatement of the multicoloured block Gauss-Seidel sweep
(MultigridSolver.hh:347-378, 408-458) that, besides updating u, accumulates the residual by PUSHING -(K_ji)^T du_j to the
neighbours i visited earlier -- compared with the oracle's own sweep followed by its computeResidual (:527-541) on the oracle's
Galerkin stencils.  Pins the three facts the kernel relies on: the stencil symmetry K_ij = K_ji^T, the colour order (parity class
(x, y, z) -> 4 px + 2 py + pz, reversed for backward sweeps) and the treatment of partially constrained nodes."""
import os

import numpy as np
import pytest

from oracle import OracleMG, OracleSim

RNG = np.random.default_rng(11)


def _sim(ne, dom, bc, data_dir):
    N = len(ne)
    s = OracleSim(np.array(ne), np.zeros(N), np.array(dom))
    s.set_isotropic(1.0, 0.3)
    s.set_interp(0, 1.0, 1e-4, 3.0, 3.0)
    s.apply_bc_file(os.path.join(data_dir, "bcs", bc))
    s.set_densities(RNG.uniform(0.05, 1.0, int(np.prod(ne))))
    return s


def _sweep_with_pushed_residual(S, nn, dmask, u0, b, forward):
    """S: [node][3^N][N][N] stencil (slot = row-major offset in {-1,0,1}^N), nn: nodes per axis.  Returns (u, r)."""
    N = len(nn)
    u = u0.copy()
    r = np.full_like(u0, np.nan)
    strides = np.array([int(np.prod(nn[a + 1:])) for a in range(N)])
    offsets = [np.array(np.unravel_index(s, (3,) * N)) - 1 for s in range(3 ** N)]
    centre = (3 ** N) // 2
    full = (1 << N) - 1
    visited = np.zeros(u.shape[0], dtype=bool)
    colours = range(2 ** N) if forward else range(2 ** N - 1, -1, -1)
    for col in colours:
        par = [(col >> (N - 1 - a)) & 1 for a in range(N)]
        grids = np.meshgrid(*[np.arange(par[a], nn[a], 2) for a in range(N)], indexing="ij")
        coords = np.stack([g.ravel() for g in grids], axis=1)
        for c in coords:
            n = int(c @ strides)
            dm = int(dmask[n])
            nbrs = []
            Ku = np.zeros(N)
            for s, d in enumerate(offsets):
                q = c + d
                if np.any(q < 0) or np.any(q >= nn):
                    continue
                j = int(q @ strides)
                nbrs.append((s, j))
                Ku += S[n, s] @ u[j]
            rhs = b[n] - Ku
            M = S[n, centre]
            du = np.zeros(N)
            if dm == 0:
                du = np.linalg.solve(M, rhs)
            elif dm != full:   # point Gauss-Seidel over the free components, in sweep direction (:358-365)
                order = range(N) if forward else range(N - 1, -1, -1)
                for i in order:
                    if not (dm >> i) & 1:
                        du[i] = (rhs[i] - M[i] @ du) / M[i, i]
            u[n] = u[n] + du
            own = rhs - M @ du
            r[n] = [0.0 if (dm >> a) & 1 else own[a] for a in range(N)]
            for s, j in nbrs:
                if s != centre and visited[j]:
                    r[j] -= S[n, s].T @ du      # -(K_nj)^T du_n = -K_jn du_n by symmetry
            visited[n] = True
    bits = (dmask[:, None] >> np.arange(N)[None, :]) & 1
    r[bits == 1] = 0.0                          # the caller's mask pass (computeResidual zeroes Dirichlet components)
    return u, r


@pytest.mark.parametrize("ne,dom,bc,levels", [((16, 8), (2.0, 1.0), "mbb_N.bc", 2), ((8, 4, 4), (2.0, 1.0, 1.0), "3D/mbb_N.bc", 2),
                                               ((8, 8, 4), (1.0, 1.0, 0.5), "3D/cantilever_flexion_E.bc", 2)])
def test_pushed_residual_equals_compute_residual(ne, dom, bc, levels, data_dir):
    N = len(ne)
    om = OracleMG(_sim(ne, dom, bc, data_dir), levels)
    om.update_stiffness()
    for l in range(1, levels):
        nn = np.array(om.get_sim(l).nn)
        S = om.stencil(l)
        # symmetry of the Galerkin operator: slot (n -> n + d) is the transpose of slot (n + d -> n)
        strides = np.array([int(np.prod(nn[a + 1:])) for a in range(N)])
        for s in range(3 ** N):
            d = np.array(np.unravel_index(s, (3,) * N)) - 1
            so = int(np.ravel_multi_index(tuple(1 - d), (3,) * N))
            idx = np.indices(tuple(nn)).reshape(N, -1).T
            ok = np.all((idx + d >= 0) & (idx + d < nn), axis=1)
            i = idx[ok] @ strides
            j = (idx[ok] + d) @ strides
            scale = np.abs(S).max()
            assert np.abs(S[i, s] - np.swapaxes(S[j, so], 1, 2)).max() < 1e-12 * scale
        dm = om.get_sim(l).dirichlet_mask()
        bits = (dm[:, None] >> np.arange(N)[None, :]) & 1
        u0 = RNG.normal(size=(om.nn(l), N)); u0[bits == 1] = 0
        b = RNG.normal(size=u0.shape)
        scale = np.abs(om.residual(l, u0, b)).max()
        for fwd in (True, False):
            u, r = _sweep_with_pushed_residual(S, nn, dm, u0, b, fwd)
            uo = om.smooth(l, u0, b, fwd)
            assert np.abs(u - uo).max() < 1e-11 * np.abs(uo).max(), (l, fwd)
            assert np.isfinite(r).all()
            assert np.abs(r - om.residual(l, uo, b)).max() < 1e-11 * scale, (l, fwd)
